#!/bin/bash
# Round-2 GPU pass B (2 GPUs): the tests that failed in pass A, the bench under torchrun at N=2 (dp_check, spiral gather,
# strong-scaling leg), the sharded-render / data-parallel check, sample_pdf occupancy variants.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/summary_b.txt
for f in tests/test_gpu_render.py tests/test_gpu_train.py; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q --timeout 300 -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  echo "$name rc=$? $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary_b.txt
  grep -E "^(FAILED|ERROR)|^E  " "gpurun_out/$name.log" | head -n 20
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 3 --warmup 3 --cpu-rays 8000 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench2 rc=$?" | tee -a gpurun_out/summary_b.txt
tail -n 5 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_2gpu.json"))
    print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "dp_check", "strong_one_frame", "spiral_120")})
    print(d["train_step"])
    print({k: v["frac"] for k, v in d["roofline_other"].items()})
except Exception as e:
    print("bench2 parse failed", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/check_sharded.py 2>&1 | tail -n 3 | tee -a gpurun_out/summary_b.txt
for mb in 8 10 12; do NERFW_PROFILE_LIB=1 NERFW_PDF_MINB=$mb python scripts/time_pdf.py 2>&1 | tail -n 1 | tee -a gpurun_out/time_pdf.txt; done
python scripts/time_pdf.py 2>&1 | tail -n 1 | tee -a gpurun_out/time_pdf.txt
