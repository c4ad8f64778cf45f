#!/bin/bash
# Round-2 GPU pass A: parity tests (one pytest process per file), smoke, default bench, sample_pdf variants.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
status=0
for f in tests/test_gpu_mlp.py tests/test_gpu_rays_sampling.py tests/test_gpu_composite.py tests/test_gpu_render.py tests/test_gpu_train.py tests/test_callers.py tests/test_gpu_effects.py tests/test_gpu_random_shapes.py tests/test_gpu_eager_baseline.py; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q --timeout 300 -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  rc=$?
  echo "$name rc=$rc $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt
  [ $rc -ne 0 ] && { status=1; grep -E "^(FAILED|ERROR)|Error|assert" "gpurun_out/$name.log" | head -n 30; }
done
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt; tail -n 6 gpurun_out/smoke.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench rc=$?" | tee -a gpurun_out/summary.txt
tail -c 6000 gpurun_out/bench_default.json; tail -n 8 gpurun_out/bench_default.err
python scripts/time_pdf.py 2>&1 | tee -a gpurun_out/time_pdf.txt
for mb in 8 10 12; do NERFW_PROFILE_LIB=1 NERFW_PDF_MINB=$mb python scripts/time_pdf.py 2>&1 | tee -a gpurun_out/time_pdf.txt; done
exit $status
