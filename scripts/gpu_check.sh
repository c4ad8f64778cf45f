#!/bin/bash
# One GPU-box pass: parity tests (one pytest process per file so a faulting kernel cannot poison the others), smoke,
# bench in each MLP mode.  Everything is logged under gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
rm -f gpurun_out/parity_errors.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
status=0
for f in tests/test_gpu_mlp.py tests/test_gpu_rays_sampling.py tests/test_gpu_composite.py tests/test_gpu_render.py; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q --timeout 300 -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  rc=$?
  echo "$name rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 25 "gpurun_out/$name.log"
  [ $rc -ne 0 ] && status=1
done
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt; tail -n 6 gpurun_out/smoke.log
for mode in ${BENCH_MODES:-fp32 bf16x3 bf16}; do
  timeout 900 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 --mlp-mode $mode > "gpurun_out/bench_$mode.json" 2> "gpurun_out/bench_$mode.err"
  echo "bench $mode rc=$?" | tee -a gpurun_out/summary.txt
  tail -c 3000 "gpurun_out/bench_$mode.json"; tail -n 5 "gpurun_out/bench_$mode.err"
done
exit $status
