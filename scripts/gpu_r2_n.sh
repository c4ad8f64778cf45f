#!/bin/bash
# bench.py under torchrun at N = $1 GPUs (default 8): one JSON line into gpurun_out/bench_${N}gpu.json
N=${1:-8}
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench N=$N rc=$?"
tail -n 4 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${N}gpu.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "dp_check", "strong_one_frame", "spiral_120")})
print(d["e2e"], d["train_step"]["ms_per_step"], d["clocks"])
PY
