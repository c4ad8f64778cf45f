#!/bin/bash
# quick pass: resampling parity tests + sample_pdf timing
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
for f in tests/test_gpu_rays_sampling.py tests/test_gpu_random_shapes.py tests/test_gpu_render.py; do
  name=$(basename "$f" .py)
  timeout 900 python -m pytest "$f" -m gpu -q --timeout 300 -p no:cacheprovider > "gpurun_out/$name.log" 2>&1
  echo "$name rc=$? $(tail -n 1 gpurun_out/$name.log)"
  grep -E "^(FAILED|ERROR)|^E  " "gpurun_out/$name.log" | head -n 20
done
python scripts/time_pdf.py 2>&1 | tail -n 1 | tee -a gpurun_out/time_pdf.txt
python scripts/time_pdf.py 2>&1 | tail -n 1 | tee -a gpurun_out/time_pdf.txt
