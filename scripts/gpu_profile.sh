#!/bin/bash
# ncu evidence for one whole-frame render: (1) every launch with its device time, (2) full-set capture of the MLP kernel
# and the HBM-bound kernels.  Plain run first; ncu only if it exits 0.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
MODE=${1:-bf16x3}
python scripts/profile_frame.py --mode $MODE > gpurun_out/profile_plain_$MODE.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$MODE.csv \
    python scripts/profile_frame.py --mode $MODE > gpurun_out/ncu_launches_$MODE.log 2>&1
echo "launch list rc=$?"
python scripts/profile_frame.py --mode $MODE --rows 200 > gpurun_out/profile_plain200_$MODE.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'mlp_tc_fwd|composite_fwd|sample_pdf|stratified|raygen' \
    -o gpurun_out/prof_$MODE -f python scripts/profile_frame.py --mode $MODE --rows 200 > gpurun_out/ncu_full_$MODE.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -20
