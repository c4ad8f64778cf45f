#!/bin/bash
# ncu evidence for profiles/: (1) launch list of bench.py itself, (2) --set full captures of every hot-path kernel on
# small fixed workloads (one 200-row crop of the frame per MLP mode, one training step).  Plain run first, ncu only if it
# exits 0 (B200_PROFILING.md).  Reports are exported to CSV on the box and deleted (gpurun_out/ may carry 64 MiB back).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export_rep () {  # $1 = report stem, $2.. = kernel regexes for source pages
  local stem=$1; shift
  ncu -i gpurun_out/$stem.ncu-rep --page raw --csv > gpurun_out/${stem}_raw.csv 2>/dev/null
  for k in "$@"; do
    ncu -i gpurun_out/$stem.ncu-rep --page source --csv --kernel-name regex:$k > gpurun_out/${stem}_src_$k.csv 2>/dev/null
  done
  rm -f gpurun_out/$stem.ncu-rep
}
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick"
$BENCH > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv $BENCH > gpurun_out/ncu_bench.log 2>&1
echo "bench launch list rc=$?"
# one step on its own (one whole 800x800 frame through the public API): the share of each kernel in a step
python scripts/profile_frame.py --mode mixed --rows 800 > gpurun_out/pf_step.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_step.csv python scripts/profile_frame.py --mode mixed --rows 800 > gpurun_out/ncu_step.log 2>&1
echo "step launch list rc=$?"
for MODE in mixed bf16; do
  python scripts/profile_frame.py --mode $MODE --rows 200 > gpurun_out/pf_$MODE.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'mlp_tc_fwd|composite_fwd|sample_pdf|stratified|raygen|normalize|merge_raw' \
      -o gpurun_out/prof_frame_$MODE -f python scripts/profile_frame.py --mode $MODE --rows 200 > gpurun_out/ncu_frame_$MODE.log 2>&1
  echo "frame $MODE rc=$?"
  export_rep prof_frame_$MODE mlp_tc_fwd
done
python scripts/time_pdf.py > gpurun_out/tp.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'sample_pdf' -c 2 \
    -o gpurun_out/prof_pdf -f python scripts/time_pdf.py > gpurun_out/ncu_pdf.log 2>&1
echo "pdf rc=$?"
export_rep prof_pdf sample_pdf
python scripts/profile_train.py mixed 1 > gpurun_out/pt.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'pass1|wgrad|composite_bwd|adam|mse|app_' \
    -o gpurun_out/prof_train -f python scripts/profile_train.py mixed 1 > gpurun_out/ncu_train_full.log 2>&1
echo "train rc=$?"
export_rep prof_train pass1 wgrad
python scripts/profile_train.py mixed 2 > gpurun_out/pt2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python scripts/profile_train.py mixed 2 > gpurun_out/ncu_train_l.log 2>&1
du -sh gpurun_out; ls -la gpurun_out | head -40
