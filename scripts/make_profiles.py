"""Rebuild the tracked summaries under profiles/ from the CSV exports scripts/gpu_profile.sh left in gpurun_out/.

    python scripts/make_profiles.py [round_tag]      # default r01
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"


def run(script, *args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), *args], capture_output=True, text=True).stdout


def write(name, text):
    with open(os.path.join(PROF, name), "w") as f:
        f.write(text)
    print("wrote", name)


CMD = "`ncu --set full --clock-control none --import-source on -k regex:... python scripts/{}` (scripts/gpu_profile.sh)"
for mode in ("mixed", "bf16x3", "bf16"):
    raw = os.path.join(OUT, f"prof_frame_{mode}_raw.csv")
    if os.path.exists(raw):
        write(f"{TAG}_ncu_full_{mode}.md",
              f"# Round {int(TAG[1:])} (final kernels) — ncu --set full, 200-row crop of the 800x800 frame (160 000 rays; coarse pass 64 depths, fine pass the 128 new depths), {mode} mode\n\n"
              f"Command: {CMD.format('profile_frame.py --mode ' + mode + ' --rows 200')}.  The two mlp_tc_fwd launches = coarse / fine pass.\n\n"
              + run("ncu_summary.py", raw))
    src = os.path.join(OUT, f"prof_frame_{mode}_src_mlp_tc_fwd.csv")
    if os.path.exists(src):
        write(f"{TAG}_hot_frame_{mode}_mlp_tc_fwd.md",
              f"# Round {int(TAG[1:])} — hottest SASS instructions, mlp_tc_fwd (frame_{mode} capture)\n\n```\n" + run("ncu_hot.py", src, "40") + "```\n")
raw = os.path.join(OUT, "prof_train_raw.csv")
if os.path.exists(raw):
    write(f"{TAG}_ncu_full_train.md",
          f"# Round {int(TAG[1:])} — ncu --set full, one training step (4096 rays, 64+128 samples): backward kernels\n\n"
          f"Command: {CMD.format('profile_train.py bf16x3 1')}.\n\n" + run("ncu_summary.py", raw))
for k in ("pass1", "wgrad"):
    src = os.path.join(OUT, f"prof_train_src_{k}.csv")
    if os.path.exists(src):
        write(f"{TAG}_hot_train_{k}.md",
              f"# Round {int(TAG[1:])} — hottest SASS instructions, {k} (training-step capture)\n\n```\n" + run("ncu_hot.py", src, "40") + "```\n")
raw = os.path.join(OUT, "prof_pdf_raw.csv")
if os.path.exists(raw):
    write(f"{TAG}_ncu_full_sample_pdf.md",
          f"# Round {int(TAG[1:])} — ncu --set full, sample_pdf_kernel at 640 000 rays, 64 + 128\n\n"
          f"Command: {CMD.format('time_pdf.py')}.\n\n" + run("ncu_summary.py", raw))

# launch lists (whole bench command; one step = one whole frame)
def launch_table(csv_name, out_name, title, blurb):
    lst = os.path.join(OUT, csv_name)
    if not os.path.exists(lst):
        return
    rows = [r for r in csv.reader(open(lst)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        if len(name) > 80:
            name = name[:34] + ".." + name[-44:]
        unit = r[hdr.index("Metric Unit")]
        t = float(r[vi].replace(",", ""))
        t_ms = t / 1e6 if unit in ("ns", "nsecond") else t / 1e3 if unit in ("us", "usecond") else t if unit in ("ms", "msecond") else t * 1e3
        agg[name][0] += 1
        agg[name][1] += t_ms
    total = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    text = (f"# Round {int(TAG[1:])} -- {title}\n\n"
            f"`ncu --metrics gpu__time_duration.sum --clock-control none`, {n} launches, {total:.1f} ms of device time (cold-cache, serialised: compare shares).\n"
            f"{blurb}\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        text += f"| `{name}` | {c} | {t:.3f} | {100 * t / total:.2f} % |\n"
    write(out_name, text)


launch_table("launches_bench.csv", f"{TAG}_launches_bench.md",
             "ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --quick` (whole command)",
             "The command contains the timed frames, the e2e frames, the per-kernel roofline loops, the other-mode frames, the 157-chunk frame, "
             "the 256+512 stress call and the training steps.")
launch_table("launches_step.csv", f"{TAG}_launches_step.md",
             "ncu launch list of ONE step: one whole 800x800 frame, 64+128, default mode (`python scripts/profile_frame.py --mode mixed --rows 800`)",
             "Default path: coarse pass `mlp_tc_fwd_kernel<1,0,0>` (bf16x3, 64 depths, full outputs), resampling, fine pass `mlp_tc_fwd_kernel<0,1,0>` "
             "(fp16, the 128 new depths), merge, compositing.  bench.py's `roofline.share_of_step` is the live counterpart of the share column.")
