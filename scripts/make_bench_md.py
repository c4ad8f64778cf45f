"""profiles/<tag>_bench.md from the bench.py JSON lines gpurun left in gpurun_out/ (bench_default.json = 1 GPU,
bench_{2,4,8}gpu.json = torchrun).      python scripts/make_bench_md.py [tag]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
files = [(1, "bench_default.json", "1 GPU, `python bench.py --steps 5 --warmup 3`")] + [
    (n, f"bench_{n}gpu.json", f"{n} GPUs, torchrun, `--steps 5 --warmup 3`") for n in (2, 4, 8)]
lines = {}
for n, f, _ in files:
    p = os.path.join(OUT, f)
    if os.path.exists(p):
        lines[n] = json.load(open(p))
o = [f"# Round {int(TAG[1:])} — bench.py on B200 (copies of gpurun_out/bench_*.json, final kernels of the round)", "",
     "Default `mixed` MLP mode with coarse re-use (192 MLP evaluations per ray), weak scaling; every line also carries the "
     "strong-scaling leg (one frame split over the ranks), the 120-frame spiral delivered to rank 0, the data-parallel "
     "gradient check and the training step.", "",
     "| GPUs | render Mrays/s (weak) | e2e Mrays/s | ms/frame-step | one frame, strong (ms / Mrays/s) | 120-frame spiral to rank 0 (s / Mrays/s) | train step ms | dp_check |",
     "|---|---|---|---|---|---|---|---|"]
STALE = {int(x) for x in os.environ.get("BENCH_MD_STALE", "").split(",") if x}   # lines measured with an earlier kernel build
for n, d in sorted(lines.items()):
    st, sp, dp = d.get("strong_one_frame", {}), d.get("spiral_120", {}), d.get("dp_check")
    o.append(f"| {n}{' (*)' if n in STALE else ''} | {d['value']:.3f} | {d['e2e']['value']:.3f} | {d['ms_per_step']:.1f} | {st.get('ms_per_frame', 0):.1f} / {st.get('value', 0):.2f} | "
             f"{sp.get('wall_s', 0):.2f} / {sp.get('value', 0):.2f} | {d['train_step']['ms_per_step']:.2f} | "
             + (f"{dp['rel_max_abs']:.1e}" if dp and 'rel_max_abs' in dp else "n/a") + " |")
if STALE:
    o += ["", "(*) measured before the mbarrier waits were given a suspend-time hint (forward tile 32.7k -> 29.4k cycles); not re-run."]
d = lines.get(1)
if d:
    r, ro = d["roofline"], d["roofline_other"]
    o += ["", "Single GPU, kernels on their own (`roofline`, `roofline_other` of the first line):", "",
          "| launch | kernel | ms | achieved | fraction of measured peak |", "|---|---|---|---|---|"]
    sm = r["step_mlp_launches"]
    o.append(f"| coarse pass (64 depths, bf16x3, full outputs) | `{r['kernel']}` | {r['kernel_ms']:.2f} | {r['achieved']:.0f} TFLOP/s algorithmic, {r['issued']:.0f} issued | {r['frac']:.3f} (issued {r['issued_frac']:.3f}) |")
    o.append(f"| both MLP launches of the step together | `mlp_tc_fwd_kernel<X3>` + `<F16>` | {sm['kernel_ms']:.2f} | {sm['achieved']:.0f} TFLOP/s algorithmic | {sm['frac']:.3f} (issued {sm['issued_frac']:.3f}) |")
    names = {"mlp_fine": "fine pass (128 new depths, fp16)", "mlp_two_pass_fine_192": "two-pass form fine launch (192 depths, fp16)",
             "mlp_bf16_192": "bf16 launch, 192 depths", "composite_fwd": "compositing fwd (192 samples)", "composite_bwd": "compositing bwd",
             "sample_pdf": "resampling as the step runs it (merged row + z_fine, 2 304 B/ray)",
             "sample_pdf_merged_row_only": "resampling, merged row only (1 792 B/ray)", "merge_raw": "merge of coarse and fine records"}
    for k, label in names.items():
        v = ro.get(k)
        if v:
            o.append(f"| {label} | `{v.get('kernel', k)}` | {v['kernel_ms']:.3f} | {v['achieved']:.0f} {v['unit']} | {v['frac']:.3f} |")
    om = d["other_modes"]
    parts = [f"{k} {v['value']:.2f} Mrays/s ({v['ms_per_step']:.1f} ms)" for k, v in om.items()]
    ch = om.get("chunked_4096_with_cpu_copy", {})
    cbc = ch.get("call_by_call", {})
    o += ["", "Other forms of the same frame on one GPU: " + "; ".join(parts) + ".  Chunk loop (the reference's 157 calls of 4096 rays with a "
          f"`.cpu()` per chunk): one library call per chunk, host enqueue {ch.get('host_us_per_call_enqueue', 0):.0f} us per call, runs "
          f"{[round(x) for x in ch.get('runs_ms', [])]} ms; the same loop call by call (9 library calls per chunk): "
          f"{cbc.get('ms_per_step', 0):.1f} ms, {cbc.get('host_us_per_call_enqueue', 0):.0f} us per call."]
for n, f, what in files:
    if n in lines:
        o += ["", f"## {f}  ({what})", "", "```json", json.dumps(lines[n], indent=1), "```"]
with open(os.path.join(ROOT, "profiles", f"{TAG}_bench.md"), "w") as fh:
    fh.write("\n".join(o) + "\n")
print("wrote", f"profiles/{TAG}_bench.md", "with", sorted(lines))
