"""Readable digest of one bench.py JSON line: python scripts/print_bench.py gpurun_out/bench_default.json"""
import json
import sys

d = json.load(open(sys.argv[1]))


def round(v, nd):   # noqa: A001  (4 significant digits, so that 4.4e-6 does not print as 0.0)
    return float(f"{v:.4g}")


print(f"value {d['value']:.3f} {d['unit']}  e2e {d['e2e']['value']:.3f}  ms/step {d['ms_per_step']:.2f}  n_gpus {d['n_gpus']}  "
      f"launches {d['gpu_launches']}  clocks {d['clocks']}")
r = d["roofline"]
print("roofline:", {k: (round(r[k], 4) if isinstance(r[k], float) else r[k]) for k in
                    ("launch", "kernel", "achieved", "frac", "issued_frac", "kernel_ms", "share_of_step")})
if "step_mlp_launches" in r:
    print("  step MLP launches:", {k: round(v, 4) for k, v in r["step_mlp_launches"].items() if isinstance(v, float)})
for k, v in d["roofline_other"].items():
    print(f"  {k}:", {kk: round(v[kk], 4) for kk in ("achieved", "frac", "kernel_ms") if kk in v})
print("train_step:", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["train_step"].items() if k != "backward"})
for k, v in d["other_modes"].items():
    print(f"  {k}:", {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk not in ("unit",)})
for k in ("stress_256_512", "config0_100x100", "eager_baseline", "strong_one_frame", "spiral_120", "dp_check", "cpu_baseline"):
    if d.get(k):
        print(f"{k}:", {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in d[k].items() if kk not in ("partition", "what", "kind", "note")})
p = d.get("parity_vs_oracle")
if p:
    print("parity:", {k: p[k] for k in ("rays", "rgb", "depth", "acc")})
    if "dense" in p:
        print("parity dense:", {k: p["dense"][k] for k in ("rays", "mean_acc", "rgb", "depth", "acc", "depth_fp32_kernel", "depth_bf16x3") if k in p["dense"]})
