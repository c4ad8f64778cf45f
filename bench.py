#!/usr/bin/env python
"""bench.py -- render throughput of the NeRF-W ray-marching hot path (BASELINE.json metric: render Mrays/s, 64+128).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mlp-mode mixed|bf16x3|fp16|bf16|fp32]

A step = one synthetic 800x800 view (640 000 rays) rendered coarse(64) + fine(64+128) with random-init NeRF-W weights
(BASELINE.json configs[1]).  `value` is timed with the rays already in HBM; `e2e` goes through the public API with the
rays in pinned HOST memory and the rgb/depth images read back to the host inside the timed region.
Under torchrun (N > 1) every rank renders its own view of the aligned spiral (weak scaling, no data-path collective);
time is the max over ranks.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200")
for _p in (PKG, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

H = W = 800
N_COARSE, N_IMPORTANCE = 64, 128
NEAR, FAR = 2.0, 6.0
FLOP_PER_SAMPLE = 1_063_936          # 2 x 531 968 MAC, un-padded
SAMPLES_PER_RAY = N_COARSE + (N_COARSE + N_IMPORTANCE)   # 64 coarse + 192 fine MLP evaluations
METRIC = "render Mrays/s (64+128 samples)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mlp-mode", default=os.environ.get("NERFW_MLP_MODE", "mixed"), choices=["mixed", "bf16x3", "fp16", "bf16", "fp32"],
                    help="mixed (default): bf16x3 coarse pass + fp16 fine pass, the cheapest arithmetic inside the fp32 parity bars")
    ap.add_argument("--cpu-rays", type=int, default=int(os.environ.get("NERFW_CPU_SAMPLE_RAYS", "24000")),
                    help="rays in the bounded CPU sample (cpu_baseline / --impl reference step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus: int):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, local, world


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms: float, world: int) -> float:
    if world == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def make_weights():
    import nerfw_oracle as orc
    sd = orc.make_state_dict(0)          # == torch.manual_seed(0); NeRF(Config()) of the reference
    emb = torch.randn(32)
    return sd, emb


def cpu_sample(sd, emb, n_rays: int, pose: np.ndarray):
    """The reference algorithm (oracle port: same torch-CPU ops, same cost) on a bounded sample of the SAME workload:
    `n_rays` rays from the centre rows of the 800x800 view, coarse 64 + fine 192, all host threads."""
    import nerfw_oracle as orc
    from nerfw.camera import blender_focal
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    focal = blender_focal(W)
    c2w = torch.from_numpy(pose)
    rows = max(1, n_rays // W)
    r0 = H // 2 - rows // 2
    # rays of the selected rows only (same formula as the full view; rays_for_view builds whole images)
    ro, rd = orc.rays_for_view(H, W, focal, c2w)
    o = ro[r0:r0 + rows].reshape(-1, 3)[:n_rays].contiguous()
    d = rd[r0:r0 + rows].reshape(-1, 3)[:n_rays].contiguous()
    torch.manual_seed(1)
    u = torch.rand(o.shape[0], N_IMPORTANCE)
    t0 = time.perf_counter()
    with torch.no_grad():
        rgb, depth, ex = orc.render_hier(sd, sd, o, d, NEAR, FAR, N_COARSE, N_IMPORTANCE, emb=emb, perturb=False, u_rand=u)
    dt = time.perf_counter() - t0
    return o.shape[0], dt, cores, (o, d, u, rgb, depth, ex["acc"])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from nerfw.camera import aligned_spiral_poses
    sd, emb = make_weights()
    pose = aligned_spiral_poses(120, 2, "x", "chair")[0]
    times = []
    rays_done = 0
    n = 0
    cores = 1
    rays_per_step = args.cpu_rays
    budget_s = 180.0          # the whole --steps K --warmup W run stays within a few minutes whatever K and W are
    for i in range(args.warmup + args.steps):
        n, dt, cores, _ = cpu_sample(sd, emb, rays_per_step, pose)
        if i >= args.warmup:
            times.append(dt)
            rays_done += n
        if i == 0:   # size the remaining steps from the measured rate (whole rows of the view, at least 2 400 rays)
            left = args.warmup + args.steps - 1
            if left > 0 and dt * left > budget_s:
                rays_per_step = max(2400, int(n * budget_s / (dt * left)) // 800 * 800)
    total = sum(times)
    val = rays_done / total / 1e6
    sample = (f"{n} rays of the 800x800 view per step (centre rows; {args.cpu_rays} in the first step), coarse 64 + fine 192, "
              "torch CPU ops == reference code path")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "800x800 view render, 64+128 samples, random-init NeRF-W (BASELINE.json configs[1])",
                   "rays_per_step": n, "note": "CPU reference path: rank 0 only, bounded sample per step (sized to ~3 min per run)"},
        "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bench_train_step(nerfw, sd, dev, world, mode, n_rays=4096, steps=6, warmup=2):
    """BASELINE.json configs[2]: 4096-ray batch per GPU, coarse+fine forward/backward + Adam, data parallel (one
    all-reduce of the flat gradient buffer per step when world > 1).  Returns ms per step (max over ranks)."""
    from config import Config
    from nerfw.train import Trainer
    m = nerfw.NeRF(Config())
    m.load_state_dict(sd, strict=True)
    m = m.to(dev)
    table = torch.nn.Parameter(torch.randn(100, 32, device=dev))
    tr = Trainer(m, table, lr=5e-4, mlp_dtype=mode)
    g = torch.Generator(device=dev).manual_seed(7)
    o = torch.tensor([0.0, 0.0, 4.0], device=dev).expand(n_rays, 3).contiguous()
    d = torch.nn.functional.normalize(torch.randn(n_rays, 3, device=dev, generator=g) * 0.3 + torch.tensor([0.0, 0.0, -1.0], device=dev), dim=-1)
    tgt = torch.rand(n_rays, 3, device=dev, generator=g)
    for _ in range(warmup):
        tr.step(o, d, tgt, 3, NEAR, FAR, N_COARSE, N_IMPORTANCE, perturb=True, generator=g)
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = tr.step(o, d, tgt, 3, NEAR, FAR, N_COARSE, N_IMPORTANCE, perturb=True, generator=g)
    e1.record()
    barrier(world)
    ms = max_over_ranks(e0.elapsed_time(e1) / steps, world)
    return {"ms_per_step": ms, "rays_per_gpu": n_rays, "n_gpus": world, "samples": "64+128 (coarse+fine fwd/bwd) + Adam",
            "forward_mode": mode, "backward": "tcgen05 bf16 MLP backward (dgrad fused with forward recompute + MN-major wgrad), composite_bwd",
            "allreduce_bytes": int(tr.flat.grad.numel() * 4) if world > 1 else 0, "final_loss": float(loss)}


def run_ours(args):
    import nerfw
    from config import Config
    from nerfw import ops
    from nerfw.camera import aligned_spiral_poses, blender_focal

    rank, local, world = dist_setup(args.gpus)
    dev = torch.device("cuda", torch.cuda.current_device())
    peaks = load_peaks()
    sd, emb = make_weights()
    model = nerfw.NeRF(Config())
    model.load_state_dict(sd, strict=True)
    model = model.to(dev).eval()
    emb_d = emb.to(dev)
    poses = aligned_spiral_poses(120, 2, "x", "chair")
    focal = blender_focal(W)
    pose = poses[(rank * 15) % 120]          # each rank renders its own view of the spiral (config 4 partitioning)
    c2w = torch.from_numpy(pose)
    mode = args.mlp_mode
    n_rays = H * W

    def render(o, d):
        with torch.no_grad():
            return nerfw.volume_render(model, o, d, NEAR, FAR, N_COARSE, N_IMPORTANCE, appearance_embedding=emb_d,
                                       perturb=False, mlp_dtype=mode, generator=gen)

    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    ro, rd = nerfw.get_rays(H, W, focal, c2w.to(dev))
    o_dev = ro.reshape(-1, 3).contiguous()
    d_dev = rd.reshape(-1, 3).contiguous()
    # host-side copies for the e2e leg (pinned)
    o_host = o_dev.cpu().pin_memory()
    d_host = d_dev.cpu().pin_memory()
    rgb_host = torch.empty((n_rays, 3), dtype=torch.float32).pin_memory()
    depth_host = torch.empty((n_rays, 1), dtype=torch.float32).pin_memory()

    def e2e_step():
        o = o_host.to(dev, non_blocking=True)
        d = d_host.to(dev, non_blocking=True)
        rgb, depth, _ = render(o, d)
        rgb_host.copy_(rgb, non_blocking=True)
        depth_host.copy_(depth, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(max(3, args.warmup)):
        render(o_dev, d_dev)
    e2e_step()
    barrier(world)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- device-resident timing ---------------------------------------------------------------------------------
    l0 = ops.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    ev0.record()
    for _ in range(args.steps):
        out = render(o_dev, d_dev)
    ev1.record()
    barrier(world)
    launches = ops.launch_count() - l0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1), world)
    # ---- end to end: pinned host rays -> device -> render -> host images -------------------------------------------
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier(world)
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3, world)
    clocks = sampler.stop() if rank == 0 else None

    # ---- dominant kernel (fine-pass MLP) on its own, CUDA events on the launching stream -------------------------
    with torch.no_grad():
        z_fine = out[2]["z_vals"].contiguous()
        dn = ops.normalize_dirs(d_dev)
        names, tensors = model.kernel_params()
        pd = {n: t.detach() for n, t in zip(names, tensors)}
        packed = model.packed_weights(names, tensors)
        emb2 = emb_d.unsqueeze(0).contiguous()
        roof = {}
        z_coarse = out[2]["z_vals_coarse"].contiguous()
        fine_mode = "fp16" if mode == "mixed" else mode          # kernel arithmetic of the fine pass (the dominant launch)
        # (label, kernel mode, depths): the fine-pass launch of the headline mode, bf16 on the same shape, and -- for the
        # mixed mode -- the bf16x3 coarse-pass launch
        roof_runs = [(mode, fine_mode, z_fine), ("bf16", "bf16", z_fine)]
        if mode == "mixed":
            roof_runs.append(("coarse_bf16x3", "bf16x3", z_coarse))
        for label, m, zz in roof_runs:
            if label in roof:
                continue
            mid = nerfw.models.resolve_mode(m)
            for _ in range(2):
                raw_m = ops.mlp_fwd(pd, packed, o_dev, dn, zz, emb2, mid)
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = max(2, args.steps)
            k0.record()
            for _ in range(reps):
                raw_m = ops.mlp_fwd(pd, packed, o_dev, dn, zz, emb2, mid)
            k1.record()
            torch.cuda.synchronize()
            if zz is z_fine:
                raw = raw_m
            kms = k0.elapsed_time(k1) / reps
            flops = FLOP_PER_SAMPLE * float(zz.numel())
            ach = flops / (kms * 1e-3) / 1e12
            # tensor-pipe FLOPs actually issued: bf16x3 runs the trunk as 3 MMAs per product (the direction layer as one),
            # heads stay on CUDA cores: 3 018 496 issued vs 1 063 936 algorithmic FLOP per sample; single-pass modes 1 054 464
            issued = ach * (3018496.0 if m == "bf16x3" else 1054464.0) / FLOP_PER_SAMPLE
            roof[label] = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                           "frac": ach / peaks["bf16_tflops_sustained"],
                           # DRAM bytes per launch: 18.9 B/sample measured by ncu --set full on a 160k-ray crop
                           # (profiles/r01_ncu_full_*.md: 141 MB read + 439 MB written for 30.7 M samples) x samples here
                           "traffic": 18.9 * float(zz.numel()),
                           "kernel": ("mlp_tc_fwd_kernel<%s>" % {"bf16x3": "X3", "fp16": "F16", "bf16": "BF16"}[m]) if m != "fp32" else "mlp_ffma_fwd_kernel",
                           "kernel_mode": m, "kernel_ms": kms, "samples_per_launch": int(zz.numel()),
                           "peak_source": f"bf16 dense sustained (fp16 and bf16 share the kind::f16 tensor-pipe rate), {peaks['source']}",
                           "issued": issued, "issued_frac": issued / peaks["bf16_tflops_sustained"],
                           "note": ("bf16x3 (fp32-parity split) issues 3 bf16 MMAs per trunk product, so the algorithmic frac is "
                                    "bounded by ~0.35; issued_frac is the tensor-pipe rate against the same peak"
                                    if m == "bf16x3" else
                                    "fine pass of the mixed mode: single fp16 MMA per product; the coarse pass (1/4 of the samples) "
                                    "runs in bf16x3, see roofline_other.mlp_coarse_bf16x3" if label == "mixed" else "")}
        # HBM-bound kernels on the same frame
        comp0, comp1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.composite_fwd(raw, z_fine)
        comp0.record()
        for _ in range(5):
            ops.composite_fwd(raw, z_fine)
        comp1.record()
        torch.cuda.synchronize()
        cms = comp0.elapsed_time(comp1) / 5
        cbytes = z_fine.numel() * 24.0 + n_rays * 20.0
        zc = out[2]["z_vals_coarse"].contiguous()
        wc = out[2]["weights_coarse"][..., 0].contiguous()
        ur = torch.rand((n_rays, N_IMPORTANCE), device=dev)
        ops.sample_pdf(zc, wc, N_IMPORTANCE, ur)
        comp0.record()
        for _ in range(5):
            ops.sample_pdf(zc, wc, N_IMPORTANCE, ur)
        comp1.record()
        torch.cuda.synchronize()
        rms = comp0.elapsed_time(comp1) / 5
        rbytes = n_rays * (3 * N_COARSE + 2 * N_IMPORTANCE) * 4.0
        g_rgb = torch.rand((n_rays, 3), device=dev)
        g_depth = torch.rand((n_rays, 1), device=dev)
        ops.composite_bwd(raw, z_fine, g_rgb, g_depth, None, None)
        comp0.record()
        for _ in range(5):
            ops.composite_bwd(raw, z_fine, g_rgb, g_depth, None, None)
        comp1.record()
        torch.cuda.synchronize()
        bms = comp0.elapsed_time(comp1) / 5
        bbytes = z_fine.numel() * 40.0
        hbm = {"composite_fwd": {"bound": "hbm", "achieved": cbytes / (cms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                                 "unit": "GB/s", "frac": cbytes / (cms * 1e-3) / 1e9 / peaks["hbm_gbs"], "kernel_ms": cms},
               "composite_bwd": {"bound": "hbm", "achieved": bbytes / (bms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                                 "unit": "GB/s", "frac": bbytes / (bms * 1e-3) / 1e9 / peaks["hbm_gbs"], "kernel_ms": bms},
               "sample_pdf": {"bound": "hbm", "achieved": rbytes / (rms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                              "unit": "GB/s", "frac": rbytes / (rms * 1e-3) / 1e9 / peaks["hbm_gbs"], "kernel_ms": rms}}

    # ---- secondary numbers: the other tensor-core mode, and the training step (BASELINE.json configs[2]) ------------
    other_modes = {}
    for m in ("bf16", "bf16x3", "mixed"):
        if m == mode:
            continue
        with torch.no_grad():
            nerfw.volume_render(model, o_dev, d_dev, NEAR, FAR, N_COARSE, N_IMPORTANCE, appearance_embedding=emb_d,
                                perturb=False, mlp_dtype=m, generator=gen)
            barrier(world)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(2):
                nerfw.volume_render(model, o_dev, d_dev, NEAR, FAR, N_COARSE, N_IMPORTANCE, appearance_embedding=emb_d,
                                    perturb=False, mlp_dtype=m, generator=gen)
            a1.record()
            barrier(world)
        ms = max_over_ranks(a0.elapsed_time(a1) / 2, world)
        other_modes[m] = {"value": world * n_rays / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms}
    # opt-in: one network for both passes -> the fine pass evaluates only the 128 new samples (192 MLP evaluations per ray)
    with torch.no_grad():
        nerfw.volume_render(model, o_dev, d_dev, NEAR, FAR, N_COARSE, N_IMPORTANCE, appearance_embedding=emb_d,
                            perturb=False, mlp_dtype=mode, generator=gen, reuse_coarse=True)
        barrier(world)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(2):
            nerfw.volume_render(model, o_dev, d_dev, NEAR, FAR, N_COARSE, N_IMPORTANCE, appearance_embedding=emb_d,
                                perturb=False, mlp_dtype=mode, generator=gen, reuse_coarse=True)
        a1.record()
        barrier(world)
    ms = max_over_ranks(a0.elapsed_time(a1) / 2, world)
    other_modes["reuse_coarse_opt_in"] = {"value": world * n_rays / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": ms,
                                          "mlp_mode": mode, "mlp_evals_per_ray": N_COARSE + N_IMPORTANCE,
                                          "note": "not the headline: the coarse samples are not re-evaluated in the fine pass"}
    # the reference's own calling pattern: 4096-ray chunks with a device->host copy per chunk
    # (render_aligned_spiral.py:136-155), one frame
    with torch.no_grad():
        barrier(world)
        t0 = time.perf_counter()
        parts = []
        for j in range(0, n_rays, 4096):
            c_rgb, c_depth, _ = nerfw.volume_render(model, o_dev[j:j + 4096], d_dev[j:j + 4096], NEAR, FAR, N_COARSE,
                                                    N_IMPORTANCE, appearance_embedding=emb_d, perturb=False,
                                                    mlp_dtype=mode, generator=gen)
            parts.append((c_rgb.cpu(), c_depth.cpu()))
        barrier(world)
        chunk_ms = max_over_ranks((time.perf_counter() - t0) * 1e3, world)
    other_modes["chunked_4096_with_cpu_copy"] = {"value": world * n_rays / (chunk_ms * 1e-3) / 1e6, "unit": "Mrays/s",
                                                 "ms_per_step": chunk_ms, "mlp_mode": mode,
                                                 "calls_per_frame": (n_rays + 4095) // 4096}
    train = bench_train_step(nerfw, sd, dev, world, mode)
    if mode != "bf16x3":
        train["bf16x3_forward_ms_per_step"] = bench_train_step(nerfw, sd, dev, world, "bf16x3")["ms_per_step"]
    if mode != "bf16":
        train["bf16_forward_ms_per_step"] = bench_train_step(nerfw, sd, dev, world, "bf16")["ms_per_step"]

    if rank != 0:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    # ---- CPU baseline on a bounded sample + parity of the same sample ----------------------------------------------
    cpu = None
    parity = None
    if not args.no_cpu_baseline:
        n_cpu, dt, cores, (oc, dc, u, rgb_o, depth_o, acc_o) = cpu_sample(sd, emb, args.cpu_rays, poses[0])
        with torch.no_grad():
            rgb_g, depth_g, ex_g = nerfw.volume_render(model, oc.to(dev), dc.to(dev), NEAR, FAR, N_COARSE, N_IMPORTANCE,
                                                       appearance_embedding=emb_d, perturb=False, mlp_dtype=mode, u_rand=u)
        parity = {"rays": n_cpu, "rgb_max_abs": float((rgb_g.cpu() - rgb_o).abs().max()),
                  "depth_max_abs": float((depth_g.cpu() - depth_o).abs().max()),
                  "acc_max_abs": float((ex_g["acc"].cpu() - acc_o).abs().max())}
        cpu = {"value": n_cpu / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} rays (centre rows of the same 800x800 view), coarse 64 + fine 192, one pass, {dt:.1f} s"}

    value = world * n_rays * args.steps / (ms_total * 1e-3) / 1e6
    e2e_val = world * n_rays * args.steps / (e2e_ms * 1e-3) / 1e6
    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"mixed": "bf16x3 coarse pass + fp16 fine pass (tcgen05 kind::f16, fp32 accumulate; fp32 parity bars)",
                  "bf16x3": "bf16x3 (fp32-parity split, fp32 accumulate)", "bf16": "bf16", "fp16": "fp16", "fp32": "f32"}[mode],
        "data": "synthetic",
        "config": {"workload": "800x800 view render, 64+128 samples, random-init NeRF-W (BASELINE.json configs[1])",
                   "rays_per_step_per_gpu": n_rays, "mlp_evals_per_ray": SAMPLES_PER_RAY, "mlp_mode": mode,
                   "call": "one whole-frame volume_render per step", "parallelism": f"rays/frames sharded x{world}, no collective",
                   "l2": "per-step working set 2.6 GB of sample buffers >> 126 MB L2 (inputs larger than L2)"},
        "e2e": {"value": e2e_val, "unit": "Mrays/s", "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(o_host.numel() * 4 + d_host.numel() * 4),
                "d2h_bytes_per_step": int(rgb_host.numel() * 4 + depth_host.numel() * 4)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof[mode],
        "roofline_other": {**{f"mlp_{k}": v for k, v in roof.items() if k != mode}, **hbm},
        "other_modes": other_modes,
        "train_step": train,
        "cpu_baseline": cpu,
        "parity_vs_oracle": parity,
    }
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    args = parse_args()
    # Exactly one JSON line may reach stdout: libraries (NCCL prints its version banner there) get stderr instead.
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device visible; the nerfw hot path has no CPU fallback "
                             "(use --impl reference for the CPU reference arm)")
        run_ours(args)


if __name__ == "__main__":
    main()
