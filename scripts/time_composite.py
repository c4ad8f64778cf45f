"""Device time of composite_fwd / composite_bwd at the frame's fine-pass size (640 000 rays x 192 samples)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200"))
import torch
from nerfw import ops
b = 640000
for n in (192, 64):
    g = torch.Generator(device="cuda").manual_seed(0)
    raw = torch.rand(b, n, 4, device="cuda", generator=g)
    z = torch.sort(torch.rand(b, n, device="cuda", generator=g) * 4 + 2, dim=-1).values
    d_rgb = torch.randn(b, 3, device="cuda", generator=g); d_dep = torch.randn(b, device="cuda", generator=g); d_acc = torch.randn(b, device="cuda", generator=g)
    def t(fn, k=10):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k
    ms = t(lambda: ops.composite_fwd(raw, z))
    print(f"N={n} composite_fwd {ms:.3f} ms, {b * (n * 24 + 20) / ms / 1e6:.0f} GB/s algorithmic")
    ms = t(lambda: ops.composite_bwd(raw, z, d_rgb, d_dep, d_acc, None))
    print(f"N={n} composite_bwd {ms:.3f} ms, {b * (n * 40 + 20) / ms / 1e6:.0f} GB/s algorithmic")
