import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200"))
import torch
from nerfw import ops
b, n, ni = 640000, 64, 128
g = torch.Generator(device="cuda").manual_seed(0)
z = torch.sort(torch.rand(b, n, device="cuda", generator=g) * 4 + 2, dim=-1).values
w = torch.rand(b, n, device="cuda", generator=g) ** 8 * 0.05
u = torch.rand(b, ni, device="cuda", generator=g)
for _ in range(3):
    out = ops.sample_pdf(z, w, ni, u)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = ops.sample_pdf(z, w, ni, u)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"sample_pdf general={os.environ.get('NERFW_RESAMPLE_GENERAL')}: {ms:.3f} ms per call, {b * (3 * n + 2 * ni) * 4 / ms / 1e6:.0f} GB/s algorithmic")
