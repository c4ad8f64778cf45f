"""sample_pdf (K8) alone at the frame's shape: 640 000 rays, 64 + 128.  With NERFW_PROFILE_LIB=1 (profiling build) the
NERFW_PDF_MINB switch selects the register / occupancy variant of the half-warp kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200"))
import torch
from nerfw import ops
b, n, ni = 640000, 64, 128
g = torch.Generator(device="cuda").manual_seed(0)
z = torch.sort(torch.rand(b, n, device="cuda", generator=g) * 4 + 2, dim=-1).values
w = torch.rand(b, n, device="cuda", generator=g) ** 8 * 0.05
u = torch.rand(b, ni, device="cuda", generator=g)
time.sleep(0.5)
for _ in range(10):
    out = ops.sample_pdf(z, w, ni, u)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = ops.sample_pdf(z, w, ni, u)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
want = ops.sample_pdf(z[:4096], w[:4096], ni, u[:4096], general_path=True)
print(f"sample_pdf minb={os.environ.get('NERFW_PDF_MINB', 'default')} profile_lib={os.environ.get('NERFW_PROFILE_LIB', '0')}: {ms:.3f} ms per call, "
      f"{b * (3 * n + 2 * ni) * 4 / ms / 1e6:.0f} GB/s algorithmic, equals general path: {bool(torch.equal(out[:4096], want))}")
