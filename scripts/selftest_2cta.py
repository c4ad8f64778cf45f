"""CTA-pair (tcgen05 cta_group::2) primitive check: D = A B^T, M = 256 over two SMs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200"))
import torch
from nerfw import ops
for mode in (0, 1):
    for n, k in ((256, 64), (256, 256), (128, 128), (64, 192)):
        g = torch.Generator().manual_seed(n + k + mode)
        a = (torch.randn(256, k, generator=g)).bfloat16().cuda()
        b = (torch.randn(n, k, generator=g)).bfloat16().cuda()
        want = a.float() @ b.float().T
        got = ops.selftest_umma_2cta(a, b, mode)
        torch.cuda.synchronize()
        err = float((got - want).abs().max())
        print(f"mode {mode} N={n} K={k}: max abs err {err:.3e}", "OK" if err < 1e-2 else "MISMATCH", flush=True)
