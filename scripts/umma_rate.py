import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200"))
import torch
from nerfw._lib import lib, check
out = torch.zeros(1, dtype=torch.int64, device="cuda")
for mode, name in ((0, "K-major SS"), (1, "K-major TS"), (2, "MN-major SS"), (3, "i8 SS (K=32)"), (4, "i8 TS (K=32)"), (5, "e4m3 SS (K=32)"), (6, "TS + kblock protocol")):
    for n in (64, 128, 256):
        for reps in (50, 500):
            check(lib().nerfw_selftest_umma_rate(mode, n, reps, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            cyc = int(out.item())
            print(f"{name:12s} N={n:3d} reps={reps:4d}: {cyc / (reps * 16):8.1f} cycles per MMA (128 x {n} x {32 if mode in (3, 4, 5) else 16})")
