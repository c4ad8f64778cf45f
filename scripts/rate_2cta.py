"""Issue rate of the CTA-pair MMA (cycles per M=256 x N x K=16 instruction), SS and TS forms."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200"))
import torch
from nerfw import ops
cyc = torch.zeros(2, dtype=torch.int64, device="cuda")
os.environ["NERFW_2CTA_CYCLES"] = str(cyc.data_ptr())
for mode in (0, 1):
    for n in (256, 128):
        k = 256
        a = torch.randn(256, k).bfloat16().cuda(); b = torch.randn(n, k).bfloat16().cuda()
        res = {}
        for reps in (1, 65):
            os.environ["NERFW_2CTA_REPS"] = str(reps)
            ops.selftest_umma_2cta(a, b, mode); torch.cuda.synchronize()
            res[reps] = int(cyc[0])
        per = (res[65] - res[1]) / (64 * 16)
        print(f"mode {'SS' if mode == 0 else 'TS'} N={n}: {per:.1f} cycles per 2-CTA MMA (M=256, K=16)  [{res}]", flush=True)
