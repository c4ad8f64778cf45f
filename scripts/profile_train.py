"""Small fixed workload for ncu: three training steps (4096 rays, 64+128 samples) through nerfw.train.Trainer."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

import nerfw  # noqa: E402
import nerfw_oracle as orc  # noqa: E402
from config import Config  # noqa: E402
from nerfw.train import Trainer  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m = nerfw.NeRF(Config())
m.load_state_dict(orc.make_state_dict(0))
m = m.cuda()
table = torch.nn.Parameter(torch.randn(100, 32, device="cuda"))
tr = Trainer(m, table, lr=5e-4, mlp_dtype=mode)
g = torch.Generator(device="cuda").manual_seed(7)
n = 4096
o = torch.tensor([0.0, 0.0, 4.0], device="cuda").expand(n, 3).contiguous()
d = torch.nn.functional.normalize(torch.randn(n, 3, device="cuda", generator=g) * 0.3 + torch.tensor([0.0, 0.0, -1.0], device="cuda"), dim=-1)
tgt = torch.rand(n, 3, device="cuda", generator=g)
for _ in range(steps):
    loss = tr.step(o, d, tgt, 3, 2.0, 6.0, 64, 128, perturb=True, generator=g)
torch.cuda.synchronize()
print("loss", float(loss))
