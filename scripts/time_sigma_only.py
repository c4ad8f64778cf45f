"""Coarse-pass-shaped bf16x3 launch (640 000 rays x 64 samples) with and without NERFW_MLP_SIGMA_ONLY."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, nerfw, nerfw_oracle as orc
from config import Config
from nerfw import ops
m = nerfw.NeRF(Config()); m.load_state_dict(orc.make_state_dict(0)); m = m.cuda()
b, n = 320000, 64
g = torch.Generator(device="cuda").manual_seed(1)
o = torch.randn(b, 3, device="cuda", generator=g)
d = torch.nn.functional.normalize(torch.randn(b, 3, device="cuda", generator=g), dim=-1)
z = torch.sort(torch.rand(b, n, device="cuda", generator=g) * 4 + 2, dim=-1).values
names, tensors = m.kernel_params()
params = {k: t.detach() for k, t in zip(names, tensors)}
packed = m.packed_weights(names, tensors)
for mode, name in ((1, "bf16x3"), (3, "fp16")):
    for flag in (False, True):
        for _ in range(2): ops.mlp_fwd(params, packed, o, d, z, None, mode, sigma_only=flag)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.mlp_fwd(params, packed, o, d, z, None, mode, sigma_only=flag)
        e1.record(); torch.cuda.synchronize()
        print(f"{name} sigma_only={flag}: {e0.elapsed_time(e1) / 5:.2f} ms")
