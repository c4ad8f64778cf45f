import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the kernels' profiling switches exist only in the profiling build (make -C csrc PROFILE=1): load that library
os.environ.setdefault("NERFW_PROFILE_LIB", "1")
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, nerfw, nerfw_oracle as orc
from config import Config
from nerfw import ops
sd = orc.make_state_dict(0); emb = torch.randn(32)
m = nerfw.NeRF(Config()); m.load_state_dict(sd); m = m.cuda()
b, n = 4096, 192
gen = torch.Generator(device="cuda").manual_seed(1)
o = torch.randn(b, 3, device="cuda", generator=gen)
d = torch.nn.functional.normalize(torch.randn(b, 3, device="cuda", generator=gen), dim=-1)
z = torch.sort(torch.rand(b, n, device="cuda", generator=gen) * 4 + 2, dim=-1).values
d_raw = torch.randn(b * n, 4, device="cuda", generator=gen)
e = emb.cuda().unsqueeze(0).contiguous()
names, tensors = m.kernel_params()
params = {k: t.detach() for k, t in zip(names, tensors)}
packed = m.packed_weights(names, tensors)
g = {k: torch.zeros_like(t) for k, t in params.items()}
de = torch.zeros(1, 32, device="cuda")
_, masks = ops.mlp_fwd(params, packed, o, d, z, e, 1, want_masks=True)
for _ in range(2):
    ops.mlp_bwd_tc(params, g, packed, o, d, z, e, d_raw, de, masks)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.mlp_bwd_tc(params, g, packed, o, d, z, e, d_raw, de, masks)
e1.record(); torch.cuda.synchronize()
print("NERFW_WGRAD_DEBUG=%s  bwd_tc total %.3f ms per call (pass1 + wgrad), %d tiles" % (os.environ.get("NERFW_WGRAD_DEBUG", "0"), e0.elapsed_time(e1) / 5, b * n // 128))
