"""Event timeline of one tile of backward pass 1 (CTA 0, third tile), from clock64 stamps (NERFW_BWD_TIMELINE)."""
import os, sys
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
ops.mlp_bwd_tc(params, g, packed, o, d, z, e, d_raw, de, masks)
tl = torch.zeros(256, dtype=torch.int64, device="cuda")
os.environ["NERFW_BWD_TIMELINE"] = str(tl.data_ptr())
ops.mlp_bwd_tc(params, g, packed, o, d, z, e, d_raw, de, masks)
torch.cuda.synchronize()
t = tl.cpu().tolist(); t0 = t[0]
ev = ["mma:acc_free", "mma:kb0", "mma:kb_last", "mma:issued", "epi:acc_full(prev step)", "epi:kb0 pub", "epi:kb3 pub", "epi:staged"]
names = [f"fwd L{i}" for i in range(8)] + ["fwd dir", "dgrad dir"] + [f"dgrad L{l}" for l in range(7, 0, -1)] + ["(after dgrad L1)"]
for step in range(18):
    row = "  ".join(f"{ev[i]}={t[16 + 8 * step + i] - t0 if t[16 + 8 * step + i] else None}" for i in range(8))
    print(f"{names[step]:12s} {row}")
