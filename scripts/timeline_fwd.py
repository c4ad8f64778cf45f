"""Event timeline of one tile of the bf16 forward kernel (CTA 0, third tile), from clock64 stamps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the kernels' profiling switches exist only in the profiling build (make -C csrc PROFILE=1): load that library
os.environ.setdefault("NERFW_PROFILE_LIB", "1")
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, nerfw, nerfw_oracle as orc
from config import Config
from nerfw import ops
m = nerfw.NeRF(Config()); m.load_state_dict(orc.make_state_dict(0)); m = m.cuda()
b, n = 40000, 192
MODE = int(sys.argv[1]) if len(sys.argv) > 1 else 2
g = torch.Generator(device="cuda").manual_seed(1)
o = torch.randn(b, 3, device="cuda", generator=g)
d = torch.nn.functional.normalize(torch.randn(b, 3, device="cuda", generator=g), dim=-1)
z = torch.sort(torch.rand(b, n, device="cuda", generator=g) * 4 + 2, dim=-1).values
names, tensors = m.kernel_params()
params = {k: t.detach() for k, t in zip(names, tensors)}
packed = m.packed_weights(names, tensors)
tl = torch.zeros(128, dtype=torch.int64, device="cuda")
ops.mlp_fwd(params, packed, o, d, z, None, MODE)
os.environ["NERFW_FWD_TIMELINE"] = str(tl.data_ptr())
ops.mlp_fwd(params, packed, o, d, z, None, MODE)
torch.cuda.synchronize()
t = tl.cpu().tolist()
t0 = t[0]
names = ["mma: acc free seen", "mma: A kb0 seen", "mma: A kb3 seen", "mma: layer issued", "epi: acc complete seen", "epi: acc in regs", "epi: kb0 published", "epi: kb3 published"]
print("cycles relative to pe_ready seen by the MMA thread (tile 3 of CTA 0)")
for layer in range(8):
    row = [(names[i], t[10 + layer * 8 + i] - t0 if t[10 + layer * 8 + i] else None) for i in range(8)]
    print(f"layer {layer}: " + "  ".join(f"{k}={v}" for k, v in row))
for slot, name in ((94, "mma: direction layer issued"), (90, "epi: encodings for next tile written"), (91, "epi: dir acc complete seen"), (92, "epi: dir acc in regs"), (93, "epi: tile written")):
    print(f"{name}: {t[slot] - t0 if t[slot] else None}")
