"""The two MLP launches of a frame on their own: coarse (640 000 x 64, bf16x3) and fine (640 000 x 128, fp16)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, nerfw, nerfw_oracle as orc
from config import Config
from nerfw import ops
m = nerfw.NeRF(Config()); m.load_state_dict(orc.make_state_dict(0)); m = m.cuda().eval()
emb = torch.randn(1, 32).cuda()
g = torch.Generator(device="cuda").manual_seed(0)
b = 640000
o = torch.tensor([0.0, 0.0, 4.0], device="cuda").expand(b, 3).contiguous()
d = torch.nn.functional.normalize(torch.randn(b, 3, device="cuda", generator=g) * 0.3 + torch.tensor([0.0, 0.0, -1.0], device="cuda"), dim=-1)
ws, packed = m.kernel_state()[2], m.packed_weights()
for mode, n in (("fp16", 128), ("bf16x3", 64), ("fp16", 128), ("bf16x3", 64), ("bf16", 128)):
    z = torch.sort(torch.rand(b, n, device="cuda", generator=g) * 4 + 2, dim=-1).values
    mid = nerfw.models.resolve_mode(mode)
    for _ in range(2):
        raw = ops.mlp_fwd(ws, packed, o, d, z, emb, mid)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        raw = ops.mlp_fwd(ws, packed, o, d, z, emb, mid)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{mode} x{n}: {ms:.2f} ms, {b * n * 1063936 / ms / 1e9:.0f} TFLOP/s algorithmic, checksum {float(raw.double().sum()):.6f}")
