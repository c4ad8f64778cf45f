"""CTA-pair forward kernel (NERFW_FWD_PAIR=1) vs the single-CTA kernel: bitwise comparison of raw and device times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, nerfw, nerfw_oracle as orc
from config import Config
from nerfw import ops
m = nerfw.NeRF(Config()); m.load_state_dict(orc.make_state_dict(0)); m = m.cuda()
b, n = int(sys.argv[1]) if len(sys.argv) > 1 else 160000, 192
g = torch.Generator(device="cuda").manual_seed(1)
o = torch.randn(b, 3, device="cuda", generator=g)
d = torch.nn.functional.normalize(torch.randn(b, 3, device="cuda", generator=g), dim=-1)
z = torch.sort(torch.rand(b, n, device="cuda", generator=g) * 4 + 2, dim=-1).values
emb = torch.randn(1, 32, device="cuda", generator=g)
names, tensors = m.kernel_params()
params = {k: t.detach() for k, t in zip(names, tensors)}
packed = m.packed_weights(names, tensors)
def run(mode, pair, reps=3):
    if pair: os.environ["NERFW_FWD_PAIR"] = "1"
    else: os.environ.pop("NERFW_FWD_PAIR", None)
    raw = ops.mlp_fwd(params, packed, o, d, z, emb, mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): raw = ops.mlp_fwd(params, packed, o, d, z, emb, mode)
    e1.record(); torch.cuda.synchronize()
    return raw, e0.elapsed_time(e1) / reps
for mode, name in ((2, "bf16"), (3, "fp16"), (1, "bf16x3")):
    ref, t0 = run(mode, False)
    got, t1 = run(mode, True)
    same = torch.equal(ref, got)
    diff = float((ref - got).abs().max())
    print(f"{name}: single {t0:.2f} ms ({b*n*1063936/t0/1e9:.0f} TF)  pair {t1:.2f} ms ({b*n*1063936/t1/1e9:.0f} TF)  bitwise equal {same}  max diff {diff:.3e}", flush=True)
