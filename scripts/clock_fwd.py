"""Average SM clock and SM cycles per 128-sample tile of the forward MLP kernel in the power-capped steady state: CTA 0 of
the profiling build stamps clock64 / globaltimer at its first and last instruction (slots 120-123 of the timeline buffer).
Compares with the clock64 timeline of the third tile of a launch (scripts/timeline_fwd.py), taken before the cap bites."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("NERFW_PROFILE_LIB", "1")
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, nerfw, nerfw_oracle as orc
from config import Config
from nerfw import ops
m = nerfw.NeRF(Config()); m.load_state_dict(orc.make_state_dict(0)); m = m.cuda().eval()
emb = torch.randn(1, 32).cuda()
g = torch.Generator(device="cuda").manual_seed(0)
ws, packed = m.kernel_state()[2], m.packed_weights()
b = 640000
o = torch.tensor([0.0, 0.0, 4.0], device="cuda").expand(b, 3).contiguous()
d = torch.nn.functional.normalize(torch.randn(b, 3, device="cuda", generator=g) * 0.3 + torch.tensor([0.0, 0.0, -1.0], device="cuda"), dim=-1)
tl = torch.zeros(128, dtype=torch.int64, device="cuda")
os.environ["NERFW_FWD_TIMELINE"] = str(tl.data_ptr())
for mode, n in (("fp16", 128), ("bf16", 128), ("bf16x3", 64)):
    z = torch.sort(torch.rand(b, n, device="cuda", generator=g) * 4 + 2, dim=-1).values
    mid = nerfw.models.resolve_mode(mode)
    for reps, label in ((1, "first launch after idle"), (25, "25th launch back to back")):
        torch.cuda.synchronize()
        import time; time.sleep(1.0)
        for _ in range(reps):
            ops.mlp_fwd(ws, packed, o, d, z, emb, mid)
        torch.cuda.synchronize()
        t = tl.cpu().tolist()
        cyc, ns = t[122] - t[120], t[123] - t[121]
        tiles = (b * n // 128 + 147) // 148
        print(f"{mode} x{n} {label}: {ns / 1e6:.2f} ms, average SM clock {cyc / ns * 1e3:.0f} MHz, {cyc / tiles:.0f} cycles per tile "
              f"({b * n * 1063936 / ns / 1e3:.0f} TFLOP/s algorithmic)")
