"""Host-side cost of one volume_render call (4096 rays, 64+128): cProfile over 300 calls, top entries by total time."""
import cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import nerfw
import nerfw_oracle as orc
from config import Config
m = nerfw.NeRF(Config()); m.load_state_dict(orc.make_state_dict(0)); m = m.cuda().eval()
emb = torch.randn(32).cuda()
g = torch.Generator(device="cuda").manual_seed(0)
o = torch.tensor([0.0, 0.0, 4.0], device="cuda").expand(4096, 3).contiguous()
d = torch.nn.functional.normalize(torch.randn(4096, 3, device="cuda", generator=g) * 0.3 + torch.tensor([0.0, 0.0, -1.0], device="cuda"), dim=-1)
def call():
    with torch.no_grad():
        return nerfw.volume_render(m, o, d, 2.0, 6.0, 64, 128, appearance_embedding=emb, perturb=False)
for _ in range(20): call()
torch.cuda.synchronize()
best = 1e9
for _ in range(10):          # bursts of 20 calls: the launch queue never fills, so this is host time only
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): call()
    best = min(best, (time.perf_counter() - t0) / 20 * 1e6)
torch.cuda.synchronize()
print(f"enqueue time per call: {best:.1f} us")
pr = cProfile.Profile()
for _ in range(15):
    torch.cuda.synchronize()
    pr.enable()
    for _ in range(20): call()
    pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
