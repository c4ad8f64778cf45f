"""Fine-pass MLP launch (fp16, 128 depths) timed (a) back to back for seconds (power-capped clocks) and (b) as short
launches with idle gaps (boost clocks, no cap), with power / SM clock sampled through NVML.  Separates 'fewer cycles'
from 'fewer joules': under the cap only the second shows up in milliseconds.  NERFW_PROFILE_LIB=1 loads the other build."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, nerfw, nerfw_oracle as orc, pynvml
from config import Config
from nerfw import ops
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
m = nerfw.NeRF(Config()); m.load_state_dict(orc.make_state_dict(0)); m = m.cuda().eval()
emb = torch.randn(1, 32).cuda()
g = torch.Generator(device="cuda").manual_seed(0)
ws, packed = m.kernel_state()[2], m.packed_weights()
mode = sys.argv[1] if len(sys.argv) > 1 else "fp16"
mid = nerfw.models.resolve_mode(mode)
n = 128
def rays(b):
    o = torch.tensor([0.0, 0.0, 4.0], device="cuda").expand(b, 3).contiguous()
    d = torch.nn.functional.normalize(torch.randn(b, 3, device="cuda", generator=g) * 0.3 + torch.tensor([0.0, 0.0, -1.0], device="cuda"), dim=-1)
    z = torch.sort(torch.rand(b, n, device="cuda", generator=g) * 4 + 2, dim=-1).values
    return o, d, z
samples = []; stop = False
def sampler():
    while not stop:
        samples.append((time.time(), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
        time.sleep(0.02)
# (b) short launches with idle gaps
o, d, z = rays(32000)
for _ in range(3): ops.mlp_fwd(ws, packed, o, d, z, emb, mid)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    time.sleep(0.08)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.mlp_fwd(ws, packed, o, d, z, emb, mid); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print(f"{mode} short (32000 x {n}, idle gaps): median {ts[10]:.3f} ms min {ts[0]:.3f} -> {32000 * n * 1063936 / ts[10] / 1e9:.0f} TFLOP/s")
# (a) sustained
o, d, z = rays(640000)
for _ in range(2): ops.mlp_fwd(ws, packed, o, d, z, emb, mid)
torch.cuda.synchronize()
th = threading.Thread(target=sampler); th.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time(); e0.record()
for _ in range(30): ops.mlp_fwd(ws, packed, o, d, z, emb, mid)
e1.record(); torch.cuda.synchronize(); t1 = time.time()
stop = True; th.join()
ms = e0.elapsed_time(e1) / 30
sel = [s for s in samples if t0 + 1.0 < s[0] < t1]
pw = sorted(s[1] for s in sel); ck = sorted(s[2] for s in sel)
print(f"{mode} sustained (640000 x {n} x 30): {ms:.2f} ms -> {640000 * n * 1063936 / ms / 1e9:.0f} TFLOP/s; power median {pw[len(pw)//2]:.0f} W (limit {pynvml.nvmlDeviceGetEnforcedPowerLimit(h)/1000:.0f} W), SM clock median {ck[len(ck)//2]} MHz")
