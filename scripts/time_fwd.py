import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the kernels' profiling switches exist only in the profiling build (make -C csrc PROFILE=1): load that library
os.environ.setdefault("NERFW_PROFILE_LIB", "1")
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, nerfw, nerfw_oracle as orc
from config import Config
from nerfw import ops
m = nerfw.NeRF(Config()); m.load_state_dict(orc.make_state_dict(0)); m = m.cuda()
b, n = 160000, 192
g = torch.Generator(device="cuda").manual_seed(1)
o = torch.randn(b, 3, device="cuda", generator=g)
d = torch.nn.functional.normalize(torch.randn(b, 3, device="cuda", generator=g), dim=-1)
z = torch.sort(torch.rand(b, n, device="cuda", generator=g) * 4 + 2, dim=-1).values
names, tensors = m.kernel_params()
params = {k: t.detach() for k, t in zip(names, tensors)}
packed = m.packed_weights(names, tensors)
for mode, name in ((2, "bf16"), (3, "fp16"), (1, "bf16x3")):
    for _ in range(2):
        ops.mlp_fwd(params, packed, o, d, z, None, mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        ops.mlp_fwd(params, packed, o, d, z, None, mode)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    print(f"skip_weights={os.environ.get('NERFW_FWD_SKIP_WEIGHTS')} {name}: {ms:.2f} ms, {b*n*1063936/ms/1e9:.0f} TFLOP/s algorithmic")
