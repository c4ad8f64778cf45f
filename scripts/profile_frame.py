"""Small fixed workload for ncu: one 800x800 coarse+fine frame (or a crop) through the public API, once.

    python scripts/profile_frame.py [--mode bf16x3] [--rows 800]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

import nerfw  # noqa: E402
import nerfw_oracle as orc  # noqa: E402
from config import Config  # noqa: E402
from nerfw.camera import aligned_spiral_poses, blender_focal  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="mixed")
ap.add_argument("--rows", type=int, default=800)
ap.add_argument("--reps", type=int, default=1)
args = ap.parse_args()
sd = orc.make_state_dict(0)
emb = torch.randn(32).cuda()
model = nerfw.NeRF(Config())
model.load_state_dict(sd)
model = model.cuda()
pose = torch.from_numpy(aligned_spiral_poses(120, 2, "x", "chair")[0]).cuda()
o, d = nerfw.get_rays(800, 800, blender_focal(800), pose)
r0 = 400 - args.rows // 2
o = o[r0:r0 + args.rows].reshape(-1, 3).contiguous()
d = d[r0:r0 + args.rows].reshape(-1, 3).contiguous()
gen = torch.Generator(device="cuda").manual_seed(0)
for _ in range(args.reps):
    with torch.no_grad():
        rgb, depth, ex = nerfw.volume_render(model, o, d, 2.0, 6.0, 64, 128, appearance_embedding=emb, perturb=False,
                                             mlp_dtype=args.mode, generator=gen)
torch.cuda.synchronize()
print("rays", o.shape[0], "rgb mean", float(rgb.mean()), "depth mean", float(depth.mean()))
