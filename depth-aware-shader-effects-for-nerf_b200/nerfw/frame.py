"""Whole-frame / whole-path rendering: the per-frame body of render_aligned_spiral.py:124-175 as one device-resident call.

The reference builds rays, walks 157 chunks of 4096 rays with a `.cpu()` sync per chunk, concatenates and quantises on
the host.  Here a frame is: raygen -> one volume_render over all H*W rays (or `chunk` rays at a time when memory is
capped) -> uint8 quantisation on the device -> one device-to-host copy.  Frames of a camera path are distributed
round-robin over the ranks.
"""
from __future__ import annotations

from typing import Iterator, Optional, Tuple

import numpy as np
import torch

from . import ops
from .camera import aligned_spiral_poses
from .parallel import frames_for_rank, world_info
from .ray_utils import get_rays
from .render import volume_render


@torch.no_grad()
def render_frame(model, height: int, width: int, focal: float, c2w, near: float, far: float, n_samples: int,
                 n_importance: int, appearance_embedding=None, chunk: Optional[int] = None,
                 mlp_dtype: Optional[str] = None, generator=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(rgb (H,W,3) float32, depth (H,W) float32, acc (H,W) float32) on the model's device, perturb=False."""
    coarse = model[0] if isinstance(model, (tuple, list)) else model
    dev = coarse.rgb_linear.weight.device
    c2w = torch.as_tensor(c2w, dtype=torch.float32)
    o, d = get_rays(height, width, focal, c2w.to(dev))
    o = o.reshape(-1, 3).contiguous()
    d = d.reshape(-1, 3)
    n = o.shape[0]
    step = n if not chunk else int(chunk)
    rgbs, depths, accs = [], [], []
    for s in range(0, n, step):
        rgb, depth, ex = volume_render(model, o[s:s + step], d[s:s + step], near, far, n_samples, n_importance,
                                       appearance_embedding=appearance_embedding, perturb=False, mlp_dtype=mlp_dtype,
                                       generator=generator)
        rgbs.append(rgb)
        depths.append(depth)
        accs.append(ex["acc"])
    cat = (lambda xs: xs[0] if len(xs) == 1 else torch.cat(xs, dim=0))
    return cat(rgbs).reshape(height, width, 3), cat(depths).reshape(height, width), cat(accs).reshape(height, width)


def quantize_frame(rgb: torch.Tensor, depth: Optional[torch.Tensor] = None):
    """uint8 images exactly as render_aligned_spiral.py:161-173: rgb -> (rgb*255).astype(uint8); depth -> min/max
    normalised to 0..255.  Returned on the device; callers copy once."""
    rgb8 = ops.quantize_u8(rgb)
    if depth is None:
        return rgb8, None
    lo, hi = depth.min(), depth.max()
    depth8 = ops.quantize_u8(((depth - lo) / (hi - lo)).contiguous())
    return rgb8, depth8


def render_path(model, poses: np.ndarray, height: int, width: int, focal: float, near: float, far: float,
                n_samples: int, n_importance: int, appearance_embedding=None, group=None, chunk: Optional[int] = None,
                mlp_dtype: Optional[str] = None) -> Iterator[Tuple[int, np.ndarray, np.ndarray]]:
    """Yields (frame index, rgb uint8 (H,W,3), depth float32 (H,W)) for the frames this rank owns (i mod world == rank)."""
    rank, world = world_info(group)
    for i in frames_for_rank(len(poses), rank, world):
        rgb, depth, _ = render_frame(model, height, width, focal, poses[i], near, far, n_samples, n_importance,
                                     appearance_embedding=appearance_embedding, chunk=chunk, mlp_dtype=mlp_dtype)
        rgb8, _ = quantize_frame(rgb)
        yield i, rgb8.cpu().numpy(), depth.cpu().numpy()


def render_aligned_spiral(model, height: int, width: int, focal: float, near: float, far: float, n_samples: int,
                          n_importance: int, appearance_embedding=None, num_frames: int = 120, loops: int = 2,
                          rotation_axis: str = "x", scene: str = "chair", **kw):
    """The frame loop of render_aligned_spiral.py:77-175 (camera path + per-frame render), without the file I/O."""
    poses = aligned_spiral_poses(num_frames, loops, rotation_axis, scene)
    return render_path(model, poses, height, width, focal, near, far, n_samples, n_importance,
                       appearance_embedding=appearance_embedding, **kw)
