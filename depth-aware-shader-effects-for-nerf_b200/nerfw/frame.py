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

import os

from . import ops
from .camera import aligned_spiral_poses, path_poses
from .io import FrameWriter, stage_to_host
from .parallel import frames_for_rank, world_info
from .render import volume_render


@torch.no_grad()
def render_frame(model, height: int, width: int, focal: float, c2w, near: float, far: float, n_samples: int,
                 n_importance: int, appearance_embedding=None, chunk: Optional[int] = None,
                 mlp_dtype: Optional[str] = None, generator=None, perturb: bool = False,
                 fine_pass: Optional[bool] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(rgb (H,W,3) float32, depth (H,W) float32, acc (H,W) float32) on the model's device."""
    coarse = model[0] if isinstance(model, (tuple, list)) else model
    dev = coarse.rgb_linear.weight.device
    c2w = torch.as_tensor(c2w, dtype=torch.float32)
    # the 12 pose floats travel as kernel arguments (nerfw_raygen takes a HOST matrix): no H2D copy, no sync
    o, d = ops.raygen(int(height), int(width), float(focal), c2w, dev, want_origins=True)
    o, d = o.reshape(-1, 3), d.reshape(-1, 3)
    n = o.shape[0]
    step = n if not chunk else int(chunk)
    rgbs, depths, accs = [], [], []
    for s in range(0, n, step):
        rgb, depth, ex = volume_render(model, o[s:s + step], d[s:s + step], near, far, n_samples, n_importance,
                                       appearance_embedding=appearance_embedding, perturb=perturb, mlp_dtype=mlp_dtype,
                                       generator=generator, fine_pass=fine_pass)
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
    """Yields (frame index, rgb uint8 (H,W,3), depth float32 (H,W)) for the frames this rank owns (i mod world == rank).
    Software-pipelined by one frame: frame i's device-to-host copy runs on a copy stream into pinned memory while frame
    i+1's kernels execute; the host only waits for a copy when it hands that frame out."""
    rank, world = world_info(group)
    coarse = model[0] if isinstance(model, (tuple, list)) else model
    copy_stream = torch.cuda.Stream(device=coarse.rgb_linear.weight.device)
    pending = None
    for i in frames_for_rank(len(poses), rank, world):
        rgb, depth, _ = render_frame(model, height, width, focal, poses[i], near, far, n_samples, n_importance,
                                     appearance_embedding=appearance_embedding, chunk=chunk, mlp_dtype=mlp_dtype)
        rgb8, _ = quantize_frame(rgb)
        staged = (i, stage_to_host(rgb8, copy_stream), stage_to_host(depth, copy_stream))
        if pending is not None:
            j, (h_rgb, e_rgb), (h_depth, e_depth) = pending
            e_depth.synchronize()
            yield j, h_rgb.numpy(), h_depth.numpy()
        pending = staged
    if pending is not None:
        j, (h_rgb, e_rgb), (h_depth, e_depth) = pending
        e_depth.synchronize()
        yield j, h_rgb.numpy(), h_depth.numpy()


def render_aligned_spiral(model, height: int, width: int, focal: float, near: float, far: float, n_samples: int,
                          n_importance: int, appearance_embedding=None, num_frames: int = 120, loops: int = 2,
                          rotation_axis: str = "x", scene: str = "chair", **kw):
    """The frame loop of render_aligned_spiral.py:77-175 (camera path + per-frame render), without the file I/O."""
    poses = aligned_spiral_poses(num_frames, loops, rotation_axis, scene)
    return render_path(model, poses, height, width, focal, near, far, n_samples, n_importance,
                       appearance_embedding=appearance_embedding, **kw)


def _dataset_view(dataset, config, device):
    """What the reference drivers read from (dataset, config): H, W, focal, near, far, sample counts, embedding row 0."""
    emb = None
    if getattr(config, "use_appearance", False):
        emb = dataset.appearance_embeddings[0].detach().to(device)        # render_aligned_spiral.py:131-133
    return int(dataset.H), int(dataset.W), float(dataset.focal), float(dataset.near), float(dataset.far), emb


def render_spiral_to_dir(model, dataset, config, output_dir, num_frames=120, fps=60, loops=2, rotation_axis="x", group=None,
                         fine_pass=None, mlp_dtype=None, writer_workers=2):
    """Drop-in for render_aligned_spiral.render_aligned_spiral(model, dataset, config, output_dir, num_frames, fps, loops,
    rotation_axis) up to the video step (render_aligned_spiral.py:15-175): same camera path, same `output/<dir>/
    frame_%04d.png` for every frame and `depth_%04d.png` (min/max-normalised) for every 10th, so apply_all_shaders.py and
    create_video.py run on the directory unchanged.  One device-resident render per frame instead of 157 chunk calls with
    a host sync each; quantisation on the device; file encoding on side threads (nerfw.io.FrameWriter) overlapping the
    next frame's kernels.  Under torch.distributed each rank renders frames i = rank (mod world) into the same directory.
    Returns the list of files this rank wrote."""
    if not output_dir.startswith("output/"):
        output_dir = os.path.join("output", output_dir)               # render_aligned_spiral.py:21-22
    coarse = model[0] if isinstance(model, (tuple, list)) else model
    dev = coarse.rgb_linear.weight.device
    h, w, focal, near, far, emb = _dataset_view(dataset, config, dev)
    poses = aligned_spiral_poses(num_frames, loops, rotation_axis, getattr(config, "scene", ""))
    rank, world = world_info(group)
    with FrameWriter(output_dir, workers=writer_workers) as wr:
        for i in frames_for_rank(num_frames, rank, world):
            rgb, depth, _ = render_frame(model, h, w, focal, poses[i], near, far, config.num_samples, config.num_importance,
                                         appearance_embedding=emb, mlp_dtype=mlp_dtype, fine_pass=fine_pass)
            rgb8, depth8 = quantize_frame(rgb, depth if i % 10 == 0 else None)
            wr.png(f"frame_{i:04d}.png", rgb8)
            if depth8 is not None:
                wr.png(f"depth_{i:04d}.png", depth8)                   # render_aligned_spiral.py:169-175
        files = list(wr.files)
    return files


def render_path_to_dir(model, dataset, config, output_dir, num_frames=120, quality="high", width=800, height=800,
                       start_frame=0, end_frame=None, save_depth=False, raw_output=False, camera_path="circle",
                       spiral_loops=2.0, height_range=(-0.5, 0.5), group=None, mlp_dtype=None, generator=None):
    """The frame loop of run.py::render_path (run.py:63-269) without the tkinter shader editor and the matplotlib depth
    figure: camera paths circle / spiral / horizontal_only / hemisphere, quality presets (preview: half the coarse samples,
    no fine pass, no jitter; medium / high: jittered, run.py:90-105), focal scaled to the requested width (run.py:198-199),
    files `rgb_%03d.png`, `raw/rgb_%03d.png` (raw_output) and `raw/depth_%03d.npy` (save_depth, fp32 -- what a
    depth-aware post-process should read instead of an 8-bit PNG).  As in the reference, every one of the `num_frames`
    poses is rendered and frame i is NAMED start_frame + i; `end_frame` only appears in the reference's progress messages
    (run.py:161-166) and is accepted for signature compatibility."""
    os.makedirs(output_dir, exist_ok=True)
    coarse = model[0] if isinstance(model, (tuple, list)) else model
    dev = coarse.rgb_linear.weight.device
    _, w0, focal0, near, far, emb = _dataset_view(dataset, config, dev)
    n_samples = config.num_samples // 2 if quality == "preview" else config.num_samples
    n_importance = 0 if quality == "preview" else config.num_importance
    perturb = quality != "preview"
    poses = path_poses(camera_path, num_frames, getattr(config, "scene", ""), spiral_loops, height_range)
    focal = focal0 * (width / w0)
    rank, world = world_info(group)
    with FrameWriter(output_dir) as wr:
        for i in range(len(poses)):
            idx = start_frame + i
            if i % world != rank:
                continue
            rgb, depth, _ = render_frame(model, height, width, focal, poses[i], near, far, n_samples, n_importance,
                                         appearance_embedding=emb, mlp_dtype=mlp_dtype, perturb=perturb, generator=generator)
            rgb8, _ = quantize_frame(rgb)
            if raw_output:
                wr.png(os.path.join("raw", f"rgb_{idx:03d}.png"), rgb8)
            if save_depth:
                wr.npy(os.path.join("raw", f"depth_{idx:03d}.npy"), depth)
            wr.png(f"rgb_{idx:03d}.png", rgb8)
        files = list(wr.files)
    return files
