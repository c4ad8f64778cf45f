"""Depth-aware effects on the renderer's fp32 buffers: fog.

The reference applies its effects to 8-bit PNGs read back from disk (apply_all_shaders.py:125-140), after the depth map
has been quantised to uint8 (render_aligned_spiral.py:171-175).  `fog` is the pointwise effect of
src/post_processor.py:451-493 evaluated directly on the device-resident float depth, as an epilogue of render_frame."""
from __future__ import annotations

import torch

from . import ops


def fog(rgb: torch.Tensor, depth: torch.Tensor, fog_start: float = 0.0, power: float = 3.0, visibility: float = 0.3,
        fog_color=(255.0, 255.0, 255.0)) -> torch.Tensor:
    """rgb (H,W,3) float in [0,1], depth (H,W) float -> uint8 (H,W,3).

    depth is normalised by its maximum when it exceeds 1 (:473-477); f = clip((d - start)/(1 - start), 0, 1)**power *
    visibility (:480-486); out = clip(rgb8 * f + fog_color * (1 - f)) (:491-493), rgb8 = (rgb*255) truncated to uint8."""
    d = depth
    if float(d.max()) > 1.0:
        d = d / d.max()
    f = (torch.clamp(d - fog_start, min=0.0) / (1.0 - fog_start)).clamp(0.0, 1.0) ** power * visibility
    rgb8 = ops.quantize_u8(rgb.contiguous()).float()
    color = torch.tensor(fog_color, dtype=torch.float32, device=rgb.device)
    out = rgb8 * f.unsqueeze(-1) + color * (1.0 - f.unsqueeze(-1))
    return out.clamp(0, 255).to(torch.uint8)
