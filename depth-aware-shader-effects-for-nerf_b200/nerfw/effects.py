"""Depth-aware effects on the renderer's device-resident buffers.

The reference applies its effects with numpy / OpenCV to 8-bit PNGs read back from disk (apply_all_shaders.py:125-140),
after the depth map has been quantised to uint8 (render_aligned_spiral.py:171-175).  Here the same effects run as CUDA
kernels (csrc/effects.cu) on the fp32 depth straight out of `render_frame`; images are uint8 (H,W,3) tensors like the
reference's effect inputs (`quantize_frame` produces them).  Only the depth-aware effects are covered: fog, toon (depth
edges) and hologram (depth-edge glow); the purely 2-D colour filters of src/post_processor.py are out of scope."""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import ops


def _image_u8(image: torch.Tensor) -> torch.Tensor:
    """uint8 (H,W,3) as is; float rgb in [0,1] is quantised like render_aligned_spiral.py:161-162."""
    if image.dtype == torch.uint8:
        return image.contiguous()
    return ops.quantize_u8(image.contiguous())


def _depth_2d(depth: torch.Tensor) -> torch.Tensor:
    if depth.dim() == 3:   # (H,W,1) from volume_render / (H,W,C): channel 0 (src/post_processor.py:410-411, :474-475)
        depth = depth[..., 0]
    return depth.contiguous().float()


def fog(image: torch.Tensor, depth: torch.Tensor, fog_start: float = 0.1, power: float = 3.0, visibility: float = 0.3,
        fog_color: Sequence[float] = (255.0, 255.0, 255.0)) -> torch.Tensor:
    """src/post_processor.py:451-493.  image (H,W,3) uint8 (or float in [0,1]), depth (H,W[,1]) float -> uint8 (H,W,3).
    fog_start defaults to the reference's effective value, params["fog_start"] = 0.1 (:45; the 0.0 at :460 is only the
    fallback of dict.get); power 3.0 and visibility 0.3 are the literals of :483 and :486."""
    img = _image_u8(image)
    d = _depth_2d(depth)
    return ops.fog(img, d, ops.max_f32(d), fog_start, power, visibility, fog_color)


def depth_edges(depth: torch.Tensor, bilateral: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Sobel gradient magnitude of the normalised depth (src/post_processor.py:72-74 / :414-416), optionally after
    cv2.bilateralFilter(depth, 9, 75, 75) (:69).  Returns (mag (H,W), mag_max (1,)) -- both stay on the device."""
    d = _depth_2d(depth)
    mag, mag_max, _ = ops.depth_edges(d, ops.max_f32(d), 9 if bilateral else 0, 75.0, 75.0)
    return mag, mag_max


def toon(image: torch.Tensor, depth: torch.Tensor, levels: int = 5, edge_strength: float = 1.0) -> torch.Tensor:
    """src/post_processor.py:64-102 with a depth map: colour quantisation + dilated depth edges."""
    img = _image_u8(image)
    mag, mag_max = depth_edges(depth, bilateral=True)
    return ops.toon(img, mag, mag_max, levels, edge_strength)


_SCANLINE_CACHE: dict = {}


def _scanline_table_on(device, height: int, num_lines: int) -> torch.Tensor:
    key = (str(device), int(height), int(num_lines))
    if key not in _SCANLINE_CACHE:
        _SCANLINE_CACHE[key] = scanline_table(height, num_lines).to(device)
    return _SCANLINE_CACHE[key]


def scanline_table(height: int, num_lines: int = 50) -> torch.Tensor:
    """Per-row factor of the hologram scanlines, built with the reference's loop (src/post_processor.py:385-393)."""
    line_height = height / num_lines
    rows = torch.ones(height, dtype=torch.float32)
    for i in range(num_lines):
        y_start = int(i * line_height)
        y_end = int(min((i + 0.7) * line_height, height))
        rows[y_start:y_end] *= 0.85
    return rows


def hologram(image: torch.Tensor, depth: Optional[torch.Tensor], num_lines: int = 50,
             noise: Optional[torch.Tensor] = None, lines: Sequence[Tuple[int, int]] = ()) -> torch.Tensor:
    """src/post_processor.py:373-449.  `noise` (H,W,3) replaces np.random.normal(0, 0.03, ...) (:399) and `lines` =
    [(x_pos, x_width), ...] the three np.random.randint draws of :443-446; both default to none (deterministic)."""
    img = _image_u8(image)
    h, w, _ = img.shape
    mag = mag_max = None
    if depth is not None:
        mag, mag_max = depth_edges(depth, bilateral=False)
    hits = None
    if lines:
        hits_host = torch.zeros(w, dtype=torch.int32)
        for x_pos, x_width in lines:
            hits_host[x_pos:min(x_pos + x_width, w)] += 1
        hits = hits_host.to(img.device)
    return ops.hologram(img, mag, mag_max, _scanline_table_on(img.device, h, num_lines), hits, noise)
