"""Camera paths of the reference's render drivers, restated: the aligned spiral of render_aligned_spiral.py:27-122.

Returns camera-to-world matrices only; the per-frame work (rays, sampling, MLP, compositing) is nerfw.frame.render_frame.
"""
from __future__ import annotations

import math

import numpy as np

_AXIS_ROT = {
    "x": lambda c, s: np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64),
    "y": lambda c, s: np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64),
    "z": lambda c, s: np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float64),
}


def look_at(cam_pos: np.ndarray, center: np.ndarray, up: np.ndarray) -> np.ndarray:
    """4x4 c2w with columns (right, camera_up, -forward, position); degenerate cases as render_aligned_spiral.py:92-115."""
    fwd = center - cam_pos
    n = np.linalg.norm(fwd)
    fwd = np.array([0.0, 0.0, -1.0]) if n < 1e-10 else fwd / n
    right = np.cross(fwd, up)
    n = np.linalg.norm(right)
    right = np.array([1.0, 0.0, 0.0]) if n < 1e-10 else right / n
    cup = np.cross(right, fwd)
    n = np.linalg.norm(cup)
    cup = up if n < 1e-10 else cup / n
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, cup, -fwd, cam_pos
    return m


def aligned_spiral_poses(num_frames: int = 120, loops: int = 2, rotation_axis: str = "x", scene: str = "chair",
                         radius: float = 4.0) -> np.ndarray:
    """(num_frames,4,4) float32 poses: theta in [0, 2 pi loops], height phi in [-0.3, 0.3], 90-degree alignment
    rotation about `rotation_axis` ('none' = identity), chair centre raised to y = 0.5 (render_aligned_spiral.py:27-76)."""
    if rotation_axis in _AXIS_ROT:
        align = _AXIS_ROT[rotation_axis](math.cos(math.pi / 2), math.sin(math.pi / 2))
    else:
        align = np.eye(3)
    center = np.array([0.0, 0.5, 0.0]) if (rotation_axis == "x" and scene == "chair") else np.zeros(3)
    up = align @ np.array([0.0, 1.0, 0.0])
    theta = np.linspace(0, 2 * math.pi * loops, num_frames)
    phi = np.linspace(-0.3, 0.3, num_frames)
    poses = np.empty((num_frames, 4, 4), dtype=np.float32)
    for i in range(num_frames):
        base = np.array([radius * math.sin(theta[i]), phi[i] * radius, radius * math.cos(theta[i])])
        poses[i] = look_at(align @ base, center, up).astype(np.float32)
    return poses


def blender_focal(width: int, camera_angle_x: float = 0.6911112070083618) -> float:
    """focal = 0.5 W / tan(0.5 camera_angle_x) (src/dataset.py:66; the Blender scenes' field of view)."""
    return 0.5 * width / np.tan(0.5 * camera_angle_x)


def path_poses(camera_path: str = "circle", num_frames: int = 120, scene: str = "lego", spiral_loops: float = 2.0,
               height_range=(-0.5, 0.5), radius: float = 4.0) -> np.ndarray:
    """(num_frames,4,4) float32 poses of run.py::render_path (run.py:113-196): 'circle', 'spiral', 'horizontal_only'
    (camera at (r sin t, height, r cos t)) and 'hemisphere' (golden-angle spiral over the upper hemisphere), looking at
    the scene's centre -- (0, 0.5, 0) with Z up for lego, (0, 0.5, 0) for chair, the origin otherwise.  Unlike
    look_at(), this driver normalises without the degenerate-case guards (run.py:182-189)."""
    center = np.array([0.0, 0.0, 0.0])
    up = np.array([0.0, 1.0, 0.0])
    if scene == "lego":
        center, up = np.array([0.0, 0.5, 0.0]), np.array([0.0, 0.0, 1.0])
    elif scene == "chair":
        center = np.array([0.0, 0.5, 0.0])
    if camera_path == "circle":
        theta = np.linspace(0, 2 * np.pi, num_frames)
        heights = np.zeros_like(theta) + (0.5 if scene == "lego" else 0.0)
        phi = np.zeros_like(theta)
    elif camera_path == "spiral":
        theta = np.linspace(0, 2 * np.pi * spiral_loops, num_frames)
        hr = [0.3, 0.7] if scene == "lego" else list(height_range)
        heights = np.linspace(hr[0], hr[1], num_frames)
        phi = np.zeros_like(theta)
    elif camera_path == "horizontal_only":
        theta = np.linspace(0, 2 * np.pi * spiral_loops, num_frames)
        heights = np.full_like(theta, 0.5)
        phi = np.zeros_like(theta)
    elif camera_path == "hemisphere":
        idx = np.arange(0, num_frames, dtype=float) + 0.5
        phi = np.arccos(1 - 2 * idx / num_frames) - np.pi / 2
        theta = np.pi * (1 + 5 ** 0.5) * idx
        heights = np.zeros_like(theta)
    else:
        raise ValueError(f"camera_path must be circle, spiral, horizontal_only or hemisphere, got {camera_path!r}")
    poses = np.empty((num_frames, 4, 4), dtype=np.float32)
    for i, angle in enumerate(theta):
        if camera_path == "hemisphere":
            pos = np.array([radius * np.cos(phi[i]) * np.sin(angle), radius * np.sin(phi[i]), radius * np.cos(phi[i]) * np.cos(angle)])
        else:
            pos = np.array([radius * np.sin(angle), heights[i], radius * np.cos(angle)])
        fwd = center - pos
        fwd = fwd / np.linalg.norm(fwd)
        right = np.cross(fwd, up)
        right = right / np.linalg.norm(right)
        cup = np.cross(right, fwd)
        cup = cup / np.linalg.norm(cup)
        m = np.eye(4)
        m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, cup, -fwd, pos
        poses[i] = m.astype(np.float32)
    return poses
