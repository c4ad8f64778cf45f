"""get_rays / sample_stratified / sample_importance with the reference's signatures (src/ray_utils.py), on sm_100a."""
from __future__ import annotations

import torch

from . import ops


def _cuda_device(*tensors) -> torch.device:
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("nerfw needs a CUDA (sm_100) device; none is visible and there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def get_rays(height, width, focal_length, c2w):
    """src/ray_utils.py:4-50.  Returns (origins, directions), both (H,W,3) on c2w's device; directions are
    bit-identical to the reference's CPU result, origins are the same stride-0 expand of c2w[:3,3] (:48)."""
    if not isinstance(c2w, torch.Tensor):
        c2w = torch.as_tensor(c2w, dtype=torch.float32)
    dev = _cuda_device(c2w)
    _, dirs = ops.raygen(int(height), int(width), float(focal_length), c2w, dev, want_origins=False)
    if not c2w.is_cuda:
        dirs = dirs.cpu()  # the reference's callers on this path expect CPU tensors (src/dataset.py:230,257)
    origins = c2w[:3, 3].to(dirs.dtype).expand(dirs.shape)
    return origins, dirs


def sample_stratified(rays_o, rays_d, near, far, n_samples, perturb=True, *, t_rand=None, generator=None):
    """src/ray_utils.py:52-88.  Returns (z_vals (...,N), pts (...,N,3)).  `t_rand` injects the uniforms of :80."""
    src_dev = rays_o.device
    dev = _cuda_device(rays_o, rays_d)
    lead = list(rays_o.shape[:-1])
    o = rays_o.to(dev, torch.float32).reshape(-1, 3).contiguous()
    d = rays_d.to(dev, torch.float32).reshape(-1, 3).contiguous()
    b = o.shape[0]
    ztab = ops.depth_table(near, far, n_samples, dev)
    tr = None
    if perturb:
        tr = t_rand.to(dev).reshape(b, n_samples) if t_rand is not None else torch.rand((b, n_samples), device=dev, generator=generator)
    z, pts = ops.stratified(o, d, ztab, tr, b, want_pts=True)
    z = z.reshape(lead + [n_samples])
    pts = pts.reshape(lead + [n_samples, 3])
    if src_dev != dev:
        z, pts = z.to(src_dev), pts.to(src_dev)
    return z, pts


def sample_importance(rays_o, rays_d, z_vals, weights, n_importance, *, u_rand=None, generator=None):
    """src/ray_utils.py:90-149 (z-gather index clamped to N-1 where the reference raises).
    Returns (z_vals_combined (...,N+NI) ascending, pts_combined (...,N+NI,3))."""
    src_dev = z_vals.device
    dev = _cuda_device(rays_o, rays_d, z_vals, weights)
    lead = list(z_vals.shape[:-1])
    n = z_vals.shape[-1]
    o = rays_o.to(dev, torch.float32).reshape(-1, 3).contiguous()
    d = rays_d.to(dev, torch.float32).reshape(-1, 3).contiguous()
    z = z_vals.detach().to(dev, torch.float32).reshape(-1, n).contiguous()
    w = weights.detach().to(dev, torch.float32).reshape(-1, n).contiguous()
    b = z.shape[0]
    ur = u_rand.to(dev).reshape(b, n_importance) if u_rand is not None else torch.rand((b, n_importance), device=dev, generator=generator)
    z_all = ops.sample_pdf(z, w, int(n_importance), ur)
    pts = ops.ray_points(o, d, z_all)
    z_all = z_all.reshape(lead + [n + n_importance])
    pts = pts.reshape(lead + [n + n_importance, 3])
    if src_dev != dev:
        z_all, pts = z_all.to(src_dev), pts.to(src_dev)
    return z_all, pts
