"""Tensor-level wrappers over the C ABI (include/nerfw.h).

Every function takes/returns torch CUDA tensors, allocates outputs with the torch caching allocator, enqueues on
torch's current stream and never synchronises.  Inputs that are not CUDA / float32 / contiguous are rejected or
made contiguous here; nothing in this module computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check, lib

_ZTAB_CACHE: dict = {}
_ULIN_CACHE: dict = {}


# Raw handles without torch's Python-level wrappers: torch.cuda.current_stream() builds a Stream object through several
# layers of device-index normalisation (~5 us), and every launch needs the handle.
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_raw_device = getattr(torch._C, "_cuda_getDevice", None)


def _stream() -> int:
    if _raw_stream is not None and _raw_device is not None:
        return _raw_stream(_raw_device())
    return torch.cuda.current_stream().cuda_stream


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def _on(device: torch.device):
    """Device guard for a launch: a no-op when `device` is already current (the common case -- entering
    torch.cuda.device() costs several microseconds per launch, which is what a 4096-ray chunk loop is made of)."""
    idx = device.index
    if idx is None or idx == (_raw_device() if _raw_device is not None else torch.cuda.current_device()):
        return _NO_GUARD
    return torch.cuda.device(device)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (got {t.device}); the nerfw kernels have no CPU path")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


_device_checked = set()


def require_device(device: torch.device) -> None:
    """Fail loudly unless `device` is an sm_100 GPU (no fallback)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx in _device_checked:
        return
    with torch.cuda.device(idx):
        check(lib().nerfw_check_device())
    _device_checked.add(idx)


def launch_count() -> int:
    return int(lib().nerfw_launch_count())


# ------------------------------------------------------------------------------------------------ rays
def raygen(height: int, width: int, focal: float, c2w: torch.Tensor, device: torch.device,
           want_origins: bool = True) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
    require_device(device)
    c = c2w.detach().to("cpu", torch.float32)
    if c.shape[-2:] not in ((4, 4), (3, 4)):
        raise ValueError(f"c2w must be (4,4) or (3,4), got {tuple(c2w.shape)}")
    host = (C.c_float * 12)(*c[:3, :4].reshape(-1).tolist())
    with _on(device):
        dirs = torch.empty((height, width, 3), dtype=torch.float32, device=device)
        origins = torch.empty((height, width, 3), dtype=torch.float32, device=device) if want_origins else None
        # `focal` arrives as a python/numpy double; the reference's tensor/scalar division rounds it to fp32 first
        check(lib().nerfw_raygen(int(height), int(width), float(torch.tensor(float(focal), dtype=torch.float32)),
                                 host, _ptr(origins), dirs.data_ptr(), _stream()))
    return origins, dirs


def normalize_dirs(d: torch.Tensor) -> torch.Tensor:
    d = _f32c(d, "rays_d")
    out = torch.empty_like(d)
    with _on(d.device):
        check(lib().nerfw_normalize_dirs(d.data_ptr(), d.numel() // 3, out.data_ptr(), _stream()))
    return out


def depth_table(near: float, far: float, n: int, device: torch.device) -> torch.Tensor:
    """near + linspace(0,1,N)*(far-near) computed with HOST torch (bit-identical to src/ray_utils.py:69-70), cached."""
    key = (float(near), float(far), int(n), device.index)
    tab = _ZTAB_CACHE.get(key)
    if tab is None:
        t = torch.linspace(0.0, 1.0, int(n))
        tab = (near + t * (far - near)).to(device)
        _ZTAB_CACHE[key] = tab
    return tab


def u_table(n_importance: int, device: torch.device) -> torch.Tensor:
    """linspace(0,1,NI+1)[:-1] from host torch (src/ray_utils.py:115), cached."""
    key = (int(n_importance), device.index)
    tab = _ULIN_CACHE.get(key)
    if tab is None:
        tab = torch.linspace(0.0, 1.0, int(n_importance) + 1)[:-1].contiguous().to(device)
        _ULIN_CACHE[key] = tab
    return tab


def stratified(rays_o: Optional[torch.Tensor], rays_d: Optional[torch.Tensor], ztab: torch.Tensor,
               t_rand: Optional[torch.Tensor], n_rays: int, want_pts: bool) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    n = ztab.numel()
    dev = ztab.device
    z = torch.empty((n_rays, n), dtype=torch.float32, device=dev)
    pts = torch.empty((n_rays, n, 3), dtype=torch.float32, device=dev) if want_pts else None
    if t_rand is not None:
        t_rand = _f32c(t_rand, "t_rand")
        if tuple(t_rand.shape) != (n_rays, n):
            raise ValueError(f"t_rand must be ({n_rays},{n}), got {tuple(t_rand.shape)}")
    with _on(dev):
        check(lib().nerfw_stratified(_ptr(rays_o), _ptr(rays_d), ztab.data_ptr(), _ptr(t_rand), n_rays, n,
                                     z.data_ptr(), _ptr(pts), _stream()))
    return z, pts


def ray_points(rays_o: torch.Tensor, rays_d: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
    b, n = z.shape
    pts = torch.empty((b, n, 3), dtype=torch.float32, device=z.device)
    with _on(z.device):
        check(lib().nerfw_ray_points(rays_o.data_ptr(), rays_d.data_ptr(), z.data_ptr(), b, n, pts.data_ptr(), _stream()))
    return pts


def sample_pdf(z_vals: torch.Tensor, weights: torch.Tensor, n_importance: int, u_rand: torch.Tensor,
               want_aux: bool = False, want_zfine: bool = False, general_path: bool = False):
    z_vals = _f32c(z_vals, "z_vals")
    weights = _f32c(weights, "weights")
    u_rand = _f32c(u_rand, "u_rand")
    b, n = z_vals.shape
    if tuple(weights.shape) != (b, n):
        raise ValueError(f"weights must be ({b},{n}), got {tuple(weights.shape)}")
    if tuple(u_rand.shape) != (b, n_importance):
        raise ValueError(f"u_rand must be ({b},{n_importance}), got {tuple(u_rand.shape)}")
    dev = z_vals.device
    out = torch.empty((b, n + n_importance), dtype=torch.float32, device=dev)
    inds = zf = cdf = None
    if want_aux:
        inds = torch.empty((b, n_importance), dtype=torch.int64, device=dev)
        zf = torch.empty((b, n_importance), dtype=torch.float32, device=dev)
        cdf = torch.empty((b, n + 1), dtype=torch.float32, device=dev)
    if want_zfine and zf is None:
        zf = torch.empty((b, n_importance), dtype=torch.float32, device=dev)
    ulin = u_table(n_importance, dev)
    fn = lib().nerfw_sample_pdf_general if general_path else lib().nerfw_sample_pdf
    with _on(dev):
        check(fn(z_vals.data_ptr(), weights.data_ptr(), ulin.data_ptr(), u_rand.data_ptr(), b, n,
                 int(n_importance), out.data_ptr(), _ptr(inds), _ptr(zf), _ptr(cdf), _stream()))
    if want_aux:
        return out, {"inds": inds, "z_fine": zf, "cdf": cdf}
    if want_zfine:
        return out, zf
    return out


def merge_raw(z_coarse: torch.Tensor, raw_coarse: torch.Tensor, z_fine: torch.Tensor, raw_fine: torch.Tensor) -> torch.Tensor:
    """(B,N) / (B*N,4) coarse and (B,NI) / (B*NI,4) fine depths and MLP outputs -> raw (B*(N+NI),4) in merged depth order."""
    b, n = z_coarse.shape
    ni = z_fine.shape[1]
    out = torch.empty((b * (n + ni), 4), dtype=torch.float32, device=z_coarse.device)
    with _on(z_coarse.device):
        check(lib().nerfw_merge_raw(_f32c(z_coarse, "z_coarse").data_ptr(), _f32c(raw_coarse, "raw_coarse").data_ptr(),
                                    _f32c(z_fine, "z_fine").data_ptr(), _f32c(raw_fine, "raw_fine").data_ptr(), b, n, ni,
                                    out.data_ptr(), _stream()))
    return out


def unmerge_raw(z_coarse: torch.Tensor, z_fine: torch.Tensor, d_merged: torch.Tensor,
                d_coarse: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Backward of merge_raw: d_merged (B*(N+NI),4) -> (d_coarse (B*N,4), d_fine (B*NI,4)).  A given d_coarse is added to."""
    b, n = z_coarse.shape
    ni = z_fine.shape[1]
    acc = d_coarse is not None
    if d_coarse is None:
        d_coarse = torch.empty((b * n, 4), dtype=torch.float32, device=z_coarse.device)
    d_fine = torch.empty((b * ni, 4), dtype=torch.float32, device=z_coarse.device)
    with _on(z_coarse.device):
        check(lib().nerfw_unmerge_raw(z_coarse.data_ptr(), z_fine.data_ptr(), _f32c(d_merged, "d_merged").data_ptr(), b, n, ni,
                                      int(acc), d_coarse.data_ptr(), d_fine.data_ptr(), _stream()))
    return d_coarse, d_fine


def posenc(x: torch.Tensor, levels: int, include_input: bool = True) -> torch.Tensor:
    x = _f32c(x, "x")
    dim = x.shape[-1]
    lead = x.shape[:-1]
    flat = x.reshape(-1, dim)
    width = dim * ((1 if include_input else 0) + 2 * levels)
    out = torch.empty((flat.shape[0], width), dtype=torch.float32, device=x.device)
    with _on(x.device):
        check(lib().nerfw_posenc(flat.data_ptr(), flat.shape[0], dim, int(levels), int(bool(include_input)),
                                 out.data_ptr(), _stream()))
    return out.reshape(*lead, width)


# ------------------------------------------------------------------------------------------------ MLP
PARAM_ORDER = ([f"pts_linears.{i}.weight" for i in range(8)] + [f"pts_linears.{i}.bias" for i in range(8)] +
               ["density_head.weight", "density_head.bias", "dir_linear.weight", "dir_linear.bias",
                "appearance_projection.weight", "appearance_projection.bias", "rgb_linear.weight", "rgb_linear.bias"])

EXPECTED_SHAPES = {
    **{f"pts_linears.{i}.weight": (256, 63 if i == 0 else (319 if i == 4 else 256)) for i in range(8)},
    **{f"pts_linears.{i}.bias": (256,) for i in range(8)},
    "density_head.weight": (1, 256), "density_head.bias": (1,),
    "dir_linear.weight": (128, 283), "dir_linear.bias": (128,),
    "appearance_projection.weight": (128, 32), "appearance_projection.bias": (128,),
    "rgb_linear.weight": (3, 128), "rgb_linear.bias": (3,),
}


def weights_struct(params) -> _lib.NerfwWeights:
    """Fill NerfwWeights/NerfwGrads from {state_dict key: contiguous fp32 CUDA tensor}.  Missing appearance -> NULL.
    An already filled struct (NeRF.kernel_state() caches one per model) is passed through."""
    if isinstance(params, _lib.NerfwWeights):
        return params
    s = _lib.NerfwWeights()
    for i in range(8):
        s.pts_w[i] = params[f"pts_linears.{i}.weight"].data_ptr()
        s.pts_b[i] = params[f"pts_linears.{i}.bias"].data_ptr()
    s.density_w = params["density_head.weight"].data_ptr()
    s.density_b = params["density_head.bias"].data_ptr()
    s.dir_w = params["dir_linear.weight"].data_ptr()
    s.dir_b = params["dir_linear.bias"].data_ptr()
    aw = params.get("appearance_projection.weight")
    ab = params.get("appearance_projection.bias")
    s.app_w = aw.data_ptr() if aw is not None else None
    s.app_b = ab.data_ptr() if ab is not None else None
    s.rgb_w = params["rgb_linear.weight"].data_ptr()
    s.rgb_b = params["rgb_linear.bias"].data_ptr()
    return s


def check_params(params: dict) -> None:
    for k, t in params.items():
        exp = EXPECTED_SHAPES.get(k)
        if exp is None:
            raise ValueError(f"unexpected parameter {k}")
        if tuple(t.shape) != exp:
            raise ValueError(
                f"{k} has shape {tuple(t.shape)}; the sm_100a kernels are specialised for the reference architecture "
                f"(hidden 256, 8 layers, skip [4], L=10/4, appearance 32) and expect {exp}.  No fallback exists.")
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError(f"{k} must be a contiguous float32 CUDA tensor (got {t.dtype} on {t.device})")


_PACKED_BYTES = None
_WS_BYTES: dict = {}


def packed_bytes() -> int:
    global _PACKED_BYTES
    if _PACKED_BYTES is None:
        _PACKED_BYTES = int(lib().nerfw_packed_bytes())
    return _PACKED_BYTES


def _mlp_workspace_bytes(n_rays: int, emb_rows: int) -> int:
    key = (n_rays, emb_rows)
    v = _WS_BYTES.get(key)
    if v is None:
        if len(_WS_BYTES) > 4096:
            _WS_BYTES.clear()
        v = _WS_BYTES[key] = int(lib().nerfw_mlp_workspace_bytes(n_rays, emb_rows))
    return v


def pack_weights(params, out: Optional[torch.Tensor] = None, device=None) -> torch.Tensor:
    dev = device if device is not None else params["rgb_linear.weight"].device
    nbytes = packed_bytes()
    if out is None:
        out = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    ws = weights_struct(params)
    with _on(dev):
        check(lib().nerfw_pack_weights(C.byref(ws), out.data_ptr(), out.numel(), _stream()))
    return out


def mlp_fwd(params: dict, packed: Optional[torch.Tensor], p: torch.Tensor, d: torch.Tensor, z: Optional[torch.Tensor],
            emb: Optional[torch.Tensor], mode: int, want_masks: bool = False, sigma_only: bool = False,
            app_ws: Optional[list] = None):
    """raw (S,4) = (r,g,b,sigma).  p,d: (B,3) rays with z (B,N), or (S,3) samples with z None.
    want_masks (tensor-core modes): also return the ReLU gate words for nerfw_mlp_bwd_tc -> (raw, masks).
    sigma_only (tensor-core modes, inference): NERFW_MLP_SIGMA_ONLY -- raw = (0, 0, 0, sigma), direction layer skipped.
    app_ws: a one-element list the caller keeps for launches that share weights AND embedding rows (the coarse and fine
    launch of one render; the chunks of one frame): the first launch stores its workspace (holding the per-row rgb-logit
    offsets) in it, later ones pass NERFW_MLP_APP_CACHED and skip the offset kernel."""
    dev = p.device
    n_rays = p.shape[0]
    n_samples = z.shape[1] if z is not None else 1
    total = n_rays * n_samples
    raw = torch.empty((total, 4), dtype=torch.float32, device=dev)
    emb_rows = 0
    if emb is not None:
        emb_rows = emb.shape[0]
    flags = int(mode) | (0x100 if sigma_only and int(mode) != 0 else 0)
    wbytes = _mlp_workspace_bytes(n_rays, emb_rows)
    stream = _stream()
    # re-use only on the stream that wrote the offsets (stream order is what makes them visible to this launch)
    if app_ws is not None and emb is not None and app_ws[0] is not None and app_ws[0][1] == stream and app_ws[0][0].numel() >= wbytes:
        ws_buf = app_ws[0][0]
        flags |= 0x200
    else:
        ws_buf = torch.empty(wbytes, dtype=torch.uint8, device=dev)
        if app_ws is not None:
            app_ws[0] = (ws_buf, stream)
    ws = weights_struct(params)
    masks = None
    if want_masks:
        masks = torch.empty(int(lib().nerfw_mlp_mask_bytes(n_rays, n_samples)), dtype=torch.uint8, device=dev)
    with _on(dev):
        check(lib().nerfw_mlp_fwd(C.byref(ws), _ptr(packed), p.data_ptr(), d.data_ptr(), _ptr(z), _ptr(emb), emb_rows,
                                  n_rays, n_samples, flags, raw.data_ptr(), _ptr(masks), ws_buf.data_ptr(), ws_buf.numel(),
                                  stream))
    if want_masks:
        return raw, masks
    return raw


def mlp_bwd(params: dict, grads: dict, p: torch.Tensor, d: torch.Tensor, z: Optional[torch.Tensor],
            emb: Optional[torch.Tensor], d_raw: torch.Tensor, d_emb: Optional[torch.Tensor]) -> None:
    """Accumulates into `grads` (same keys as params) and d_emb."""
    dev = p.device
    n_rays = p.shape[0]
    n_samples = z.shape[1] if z is not None else 1
    emb_rows = emb.shape[0] if emb is not None else 0
    wbytes = int(lib().nerfw_mlp_bwd_workspace_bytes(n_rays, n_samples, emb_rows))
    ws_buf = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    ws = weights_struct(params)
    gs = weights_struct(grads)
    with _on(dev):
        check(lib().nerfw_mlp_bwd(C.byref(ws), p.data_ptr(), d.data_ptr(), _ptr(z), _ptr(emb), emb_rows, n_rays, n_samples,
                                  d_raw.data_ptr(), C.byref(gs), _ptr(d_emb), ws_buf.data_ptr(), wbytes, _stream()))


def mlp_bwd_tc(params: dict, grads: dict, packed: torch.Tensor, p: torch.Tensor, d: torch.Tensor,
               z: Optional[torch.Tensor], emb: Optional[torch.Tensor], d_raw: torch.Tensor,
               d_emb: Optional[torch.Tensor], masks: Optional[torch.Tensor] = None) -> None:
    """Tensor-core (bf16) backward; accumulates into `grads` and d_emb.  emb: None, (1,32) shared or (n_rays,32).
    masks: the ReLU gates returned by mlp_fwd(want_masks=True)."""
    dev = p.device
    n_rays = p.shape[0]
    n_samples = z.shape[1] if z is not None else 1
    emb_rows = emb.shape[0] if emb is not None else 0
    wbytes = int(lib().nerfw_mlp_bwd_tc_workspace_bytes(n_rays, n_samples))
    ws_buf = torch.empty(wbytes, dtype=torch.uint8, device=dev)
    ws = weights_struct(params)
    gs = weights_struct(grads)
    with _on(dev):
        check(lib().nerfw_mlp_bwd_tc(C.byref(ws), packed.data_ptr(), p.data_ptr(), d.data_ptr(), _ptr(z), _ptr(emb), emb_rows,
                                     n_rays, n_samples, d_raw.data_ptr(), _ptr(masks), C.byref(gs), _ptr(d_emb),
                                     ws_buf.data_ptr(), wbytes, _stream()))


# ------------------------------------------------------------------------------------------------ compositing
def composite_fwd(raw: torch.Tensor, z: torch.Tensor, want_weights: bool = True):
    b, n = z.shape
    dev = z.device
    rgb = torch.empty((b, 3), dtype=torch.float32, device=dev)
    depth = torch.empty((b, 1), dtype=torch.float32, device=dev)
    acc = torch.empty((b, 1), dtype=torch.float32, device=dev)
    w = torch.empty((b, n), dtype=torch.float32, device=dev) if want_weights else None
    with _on(dev):
        check(lib().nerfw_composite_fwd(raw.data_ptr(), z.data_ptr(), b, n, rgb.data_ptr(), depth.data_ptr(),
                                        acc.data_ptr(), _ptr(w), _stream()))
    return rgb, depth, acc, w


def composite_bwd(raw: torch.Tensor, z: torch.Tensor, d_rgb: torch.Tensor, d_depth: Optional[torch.Tensor],
                  d_acc: Optional[torch.Tensor], d_w: Optional[torch.Tensor]) -> torch.Tensor:
    b, n = z.shape
    d_raw = torch.empty((b * n, 4), dtype=torch.float32, device=z.device)
    with _on(z.device):
        check(lib().nerfw_composite_bwd(raw.data_ptr(), z.data_ptr(), b, n, d_rgb.data_ptr(), _ptr(d_depth), _ptr(d_acc),
                                        _ptr(d_w), d_raw.data_ptr(), _stream()))
    return d_raw


# ------------------------------------------------------------------------------------------------ whole inference path
_VR_WS_BYTES: dict = {}


def _pad4(n: int) -> int:
    return (n + 3) & ~3


def volume_render_fused(params, packed: Optional[torch.Tensor], rays_o: torch.Tensor, rays_d: torch.Tensor,
                        ztab: torch.Tensor, t_rand: Optional[torch.Tensor], n_importance: int,
                        u_rand: Optional[torch.Tensor], emb: Optional[torch.Tensor], mode_coarse: int, mode_fine: int) -> dict:
    """nerfw_volume_render: the inference form of src/render.py:5-97 in one library call (same kernels, same order, same
    bits as the call-by-call path).  rays_o / rays_d (B,3) fp32 contiguous CUDA, rays_d not normalised.  Returns the
    outputs as views of ONE allocation: rgb, depth, acc, weights, z_vals [+ *_coarse, z_coarse when n_importance > 0]."""
    dev = rays_o.device
    b = rays_o.shape[0]
    n = ztab.numel()
    ni = int(n_importance)
    emb_rows = emb.shape[0] if emb is not None else 0
    key = (b, n, ni, emb_rows)
    wbytes = _VR_WS_BYTES.get(key)
    if wbytes is None:
        if len(_VR_WS_BYTES) > 4096:
            _VR_WS_BYTES.clear()
        wbytes = _VR_WS_BYTES[key] = int(lib().nerfw_volume_render_workspace_bytes(b, n, ni, emb_rows))
    # one output allocation; every section starts on a 16-byte boundary
    sizes = [("rgb", 3 * b, (b, 3)), ("depth", b, (b, 1)), ("acc", b, (b, 1)), ("weights", b * (n + ni), (b, n + ni)),
             ("z_vals", b * (n + ni), (b, n + ni))]
    if ni > 0:
        sizes += [("rgb_coarse", 3 * b, (b, 3)), ("depth_coarse", b, (b, 1)), ("acc_coarse", b, (b, 1)),
                  ("weights_coarse", b * n, (b, n)), ("z_coarse", b * n, (b, n))]
    total = 0
    offs = []
    for _, cnt, _shape in sizes:
        offs.append(total)
        total += _pad4(cnt)
    flat = torch.empty(max(total, 1), dtype=torch.float32, device=dev)
    ws_buf = torch.empty(max(wbytes, 16), dtype=torch.uint8, device=dev)
    base = flat.data_ptr()
    out = _lib.NerfwRenderOut()
    for (name, _cnt, _shape), off in zip(sizes, offs):
        setattr(out, name, base + 4 * off)
    ws = weights_struct(params)
    ulin = u_table(ni, dev) if ni > 0 else None
    with _on(dev):
        check(lib().nerfw_volume_render(C.byref(ws), _ptr(packed), rays_o.data_ptr(), rays_d.data_ptr(), b, ztab.data_ptr(),
                                        _ptr(t_rand), n, _ptr(ulin), _ptr(u_rand), ni, _ptr(emb), emb_rows, int(mode_coarse),
                                        int(mode_fine), C.byref(out), ws_buf.data_ptr(), ws_buf.numel(), _stream()))
    # views are cut after the launches are queued (host time that overlaps the GPU's)
    return {name: flat[off:off + cnt].view(shape) for (name, cnt, shape), off in zip(sizes, offs)}


# ------------------------------------------------------------------------------------------------ training helpers
def adam_step(param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int,
              lr: float, betas=(0.9, 0.999), eps: float = 1e-8, grad_scale: float = 1.0) -> None:
    with _on(param.device):
        check(lib().nerfw_adam(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(),
                               param.numel(), lr, betas[0], betas[1], eps, int(step), grad_scale, _stream()))


def mse(rgb: torch.Tensor, target: torch.Tensor, loss_scale: float = 1.0, want_grad: bool = True):
    rgb = _f32c(rgb, "rgb")
    target = _f32c(target, "target")
    loss = torch.empty(1, dtype=torch.float32, device=rgb.device)
    d = torch.empty_like(rgb) if want_grad else None
    with _on(rgb.device):
        check(lib().nerfw_mse(rgb.data_ptr(), target.data_ptr(), rgb.numel(), loss_scale, loss.data_ptr(), _ptr(d), _stream()))
    return loss, d


def quantize_u8(x: torch.Tensor) -> torch.Tensor:
    x = _f32c(x, "x")
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    with _on(x.device):
        check(lib().nerfw_quantize_u8(x.data_ptr(), x.numel(), out.data_ptr(), _stream()))
    return out


def selftest_umma_mn(at: torch.Tensor, bt: torch.Tensor) -> torch.Tensor:
    """D = At^T Bt for At (K,128), Bt (K,N) bf16: both UMMA operands MN-major (the wgrad form)."""
    assert at.dtype == torch.bfloat16 and bt.dtype == torch.bfloat16 and at.shape[1] == 128 and at.shape[0] == bt.shape[0]
    at = at.contiguous()
    bt = bt.contiguous()
    d = torch.empty((128, bt.shape[1]), dtype=torch.float32, device=at.device)
    with _on(at.device):
        check(lib().nerfw_selftest_umma_mn(at.data_ptr(), bt.data_ptr(), bt.shape[1], at.shape[0], d.data_ptr(), _stream()))
    return d


def selftest_umma(a: torch.Tensor, b: torch.Tensor, mode: int) -> torch.Tensor:
    """D = A B^T for A (128,K) bf16, B (N,K) bf16 through one tcgen05 tile (tests)."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.shape[0] == 128
    a = a.contiguous()
    b = b.contiguous()
    d = torch.empty((128, b.shape[0]), dtype=torch.float32, device=a.device)
    with _on(a.device):
        check(lib().nerfw_selftest_umma(a.data_ptr(), b.data_ptr(), b.shape[0], a.shape[1], int(mode), d.data_ptr(), _stream()))
    return d


# ---- depth-aware effects --------------------------------------------------------------------------
def _u8c(x: torch.Tensor, name: str) -> torch.Tensor:
    if x.dtype != torch.uint8 or not x.is_cuda:
        raise ValueError(f"{name} must be a CUDA uint8 tensor")
    return x.contiguous()


def max_f32(x: torch.Tensor) -> torch.Tensor:
    """Device scalar max(x) (depth.max() of src/post_processor.py:64,408,476); no host sync."""
    x = _f32c(x, "x")
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    with _on(x.device):
        check(lib().nerfw_max_f32(x.data_ptr(), x.numel(), out.data_ptr(), _stream()))
    return out


def fog(image: torch.Tensor, depth: torch.Tensor, depth_max: torch.Tensor, fog_start: float, power: float,
        visibility: float, fog_color) -> torch.Tensor:
    image = _u8c(image, "image")
    depth = _f32c(depth, "depth")
    n = depth.numel()
    if image.numel() != 3 * n:
        raise ValueError(f"image must hold 3 values per depth pixel, got {tuple(image.shape)} vs {tuple(depth.shape)}")
    out = torch.empty_like(image)
    color = (C.c_float * 3)(*[float(c) for c in fog_color])
    with _on(depth.device):
        check(lib().nerfw_fog(image.data_ptr(), depth.data_ptr(), depth_max.data_ptr(), n, float(fog_start), float(power),
                              float(visibility), color, out.data_ptr(), _stream()))
    return out


def depth_edges(depth: torch.Tensor, depth_max: torch.Tensor, bilateral_d: int = 0, sigma_color: float = 75.0,
                sigma_space: float = 75.0):
    """-> (mag (H,W), mag_max (1,), filtered (H,W) or None)."""
    depth = _f32c(depth, "depth")
    if depth.dim() != 2:
        raise ValueError(f"depth must be (H,W), got {tuple(depth.shape)}")
    h, w = depth.shape
    mag = torch.empty_like(depth)
    mag_max = torch.empty(1, dtype=torch.float32, device=depth.device)
    filtered = torch.empty_like(depth) if bilateral_d else None
    with _on(depth.device):
        check(lib().nerfw_depth_edges(depth.data_ptr(), depth_max.data_ptr(), h, w, int(bilateral_d), float(sigma_color),
                                      float(sigma_space), _ptr(filtered), mag.data_ptr(), mag_max.data_ptr(), _stream()))
    return mag, mag_max, filtered


def toon(image: torch.Tensor, mag: torch.Tensor, mag_max: torch.Tensor, levels: int, edge_strength: float) -> torch.Tensor:
    image = _u8c(image, "image")
    h, w = mag.shape
    if tuple(image.shape) != (h, w, 3):
        raise ValueError(f"image must be ({h},{w},3), got {tuple(image.shape)}")
    out = torch.empty_like(image)
    with _on(image.device):
        check(lib().nerfw_toon(image.data_ptr(), mag.data_ptr(), mag_max.data_ptr(), h, w, int(levels), float(edge_strength),
                               out.data_ptr(), _stream()))
    return out


def hologram(image: torch.Tensor, mag, mag_max, row_scale: torch.Tensor, col_hits, noise) -> torch.Tensor:
    image = _u8c(image, "image")
    h, w, _ = image.shape
    row_scale = _f32c(row_scale, "row_scale")
    if row_scale.numel() != h:
        raise ValueError(f"row_scale must have {h} entries")
    if col_hits is not None and (col_hits.dtype != torch.int32 or col_hits.numel() != w):
        raise ValueError(f"col_hits must be int32 with {w} entries")
    if noise is not None:
        noise = _f32c(noise, "noise")
        if tuple(noise.shape) != (h, w, 3):
            raise ValueError(f"noise must be ({h},{w},3)")
    out = torch.empty_like(image)
    with _on(image.device):
        check(lib().nerfw_hologram(image.data_ptr(), _ptr(mag), _ptr(mag_max), row_scale.data_ptr(), _ptr(col_hits),
                                   _ptr(noise), h, w, out.data_ptr(), _stream()))
    return out
