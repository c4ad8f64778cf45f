"""volume_render with the reference's signature (src/render.py:5-97): sampling -> MLP -> compositing on sm_100a.

Differences from the reference, all additive:
  * `n_importance > 0` really runs the hierarchical fine pass (the reference's branch body is `pass`, F1); set
    `fine_pass=False` (or NERFW_COARSE_ONLY=1) to reproduce the reference's coarse-only output exactly.
  * extras gains 'acc' (sum of weights, F4) and, for the fine pass, the coarse outputs.
  * `model` may be one NeRF (shared coarse/fine weights, F3) or a (coarse, fine) pair.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import ops
from .autograd import RenderFn, ReuseRenderFn
from .models import NeRF, resolve_mode


def _shade(model: NeRF, o, d, z, emb, mode, role="single", sigma_only=False):
    """(rgb (B,3), depth (B,1), acc (B,1), weights (B,N)) for given depths -- src/render.py:29-80."""
    names, tensors, ws = model.kernel_state()[:3]
    training = torch.is_grad_enabled() and (any(t.requires_grad for t in tensors) or (emb is not None and emb.requires_grad))
    mode_id = resolve_mode(mode or model.mlp_mode, role, training)
    packed = model.packed_weights() if mode_id != 0 else None
    if training:
        return RenderFn.apply(mode_id, names, o, d, z, emb, packed, *tensors)
    raw = ops.mlp_fwd(ws, packed, o, d, z, None if emb is None else emb.detach(), mode_id, sigma_only=sigma_only,
                      app_ws=model.app_workspace(emb, packed))
    return ops.composite_fwd(raw, z, want_weights=True)


def _render_reusing_coarse(model: NeRF, o, d, z, emb, mode, n_importance, u_rand, generator, orig_shape, src_dev):
    """Hierarchical render with one network: coarse pass on z (full outputs), resampling, fine pass on the NEW depths
    only, merge of both sets of (r,g,b,sigma) records into depth order, compositing of the merged row.  Differentiable
    (ReuseRenderFn) when gradients are being recorded."""
    dev, b = o.device, o.shape[0]
    names, tensors, params = model.kernel_state()[:3]
    training = torch.is_grad_enabled() and (any(t.requires_grad for t in tensors) or (emb is not None and emb.requires_grad))
    mode_c, mode_f = resolve_mode(mode or model.mlp_mode, "coarse", training), resolve_mode(mode or model.mlp_mode, "fine", training)
    packed = model.packed_weights() if (mode_c != 0 or mode_f != 0) else None
    ur = u_rand.to(dev).reshape(b, n_importance) if u_rand is not None else torch.rand((b, n_importance), device=dev, generator=generator)
    if training:
        rgb, depth, acc, w, rgb_c, depth_c, acc_c, w_c, z_all = ReuseRenderFn.apply(mode_c, mode_f, names, o, d, z, emb, packed,
                                                                                   ur.float().contiguous(), *tensors)
    else:
        e = None if emb is None else emb.detach()
        app_ws = model.app_workspace(emb, packed)
        raw_c = ops.mlp_fwd(params, packed, o, d, z, e, mode_c, app_ws=app_ws)
        rgb_c, depth_c, acc_c, w_c = ops.composite_fwd(raw_c, z, want_weights=True)
        z_all, z_new = ops.sample_pdf(z, w_c, n_importance, ur, want_zfine=True)       # src/ray_utils.py:90-149
        raw_f = ops.mlp_fwd(params, packed, o, d, z_new, e, mode_f, app_ws=app_ws)
        raw = ops.merge_raw(z, raw_c, z_new, raw_f)
        rgb, depth, acc, w = ops.composite_fwd(raw, z_all, want_weights=True)
    extras = {"rgb_coarse": rgb_c.reshape(*orig_shape[:-1], 3), "depth_coarse": depth_c.reshape(*orig_shape[:-1], 1),
              "acc_coarse": acc_c, "weights_coarse": w_c.unsqueeze(-1), "z_vals_coarse": z,
              "weights": w.unsqueeze(-1), "z_vals": z_all, "acc": acc}
    rgb_map = rgb.reshape(*orig_shape[:-1], 3)
    depth_map = depth.reshape(*orig_shape[:-1], 1)
    if src_dev != dev:
        rgb_map, depth_map = rgb_map.to(src_dev), depth_map.to(src_dev)
        extras = {k: v.to(src_dev) for k, v in extras.items()}
    return rgb_map, depth_map, extras


def _records_gradients(model: NeRF, emb) -> bool:
    if not torch.is_grad_enabled():
        return False
    tensors = model.kernel_state()[1]
    return any(t.requires_grad for t in tensors) or (emb is not None and emb.requires_grad)


def _render_fused(model: NeRF, o, d, ztab, tr, n_importance, u_rand, generator, emb, mode, orig_shape, src_dev):
    """Inference through nerfw_volume_render (one call): single pass (n_importance == 0) or coarse + fine with re-use."""
    dev, b = o.device, o.shape[0]
    params = model.kernel_state()[2]
    hier = n_importance > 0
    m = mode or model.mlp_mode
    mode_c = resolve_mode(m, "coarse" if hier else "single", False)
    mode_f = resolve_mode(m, "fine", False) if hier else mode_c
    packed = model.packed_weights() if (mode_c != 0 or mode_f != 0) else None
    ur = None
    if hier:
        ur = u_rand.to(dev).reshape(b, n_importance) if u_rand is not None else torch.rand((b, n_importance), device=dev, generator=generator)
        ur = ur.float().contiguous()
    if tr is not None:
        tr = tr.float().contiguous()
    r = ops.volume_render_fused(params, packed, o, d, ztab, tr, n_importance, ur, None if emb is None else emb.detach(),
                                mode_c, mode_f)
    extras = {"weights": r["weights"].unsqueeze(-1), "z_vals": r["z_vals"], "acc": r["acc"]}
    if hier:
        extras.update({"rgb_coarse": r["rgb_coarse"].reshape(*orig_shape[:-1], 3),
                       "depth_coarse": r["depth_coarse"].reshape(*orig_shape[:-1], 1), "acc_coarse": r["acc_coarse"],
                       "weights_coarse": r["weights_coarse"].unsqueeze(-1), "z_vals_coarse": r["z_coarse"]})
    rgb_map = r["rgb"].reshape(*orig_shape[:-1], 3)
    depth_map = r["depth"].reshape(*orig_shape[:-1], 1)
    if src_dev != dev:
        rgb_map, depth_map = rgb_map.to(src_dev), depth_map.to(src_dev)
        extras = {k: v.to(src_dev) for k, v in extras.items()}
    return rgb_map, depth_map, extras


def volume_render(model, rays_o, rays_d, near, far, n_samples, n_importance,
                  appearance_embedding=None, background_color=None, perturb=True, *,
                  mlp_dtype: Optional[str] = None, fine_pass: Optional[bool] = None, generator=None,
                  t_rand=None, u_rand=None, coarse_rgb: Optional[bool] = None, reuse_coarse: Optional[bool] = None,
                  fused: Optional[bool] = None):
    """Returns (rgb_map (...,3), depth_map (...,1), extras) like src/render.py:92-97.

    `background_color` is accepted and ignored, as in the reference (src/render.py:6).  `t_rand` (B,N) / `u_rand`
    (B,NI) inject the uniforms the reference draws at src/ray_utils.py:80 and :119 (otherwise torch.rand on the device,
    in that order).  `coarse_rgb` (two-pass form only): whether the coarse pass of a hierarchical render also evaluates
    colour (extras['rgb_coarse']); default: only when gradients are recorded (the training loss uses it) -- in two-pass
    inference the coarse pass only has to place the fine samples, so its direction layer and rgb head are skipped.  With
    coarse re-use (below) the coarse records are part of the result, so colour is always evaluated and returned.
    `reuse_coarse` (default ON for inference with ONE network for both passes -- the only case the reference has,
    src/train.py:30; NERFW_REUSE_COARSE=0 or reuse_coarse=False turns it off): the network's value at a depth does not
    depend on which pass asks, so the fine pass evaluates only the n_importance NEW depths and the coarse pass's
    (r,g,b,sigma) records are merged in at the n_samples coarse depths: 192 instead of 256 MLP evaluations per ray at
    64 + 128.  Bit-identical to the two-pass form in fp32 / bf16x3; in "mixed" the coarse records keep their bf16x3
    values (closer to fp32 than the fp16 re-evaluation they replace).  Under autograd the same holds for the backward
    (ReuseRenderFn: every depth goes through the MLP backward once; the gradients equal the two-pass form's, where the
    coarse depths are evaluated twice and the two contributions summed).  (coarse, fine) pairs always evaluate every
    depth in the fine pass.
    `fused` (default ON; NERFW_FUSED_RENDER=0 or fused=False turns it off): when no gradient is recorded, a single-pass or
    re-use render is ONE library call (nerfw_volume_render: the same launches in the same order, bit-identical outputs,
    a sixth of the host time per call -- what the reference's 4096-ray chunk loop is made of)."""
    coarse, fine = (model if isinstance(model, (tuple, list)) else (model, model))
    if fine_pass is None:
        fine_pass = os.environ.get("NERFW_COARSE_ONLY", "0") != "1"
    dev = coarse.rgb_linear.weight.device
    if dev.type != "cuda":
        raise RuntimeError(f"volume_render: the model is on {dev}; move it to a CUDA (sm_100) device -- there is no CPU path")
    src_dev = rays_o.device
    orig_shape = rays_o.shape
    o = rays_o.to(dev, torch.float32).reshape(-1, 3).contiguous()
    d = rays_d.to(dev, torch.float32).reshape(-1, 3).contiguous()
    b = o.shape[0]
    emb = coarse._prep_emb(appearance_embedding, b, dev)                   # src/render.py:33-46
    ztab = ops.depth_table(near, far, n_samples, dev)
    tr = None
    if perturb:
        tr = t_rand.to(dev).reshape(b, n_samples) if t_rand is not None else torch.rand((b, n_samples), device=dev, generator=generator)
    hier = n_importance > 0 and fine_pass
    if coarse_rgb is None:
        coarse_rgb = torch.is_grad_enabled()
    if reuse_coarse is None:
        reuse_coarse = os.environ.get("NERFW_REUSE_COARSE", "1") != "0"
    reuse_coarse = bool(reuse_coarse) and hier and coarse is fine
    if fused is None:
        fused = os.environ.get("NERFW_FUSED_RENDER", "1") != "0"
    if fused and b > 0 and (reuse_coarse or not hier) and not _records_gradients(coarse, emb):
        return _render_fused(coarse, o, d, ztab, tr, int(n_importance) if hier else 0, u_rand, generator, emb, mlp_dtype,
                             orig_shape, src_dev)
    d = ops.normalize_dirs(d)                                              # src/render.py:19
    z, _ = ops.stratified(None, None, ztab, tr, b, want_pts=False)         # src/render.py:22 (pts never materialised)
    if reuse_coarse:
        return _render_reusing_coarse(coarse, o, d, z, emb, mlp_dtype, int(n_importance), u_rand, generator, orig_shape, src_dev)
    sigma_only = hier and not coarse_rgb and not torch.is_grad_enabled()
    rgb, depth, acc, w = _shade(coarse, o, d, z, emb, mlp_dtype, "coarse" if hier else "single", sigma_only)
    extras = {}
    if hier:
        ur = u_rand.to(dev).reshape(b, n_importance) if u_rand is not None else torch.rand((b, n_importance), device=dev, generator=generator)
        z_all = ops.sample_pdf(z, w.detach(), int(n_importance), ur)       # src/ray_utils.py:90-149
        extras.update({"depth_coarse": depth.reshape(*orig_shape[:-1], 1),
                       "acc_coarse": acc, "weights_coarse": w.unsqueeze(-1), "z_vals_coarse": z})
        if not sigma_only:
            extras["rgb_coarse"] = rgb.reshape(*orig_shape[:-1], 3)
        z = z_all
        rgb, depth, acc, w = _shade(fine, o, d, z, emb, mlp_dtype, "fine")
    extras.update({"weights": w.unsqueeze(-1), "z_vals": z, "acc": acc})
    rgb_map = rgb.reshape(*orig_shape[:-1], 3)
    depth_map = depth.reshape(*orig_shape[:-1], 1)
    if src_dev != dev:
        rgb_map, depth_map = rgb_map.to(src_dev), depth_map.to(src_dev)
        extras = {k: v.to(src_dev) for k, v in extras.items()}
    return rgb_map, depth_map, extras
