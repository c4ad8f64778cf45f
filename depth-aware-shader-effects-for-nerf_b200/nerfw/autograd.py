"""torch.autograd glue: the MLP alone (NeRF.forward) and MLP + compositing (volume_render) as autograd Functions.

Forward runs in the requested MLP mode; backward is composite_bwd (K7) followed by the MLP backward (K5): the tcgen05
kernels after a tensor-core forward (bf16 operands, gates from the forward's ReLU masks), the fp32 CUDA-core kernel after
an fp32 forward.  Both recompute the activations tile by tile instead of storing them.
"""
from __future__ import annotations

from typing import Optional

import torch

import os

from . import ops
from ._lib import MLP_FP32


def backward_uses_tensor_cores(mode: int, emb) -> bool:
    """bf16 tcgen05 backward whenever the forward ran in a tensor-core mode (no, one shared, or one embedding row per
    ray / sample); NERFW_BWD_MODE=fp32 forces the fp32 CUDA-core backward."""
    if os.environ.get("NERFW_BWD_MODE", "bf16").lower() == "fp32":
        return False
    return mode != MLP_FP32


# Optional gradient sink: {parameter name: tensor}.  While set (nerfw.train.Trainer does, with views into its flat gradient
# buffer), the backward kernels accumulate straight into these tensors and autograd receives None for the parameters,
# which removes ~100 zero-fill / add launches per training step.  Plain `loss.backward()` users never see it.
_GRAD_SINK = None


class grad_sink:
    def __init__(self, sink: dict):
        self.sink = sink

    def __enter__(self):
        global _GRAD_SINK
        self.prev, _GRAD_SINK = _GRAD_SINK, self.sink
        return self

    def __exit__(self, *exc):
        global _GRAD_SINK
        _GRAD_SINK = self.prev
        return False


def _param_dict(names, tensors):
    return {n: t for n, t in zip(names, tensors)}


def _grad_targets(params):
    """Where the MLP backward kernels accumulate: straight into the Trainer's flat buffer (grad sink) or into fresh zeros."""
    sink = _GRAD_SINK
    direct = sink is not None and all(n in sink for n in params)
    grads = {n: sink[n] for n in params} if direct else {n: torch.zeros_like(t) for n, t in params.items()}
    return grads, direct


def _mlp_backward_into(params, grads, d_emb, packed, mode, p, d, z, emb, d_raw, masks):
    """One MLP backward launch (K5) accumulating into `grads` / `d_emb`."""
    if packed is not None and backward_uses_tensor_cores(mode, emb):
        ops.mlp_bwd_tc(params, grads, packed, p, d, z, emb, d_raw.contiguous(), d_emb, masks)
    else:
        ops.mlp_bwd(params, grads, p, d, z, emb, d_raw.contiguous(), d_emb)


def _mlp_backward(ctx, d_raw):
    names = ctx.names
    saved = ctx.saved_tensors
    p, d = saved[0], saved[1]
    k = 2
    z = None
    if ctx.has_z:
        z = saved[k]
        k += 1
    emb = None
    if ctx.has_emb:
        emb = saved[k]
        k += 1
    params = _param_dict(names, saved[k:k + len(names)])
    grads, direct = _grad_targets(params)
    d_emb = torch.zeros_like(emb) if emb is not None else None
    _mlp_backward_into(params, grads, d_emb, ctx.packed, ctx.mode, p, d, z, emb, d_raw, ctx.masks)
    if direct:
        grads = {n: None for n in params}   # already accumulated in place
    return grads, d_emb


class MlpFn(torch.autograd.Function):
    """raw (S,4) = NeRF MLP on samples.  args: mode, names, p, d, z|None, emb|None, packed|None, *params"""

    @staticmethod
    def forward(ctx, mode, names, p, d, z, emb, packed, *params):
        pd = _param_dict(names, params)
        ctx.masks = None
        if packed is not None and backward_uses_tensor_cores(mode, emb):
            raw, ctx.masks = ops.mlp_fwd(pd, packed, p, d, z, emb, mode, want_masks=True)
        else:
            raw = ops.mlp_fwd(pd, packed, p, d, z, emb, mode)
        ctx.names = names
        ctx.mode, ctx.packed = mode, packed
        ctx.has_z = z is not None
        ctx.has_emb = emb is not None
        keep = [p, d] + ([z] if z is not None else []) + ([emb] if emb is not None else []) + list(params)
        ctx.save_for_backward(*keep)
        return raw

    @staticmethod
    def backward(ctx, d_raw):
        grads, d_emb = _mlp_backward(ctx, d_raw)
        return (None, None, None, None, None, d_emb, None) + tuple(grads[n] for n in ctx.names)


class RenderFn(torch.autograd.Function):
    """(rgb_map, depth, acc, weights) = composite(MLP(o + d z)).  args: mode, names, o, d, z, emb|None, packed|None, *params"""

    @staticmethod
    def forward(ctx, mode, names, o, d, z, emb, packed, *params):
        pd = _param_dict(names, params)
        ctx.masks = None
        if packed is not None and backward_uses_tensor_cores(mode, emb):
            raw, ctx.masks = ops.mlp_fwd(pd, packed, o, d, z, emb, mode, want_masks=True)
        else:
            raw = ops.mlp_fwd(pd, packed, o, d, z, emb, mode)
        rgb, depth, acc, w = ops.composite_fwd(raw, z, want_weights=True)
        ctx.names = names
        ctx.mode, ctx.packed = mode, packed
        ctx.has_z = True
        ctx.has_emb = emb is not None
        ctx.set_materialize_grads(False)
        keep = [o, d, z] + ([emb] if emb is not None else []) + list(params) + [raw]
        ctx.save_for_backward(*keep)
        return rgb, depth, acc, w

    @staticmethod
    def backward(ctx, d_rgb, d_depth, d_acc, d_w):
        saved = ctx.saved_tensors
        raw = saved[-1]
        z = saved[2]
        if d_rgb is None:
            d_rgb = torch.zeros((z.shape[0], 3), dtype=torch.float32, device=z.device)
        cont = lambda t: None if t is None else t.contiguous().float()
        d_raw = ops.composite_bwd(raw, z, cont(d_rgb), cont(d_depth), cont(d_acc), cont(d_w))
        # reuse the MLP backward on everything but the trailing `raw`
        class _Ctx:
            pass
        c = _Ctx()
        c.names, c.has_z, c.has_emb = ctx.names, True, ctx.has_emb
        c.mode, c.packed, c.masks = ctx.mode, ctx.packed, ctx.masks
        c.saved_tensors = saved[:-1]
        grads, d_emb = _mlp_backward(c, d_raw)
        return (None, None, None, None, None, d_emb, None) + tuple(grads[n] for n in ctx.names)


class ReuseRenderFn(torch.autograd.Function):
    """Hierarchical render with ONE network for both passes, differentiable: the fine pass evaluates only the NI new
    depths; the coarse pass's (r,g,b,sigma) records are merged in at the N coarse depths (nerfw_merge_raw) and the
    backward scatters the merged row's gradient back to the two lists (nerfw_unmerge_raw), so every depth goes through
    the MLP forward and backward exactly once -- N + NI evaluations per ray instead of 2N + NI, the same gradients (the
    two-pass form evaluates the same function at the coarse depths twice and sums the two contributions).

    args: mode_c, mode_f, names, o, d, z (B,N), emb|None, packed|None, u_rand (B,NI), *params
    returns (rgb, depth, acc, w, rgb_c, depth_c, acc_c, w_c, z_all); the resampling sees detached coarse weights."""

    @staticmethod
    def forward(ctx, mode_c, mode_f, names, o, d, z, emb, packed, u_rand, *params):
        pd = _param_dict(names, params)
        tc_c = packed is not None and backward_uses_tensor_cores(mode_c, emb)
        tc_f = packed is not None and backward_uses_tensor_cores(mode_f, emb)
        ctx.masks_c = ctx.masks_f = None
        if tc_c:
            raw_c, ctx.masks_c = ops.mlp_fwd(pd, packed, o, d, z, emb, mode_c, want_masks=True)
        else:
            raw_c = ops.mlp_fwd(pd, packed, o, d, z, emb, mode_c)
        rgb_c, depth_c, acc_c, w_c = ops.composite_fwd(raw_c, z, want_weights=True)
        z_all, z_new = ops.sample_pdf(z, w_c, u_rand.shape[1], u_rand, want_zfine=True)
        if tc_f:
            raw_f, ctx.masks_f = ops.mlp_fwd(pd, packed, o, d, z_new, emb, mode_f, want_masks=True)
        else:
            raw_f = ops.mlp_fwd(pd, packed, o, d, z_new, emb, mode_f)
        raw = ops.merge_raw(z, raw_c, z_new, raw_f)
        rgb, depth, acc, w = ops.composite_fwd(raw, z_all, want_weights=True)
        ctx.names, ctx.packed, ctx.modes, ctx.has_emb = names, packed, (mode_c, mode_f), emb is not None
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(z_all)
        ctx.save_for_backward(*([o, d, z, z_new, z_all, raw_c, raw] + ([emb] if emb is not None else []) + list(params)))
        return rgb, depth, acc, w, rgb_c, depth_c, acc_c, w_c, z_all

    @staticmethod
    def backward(ctx, d_rgb, d_depth, d_acc, d_w, d_rgb_c, d_depth_c, d_acc_c, d_w_c, _d_z):
        saved = ctx.saved_tensors
        o, d, z, z_new, z_all, raw_c, raw = saved[:7]
        k = 7
        emb = None
        if ctx.has_emb:
            emb = saved[k]
            k += 1
        params = _param_dict(ctx.names, saved[k:k + len(ctx.names)])
        b = z.shape[0]
        cont = lambda t: None if t is None else t.contiguous().float()
        zeros3 = lambda: torch.zeros((b, 3), dtype=torch.float32, device=z.device)
        d_raw_c = None
        if any(t is not None for t in (d_rgb_c, d_depth_c, d_acc_c, d_w_c)):      # the coarse outputs were used too
            d_raw_c = ops.composite_bwd(raw_c, z, cont(d_rgb_c) if d_rgb_c is not None else zeros3(), cont(d_depth_c),
                                        cont(d_acc_c), cont(d_w_c))
        d_raw_m = ops.composite_bwd(raw, z_all, cont(d_rgb) if d_rgb is not None else zeros3(), cont(d_depth), cont(d_acc), cont(d_w))
        d_raw_c, d_raw_f = ops.unmerge_raw(z, z_new, d_raw_m, d_raw_c)
        grads, direct = _grad_targets(params)
        d_emb = torch.zeros_like(emb) if emb is not None else None
        mode_c, mode_f = ctx.modes
        _mlp_backward_into(params, grads, d_emb, ctx.packed, mode_c, o, d, z, emb, d_raw_c, ctx.masks_c)
        _mlp_backward_into(params, grads, d_emb, ctx.packed, mode_f, o, d, z_new, emb, d_raw_f, ctx.masks_f)
        if direct:
            grads = {n: None for n in params}
        return (None, None, None, None, None, None, d_emb, None, None) + tuple(grads[n] for n in ctx.names)
