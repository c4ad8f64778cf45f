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
    sink = _GRAD_SINK
    direct = sink is not None and all(n in sink for n in params)
    grads = {n: sink[n] for n in params} if direct else {n: torch.zeros_like(t) for n, t in params.items()}
    d_emb = torch.zeros_like(emb) if emb is not None else None
    if ctx.packed is not None and backward_uses_tensor_cores(ctx.mode, emb):
        ops.mlp_bwd_tc(params, grads, ctx.packed, p, d, z, emb, d_raw.contiguous(), d_emb, ctx.masks)
    else:
        ops.mlp_bwd(params, grads, p, d, z, emb, d_raw.contiguous(), d_emb)
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
