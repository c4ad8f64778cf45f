"""The training step of src/train.py:54-92 on the sm_100a kernels, single GPU or ray-batch data parallel.

    trainer = Trainer(model, appearance_embeddings, lr=5e-4)
    loss = trainer.step(rays_o, rays_d, target_rgb, img_idx, near, far, n_samples, n_importance)

Per step: volume_render(perturb=True) -> mse -> backward (composite_bwd + fp32 MLP backward) -> one all-reduce of the
flat gradient buffer (world > 1) -> fused Adam over the flat parameter buffer.  The optimizer is torch.optim.Adam's
default update (src/train.py:33-39).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .autograd import grad_sink
from .parallel import FlatParams, shard_bounds, world_info
from .render import volume_render


class Trainer:
    def __init__(self, model, appearance_embeddings: Optional[torch.nn.Parameter] = None, lr: float = 5e-4,
                 betas=(0.9, 0.999), eps: float = 1e-8, group=None, mlp_dtype: Optional[str] = None,
                 coarse_loss: bool = True):
        self.model = model
        self.emb_table = appearance_embeddings
        self.group = group
        self.mlp_dtype = mlp_dtype
        self.coarse_loss = coarse_loss
        self.lr, self.betas, self.eps = lr, betas, eps
        params = [p for p in model.parameters()]
        if appearance_embeddings is not None:
            dev = params[0].device
            if appearance_embeddings.device != dev:
                appearance_embeddings.data = appearance_embeddings.data.to(dev)
            params.append(appearance_embeddings)
        self.flat = FlatParams(params)
        self.exp_avg = torch.zeros_like(self.flat.param)
        self.exp_avg_sq = torch.zeros_like(self.flat.param)
        self.step_count = 0

    def step(self, rays_o, rays_d, target, img_idx: Optional[int], near, far, n_samples, n_importance,
             perturb: bool = True, shard: bool = False, generator=None):
        """One optimisation step on this rank's rays.  shard=True: the arguments hold the GLOBAL batch and this rank
        takes its contiguous slice; otherwise they are already local.  Returns the (local) loss as a 1-element CUDA
        tensor without synchronising."""
        rank, world = world_info(self.group)
        o = rays_o.reshape(-1, 3)
        d = rays_d.reshape(-1, 3)
        t = target.reshape(-1, 3)
        if shard and world > 1:
            s, e = shard_bounds(o.shape[0], rank, world)
            o, d, t = o[s:e], d[s:e], t[s:e]
        dev = self.flat.param.device
        emb = None
        if self.emb_table is not None and img_idx is not None:
            emb = self.emb_table[img_idx]
        self.flat.zero_grad()
        rgb, _, extras = volume_render(self.model, o, d, near, far, n_samples, n_importance, appearance_embedding=emb,
                                       perturb=perturb, mlp_dtype=self.mlp_dtype, generator=generator)
        loss, d_rgb = ops.mse(rgb, t.to(dev))
        outs, grads = [rgb], [d_rgb]
        if self.coarse_loss and "rgb_coarse" in extras:     # canonical NeRF: the coarse network is trained too
            loss_c, d_c = ops.mse(extras["rgb_coarse"], t.to(dev))
            outs.append(extras["rgb_coarse"])
            grads.append(d_c)
            loss = loss + loss_c
        # the MLP backward kernels accumulate straight into the flat gradient buffer (views keyed by parameter name)
        with grad_sink({n: p.grad for n, p in self.model.named_parameters()}):
            torch.autograd.backward(outs, grads)
        self.flat.all_reduce(self.group)
        self.step_count += 1
        ops.adam_step(self.flat.param, self.flat.grad, self.exp_avg, self.exp_avg_sq, self.step_count, self.lr,
                      self.betas, self.eps, grad_scale=1.0 / world)
        self.model._packed_key = None   # parameters changed behind autograd's back: rebuild the bf16 image lazily
        return loss
