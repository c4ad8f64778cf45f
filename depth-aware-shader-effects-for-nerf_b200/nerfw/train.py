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

    def state_dict(self) -> dict:
        """Adam state for checkpoints (tensors and numbers only; goes under 'optimizer_state_dict')."""
        return {"kind": "nerfw.flat_adam", "step": int(self.step_count), "lr": float(self.lr), "betas": tuple(self.betas),
                "eps": float(self.eps), "exp_avg": self.exp_avg.detach().cpu(), "exp_avg_sq": self.exp_avg_sq.detach().cpu(),
                "offsets": list(self.flat.offsets)}

    def load_state_dict(self, state: dict) -> None:
        if state.get("kind") != "nerfw.flat_adam":
            raise ValueError("not a nerfw Trainer state (a torch.optim.Adam state_dict belongs to torch.optim.Adam)")
        if list(state["offsets"]) != list(self.flat.offsets):
            raise ValueError("optimizer state was saved for a different parameter layout")
        self.exp_avg.copy_(state["exp_avg"])
        self.exp_avg_sq.copy_(state["exp_avg_sq"])
        self.step_count = int(state["step"])
        self.lr, self.betas, self.eps = float(state["lr"]), tuple(state["betas"]), float(state["eps"])

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
            if isinstance(img_idx, torch.Tensor) and img_idx.dim() > 0:      # one image index per ray (RayBank cross_image)
                idx = img_idx.reshape(-1).to(dev)
                if shard and world > 1:
                    idx = idx[s:e]
                emb = self.emb_table[idx]                                    # (B,32): per-ray rows, src/render.py:39-44
            else:
                emb = self.emb_table[int(img_idx)]
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
        self.model.invalidate_packed()   # the fused Adam kernel wrote the parameters through data_ptr: no _version bump
        return loss
