"""Checkpoint files with the reference's schema.

save: src/train.py:113-125 -- {'model_state_dict', 'optimizer_state_dict', 'loss', 'psnr', 'iteration',
'appearance_embeddings'} as `checkpoint_{i:06d}.pt`; load: run.py:361-366 / render_aligned_spiral.py:242-248.
Because NeRF keeps the reference parameter layout, files written by either code base load in the other."""
from __future__ import annotations

import os
from typing import Optional

import torch


def save_checkpoint(save_dir: str, iteration: int, model, appearance_embeddings=None, optimizer_state: Optional[dict] = None,
                    loss: float = float("nan"), psnr: float = float("nan")) -> str:
    os.makedirs(save_dir, exist_ok=True)
    ckpt = {"model_state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
            "optimizer_state_dict": optimizer_state or {}, "loss": float(loss), "psnr": float(psnr), "iteration": int(iteration)}
    if appearance_embeddings is not None:
        ckpt["appearance_embeddings"] = appearance_embeddings.detach().cpu()
    path = os.path.join(save_dir, f"checkpoint_{iteration:06d}.pt")
    torch.save(ckpt, path)
    return path


def load_checkpoint(path: str, model, appearance_embeddings=None, map_location=None) -> dict:
    ckpt = torch.load(path, map_location=map_location or "cpu", weights_only=False)
    model.load_state_dict(ckpt["model_state_dict"], strict=True)
    if appearance_embeddings is not None and "appearance_embeddings" in ckpt:
        with torch.no_grad():
            appearance_embeddings.data = ckpt["appearance_embeddings"].to(appearance_embeddings.device)
    return ckpt
