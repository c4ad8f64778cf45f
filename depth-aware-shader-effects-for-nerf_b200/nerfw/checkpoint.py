"""Checkpoint files with the reference's schema.

save: src/train.py:113-125 -- {'model_state_dict', 'optimizer_state_dict', 'loss', 'psnr', 'iteration',
'appearance_embeddings'} as `checkpoint_{i:06d}.pt`; load: run.py:361-366 / render_aligned_spiral.py:242-248.
Because NeRF keeps the reference parameter layout, files written by either code base load in the other."""
from __future__ import annotations

import os
from typing import Optional

import torch


def save_checkpoint(save_dir: str, iteration: int, model, appearance_embeddings=None, optimizer_state: Optional[dict] = None,
                    loss: float = float("nan"), psnr: float = float("nan"), name: Optional[str] = None) -> str:
    """`optimizer_state`: torch.optim.Adam.state_dict() or nerfw.train.Trainer.state_dict() (tensors and numbers only)."""
    os.makedirs(save_dir, exist_ok=True)
    ckpt = {"model_state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()},
            "optimizer_state_dict": optimizer_state or {}, "loss": float(loss), "psnr": float(psnr), "iteration": int(iteration)}
    if appearance_embeddings is not None:
        ckpt["appearance_embeddings"] = appearance_embeddings.detach().cpu()
    path = os.path.join(save_dir, name or f"checkpoint_{iteration:06d}.pt")   # 'checkpoint_final.pt': src/train.py:182
    torch.save(ckpt, path)
    return path


def load_checkpoint(path: str, model, appearance_embeddings=None, map_location=None, trainer=None) -> dict:
    """Loads model weights (strict) and, if given, the embedding table IN PLACE (`copy_`, so a Trainer's flat-buffer views
    stay attached) after a shape check.  The file is read with weights_only=True: the schema holds tensors and numbers
    only, so nothing is unpickled as code.  `trainer`: also restores its Adam moments / step count when the file has them."""
    ckpt = torch.load(path, map_location=map_location or "cpu", weights_only=True)
    sd = ckpt["model_state_dict"]
    with torch.no_grad():
        own = model.state_dict()
        missing = [k for k in own if k not in sd]
        unexpected = [k for k in sd if k not in own]
        if missing or unexpected:
            raise RuntimeError(f"checkpoint does not match the model: missing {missing}, unexpected {unexpected}")
        for k, v in own.items():
            if tuple(v.shape) != tuple(sd[k].shape):
                raise RuntimeError(f"checkpoint tensor {k} has shape {tuple(sd[k].shape)}, the model expects {tuple(v.shape)}")
            v.copy_(sd[k])                      # in place: parameters keep their storage (and any flat-buffer aliasing)
        if appearance_embeddings is not None and ckpt.get("appearance_embeddings") is not None:
            src = ckpt["appearance_embeddings"]
            if tuple(src.shape) != tuple(appearance_embeddings.shape):
                raise RuntimeError(f"checkpoint embedding table is {tuple(src.shape)}, expected {tuple(appearance_embeddings.shape)}")
            appearance_embeddings.copy_(src)
    if hasattr(model, "invalidate_packed"):
        model.invalidate_packed()
    if trainer is not None and ckpt.get("optimizer_state_dict"):
        trainer.load_state_dict(ckpt["optimizer_state_dict"])
    return ckpt
