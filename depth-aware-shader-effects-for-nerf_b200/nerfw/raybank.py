"""Device-resident training ray bank.

The reference decodes a PNG, builds the full image's rays on the CPU and draws `np.random.choice` pixels on every step
(src/dataset.py:206-277).  Here all rays of all images are generated once on the GPU (raygen kernel, bit-identical
directions) and a step's batch is a device-side gather: no host work per step.

    bank = RayBank(images, poses, focal)            # images (n,H,W,3|4) in [0,1] or uint8, poses (n,4,4)
    batch = bank.sample(4096)                       # dict with the keys of dataset.get_rays()
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


class RayBank:
    def __init__(self, images: torch.Tensor, poses: torch.Tensor, focal: float, device: Optional[torch.device] = None):
        if images.dim() != 4 or images.shape[-1] not in (3, 4):
            raise ValueError(f"images must be (n,H,W,3|4), got {tuple(images.shape)}")
        if poses.shape[0] != images.shape[0]:
            raise ValueError("one pose per image expected")
        dev = device or (images.device if images.is_cuda else torch.device("cuda", torch.cuda.current_device()))
        self.device = dev
        n, h, w, c = images.shape
        self.n_images, self.H, self.W, self.focal = n, h, w, float(focal)
        img = images.to(dev)
        if img.dtype == torch.uint8:
            img = img.float() / 255.0       # torchvision ToTensor scaling (src/dataset.py:142-150)
        self.rgb = img[..., :3].reshape(n, h * w, 3).contiguous()
        self.alpha = img[..., 3:4].reshape(n, h * w, 1).contiguous() if c == 4 else None
        self.origins = poses[:, :3, 3].to(dev, torch.float32).contiguous()              # (n,3): stride-0 per image
        dirs = []
        for i in range(n):
            _, d = ops.raygen(h, w, self.focal, poses[i], dev, want_origins=False)
            dirs.append(d.reshape(h * w, 3))
        self.dirs = torch.stack(dirs)                                                   # (n,H*W,3)

    def __len__(self) -> int:
        return self.n_images

    def image_rays(self, idx: int) -> dict:
        """All rays of image idx (dataset.get_rays(idx=...), src/dataset.py:221-246)."""
        return {"rays_o": self.origins[idx].expand(self.H * self.W, 3), "rays_d": self.dirs[idx], "rgb": self.rgb[idx],
                "alpha": None if self.alpha is None else self.alpha[idx], "appearance_idx": idx, "img_idx": idx}

    def sample(self, batch_size: int, generator: Optional[torch.Generator] = None, img_idx: Optional[int] = None,
               cross_image: bool = False) -> dict:
        """A training batch with the keys of dataset.get_rays() (src/dataset.py:248-277).

        cross_image=False (the reference's sampling): `batch_size` distinct random pixels of ONE random image;
        'appearance_idx' is that image's index (an int), so the step uses one shared embedding row.
        cross_image=True: `batch_size` random (image, pixel) pairs over the whole bank -- decorrelated batches, which the
        reference's one-PNG-per-step loader cannot produce; 'appearance_idx' is then an int64 tensor (B,) and the step
        gathers one embedding row per ray (the (B,D) case of src/render.py:39-44; forward and tcgen05 backward take
        emb_rows == n_rays)."""
        hw = self.H * self.W
        if cross_image:
            flat = torch.randint(0, self.n_images * hw, (batch_size,), generator=generator, device=self.device)
            img = torch.div(flat, hw, rounding_mode="floor")
            return {"rays_o": self.origins[img], "rays_d": self.dirs.reshape(-1, 3)[flat], "rgb": self.rgb.reshape(-1, 3)[flat],
                    "alpha": None if self.alpha is None else self.alpha.reshape(-1, 1)[flat],
                    "appearance_idx": img, "img_idx": img}
        if img_idx is None:
            img_idx = int(torch.randint(0, self.n_images, (1,), generator=generator, device=self.device))
        sel = torch.randperm(hw, generator=generator, device=self.device)[:batch_size]
        return {"rays_o": self.origins[img_idx].expand(sel.numel(), 3), "rays_d": self.dirs[img_idx][sel],
                "rgb": self.rgb[img_idx][sel], "alpha": None if self.alpha is None else self.alpha[img_idx][sel],
                "appearance_idx": img_idx, "img_idx": img_idx}
