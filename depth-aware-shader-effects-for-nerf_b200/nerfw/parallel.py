"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed (NCCL over NVLink on the box, gloo in
the CPU tests).

Rendering shards rays (or frames) across ranks with no data-path collective; the only exchange is the optional gather
of the finished (rgb, depth, acc) records.  Training is ray-batch data parallel: one all-reduce of a single flat fp32
gradient buffer (534 276 model gradients + the embedding table) per step, then a replicated fused Adam update.
The reference has no distributed code at all; the partitioning follows section 8(e).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def world_info(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced partition of n items: the first n % world ranks own one extra item."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def frames_for_rank(n_frames: int, rank: int, world: int) -> List[int]:
    """Round-robin frame ownership (frame i -> rank i mod world), the config-4 partitioning of a camera path."""
    return list(range(rank, n_frames, world))


def gather_rows(local: torch.Tensor, n_total: int, group=None, dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """Concatenate per-rank row blocks produced under shard_bounds(n_total, ...).  dst=None -> every rank gets the result
    (all_gather), else only rank dst (others return None).  Blocks are padded to the largest shard for the collective."""
    rank, world = world_info(group)
    if world == 1:
        return local
    width = local.shape[1:]
    max_rows = (n_total + world - 1) // world
    pad = torch.zeros((max_rows,) + tuple(width), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    if dst is not None and rank != dst:
        return None
    parts = []
    for r in range(world):
        s, e = shard_bounds(n_total, r, world)
        parts.append(bufs[r][: e - s])
    return torch.cat(parts, dim=0)


def render_sharded(render_fn: Callable[[torch.Tensor, torch.Tensor], Tuple[torch.Tensor, torch.Tensor, dict]],
                   rays_o: torch.Tensor, rays_d: torch.Tensor, group=None, dst: Optional[int] = None):
    """Render one batch of rays split across the ranks.  `render_fn(o, d)` is any volume_render closure returning
    (rgb (n,3), depth (n,1), extras with 'acc' (n,1)).  Returns (rgb, depth, acc) for ALL rays (None on non-dst ranks)."""
    rank, world = world_info(group)
    o = rays_o.reshape(-1, 3)
    d = rays_d.reshape(-1, 3)
    n = o.shape[0]
    s, e = shard_bounds(n, rank, world)
    rgb, depth, extras = render_fn(o[s:e], d[s:e])
    rec = torch.cat([rgb.reshape(-1, 3), depth.reshape(-1, 1), extras["acc"].reshape(-1, 1)], dim=1)  # 20 B per ray
    full = gather_rows(rec, n, group=group, dst=dst)
    if full is None:
        return None
    return full[:, :3], full[:, 3:4], full[:, 4:5]


class FlatParams:
    """One contiguous fp32 buffer holding every trainable tensor (and one for the gradients), so the data-parallel
    exchange is a single all-reduce and the optimizer a single kernel.  Parameters become views into the buffer; their
    state_dict layout is unchanged."""

    def __init__(self, tensors: Sequence[torch.nn.Parameter]):
        self.tensors = list(tensors)
        dev = self.tensors[0].device
        sizes = [t.numel() for t in self.tensors]
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + ((s + 3) // 4) * 4)   # keep every view 16-byte aligned
        total = self.offsets[-1]
        self.param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        for t, off in zip(self.tensors, self.offsets):
            n = t.numel()
            self.param[off:off + n].copy_(t.detach().reshape(-1))
            t.data = self.param[off:off + n].view(t.shape)
            t.grad = self.grad[off:off + n].view(t.shape)

    def zero_grad(self):
        self.grad.zero_()
        for t, off in zip(self.tensors, self.offsets):   # re-attach in case something replaced .grad
            n = t.numel()
            if t.grad is None or t.grad.data_ptr() != self.grad.data_ptr() + off * 4:
                t.grad = self.grad[off:off + n].view(t.shape)

    def all_reduce(self, group=None):
        _, world = world_info(group)
        if world > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
        return world
