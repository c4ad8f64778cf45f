"""Pure-write and read+write HBM bandwidth on this GPU (context for the store-heavy backward pass 1)."""
import torch
x = torch.empty(1 << 30, dtype=torch.float32, device="cuda")   # 4 GiB
y = torch.empty_like(x)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = t(lambda: x.zero_()); print(f"memset 4 GiB: {ms:.3f} ms, {x.numel() * 4 / ms / 1e6:.0f} GB/s written")
ms = t(lambda: x.fill_(1.5)); print(f"fill   4 GiB: {ms:.3f} ms, {x.numel() * 4 / ms / 1e6:.0f} GB/s written")
ms = t(lambda: y.copy_(x)); print(f"copy   4 GiB: {ms:.3f} ms, {2 * x.numel() * 4 / ms / 1e6:.0f} GB/s read+written")
ms = t(lambda: x.sum()); print(f"sum    4 GiB: {ms:.3f} ms, {x.numel() * 4 / ms / 1e6:.0f} GB/s read")
