"""Frame files with the reference drivers' names, written off the render thread.

render_aligned_spiral.py:158-175 writes `frame_%04d.png` for every frame and `depth_%04d.png` (min/max-normalised uint8)
for every 10th; run.py:233-269 writes `rgb_%03d.png`, and under `raw/` `rgb_%03d.png` / `depth_%03d.npy`.  Those names
are what apply_all_shaders.py:13-27 and create_video.py glob for, so a directory written here is consumed by the
untouched post-processing scripts.

The renderer hands over DEVICE tensors: the writer starts one asynchronous device-to-host copy into a pinned staging
buffer on its own CUDA stream (ordered after the producing kernels by an event) and a worker thread encodes the file
once the copy has landed, so frame i's transfer and PNG compression overlap frame i+1's kernels.
"""
from __future__ import annotations

import os
import queue
import threading
from typing import Optional

import numpy as np
import torch


def stage_to_host(t: torch.Tensor, copy_stream: "torch.cuda.Stream"):
    """Start an asynchronous device-to-host copy of `t` into a fresh pinned buffer on `copy_stream`, ordered after the
    work already queued on the current stream.  Returns (pinned host tensor, event that fires when the copy has landed)."""
    ready = torch.cuda.Event()
    ready.record(torch.cuda.current_stream(t.device))
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    with torch.cuda.stream(copy_stream):
        copy_stream.wait_event(ready)
        host.copy_(t, non_blocking=True)
        t.record_stream(copy_stream)
        done = torch.cuda.Event()
        done.record(copy_stream)
    return host, done


class FrameWriter:
    def __init__(self, output_dir: str, workers: int = 2, max_pending: int = 4, png_compress_level: int = 6):
        self.output_dir = output_dir
        os.makedirs(output_dir, exist_ok=True)
        self.compress_level = int(png_compress_level)
        self._q: "queue.Queue" = queue.Queue(maxsize=max_pending)      # back-pressure: bounded pinned memory
        self._err: Optional[BaseException] = None
        self._threads = [threading.Thread(target=self._work, daemon=True) for _ in range(max(1, workers))]
        self._copy_stream = None
        self.files = []
        for t in self._threads:
            t.start()

    # ---- producer side (render thread) ---------------------------------------------------------------------
    def _stage(self, t: torch.Tensor):
        """Device tensor -> (pinned host tensor, event) with the copy in flight; host tensors pass through."""
        if not isinstance(t, torch.Tensor):
            return torch.as_tensor(t), None
        if not t.is_cuda:
            return t, None
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=t.device)
        return stage_to_host(t, self._copy_stream)

    def _put(self, kind: str, rel: str, t):
        if self._err is not None:
            raise RuntimeError("frame writer failed") from self._err
        path = os.path.join(self.output_dir, rel)
        host, done = self._stage(t)
        self.files.append(path)
        self._q.put((kind, path, host, done))

    def png(self, rel: str, image_u8):
        """uint8 (H,W,3) or (H,W) -> PNG (PIL, as render_aligned_spiral.py:166,175 / run.py:242,266)."""
        self._put("png", rel, image_u8)

    def npy(self, rel: str, array):
        """float32 array -> .npy (run.py:246)."""
        self._put("npy", rel, array)

    # ---- worker side -------------------------------------------------------------------------------------
    def _work(self):
        from PIL import Image
        while True:
            item = self._q.get()
            if item is None:
                self._q.task_done()
                return
            kind, path, host, done = item
            try:
                if done is not None:
                    done.synchronize()
                arr = host.numpy()
                os.makedirs(os.path.dirname(path), exist_ok=True)
                if kind == "png":
                    Image.fromarray(arr).save(path, compress_level=self.compress_level)
                else:
                    np.save(path, arr)
            except BaseException as e:  # surfaced on the render thread at the next call / close()
                self._err = e
            finally:
                self._q.task_done()

    def close(self):
        for _ in self._threads:
            self._q.put(None)
        for t in self._threads:
            t.join()
        if self._err is not None:
            raise RuntimeError("frame writer failed") from self._err

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False
