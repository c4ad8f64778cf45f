"""nerfw.io.FrameWriter on host tensors (no GPU): the reference drivers' file names and formats, written off the caller's
thread (render_aligned_spiral.py:158-175, run.py:233-269), errors surfaced to the caller."""
import os

import numpy as np
import pytest
import torch


def test_frame_writer_names_and_contents(tmp_path):
    from PIL import Image
    from nerfw.io import FrameWriter
    rgb = (torch.rand(3, 20, 24, 3) * 255).to(torch.uint8)
    depth8 = (torch.rand(20, 24) * 255).to(torch.uint8)
    depth = torch.rand(20, 24) * 4 + 2
    out = str(tmp_path / "output" / "spiral")
    with FrameWriter(out, workers=2, max_pending=2) as wr:
        for i in range(3):
            wr.png(f"frame_{i:04d}.png", rgb[i])
        wr.png("depth_0000.png", depth8)
        wr.npy(os.path.join("raw", "depth_000.npy"), depth)
        files = list(wr.files)
    assert sorted(os.listdir(out)) == ["depth_0000.png", "frame_0000.png", "frame_0001.png", "frame_0002.png", "raw"]
    assert [os.path.basename(f) for f in files] == ["frame_0000.png", "frame_0001.png", "frame_0002.png", "depth_0000.png", "depth_000.npy"]
    for i in range(3):
        assert np.array_equal(np.asarray(Image.open(os.path.join(out, f"frame_{i:04d}.png"))), rgb[i].numpy())
    assert np.array_equal(np.asarray(Image.open(os.path.join(out, "depth_0000.png"))), depth8.numpy())
    assert np.array_equal(np.load(os.path.join(out, "raw", "depth_000.npy")), depth.numpy())
    # what apply_all_shaders.py:13-27 globs for
    assert sorted(f for f in os.listdir(out) if f.startswith("frame_") and f.endswith(".png")) == [f"frame_{i:04d}.png" for i in range(3)]
    assert [f.split("_")[1].split(".")[0] for f in os.listdir(out) if f.startswith("depth_") and f.endswith(".png")] == ["0000"]


def test_frame_writer_surfaces_worker_errors(tmp_path):
    from nerfw.io import FrameWriter
    wr = FrameWriter(str(tmp_path / "o"))
    wr.png("bad.png", torch.rand(4, 4, 7))          # 7 channels: PIL cannot encode it
    with pytest.raises(RuntimeError, match="frame writer failed"):
        wr.close()
