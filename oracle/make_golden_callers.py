"""Golden vectors from the reference's UNMODIFIED CALLERS (TEST INFRASTRUCTURE; run here, where /root/reference exists).

    python oracle/make_golden_callers.py

The drop-in claim is about callers, not only callees: `render_aligned_spiral.py` (camera path, 4096-ray chunk loop, PNG
naming), `run.py`'s model smoke test and `src/train.py::train_nerf` (the optimisation step) must keep working when this
repo's `src/` shadows the reference's.  Those drivers cannot run on the GPU box (no reference tree there) and this
container has no GPU, so this script executes the reference drivers AS THEY ARE -- their own `src/`, CPU, stub dataset,
`matplotlib` stubbed (absent here) -- and records what they computed:

  callers_poses.npz    every c2w `render_aligned_spiral` hands to `get_rays` (120 frames, rotation axes x / y / z / none,
                       scene chair / lego): pins nerfw.camera.aligned_spiral_poses (render_aligned_spiral.py:27-122)
  callers_spiral.npz   the PNG files the driver wrote for a 3-frame 24x24 render (names + decoded pixels)
  callers_train.npz    3 iterations of train_nerf on a fixed 64-ray batch: the uniforms it drew, loss per iteration,
                       parameter norms and the trained embedding row afterwards (src/train.py:54-92)

tests/test_gpu_callers.py replays the same drivers' loops through this repo's `src/` shim on the GPU and compares.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("NERFW_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)
sys.path.insert(0, REF)

import nerfw_oracle as orc  # noqa: E402


def stub_matplotlib():
    """src/train.py imports matplotlib.pyplot at module level; it is only used at iteration 1000 and at the end."""
    if "matplotlib" in sys.modules:
        return
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    for name in ("figure", "subplot", "plot", "title", "xlabel", "ylabel", "savefig", "close", "imshow", "axis", "colorbar"):
        setattr(plt, name, lambda *a, **k: None)
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt


class StubDataset:
    """What the drivers read from NeRFDataset (src/dataset.py:60-85, :206-277): H, W, focal, near, far, the embedding
    table and get_rays() batches.  Rays come from the golden camera so the values are reproducible anywhere."""

    def __init__(self, h, w, n_images=3, seed=1):
        self.H, self.W = h, w
        self.focal = 0.5 * w / np.tan(0.5 * 0.6911112070083618)
        self.near, self.far = 2.0, 6.0
        g = torch.Generator().manual_seed(seed)
        self.appearance_embeddings = torch.nn.Parameter(torch.randn(n_images, 32, generator=g))
        _, _, focal, c2w = orc.golden_camera(40, 40)
        ro, rd = orc.rays_for_view(40, 40, focal, c2w)
        sel = torch.arange(0, 1600, 25)[:64]
        self.batch = {"rays_o": ro.reshape(-1, 3)[sel].contiguous(), "rays_d": rd.reshape(-1, 3)[sel].contiguous(),
                      "rgb": torch.rand(64, 3, generator=g), "alpha": None, "appearance_idx": 1, "img_idx": 1}

    def __len__(self):
        return self.appearance_embeddings.shape[0]

    def get_rays(self, idx=None, batch_size=None):
        return self.batch


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def main():
    stub_matplotlib()
    with quiet():
        from config import Config
        import render_aligned_spiral as ras          # the unmodified driver, bound to the reference's own src/
        from src import train as ref_train
    from PIL import Image
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    os.makedirs(OUT, exist_ok=True)
    cwd = os.getcwd()

    # ---- camera path: capture the c2w the driver passes to get_rays ------------------------------------------------
    poses = {}
    real_get_rays, real_render = ras.get_rays, ras.volume_render
    captured = []

    def spy_get_rays(h, w, focal, c2w):
        captured.append(c2w.detach().cpu().numpy().copy())
        return real_get_rays(h, w, focal, c2w)

    def fake_render(model, o, d, **kw):
        return torch.zeros(o.shape[0], 3), torch.ones(o.shape[0], 1) * torch.arange(o.shape[0])[:, None], {}

    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            ras.get_rays, ras.volume_render = spy_get_rays, fake_render
            for scene in ("chair", "lego"):
                for axis in ("x", "y", "z", "none"):
                    cfg = Config()
                    cfg.device, cfg.scene, cfg.use_appearance = torch.device("cpu"), scene, False
                    captured.clear()
                    with quiet(), contextlib.redirect_stderr(io.StringIO()):
                        ras.render_aligned_spiral(torch.nn.Identity(), StubDataset(2, 2), cfg, "o", num_frames=120, loops=2,
                                                  rotation_axis=axis)
                    poses[f"{scene}_{axis}"] = np.stack(captured).astype(np.float32)
            cfg = Config()
            cfg.device, cfg.scene, cfg.use_appearance = torch.device("cpu"), "chair", False
            captured.clear()
            with quiet(), contextlib.redirect_stderr(io.StringIO()):
                ras.render_aligned_spiral(torch.nn.Identity(), StubDataset(2, 2), cfg, "o", num_frames=7, loops=1, rotation_axis="x")
            poses["chair_x_7f_1loop"] = np.stack(captured).astype(np.float32)
        finally:
            ras.get_rays, ras.volume_render = real_get_rays, real_render
            os.chdir(cwd)
    # ---- run.py::render_path camera paths (run.py:113-196): same capture through the unmodified driver --------------
    with quiet():
        import run as ref_run
        import src.ray_utils as ref_ray_utils
    real_gr, real_vr = ref_ray_utils.get_rays, ref_run.volume_render

    def spy_get_rays2(h, w, focal, c2w):
        captured.append(c2w.detach().cpu().numpy().copy())
        return real_gr(h, w, focal, c2w)

    with tempfile.TemporaryDirectory() as tmp:
        try:
            ref_ray_utils.get_rays, ref_run.volume_render = spy_get_rays2, fake_render
            for scene in ("lego", "chair", "hotdog"):
                for path in ("circle", "spiral", "horizontal_only", "hemisphere"):
                    cfg = Config()
                    cfg.device, cfg.scene, cfg.use_appearance = torch.device("cpu"), scene, False
                    captured.clear()
                    with quiet(), contextlib.redirect_stderr(io.StringIO()):
                        ref_run.render_path(torch.nn.Identity(), StubDataset(2, 2), cfg, os.path.join(tmp, f"{scene}_{path}"),
                                            num_frames=24, quality="preview", width=2, height=2, camera_path=path,
                                            spiral_loops=1.5, height_range=[-0.4, 0.6])
                    poses[f"run_{scene}_{path}"] = np.stack(captured).astype(np.float32)
            run_files = sorted(os.listdir(os.path.join(tmp, "lego_circle")))[:4]
        finally:
            ref_ray_utils.get_rays, ref_run.volume_render = real_gr, real_vr
    poses["run_file_names"] = np.array(run_files)
    np.savez_compressed(os.path.join(OUT, "callers_poses.npz"), **poses)

    # ---- the driver end to end: 3 frames of 24x24, reference model (seed 0), files it writes ---------------------
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            cfg = Config()
            cfg.device, cfg.scene = torch.device("cpu"), "chair"
            cfg.num_samples, cfg.num_importance = 64, 128
            torch.manual_seed(0)
            with quiet():
                model = ras.NeRF(cfg)
            ds = StubDataset(24, 24)
            with quiet(), contextlib.redirect_stderr(io.StringIO()):
                ras.render_aligned_spiral(model, ds, cfg, "spiral", num_frames=3, loops=1, rotation_axis="x")
            files = sorted(f for f in os.listdir(os.path.join("output", "spiral")) if f.endswith(".png"))
            imgs = {f.replace(".", "_"): np.asarray(Image.open(os.path.join("output", "spiral", f))) for f in files}
        finally:
            os.chdir(cwd)
    np.savez_compressed(os.path.join(OUT, "callers_spiral.npz"), files=np.array(files),
                        emb_table=ds.appearance_embeddings.detach().numpy(), **imgs)

    # ---- train_nerf: 3 iterations (all on the 64-ray start-up batch, src/train.py:26,55-57) ------------------------
    cfg = Config()
    cfg.device = torch.device("cpu")
    cfg.num_iterations, cfg.num_samples, cfg.num_importance = 3, 64, 128
    ds = StubDataset(24, 24)
    emb0 = ds.appearance_embeddings.detach().clone()
    drawn = []
    real_rand = torch.rand

    def spy_rand(*a, **k):
        t = real_rand(*a, **k)
        drawn.append(t.clone())
        return t

    losses = []
    real_mse = torch.nn.functional.mse_loss

    def spy_mse(a, b, *r, **k):
        v = real_mse(a, b, *r, **k)
        losses.append(float(v))
        return v

    with tempfile.TemporaryDirectory() as tmp:
        torch.manual_seed(0)
        torch.rand, torch.nn.functional.mse_loss = spy_rand, spy_mse
        try:
            with quiet(), contextlib.redirect_stderr(io.StringIO()):
                trained = ref_train.train_nerf(cfg, ds, save_dir=tmp)
            ck = torch.load(os.path.join(tmp, "checkpoint_final.pt"), weights_only=False)
        finally:
            torch.rand, torch.nn.functional.mse_loss = real_rand, real_mse
    assert len(drawn) == 3 and all(t.shape == (64, 64) for t in drawn), [t.shape for t in drawn]
    assert sorted(ck) == ["appearance_embeddings", "iteration", "loss", "model_state_dict", "optimizer_state_dict", "psnr"]
    sd = trained.state_dict()
    np.savez_compressed(os.path.join(OUT, "callers_train.npz"),
                        t_rand=torch.stack(drawn).numpy(), losses=np.array(losses, dtype=np.float64),
                        rays_o=ds.batch["rays_o"].numpy(), rays_d=ds.batch["rays_d"].numpy(), target=ds.batch["rgb"].numpy(),
                        emb_table0=emb0.numpy(), emb_table=ds.appearance_embeddings.detach().numpy(),
                        param_names=np.array(list(sd)), param_norms=np.array([float(v.double().norm()) for v in sd.values()]),
                        rgb_bias=sd["rgb_linear.bias"].numpy(), density_w=sd["density_head.weight"].numpy(),
                        layer7_bias=sd["pts_linears.7.bias"].numpy(), ckpt_keys=np.array(sorted(ck)))
    print("callers goldens written:", {k: v.shape for k, v in poses.items() if k.endswith("_x")}, files, losses)


if __name__ == "__main__":
    main()
