"""The CPU oracle restatement against the golden vectors generated from the real reference (oracle/make_golden.py).
Runs everywhere (no GPU, no /root/reference)."""
import hashlib

import numpy as np
import torch


def sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def test_state_dict_matches_reference_init(oracle, manifest):
    sd = oracle.make_state_dict(0)
    emb = torch.randn(32)
    assert sorted(sd) == sorted(manifest["state_dict"])
    for k, v in sd.items():
        assert list(v.shape) == manifest["state_dict"][k]["shape"]
        assert sha(v) == manifest["state_dict"][k]["sha256"], k
    assert sha(emb) == manifest["emb_sha256"]
    assert sum(v.numel() for v in sd.values()) == manifest["n_params"] == 534276


def test_rays_bit_exact(oracle, golden, manifest):
    g = golden("rays_100")
    o, d = oracle.rays_for_view(100, 100, float(g["focal"]), torch.from_numpy(g["c2w"]))
    assert torch.equal(d, torch.from_numpy(g["rays_d"]))
    assert sha(d) == manifest["cases"]["rays_100"]["sha256_d"]
    assert o.stride()[:2] == (0, 0)  # expanded view, src/ray_utils.py:48
    for name, case in manifest["cases"]["rays_rot"].items():
        o, d = oracle.rays_for_view(case["h"], case["w"], case["focal"], torch.tensor(case["c2w"]))
        assert sha(d) == case["sha256_d"], name


def test_stratified_bit_exact(oracle, golden):
    g = golden("stratified")
    o, d = torch.from_numpy(g["o"]), torch.from_numpy(g["d"])
    z, _ = oracle.stratified_depths(o, d, 2.0, 6.0, 64, perturb=False)
    assert torch.equal(z[0], torch.from_numpy(g["z_det"]))
    z, pts = oracle.stratified_depths(o, d, 2.0, 6.0, 64, perturb=True, t_rand=torch.from_numpy(g["t_rand"]))
    assert torch.equal(z, torch.from_numpy(g["z_perturb"]))
    assert torch.equal(pts, torch.from_numpy(g["pts_perturb"]))
    for n in (2, 64, 128, 192, 256):
        assert torch.equal(oracle.depth_table(2.0, 6.0, n), torch.from_numpy(g[f"ztab_{n}"]))


def test_mlp_bit_exact(oracle, golden, state_dict):
    sd, emb = state_dict
    g = golden("mlp_64")
    assert torch.equal(emb, torch.from_numpy(g["emb"]))
    x, d = torch.from_numpy(g["x"]), torch.from_numpy(g["d"])
    assert torch.equal(oracle.encode(x, 10), torch.from_numpy(g["pe"]))
    rgb, sig = oracle.mlp_forward(sd, x, d, emb.unsqueeze(0))
    np.testing.assert_allclose(rgb.numpy(), g["rgb"], rtol=0, atol=1e-6)  # MKL sgemm blocking may differ across hosts
    np.testing.assert_allclose(sig.numpy(), g["sigma"], rtol=0, atol=1e-6)
    rgb, sig = oracle.mlp_forward(sd, x, d, None)
    np.testing.assert_allclose(rgb.numpy(), g["rgb_noemb"], rtol=0, atol=1e-6)


def test_view100_coarse(oracle, golden, state_dict, manifest):
    sd, emb = state_dict
    h, w, focal, c2w = oracle.golden_camera()
    o, d = oracle.rays_for_view(h, w, focal, c2w)
    sel = slice(45, 55)
    with torch.no_grad():
        rgb, depth, ex = oracle.render_coarse(sd, o[sel, sel], d[sel, sel], 2.0, 6.0, 64, emb=emb, perturb=False)
    g = golden("view100_coarse")
    np.testing.assert_allclose(rgb.numpy(), g["rgb"][sel, sel], rtol=0, atol=2e-6)
    np.testing.assert_allclose(depth.numpy(), g["depth"][sel, sel], rtol=0, atol=2e-5)
    m = manifest["cases"]["view100_coarse"]
    np.testing.assert_allclose(g["rgb"][50, 50], m["rgb_50_50"], atol=1e-7)
    assert abs(float(g["rgb"].astype(np.float64).sum()) - m["sum_rgb"]) < 1e-6


def test_resample_matches_reference(oracle, golden, manifest):
    g = golden("resample")
    assert manifest["cases"]["resample_safe"]["reference_ran"]       # unpatched reference ran on this input
    assert manifest["cases"]["resample_generic"]["reference_raised"]  # and raises on generic input
    o = torch.zeros(8, 3)
    d = torch.ones(8, 3)
    z, _, aux = oracle.resample_pdf(o, d, torch.from_numpy(g["z_safe"]), torch.from_numpy(g["w_safe"]), 128,
                                    u_rand=torch.from_numpy(g["u_rand_safe"]), return_aux=True)
    assert torch.equal(z, torch.from_numpy(g["out_safe"]))
    assert torch.equal(aux["inds"], torch.from_numpy(g["inds_safe"]))
    o = torch.zeros(256, 3)
    d = torch.ones(256, 3)
    z, _, aux = oracle.resample_pdf(o, d, torch.from_numpy(g["z_gen"]), torch.from_numpy(g["w_gen"]), 128,
                                    u_rand=torch.from_numpy(g["u_rand_gen"]), return_aux=True)
    assert torch.equal(aux["inds"], torch.from_numpy(g["inds_gen"]))
    assert torch.equal(z, torch.from_numpy(g["out_gen"]))
    assert bool((z[:, 1:] >= z[:, :-1]).all())


def test_composite_properties(oracle):
    torch.manual_seed(3)
    sigma = torch.rand(16, 64, 1) * 5
    rgb = torch.rand(16, 64, 3)
    z = torch.sort(torch.rand(16, 64) * 4 + 2, dim=-1).values
    c, depth, w = oracle.composite(sigma, rgb, z)
    assert bool((w >= 0).all()) and bool((w.sum(1) <= 1 + 1e-5).all())
    assert bool((depth >= 2 - 1e-4).all()) and bool((depth <= 6 + 1e-4).all())
    # zero density -> nothing accumulates
    c0, d0, w0 = oracle.composite(torch.zeros(4, 8, 1), torch.rand(4, 8, 3), torch.linspace(2, 6, 8).expand(4, 8))
    assert float(w0.abs().max()) == 0.0 and float(c0.abs().max()) == 0.0
