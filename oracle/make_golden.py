"""Pin the oracle against the real reference and write tests/golden/.

Run HERE (the container that has /root/reference); the GPU box only reads the committed output.

    python oracle/make_golden.py

For every case it runs the UNMODIFIED reference (imported from /root/reference, stdout silenced)
and the oracle restatement on the same seeded inputs, requires bit-equality between the two, and
stores the reference's outputs.  Nothing is copied from the reference tree: only its outputs.
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("NERFW_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)
sys.path.insert(0, REF)

import nerfw_oracle as orc  # noqa: E402


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def main():
    with quiet():
        from config import Config
        from src.models import NeRF, PositionalEncoding
        from src.ray_utils import get_rays, sample_importance, sample_stratified
        from src.render import volume_render
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    os.makedirs(OUT, exist_ok=True)
    man = {"torch": torch.__version__, "cases": {}}

    # ---- weights: reference init under seed 0 == oracle.make_state_dict(0)
    torch.manual_seed(0)
    with quiet():
        model = NeRF(Config())
    emb = torch.randn(32)
    sd_ref = {k: v.detach().clone() for k, v in model.state_dict().items()}
    sd = orc.make_state_dict(0)
    emb_o = torch.randn(32)
    assert list(sd_ref) == list(sd), (list(sd_ref), list(sd))
    for k in sd:
        assert torch.equal(sd[k], sd_ref[k]), k
    assert torch.equal(emb, emb_o)
    man["state_dict"] = {k: {"shape": list(v.shape), "sha256": sha(v)} for k, v in sd_ref.items()}
    man["emb_sha256"] = sha(emb)
    man["n_params"] = int(sum(v.numel() for v in sd_ref.values()))

    # ---- rays (bit-exact), 100x100 identity camera + three random rotations at 37x53 and 800x800 hashes
    H, W, focal, c2w = orc.golden_camera()
    ro, rd = get_rays(H, W, focal, c2w)
    ro2, rd2 = orc.rays_for_view(H, W, focal, c2w)
    assert torch.equal(rd, rd2) and torch.equal(ro, ro2)
    np.savez_compressed(os.path.join(OUT, "rays_100.npz"), c2w=c2w.numpy(), focal=np.float64(focal),
                        rays_d=rd.numpy(), origin=c2w[:3, 3].numpy())
    man["cases"]["rays_100"] = {"sha256_d": sha(rd), "d00": rd[0, 0].tolist(), "d9999": rd[99, 99].tolist()}
    rot_cases = {}
    g = torch.Generator().manual_seed(7)
    for n, (h, w) in enumerate([(37, 53), (800, 800), (64, 48)]):
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
        m = torch.eye(4)
        m[:3, :3] = q
        m[:3, 3] = torch.randn(3, generator=g)
        f = 0.5 * w / np.tan(0.5 * 0.6911112070083618)
        a = get_rays(h, w, f, m)
        b = orc.rays_for_view(h, w, f, m)
        assert torch.equal(a[1], b[1]) and torch.equal(a[0], b[0])
        rot_cases[f"rot{n}"] = {"h": h, "w": w, "focal": float(f), "c2w": m.numpy().tolist(),
                                "sha256_d": sha(a[1]), "sha256_o": sha(a[0].contiguous())}
        if h * w < 5000:
            np.savez_compressed(os.path.join(OUT, f"rays_rot{n}.npz"), c2w=m.numpy(), focal=np.float64(f),
                                rays_d=a[1].numpy())
    man["cases"]["rays_rot"] = rot_cases

    # ---- stratified depths
    o4 = ro.reshape(-1, 3)[:4].contiguous()
    d4 = rd.reshape(-1, 3)[:4].contiguous()
    zt, _ = sample_stratified(o4, d4, 2.0, 6.0, 64, perturb=False)
    zt2, _ = orc.stratified_depths(o4, d4, 2.0, 6.0, 64, perturb=False)
    assert torch.equal(zt, zt2)
    torch.manual_seed(123)
    zp, pp = sample_stratified(o4, d4, 2.0, 6.0, 64, perturb=True)
    torch.manual_seed(123)
    zp2, pp2 = orc.stratified_depths(o4, d4, 2.0, 6.0, 64, perturb=True)
    assert torch.equal(zp, zp2) and torch.equal(pp, pp2)
    torch.manual_seed(123)
    tr = torch.rand(4, 64)
    tables = {f"ztab_{n}": orc.depth_table(2.0, 6.0, n).numpy() for n in (2, 64, 128, 192, 256)}
    np.savez_compressed(os.path.join(OUT, "stratified.npz"), o=o4.numpy(), d=d4.numpy(), z_det=zt[0].numpy(),
                        t_rand=tr.numpy(), z_perturb=zp.numpy(), pts_perturb=pp.numpy(), **tables)
    man["cases"]["stratified"] = {"ztab64_sha256": sha(zt[0].contiguous()), "z_perturb_0_3": zp[0, :3].tolist()}

    # ---- positional encoding + MLP forward (fp tolerance cases)
    g = torch.Generator().manual_seed(11)
    x = (torch.rand(64, 3, generator=g) - 0.5) * 8.0
    dd = torch.nn.functional.normalize(torch.randn(64, 3, generator=g), dim=-1)
    with quiet():
        pe = PositionalEncoding(10)(x)
        rgb_m, sig_m = model(x, dd, emb.unsqueeze(0))
        rgb_n, sig_n = model(x, dd, None)
    assert torch.equal(pe, orc.encode(x, 10))
    r2, s2 = orc.mlp_forward(sd, x, dd, emb.unsqueeze(0))
    assert torch.equal(rgb_m, r2) and torch.equal(sig_m, s2)
    r3, s3 = orc.mlp_forward(sd, x, dd, None)
    assert torch.equal(rgb_n, r3) and torch.equal(sig_n, s3)
    np.savez_compressed(os.path.join(OUT, "mlp_64.npz"), x=x.numpy(), d=dd.numpy(), emb=emb.numpy(),
                        pe=pe.detach().numpy(), rgb=rgb_m.detach().numpy(), sigma=sig_m.detach().numpy(),
                        rgb_noemb=rgb_n.detach().numpy(), sigma_noemb=sig_n.detach().numpy())

    # ---- config 1: 100x100 view, 64 samples, reference volume_render as-is (coarse only)
    with torch.no_grad(), quiet():
        rgb, depth, ex = volume_render(model, ro, rd, 2.0, 6.0, 64, 128, appearance_embedding=emb, perturb=False)
    with torch.no_grad():
        rgb_o, depth_o, ex_o = orc.render_coarse(sd, ro, rd, 2.0, 6.0, 64, emb=emb, perturb=False)
    assert torch.equal(rgb, rgb_o) and torch.equal(depth, depth_o) and torch.equal(ex["weights"], ex_o["weights"])
    acc = ex["weights"].sum(1)
    np.savez_compressed(os.path.join(OUT, "view100_coarse.npz"), rgb=rgb.numpy(), depth=depth.numpy(),
                        acc=acc.numpy(), weights_5050=ex["weights"][5050, :, 0].numpy())
    man["cases"]["view100_coarse"] = {
        "rgb_50_50": rgb[50, 50].tolist(), "depth_50_50": float(depth[50, 50, 0]), "acc_5050": float(acc[5050, 0]),
        "sum_rgb": float(rgb.double().sum()), "sum_depth": float(depth.double().sum()),
        "sum_weights": float(ex["weights"].double().sum()),
    }

    # ---- dense-ish scene variant (density head x200, +1 bias): acc ~ 1, the depth-sensitive case
    sd_dense = {k: v.clone() for k, v in sd.items()}
    sd_dense["density_head.weight"] *= 200.0
    sd_dense["density_head.bias"] += 1.0
    with torch.no_grad():
        rgb_d, depth_d, ex_d = orc.render_coarse(sd_dense, ro[40:60, 40:60], rd[40:60, 40:60], 2.0, 6.0, 64, emb=emb, perturb=False)
    model_d = NeRF.__new__(NeRF)
    with quiet():
        model_d = NeRF(Config())
    model_d.load_state_dict(sd_dense)
    with torch.no_grad(), quiet():
        rgb_dr, depth_dr, ex_dr = volume_render(model_d, ro[40:60, 40:60], rd[40:60, 40:60], 2.0, 6.0, 64, 0,
                                                appearance_embedding=emb, perturb=False)
    assert torch.equal(rgb_d, rgb_dr) and torch.equal(depth_d, depth_dr)
    np.savez_compressed(os.path.join(OUT, "crop20_dense.npz"), rgb=rgb_d.numpy(), depth=depth_d.numpy(),
                        acc=ex_d["acc"].numpy())
    man["cases"]["crop20_dense"] = {"mean_acc": float(ex_d["acc"].mean())}

    # ---- perturbed coarse render on 256 rays (RNG-stream parity)
    sel = torch.arange(0, 10000, 39)[:256]
    o256 = ro.reshape(-1, 3)[sel].contiguous()
    d256 = rd.reshape(-1, 3)[sel].contiguous()
    torch.manual_seed(5)
    with torch.no_grad(), quiet():
        rgb_p, depth_p, ex_p = volume_render(model, o256, d256, 2.0, 6.0, 64, 0, appearance_embedding=emb, perturb=True)
    torch.manual_seed(5)
    with torch.no_grad():
        rgb_po, depth_po, ex_po = orc.render_coarse(sd, o256, d256, 2.0, 6.0, 64, emb=emb, perturb=True)
    assert torch.equal(rgb_p, rgb_po) and torch.equal(ex_p["z_vals"], ex_po["z_vals"])
    np.savez_compressed(os.path.join(OUT, "rays256_perturb.npz"), sel=sel.numpy(), rgb=rgb_p.numpy(),
                        depth=depth_p.numpy(), z_vals=ex_p["z_vals"].numpy(), weights=ex_p["weights"][..., 0].numpy())

    # ---- sample_importance: (a) a case the unpatched reference survives, bit-compared; (b) generic case
    #      where the reference raises (F2) -- recorded as such; oracle output stored.
    w_c = ex_p["weights"][..., 0].contiguous()
    z_c = ex_p["z_vals"].contiguous()
    # (a) concentrate all mass in the first bins so no u falls in the last bin
    w_safe = torch.zeros(8, 64)
    w_safe[:, 3:9] = torch.rand(8, 6, generator=g) + 0.5
    w_safe = w_safe * 1e4  # the +1e-5 floor leaves ~1e-8 of mass per trailing bin
    torch.manual_seed(21)
    try:
        zs_ref, ps_ref = sample_importance(o256[:8], d256[:8], z_c[:8], w_safe, 128)
        ref_ok = True
    except (RuntimeError, IndexError):
        ref_ok = False
    torch.manual_seed(21)
    zs_o, ps_o, aux_s = orc.resample_pdf(o256[:8], d256[:8], z_c[:8], w_safe, 128, return_aux=True)
    if ref_ok:
        assert torch.equal(zs_ref, zs_o) and torch.equal(ps_ref, ps_o)
    man["cases"]["resample_safe"] = {"reference_ran": ref_ok}
    torch.manual_seed(21)
    u_safe = torch.rand(8, 128)
    # (b) generic: the coarse weights of the perturbed render
    raised = False
    torch.manual_seed(22)
    try:
        sample_importance(o256, d256, z_c, w_c, 128)
    except (RuntimeError, IndexError):
        raised = True
    torch.manual_seed(22)
    u_gen = torch.rand(256, 128)
    zg_o, _, aux_g = orc.resample_pdf(o256, d256, z_c, w_c, 128, u_rand=u_gen, return_aux=True)
    man["cases"]["resample_generic"] = {"reference_raised": raised,
                                        "frac_last_bin": float((aux_g["inds"] >= 64).float().mean())}
    np.savez_compressed(os.path.join(OUT, "resample.npz"),
                        z_safe=z_c[:8].numpy(), w_safe=w_safe.numpy(), u_rand_safe=u_safe.numpy(),
                        out_safe=zs_o.numpy(), inds_safe=aux_s["inds"].numpy(), zfine_safe=aux_s["z_fine"].numpy(),
                        z_gen=z_c.numpy(), w_gen=w_c.numpy(), u_rand_gen=u_gen.numpy(),
                        out_gen=zg_o.numpy(), inds_gen=aux_g["inds"].numpy(), zfine_gen=aux_g["z_fine"].numpy())

    # ---- composed coarse+fine oracle on the 100x100 view (fixed u), centre 32x32 crop stored
    torch.manual_seed(33)
    u_fix = torch.rand(32 * 32, 128)
    oc = ro[34:66, 34:66].reshape(-1, 3).contiguous()
    dc = rd[34:66, 34:66].reshape(-1, 3).contiguous()
    with torch.no_grad():
        rgb_h, depth_h, ex_h = orc.render_hier(sd, sd, oc, dc, 2.0, 6.0, 64, 128, emb=emb, perturb=False, u_rand=u_fix)
    np.savez_compressed(os.path.join(OUT, "crop32_hier.npz"), rgb=rgb_h.numpy(), depth=depth_h.numpy(),
                        acc=ex_h["acc"].numpy(), z_vals=ex_h["z_vals"].numpy(), u_rand=u_fix.numpy(),
                        inds=ex_h["inds"].numpy().astype(np.int16))
    man["cases"]["crop32_hier"] = {"sum_rgb": float(rgb_h.double().sum()), "sum_depth": float(depth_h.double().sum())}

    # ---- DENSE hierarchical case (density head x200, bias +1: acc ~ 1 like a trained, opaque scene), centre 24x24 crop,
    #      64 + 128, fixed u.  The coarse half is the reference's own volume_render (asserted bit-equal here);
    #      resampling + fine pass are the composed oracle of SURVEY.md section 8c.
    torch.manual_seed(34)
    u_d = torch.rand(24 * 24, 128)
    od = ro[38:62, 38:62].reshape(-1, 3).contiguous()
    dd_ = rd[38:62, 38:62].reshape(-1, 3).contiguous()
    with torch.no_grad(), quiet():
        rgb_c_ref, depth_c_ref, ex_c_ref = volume_render(model_d, od, dd_, 2.0, 6.0, 64, 0, appearance_embedding=emb, perturb=False)
    with torch.no_grad():
        rgb_hd, depth_hd, ex_hd = orc.render_hier(sd_dense, sd_dense, od, dd_, 2.0, 6.0, 64, 128, emb=emb, perturb=False, u_rand=u_d)
    assert torch.equal(ex_hd["rgb_coarse"], rgb_c_ref) and torch.equal(ex_hd["depth_coarse"], depth_c_ref)
    assert torch.equal(ex_hd["weights_coarse"], ex_c_ref["weights"])
    np.savez_compressed(os.path.join(OUT, "crop24_dense_hier.npz"), rgb=rgb_hd.numpy(), depth=depth_hd.numpy(),
                        acc=ex_hd["acc"].numpy(), z_vals=ex_hd["z_vals"].numpy(), u_rand=u_d.numpy(),
                        inds=ex_hd["inds"].numpy().astype(np.int16), depth_coarse=depth_c_ref.numpy(),
                        weights_coarse=ex_hd["weights_coarse"][..., 0].numpy())
    man["cases"]["crop24_dense_hier"] = {"mean_acc": float(ex_hd["acc"].mean()), "sum_depth": float(depth_hd.double().sum()),
                                         "frac_last_bin": float((ex_hd["inds"] >= 64).float().mean())}

    # ---- gradients of the coarse path (autograd through the reference), 64 rays
    sel64 = torch.arange(0, 10000, 157)[:64]
    o64 = ro.reshape(-1, 3)[sel64].contiguous()
    d64 = rd.reshape(-1, 3)[sel64].contiguous()
    tgt = torch.rand(64, 3, generator=g)
    emb_p = emb.clone().requires_grad_(True)
    model.zero_grad()
    torch.manual_seed(9)
    with quiet():
        rgb_g, _, _ = volume_render(model, o64, d64, 2.0, 6.0, 64, 0, appearance_embedding=emb_p, perturb=True)
    loss = torch.nn.functional.mse_loss(rgb_g, tgt)
    loss.backward()
    grads = {k.replace(".", "__"): p.grad.numpy() for k, p in model.named_parameters()
             if k.startswith(("density_head", "rgb_linear", "appearance_projection")) or k.endswith("bias")}
    gnorm = {k: float(p.grad.double().norm()) for k, p in model.named_parameters()}
    np.savez_compressed(os.path.join(OUT, "grads_64.npz"), sel=sel64.numpy(), target=tgt.numpy(),
                        loss=np.float32(loss.item()), emb_grad=emb_p.grad.numpy(), **grads)
    man["cases"]["grads_64"] = {"loss": float(loss), "grad_norms": gnorm}

    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(man, f, indent=1, sort_keys=True)
    tot = sum(os.path.getsize(os.path.join(OUT, p)) for p in os.listdir(OUT))
    print(f"golden vectors written to {OUT} ({tot/1e6:.2f} MB); oracle == reference on every case")
    print(json.dumps(man["cases"]["view100_coarse"], indent=1))


if __name__ == "__main__":
    main()
