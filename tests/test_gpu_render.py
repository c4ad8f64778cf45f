"""volume_render end to end (config 0 of BASELINE.json and the composed coarse+fine path) against golden vectors
generated from the reference, plus training-step gradients."""
import numpy as np
import pytest
import torch

from gpu_util import maxabs, record

pytestmark = pytest.mark.gpu

# (rgb, depth, acc) max-abs bounds per MLP mode.  fp32 / bf16x3: the north-star 1e-3 bound (measured far below);
# bf16: the stated looser tensor-core bounds.
# "mixed" (the default: bf16x3 coarse pass + fp16 fine pass; a single pass resolves to bf16x3) carries the fp32 bars
TOL = {"fp32": (1e-4, 1e-3, 1e-4), "bf16x3": (1e-3, 1e-3, 1e-3), "mixed": (1e-3, 1e-3, 1e-3), "bf16": (1e-2, 5e-2, 1e-2),
       "fp16": (1e-3, 1e-2, 1e-3)}


def view(oracle):
    h, w, focal, c2w = oracle.golden_camera()
    import nerfw
    return nerfw.get_rays(h, w, focal, c2w.cuda())


@pytest.mark.parametrize("mode", ["fp32", "bf16x3", "bf16"])
def test_view100_coarse_golden(cuda_model, oracle, golden, mode):
    """BASELINE.json configs[0]: the 100x100 view exactly as the reference executes it (coarse only, F1)."""
    import nerfw
    model, emb = cuda_model
    o, d = view(oracle)
    g = golden("view100_coarse")
    with torch.no_grad():
        rgb, depth, ex = nerfw.volume_render(model, o, d, 2.0, 6.0, 64, 128, appearance_embedding=emb, perturb=False,
                                             mlp_dtype=mode, fine_pass=False)
    assert rgb.shape == (100, 100, 3) and depth.shape == (100, 100, 1)
    assert ex["weights"].shape == (10000, 64, 1) and ex["z_vals"].shape == (10000, 64) and ex["acc"].shape == (10000, 1)
    e = dict(rgb=maxabs(rgb, g["rgb"]), depth=maxabs(depth, g["depth"]), acc=maxabs(ex["acc"], g["acc"]),
             w5050=maxabs(ex["weights"][5050, :, 0], g["weights_5050"]))
    record(f"view100_coarse_{mode}", **e)
    t = TOL[mode]
    assert e["rgb"] <= t[0] and e["depth"] <= t[1] and e["acc"] <= t[2], e


def dense_model(state_dict):
    """density head x200, bias +1 (acc ~ 1, like a trained opaque scene): the variant oracle/make_golden.py renders."""
    import nerfw
    from config import Config
    sd, emb = state_dict
    sd = {k: v.clone() for k, v in sd.items()}
    sd["density_head.weight"] *= 200.0
    sd["density_head.bias"] += 1.0
    m = nerfw.NeRF(Config())
    m.load_state_dict(sd)
    return m.cuda(), emb.cuda()


def index_mismatch_rate(ex, u_rand, want_inds):
    """End-to-end sample-index mismatch rate: searchsorted indices of the resampling step as the
    GPU path computed them (its own coarse weights -> cdf -> inds) against the oracle's, same uniforms."""
    from nerfw import ops
    w = ex["weights_coarse"].squeeze(-1).contiguous()
    _, aux = ops.sample_pdf(ex["z_vals_coarse"], w, u_rand.shape[1], u_rand.cuda(), want_aux=True)
    return float((aux["inds"].cpu() != torch.from_numpy(want_inds).long()).float().mean())


@pytest.mark.parametrize("mode", ["fp32", "bf16x3", "mixed", "fp16", "bf16"])
def test_dense_scene_crop(cuda_model, oracle, golden, state_dict, mode):
    """density head x200, bias +1: acc ~ 1, the depth-sensitive case (SURVEY.md hard part 3).  Single pass: "mixed"
    resolves to bf16x3; the bars are the same 1x bars as everywhere else."""
    import nerfw
    m, emb_d = dense_model(state_dict)
    _, emb = state_dict
    o, d = view(oracle)
    g = golden("crop20_dense")
    with torch.no_grad():
        rgb, depth, ex = nerfw.volume_render(m, o[40:60, 40:60], d[40:60, 40:60], 2.0, 6.0, 64, 0,
                                             appearance_embedding=emb.cuda(), perturb=False, mlp_dtype=mode)
    e = dict(rgb=maxabs(rgb, g["rgb"]), depth=maxabs(depth, g["depth"]), acc=maxabs(ex["acc"], g["acc"]))
    record(f"crop20_dense_{mode}", **e)
    t = TOL[mode]
    assert e["rgb"] <= t[0] and e["depth"] <= t[1] and e["acc"] <= t[2], e


@pytest.mark.parametrize("reuse", [False, True])
@pytest.mark.parametrize("mode", ["fp32", "bf16x3", "mixed", "fp16", "bf16"])
def test_dense_hierarchical_golden(oracle, golden, state_dict, mode, reuse):
    """The opaque-scene case of the headline path: density x200 (mean acc 0.99), coarse 64 + fine 64+128, fixed uniforms,
    both the two-pass form and the default reuse_coarse form, against the composed oracle whose coarse half is asserted
    bit-equal to the reference's own volume_render (oracle/make_golden.py).  1x bars."""
    import nerfw
    m, emb = dense_model(state_dict)
    o, d = view(oracle)
    g = golden("crop24_dense_hier")
    oc, dc = o[38:62, 38:62].reshape(-1, 3), d[38:62, 38:62].reshape(-1, 3)
    u = torch.from_numpy(g["u_rand"])
    with torch.no_grad():
        rgb, depth, ex = nerfw.volume_render(m, oc, dc, 2.0, 6.0, 64, 128, appearance_embedding=emb, perturb=False,
                                             mlp_dtype=mode, u_rand=u, reuse_coarse=reuse)
    assert ex["z_vals"].shape == (576, 192)
    de = (depth.cpu() - torch.from_numpy(g["depth"])).abs().reshape(-1)
    e = dict(rgb=maxabs(rgb, g["rgb"]), depth=float(de.max()), acc=maxabs(ex["acc"], g["acc"]),
             depth_coarse=maxabs(ex["depth_coarse"], g["depth_coarse"]),
             w_coarse=maxabs(ex["weights_coarse"].squeeze(-1), g["weights_coarse"]),
             depth_p99=float(de.kthvalue(int(0.99 * de.numel())).values), rays_over_bar=float((de > 1e-3).sum()),
             ind_mismatch_rate=index_mismatch_rate(ex, u, g["inds"]), mean_acc=float(ex["acc"].mean()))
    record(f"crop24_dense_hier_{mode}{'_reuse' if reuse else ''}", **e)
    t = TOL[mode]
    assert e["rgb"] <= t[0] and e["depth"] <= t[1] and e["acc"] <= t[2], e
    if mode in ("fp32", "bf16x3", "mixed"):     # the coarse pass keeps fp32-level weights in these modes
        assert e["ind_mismatch_rate"] <= 1e-4, e


def test_perturbed_render_rng_parity(cuda_model, oracle, golden):
    """perturb=True with the reference's own uniforms (torch.manual_seed(5); rand(256,64)): z_vals bit-exact."""
    import nerfw
    model, emb = cuda_model
    o, d = view(oracle)
    g = golden("rays256_perturb")
    sel = torch.from_numpy(g["sel"]).cuda()
    torch.manual_seed(5)
    t_rand = torch.rand(256, 64)
    with torch.no_grad():
        rgb, depth, ex = nerfw.volume_render(model, o.reshape(-1, 3)[sel], d.reshape(-1, 3)[sel], 2.0, 6.0, 64, 0,
                                             appearance_embedding=emb, perturb=True, mlp_dtype="fp32", t_rand=t_rand)
    assert torch.equal(ex["z_vals"].cpu(), torch.from_numpy(g["z_vals"]))
    e = dict(rgb=maxabs(rgb, g["rgb"]), depth=maxabs(depth, g["depth"]), w=maxabs(ex["weights"][..., 0], g["weights"]))
    record("rays256_perturb_fp32", **e)
    assert e["rgb"] <= 1e-5 and e["depth"] <= 1e-4 and e["w"] <= 1e-6, e


@pytest.mark.parametrize("mode", ["fp32", "bf16x3", "mixed", "bf16", "fp16"])
def test_hierarchical_golden(cuda_model, oracle, golden, mode):
    """coarse 64 + fine 64+128 with the composed oracle's uniforms."""
    import nerfw
    model, emb = cuda_model
    o, d = view(oracle)
    g = golden("crop32_hier")
    oc = o[34:66, 34:66].reshape(-1, 3)
    dc = d[34:66, 34:66].reshape(-1, 3)
    with torch.no_grad():
        rgb, depth, ex = nerfw.volume_render(model, oc, dc, 2.0, 6.0, 64, 128, appearance_embedding=emb, perturb=False,
                                             mlp_dtype=mode, u_rand=torch.from_numpy(g["u_rand"]))
    assert ex["z_vals"].shape == (1024, 192) and ex["weights"].shape == (1024, 192, 1)
    zmis = float((ex["z_vals"].cpu() != torch.from_numpy(g["z_vals"])).float().mean())
    e = dict(rgb=maxabs(rgb, g["rgb"]), depth=maxabs(depth, g["depth"]), acc=maxabs(ex["acc"], g["acc"]),
             z=maxabs(ex["z_vals"], g["z_vals"]), z_mismatch_rate=zmis,
             ind_mismatch_rate=index_mismatch_rate(ex, torch.from_numpy(g["u_rand"]), g["inds"]))
    record(f"crop32_hier_{mode}", **e)
    t = TOL[mode]
    assert e["rgb"] <= t[0] and e["depth"] <= t[1] and e["acc"] <= t[2], e
    assert bool((ex["z_vals"][:, 1:] >= ex["z_vals"][:, :-1]).all())
    if mode in ("fp32", "bf16x3", "mixed"):
        # north star: "sample indices for fixed seeds, bit-exact" holds given identical cdf / u (tests/
        # test_gpu_rays_sampling.py); end to end the cdf carries the coarse weights' last-bit differences, which flip an
        # index whenever a u lands within ~1e-7 of a cdf entry.  At random init the pdf is nearly flat (64 evenly spaced cdf
        # entries against 128 strata), the worst case for such ties: measured 1.1e-4 with the all-fp32 kernel, 1.5e-4 with
        # bf16x3 / mixed (SURVEY.md a-5 (ii) guessed <= 1e-5); 0 / 2.7e-5 on the dense golden.  Stated bound: 5e-4.
        assert e["ind_mismatch_rate"] <= 5e-4, e


def test_chunked_equals_whole_and_cpu_inputs(cuda_model, oracle):
    """4096-ray chunks (the reference's calling pattern, render_aligned_spiral.py:136-152) == one whole call; CPU ray
    tensors are accepted and results come back on the CPU like any torch op would."""
    import nerfw
    model, emb = cuda_model
    o, d = view(oracle)
    o, d = o.reshape(-1, 3).contiguous(), d.reshape(-1, 3)
    with torch.no_grad():
        whole = nerfw.volume_render(model, o, d, 2.0, 6.0, 64, 0, appearance_embedding=emb, perturb=False)
        parts = [nerfw.volume_render(model, o[i:i + 4096], d[i:i + 4096], 2.0, 6.0, 64, 0, appearance_embedding=emb,
                                     perturb=False) for i in range(0, 10000, 4096)]
        cpu = nerfw.volume_render(model, o[:100].cpu(), d[:100].cpu(), 2.0, 6.0, 64, 0, appearance_embedding=emb.cpu(),
                                  perturb=False)
    assert torch.equal(torch.cat([p[0] for p in parts]), whole[0])
    assert torch.equal(torch.cat([p[1] for p in parts]), whole[1])
    assert not cpu[0].is_cuda and torch.equal(cpu[0], whole[0][:100].cpu())
    empty = nerfw.volume_render(model, o[:0], d[:0], 2.0, 6.0, 64, 0, perturb=False)
    assert empty[0].shape == (0, 3) and empty[1].shape == (0, 1)


@pytest.mark.parametrize("mode", ["fp32", None])
def test_training_gradients_golden(cuda_model, oracle, golden, manifest, mode):
    """loss.backward() through volume_render (src/train.py:77-91) on 64 rays: loss, embedding gradient, every bias /
    head gradient stored in the golden file, and all 24 gradient norms, vs autograd through the REFERENCE -- in the fp32
    mode and in the default mode (None: bf16x3 forward + tcgen05 bf16 backward, what a training script gets).
    Bounds: fp32 5e-4 of each tensor's max gradient, norms 1e-3, loss 1e-6; default 3e-2 / 1e-2 / 1e-5."""
    import nerfw
    model, emb = cuda_model
    o, d = view(oracle)
    g = golden("grads_64")
    sel = torch.from_numpy(g["sel"]).cuda()
    torch.manual_seed(9)
    t_rand = torch.rand(64, 64)
    emb_p = emb.clone().requires_grad_(True)
    model.zero_grad()
    tol, tol_norm, tol_loss = (5e-4, 1e-3, 1e-6) if mode == "fp32" else (3e-2, 1e-2, 1e-5)
    rgb, _, _ = nerfw.volume_render(model, o.reshape(-1, 3)[sel], d.reshape(-1, 3)[sel], 2.0, 6.0, 64, 0,
                                    appearance_embedding=emb_p, perturb=True, mlp_dtype=mode, t_rand=t_rand)
    loss = torch.nn.functional.mse_loss(rgb, torch.from_numpy(g["target"]).cuda())
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) <= tol_loss
    worst, worst_norm = 0.0, 0.0
    for k, p in model.named_parameters():
        key = k.replace(".", "__")
        ref_norm = manifest["cases"]["grads_64"]["grad_norms"][k]
        rel_n = abs(float(p.grad.double().norm()) - ref_norm) / (ref_norm + 1e-12)
        worst_norm = max(worst_norm, rel_n)
        assert rel_n <= tol_norm, (k, rel_n)
        if key in g.files:
            rel = maxabs(p.grad, g[key]) / (float(np.abs(g[key]).max()) + 1e-12)
            worst = max(worst, rel)
            assert rel <= tol, (k, rel)
    rel_e = maxabs(emb_p.grad, g["emb_grad"]) / (float(np.abs(g["emb_grad"]).max()) + 1e-12)
    record(f"train_grads_64_{mode or 'default'}", worst_rel=worst, worst_norm_rel=worst_norm, emb_rel=rel_e, loss=float(loss.detach()))
    assert rel_e <= tol
    model.zero_grad()


def test_render_frame_matches_chunked_reference_pattern(cuda_model, oracle):
    """Whole-frame call == the reference's chunk loop (render_aligned_spiral.py:136-158), and the uint8 image equals
    (rgb * 255).astype(uint8) of render_aligned_spiral.py:161."""
    import nerfw
    from nerfw.frame import quantize_frame, render_frame, render_path
    from nerfw.camera import aligned_spiral_poses, blender_focal
    model, emb = cuda_model
    pose = aligned_spiral_poses(120, 2, "x", "chair")[7]
    hh = ww = 48
    focal = blender_focal(ww)
    rgb, depth, acc = render_frame(model, hh, ww, focal, pose, 2.0, 6.0, 64, 0, appearance_embedding=emb)
    o, d = nerfw.get_rays(hh, ww, focal, torch.from_numpy(pose).cuda())
    chunks = []
    with torch.no_grad():
        for j in range(0, hh * ww, 1000):
            c, _, _ = nerfw.volume_render(model, o.reshape(-1, 3)[j:j + 1000], d.reshape(-1, 3)[j:j + 1000], 2.0, 6.0, 64, 0,
                                          appearance_embedding=emb, perturb=False)
            chunks.append(c.cpu())
    want = torch.cat(chunks).reshape(hh, ww, 3)
    assert torch.equal(rgb.cpu(), want)
    rgb8, depth8 = quantize_frame(rgb, depth)
    assert np.array_equal(rgb8.cpu().numpy(), (want * 255).numpy().astype(np.uint8))
    dn = depth.cpu().numpy()
    assert np.array_equal(depth8.cpu().numpy(), ((dn - dn.min()) / (dn.max() - dn.min()) * 255).astype(np.uint8))
    frames = list(render_path(model, aligned_spiral_poses(3, 1), 16, 16, blender_focal(16), 2.0, 6.0, 16, 16, appearance_embedding=emb))
    assert [f[0] for f in frames] == [0, 1, 2] and frames[0][1].shape == (16, 16, 3) and frames[0][1].dtype == np.uint8


def test_ray_bank_checkpoint_and_fog(cuda_model, oracle, tmp_path):
    """Rows N2-N4 of SURVEY.md section 8f: device ray bank == per-image get_rays; checkpoint schema round trip with the
    reference's keys; fog on float buffers == the reference's numpy formula (src/post_processor.py:451-493)."""
    import nerfw
    from nerfw.checkpoint import load_checkpoint, save_checkpoint
    from nerfw.effects import fog
    from nerfw.raybank import RayBank
    from nerfw.camera import aligned_spiral_poses, blender_focal
    from config import Config
    model, emb = cuda_model
    poses = torch.from_numpy(aligned_spiral_poses(3, 1))
    imgs = torch.rand(3, 20, 24, 4)
    focal = blender_focal(24)
    bank = RayBank(imgs, poses, focal)
    _, want = oracle.rays_for_view(20, 24, focal, poses[1])
    full = bank.image_rays(1)
    assert torch.equal(full["rays_d"].cpu(), want.reshape(-1, 3)) and full["rgb"].shape == (480, 3)
    g = torch.Generator(device="cuda").manual_seed(0)
    b = bank.sample(100, generator=g)
    assert b["rays_d"].shape == (100, 3) and b["rgb"].shape == (100, 3) and b["alpha"].shape == (100, 1)
    assert 0 <= b["img_idx"] < 3 and torch.equal(b["rays_o"][0].cpu(), poses[b["img_idx"], :3, 3])
    # checkpoint
    table = torch.nn.Parameter(torch.randn(3, 32))
    path = save_checkpoint(str(tmp_path), 1000, model, table, loss=0.5, psnr=3.0)
    assert path.endswith("checkpoint_001000.pt")
    ck = torch.load(path, weights_only=True)
    assert set(ck) == {"model_state_dict", "optimizer_state_dict", "loss", "psnr", "iteration", "appearance_embeddings"}
    m2 = nerfw.NeRF(Config())
    t2 = torch.nn.Parameter(torch.zeros(3, 32))
    load_checkpoint(path, m2, t2)
    assert all(torch.equal(a.cpu(), b_) for a, b_ in zip(model.state_dict().values(), m2.state_dict().values()))
    assert torch.equal(t2.data, table.data)
    # fog
    rgb = torch.rand(16, 16, 3, device="cuda")
    depth = torch.rand(16, 16, device="cuda") * 4 + 2
    got = fog(rgb, depth, fog_start=0.0).cpu().numpy()
    img8 = (rgb.cpu() * 255).numpy().astype(np.uint8)
    dn = depth.cpu().numpy()
    dn = dn / dn.max()
    adj = np.clip(np.maximum(dn - 0.0, 0.0) / 1.0, 0.0, 1.0) ** 3.0 * 0.3
    f3 = np.stack([adj] * 3, axis=2)
    ref = np.clip(img8.astype(np.float32) * f3 + np.array([255, 255, 255], np.float32) * (1.0 - f3), 0, 255).astype(np.uint8)
    assert int(np.abs(got.astype(int) - ref.astype(int)).max()) <= 1   # see tests/test_gpu_effects.py for the reference-generated vectors


@pytest.mark.parametrize("mode", ["bf16x3", "mixed", "bf16"])
def test_high_sample_stress_config(cuda_model, oracle, state_dict, mode):
    """BASELINE.json configs[4]: 256 + 512 samples (fine pass on 768), depth output, vs the composed oracle on 48 rays."""
    import nerfw
    model, emb = cuda_model
    sd, emb_cpu = state_dict
    o, d = view(oracle)
    sel = torch.arange(0, 10000, 209, device="cuda")[:48]
    oc, dc = o.reshape(-1, 3)[sel].contiguous(), d.reshape(-1, 3)[sel].contiguous()
    torch.manual_seed(17)
    u = torch.rand(48, 512)
    with torch.no_grad():
        rgb_o, depth_o, ex_o = oracle.render_hier(sd, sd, oc.cpu(), dc.cpu(), 2.0, 6.0, 256, 512, emb=emb_cpu, perturb=False, u_rand=u)
        rgb, depth, ex = nerfw.volume_render(model, oc, dc, 2.0, 6.0, 256, 512, appearance_embedding=emb, perturb=False,
                                             mlp_dtype=mode, u_rand=u)
    assert ex["z_vals"].shape == (48, 768)
    e = dict(rgb=maxabs(rgb, rgb_o), depth=maxabs(depth, depth_o), acc=maxabs(ex["acc"], ex_o["acc"]))
    record(f"stress_256_512_{mode}", **e)
    t = TOL[mode]
    assert e["rgb"] <= t[0] and e["depth"] <= t[1] and e["acc"] <= t[2], e


def test_full_frame_properties(cuda_model):
    """BASELINE.json configs[1] at full size (800x800, 64+128) through the whole-frame call: value ranges that hold for any
    weights, determinism, and independence of where a ray sits in the batch (tile / CTA assignment)."""
    import nerfw
    from nerfw.camera import aligned_spiral_poses, blender_focal
    from nerfw.frame import render_frame
    model, emb = cuda_model
    pose = aligned_spiral_poses(120, 2, "x", "chair")[0]
    g = torch.Generator(device="cuda").manual_seed(11)
    rgb, depth, acc = render_frame(model, 800, 800, blender_focal(800), pose, 2.0, 6.0, 64, 128, appearance_embedding=emb, generator=g)
    assert rgb.shape == (800, 800, 3) and depth.shape == (800, 800) and acc.shape == (800, 800)
    assert bool(torch.isfinite(rgb).all()) and bool(torch.isfinite(depth).all())
    assert float(rgb.min()) >= 0.0 and float(rgb.max()) <= 1.0 + 1e-5        # sum w * sigmoid with sum w <= 1
    assert float(acc.min()) >= 0.0 and float(acc.max()) <= 1.0 + 1e-5
    hit = acc > 1e-6
    assert float(depth[hit].min()) >= 2.0 - 1e-3 and float(depth[hit].max()) <= 6.0 + 1e-3
    g2 = torch.Generator(device="cuda").manual_seed(11)
    rgb2, depth2, _ = render_frame(model, 800, 800, blender_focal(800), pose, 2.0, 6.0, 64, 128, appearance_embedding=emb, generator=g2)
    assert torch.equal(rgb, rgb2) and torch.equal(depth, depth2)
    # coarse pass (no randomness): a 4096-ray slice rendered alone equals the same rays inside the whole frame
    o, d = nerfw.get_rays(800, 800, blender_focal(800), torch.from_numpy(pose).cuda())
    o, d = o.reshape(-1, 3).contiguous(), d.reshape(-1, 3)
    with torch.no_grad():
        whole = nerfw.volume_render(model, o, d, 2.0, 6.0, 64, 0, appearance_embedding=emb, perturb=False)
        part = nerfw.volume_render(model, o[333333:337429], d[333333:337429], 2.0, 6.0, 64, 0, appearance_embedding=emb, perturb=False)
    assert torch.equal(whole[0][333333:337429], part[0]) and torch.equal(whole[1][333333:337429], part[1])


@pytest.mark.parametrize("n,ni", [(37, 53), (8, 200), (2, 1), (130, 64)])
def test_odd_sample_counts_end_to_end(cuda_model, oracle, state_dict, n, ni):
    """Sample counts that are multiples of nothing (tile / warp / chunk remainders everywhere) vs the composed oracle."""
    import nerfw
    model, emb = cuda_model
    sd, emb_cpu = state_dict
    o, d = view(oracle)
    sel = torch.arange(3, 10000, 131, device="cuda")[:70]
    oc, dc = o.reshape(-1, 3)[sel].contiguous(), d.reshape(-1, 3)[sel].contiguous()
    torch.manual_seed(n * 1000 + ni)
    u = torch.rand(70, ni)
    t = torch.rand(70, n)
    with torch.no_grad():
        rgb_o, depth_o, ex_o = oracle.render_hier(sd, sd, oc.cpu(), dc.cpu(), 2.0, 6.0, n, ni, emb=emb_cpu, perturb=True,
                                                  t_rand=t, u_rand=u)
        rgb, depth, ex = nerfw.volume_render(model, oc, dc, 2.0, 6.0, n, ni, appearance_embedding=emb, perturb=True,
                                             mlp_dtype="bf16x3", t_rand=t, u_rand=u)
    assert ex["z_vals"].shape == (70, n + ni) and torch.equal(ex["z_vals_coarse"].cpu(), ex_o["z_vals_coarse"])
    e = dict(rgb=maxabs(rgb, rgb_o), depth=maxabs(depth, depth_o), acc=maxabs(ex["acc"], ex_o["acc"]))
    record(f"odd_counts_{n}_{ni}", **e)
    assert e["rgb"] <= 1e-3 and e["depth"] <= 2e-3 and e["acc"] <= 1e-3, e


def test_reuse_coarse_equals_two_pass(cuda_model, oracle, golden):
    """reuse_coarse (the inference default with one network for both passes): the fine pass evaluates only the 128 new
    samples and the coarse outputs are merged in by depth.  In bf16x3 mode every sample is evaluated by the same kernel either way, so the
    result is bit-identical to the default path; in the default mixed mode it stays inside the fp32 parity bars."""
    import nerfw
    model, emb = cuda_model
    o, d = view(oracle)
    g = golden("crop32_hier")
    oc = o[34:66, 34:66].reshape(-1, 3)
    dc = d[34:66, 34:66].reshape(-1, 3)
    u = torch.from_numpy(g["u_rand"])
    kw = dict(appearance_embedding=emb, perturb=False, u_rand=u)
    with torch.no_grad():
        ref = nerfw.volume_render(model, oc, dc, 2.0, 6.0, 64, 128, mlp_dtype="bf16x3", reuse_coarse=False, **kw)
        got = nerfw.volume_render(model, oc, dc, 2.0, 6.0, 64, 128, mlp_dtype="bf16x3", **kw)   # default: reuse
        assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1]) and torch.equal(got[2]["weights"], ref[2]["weights"])
        assert torch.equal(got[2]["z_vals"], ref[2]["z_vals"])
        mix = nerfw.volume_render(model, oc, dc, 2.0, 6.0, 64, 128, mlp_dtype="mixed", reuse_coarse=True, **kw)
        e = dict(rgb=maxabs(mix[0], g["rgb"]), depth=maxabs(mix[1], g["depth"]), acc=maxabs(mix[2]["acc"], g["acc"]))
        record("crop32_hier_mixed_reuse_coarse", **e)
        assert e["rgb"] <= 1e-3 and e["depth"] <= 1e-3 and e["acc"] <= 1e-3, e
        # injected uniforms outside [0,1): the fine list is not sorted, the merge ranks by counting
        u_bad = u.clone()
        u_bad[:64] *= 3.0
        kw_bad = dict(appearance_embedding=emb, perturb=False, u_rand=u_bad)
        ref_b = nerfw.volume_render(model, oc, dc, 2.0, 6.0, 64, 128, mlp_dtype="bf16x3", reuse_coarse=False, **kw_bad)
        got_b = nerfw.volume_render(model, oc, dc, 2.0, 6.0, 64, 128, mlp_dtype="bf16x3", reuse_coarse=True, **kw_bad)
        assert torch.equal(got_b[0], ref_b[0]) and torch.equal(got_b[1], ref_b[1])
        # fp32 CUDA-core mode: same property; and a (coarse, fine) pair or a recorded graph never takes the reuse path
        r32 = nerfw.volume_render(model, oc, dc, 2.0, 6.0, 64, 128, mlp_dtype="fp32", reuse_coarse=False, **kw)
        g32 = nerfw.volume_render(model, oc, dc, 2.0, 6.0, 64, 128, mlp_dtype="fp32", **kw)
        assert torch.equal(g32[0], r32[0]) and torch.equal(g32[1], r32[1])
        pair = nerfw.volume_render((model, model), oc, dc, 2.0, 6.0, 64, 128, mlp_dtype="bf16x3", **kw)
        assert torch.equal(pair[0], ref[0])
    # N + NI = 4096: the merge kernel's 64 KB of dynamic shared memory (above the 48 KB default limit)
    with torch.no_grad():
        ub = torch.rand(8, 3584, generator=torch.Generator().manual_seed(3))
        big = nerfw.volume_render(model, oc[:8], dc[:8], 2.0, 6.0, 512, 3584, appearance_embedding=emb, perturb=False,
                                  mlp_dtype="bf16x3", u_rand=ub)
        big2 = nerfw.volume_render(model, oc[:8], dc[:8], 2.0, 6.0, 512, 3584, appearance_embedding=emb, perturb=False,
                                   mlp_dtype="bf16x3", u_rand=ub, reuse_coarse=False)
    assert big[2]["z_vals"].shape == (8, 4096) and bool(torch.isfinite(big[0]).all())
    assert torch.equal(big[0], big2[0]) and torch.equal(big[1], big2[1])


@pytest.mark.parametrize("mode", ["fp32", "bf16x3", "mixed", "bf16"])
def test_fused_call_equals_call_by_call(cuda_model, oracle, mode):
    """nerfw_volume_render (one library call per render, the inference default) launches the same kernels in the same order
    as the call-by-call path: every output is bit-identical -- hierarchical with re-use and single pass, shared / per-ray /
    no embedding, jittered depths, ragged ray counts."""
    import nerfw
    model, emb = cuda_model
    o, d = view(oracle)
    gen = torch.Generator().manual_seed(11)
    for b, rows in ((1000, 1), (129, 0), (777, 777), (1, 1)):
        oc = o.reshape(-1, 3)[1234:1234 + b]
        dc = d.reshape(-1, 3)[1234:1234 + b] * 1.7          # not normalised: the call normalises like src/render.py:19
        e = None if rows == 0 else (emb if rows == 1 else torch.randn(rows, 32, generator=gen).cuda())
        t = torch.rand(b, 64, generator=gen)
        u = torch.rand(b, 128, generator=gen)
        for ni, fine in ((128, True), (128, False), (0, True)):
            for perturb in (False, True):
                kw = dict(appearance_embedding=e, perturb=perturb, t_rand=t if perturb else None, u_rand=u, mlp_dtype=mode,
                          fine_pass=fine)
                with torch.no_grad():
                    a = nerfw.volume_render(model, oc, dc, 2.0, 6.0, 64, ni, fused=False, **kw)
                    f = nerfw.volume_render(model, oc, dc, 2.0, 6.0, 64, ni, fused=True, **kw)
                assert torch.equal(a[0], f[0]) and torch.equal(a[1], f[1])
                assert set(a[2]) == set(f[2]), (sorted(a[2]), sorted(f[2]))
                for k in a[2]:
                    assert a[2][k].shape == f[2][k].shape and torch.equal(a[2][k], f[2][k]), (k, b, rows, ni, fine, perturb)
    # gradients recorded -> never the fused call (it has no backward); CPU rays come back on the CPU
    with torch.enable_grad():
        rgb, _, _ = nerfw.volume_render(model, o.reshape(-1, 3)[:64], d.reshape(-1, 3)[:64], 2.0, 6.0, 64, 128,
                                        appearance_embedding=emb, mlp_dtype=mode)
        assert rgb.requires_grad
    with torch.no_grad():
        rgb, depth, ex = nerfw.volume_render(model, o.reshape(-1, 3)[:64].cpu(), d.reshape(-1, 3)[:64].cpu(), 2.0, 6.0, 64, 128,
                                             appearance_embedding=emb, perturb=False, mlp_dtype=mode)
    assert rgb.device.type == "cpu" and depth.device.type == "cpu" and ex["z_vals"].device.type == "cpu"


def test_volume_render_abi_errors(cuda_model):
    """Raw C-ABI behaviour of nerfw_volume_render: too small a workspace, missing coarse outputs, zero rays."""
    import ctypes as C
    from nerfw import _lib, ops
    model, emb = cuda_model
    lib = _lib.lib()
    b, n, ni = 256, 64, 128
    need = lib.nerfw_volume_render_workspace_bytes(b, n, ni, 1)
    assert need >= b * (n + ni + ni + n) * 16 and lib.nerfw_volume_render_workspace_bytes(-1, n, ni, 1) == 0
    ws = model.kernel_state()[2]
    packed = model.packed_weights()
    o = torch.zeros(b, 3, device="cuda"); o[:, 2] = 4.0
    d = torch.nn.functional.normalize(torch.randn(b, 3, device="cuda") * 0.2 + torch.tensor([0.0, 0.0, -1.0], device="cuda"), dim=-1)
    ztab = ops.depth_table(2.0, 6.0, n, o.device)
    ulin = ops.u_table(ni, o.device)
    u = torch.rand(b, ni, device="cuda")
    buf = torch.empty(need, dtype=torch.uint8, device="cuda")
    outs = {k: torch.empty(b * (n + ni), device="cuda") for k, _ in _lib.NerfwRenderOut._fields_}
    out = _lib.NerfwRenderOut()
    for k, t in outs.items():
        setattr(out, k, t.data_ptr())
    args = lambda o_struct, wsb: (C.byref(ws), packed.data_ptr(), o.data_ptr(), d.data_ptr(), b, ztab.data_ptr(), None, n,
                                  ulin.data_ptr(), u.data_ptr(), ni, emb.reshape(1, 32).contiguous().data_ptr(), 1, 1, 3,
                                  C.byref(o_struct), buf.data_ptr(), wsb, None)
    assert lib.nerfw_volume_render(*args(out, need)) == 0
    assert lib.nerfw_volume_render(*args(out, need - 16)) == -4 and b"workspace" in lib.nerfw_last_error()
    bad = _lib.NerfwRenderOut()
    for k, t in outs.items():
        setattr(bad, k, t.data_ptr())
    bad.weights_coarse = None
    assert lib.nerfw_volume_render(*args(bad, need)) == -1 and b"coarse output" in lib.nerfw_last_error()
    a0 = list(args(out, need)); a0[4] = 0
    assert lib.nerfw_volume_render(*a0) == 0
    torch.cuda.synchronize()
