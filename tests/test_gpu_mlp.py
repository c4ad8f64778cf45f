"""K4/K5: the fused MLP forward (fp32 FFMA, bf16x3 and bf16 tcgen05 modes) and backward against the reference's
NeRF.forward (golden) and the oracle / its autograd.  Tolerances are written next to each assertion."""
import pytest
import torch

from gpu_util import maxabs, record

pytestmark = pytest.mark.gpu

# max-abs bounds on (rgb, sigma) of one MLP evaluation at random init (|sigma| ~ 0.1, rgb ~ 0.5).  In bf16x3 inference the
# direction layer (which only feeds the rgb sigmoid) runs as a single bf16 MMA: per-sample rgb 7e-5, sigma stays 1e-7.
FWD_TOL = {"fp32": (2e-6, 2e-6), "bf16x3": (5e-4, 2e-5), "bf16": (8e-3, 8e-3), "fp16": (1.5e-3, 1.5e-3)}


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("n,k", [(256, 64), (256, 256), (128, 128), (16, 64), (48, 192)])
def test_umma_primitives(mode, n, k):
    """One tcgen05 tile D = A B^T: pins the smem descriptor + 128B swizzle (mode 0) and the TMEM A layout (mode 1)."""
    from nerfw import ops
    g = torch.Generator().manual_seed(n + k + mode)
    a = torch.randn(128, k, generator=g).bfloat16()
    b = torch.randn(n, k, generator=g).bfloat16()
    want = a.float() @ b.float().t()
    got = ops.selftest_umma(a.cuda(), b.cuda(), mode)
    err = maxabs(got, want)
    record(f"umma_{'ts' if mode else 'ss'}_{n}x{k}", maxabs=err)
    assert err <= 1e-3, err   # exact bf16 products, fp32 accumulation order only


@pytest.mark.parametrize("n,k", [(64, 64), (256, 128), (128, 256), (256, 16), (192, 48)])
def test_umma_mn_major_primitive(n, k):
    """Both operands MN-major (the wgrad form dW = dZ^T X): D = At^T Bt."""
    from nerfw import ops
    g = torch.Generator().manual_seed(n * 7 + k)
    at = torch.randn(k, 128, generator=g).bfloat16()
    bt = torch.randn(k, n, generator=g).bfloat16()
    want = at.float().t() @ bt.float()
    got = ops.selftest_umma_mn(at.cuda(), bt.cuda())
    err = maxabs(got, want)
    record(f"umma_mn_{n}x{k}", maxabs=err)
    assert err <= 1e-3, err


@pytest.mark.parametrize("mode", ["fp32", "bf16x3", "bf16", "fp16"])
def test_mlp_forward_golden(cuda_model, golden, mode):
    model, emb = cuda_model
    g = golden("mlp_64")
    x, d = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["d"]).cuda()
    model.mlp_mode = mode
    try:
        with torch.no_grad():
            rgb, sigma = model(x, d, emb.unsqueeze(0))
            rgb_n, sigma_n = model(x, d, None)
    finally:
        model.mlp_mode = None
    assert rgb.shape == (64, 3) and sigma.shape == (64, 1)
    e = dict(rgb=maxabs(rgb, g["rgb"]), sigma=maxabs(sigma, g["sigma"]), rgb_noemb=maxabs(rgb_n, g["rgb_noemb"]),
             sigma_noemb=maxabs(sigma_n, g["sigma_noemb"]))
    record(f"mlp_fwd_golden_{mode}", **e)
    tr, ts = FWD_TOL[mode]
    assert e["rgb"] <= tr and e["rgb_noemb"] <= tr and e["sigma"] <= ts and e["sigma_noemb"] <= ts, e


@pytest.mark.parametrize("mode", ["fp32", "bf16x3", "bf16", "fp16"])
def test_mlp_forward_shapes_and_embeddings(cuda_model, oracle, state_dict, mode):
    """ragged sizes (not multiples of the 64/128-sample tiles), (D,), (1,D) and per-sample (S,D) embeddings."""
    model, emb = cuda_model
    sd, _ = state_dict
    gen = torch.Generator().manual_seed(77)
    tr, ts = FWD_TOL[mode]
    model.mlp_mode = mode
    try:
        for s in (1, 10, 63, 129, 1000):
            x = (torch.rand(s, 3, generator=gen) - 0.5) * 8
            d = torch.nn.functional.normalize(torch.randn(s, 3, generator=gen), dim=-1)
            e_rows = torch.randn(s, 32, generator=gen)
            for e in (None, e_rows[0], e_rows[:1], e_rows):
                want_rgb, want_sigma = oracle.mlp_forward(sd, x, d, e)
                with torch.no_grad():
                    rgb, sigma = model(x.cuda(), d.cuda(), None if e is None else e.cuda())
                assert maxabs(rgb, want_rgb) <= tr * 2 and maxabs(sigma, want_sigma) <= ts * 2, (s, None if e is None else e.shape)
    finally:
        model.mlp_mode = None


def test_mlp_ray_form_equals_sample_form(cuda_model):
    from nerfw import ops
    model, emb = cuda_model
    gen = torch.Generator().manual_seed(5)
    b, n = 37, 64
    o = torch.randn(b, 3, generator=gen).cuda()
    d = torch.nn.functional.normalize(torch.randn(b, 3, generator=gen), dim=-1).cuda()
    z = torch.sort(torch.rand(b, n, generator=gen) * 4 + 2, dim=-1).values.cuda()
    pts = ops.ray_points(o, d, z).reshape(-1, 3)
    dirs = d.unsqueeze(1).expand(-1, n, -1).reshape(-1, 3).contiguous()
    e = emb.unsqueeze(0)
    for mode in ("fp32", "bf16x3", "bf16"):
        with torch.no_grad():
            a = model.run_mlp(o, d, z, e, mode)
            b_ = model.run_mlp(pts, dirs, None, e, mode)
        assert torch.equal(a, b_), mode   # identical arithmetic, only the addressing differs


@pytest.mark.parametrize("mode,tol", [("bf16x3", 5e-4), ("bf16", 1.5e-2)])
def test_tensor_core_mlp_vs_fp32_kernel_large(cuda_model, mode, tol):
    """200k samples (1563 tiles over 148 persistent CTAs): exercises the weight ring, phase wrap-around and tile loop."""
    model, emb = cuda_model
    gen = torch.Generator(device="cuda").manual_seed(3)
    b, n = 3125, 64
    o = torch.randn(b, 3, device="cuda", generator=gen)
    d = torch.nn.functional.normalize(torch.randn(b, 3, device="cuda", generator=gen), dim=-1)
    z = torch.sort(torch.rand(b, n, device="cuda", generator=gen) * 4 + 2, dim=-1).values
    e = emb.unsqueeze(0)
    with torch.no_grad():
        ref = model.run_mlp(o, d, z, e, "fp32")
        got = model.run_mlp(o, d, z, e, mode)
        again = model.run_mlp(o, d, z, e, mode)
    err = maxabs(got, ref)
    err_sigma = maxabs(got[:, 3], ref[:, 3])
    record(f"mlp_tc_vs_ffma_{mode}", maxabs=err, sigma=err_sigma)
    assert err <= tol, err
    if mode == "bf16x3":
        assert err_sigma <= 3e-5, err_sigma   # the trunk (density) keeps the 3-term split
    assert torch.equal(got, again)   # deterministic


def test_packed_cache_tracks_parameter_updates(cuda_model):
    model, emb = cuda_model
    x = torch.rand(130, 3, device="cuda")
    d = torch.nn.functional.normalize(torch.randn(130, 3, device="cuda"), dim=-1)
    with torch.no_grad():
        a = model.run_mlp(x, d, None, None, "bf16x3")
        saved = model.rgb_linear.bias.clone()
        model.rgb_linear.bias.add_(1.0)          # bumps ._version -> the bf16 image must be rebuilt
        b = model.run_mlp(x, d, None, None, "bf16x3")
        model.rgb_linear.bias.copy_(saved)
        c = model.run_mlp(x, d, None, None, "bf16x3")
    assert float((b[:, :3] - a[:, :3]).min()) > 0.0
    assert torch.equal(a, c)


# Stated bounds of the backward kernels against float64 autograd through the oracle: (max-abs error / max-abs gradient per
# tensor, cosine per tensor).  fp32 CUDA-core kernel: summation order only.  Default path (forward bf16x3, tcgen05 backward
# with bf16 operands and fp32 accumulation, gates from the forward): bf16 operand rounding, 2^-9 per product.
BWD_TOL = {"fp32": (2e-4, 0.9999999), None: (3e-2, 0.9999)}


@pytest.mark.parametrize("mode", ["fp32", None])
def test_mlp_backward_vs_oracle_autograd(cuda_model, oracle, state_dict, mode):
    """All 24 parameter gradients + the embedding gradient of sum(raw * cotangent), vs autograd through the oracle
    in float64 on the same inputs -- for the fp32 kernel AND for the default path (mode None: what `loss.backward()`
    runs unless told otherwise, i.e. the tcgen05 backward), shared and per-sample embeddings."""
    model, emb = cuda_model
    sd, _ = state_dict
    gen = torch.Generator().manual_seed(21)
    s = 200
    x = (torch.rand(s, 3, generator=gen) - 0.5) * 6
    d = torch.nn.functional.normalize(torch.randn(s, 3, generator=gen), dim=-1)
    cot = torch.randn(s, 4, generator=gen)
    tol, tol_cos = BWD_TOL[mode]
    for rows in (1, s):
        e = torch.randn(rows, 32, generator=gen)
        sd64 = {k: v.double().requires_grad_(True) for k, v in sd.items()}
        e64 = e.double().requires_grad_(True)
        rgb, sigma = oracle.mlp_forward(sd64, x.double(), d.double(), e64)
        (torch.cat([rgb, sigma], dim=-1) * cot.double()).sum().backward()
        model.zero_grad()
        eg = e.cuda().requires_grad_(True)
        model.mlp_mode = mode
        try:
            rgb_g, sigma_g = model(x.cuda(), d.cuda(), eg)
        finally:
            model.mlp_mode = None
        (torch.cat([rgb_g, sigma_g], dim=-1) * cot.cuda()).sum().backward()
        worst, worst_cos = 0.0, 1.0
        for k, p in model.named_parameters():
            ref = sd64[k].grad
            rel = maxabs(p.grad, ref) / (float(ref.abs().max()) + 1e-12)
            cos = float((p.grad.double().cpu() * ref).sum() / (p.grad.double().norm().cpu() * ref.norm() + 1e-30))
            worst, worst_cos = max(worst, rel), min(worst_cos, cos)
            assert rel <= tol and cos >= tol_cos, (k, rel, cos, rows, mode)
        rel_e = maxabs(eg.grad, e64.grad) / (float(e64.grad.abs().max()) + 1e-12)
        record(f"mlp_bwd_vs_oracle_{mode or 'default'}_rows{rows}", worst_param_rel=worst, worst_cos=worst_cos, emb_rel=rel_e)
        assert rel_e <= tol, rel_e
    model.zero_grad()


@pytest.mark.parametrize("with_emb", [True, False])
@pytest.mark.parametrize("b,n", [(3, 64), (61, 192), (700, 64)])
def test_tensor_core_backward_vs_fp32_backward(cuda_model, b, n, with_emb):
    """tcgen05 backward (bf16 operands, fp32 accumulate, ReLU gates taken from the bf16x3 forward) against the fp32
    CUDA-core backward on the same inputs and cotangent.  Stated bf16 bounds: every parameter gradient within 3e-2 of
    that tensor's max-abs gradient and cosine >= 0.9999 (measured 1.5e-2 / 0.99996)."""
    from nerfw import ops
    model, emb = cuda_model
    gen = torch.Generator(device="cuda").manual_seed(b * 13 + n)
    o = torch.randn(b, 3, device="cuda", generator=gen)
    d = torch.nn.functional.normalize(torch.randn(b, 3, device="cuda", generator=gen), dim=-1)
    z = torch.sort(torch.rand(b, n, device="cuda", generator=gen) * 4 + 2, dim=-1).values
    d_raw = torch.randn(b * n, 4, device="cuda", generator=gen)
    e = emb.unsqueeze(0).contiguous() if with_emb else None
    names, tensors = model.kernel_params()
    params = {k: t.detach() for k, t in zip(names, tensors)}
    packed = model.packed_weights(names, tensors)
    g_ref = {k: torch.zeros_like(t) for k, t in params.items()}
    g_tc = {k: torch.zeros_like(t) for k, t in params.items()}
    de_ref = torch.zeros(1, 32, device="cuda") if with_emb else None
    de_tc = torch.zeros(1, 32, device="cuda") if with_emb else None
    ops.mlp_bwd(params, g_ref, o, d, z, e, d_raw, de_ref)
    _, masks = ops.mlp_fwd(params, packed, o, d, z, e, 1, want_masks=True)   # ReLU gates of the bf16x3 forward
    ops.mlp_bwd_tc(params, g_tc, packed, o, d, z, e, d_raw, de_tc, masks)
    worst, worst_cos = 0.0, 1.0
    for k in params:
        if not with_emb and k.startswith("appearance"):
            assert float(g_tc[k].abs().max()) == 0.0
            continue
        ref, got = g_ref[k].double(), g_tc[k].double()
        rel = float((ref - got).abs().max() / (ref.abs().max() + 1e-20))
        cos = float((ref * got).sum() / (ref.norm() * got.norm() + 1e-30))
        worst, worst_cos = max(worst, rel), min(worst_cos, cos)
        assert rel <= 3e-2 and cos >= 0.9999, (k, rel, cos)
    if with_emb:
        rel = float((de_ref - de_tc).abs().max() / (de_ref.abs().max() + 1e-20))
        assert rel <= 1e-2, rel
    record(f"mlp_bwd_tc_{b}x{n}_{'emb' if with_emb else 'noemb'}", worst_rel=worst, worst_cos=worst_cos)


def test_tensor_core_backward_without_forward_gates(cuda_model):
    """relu_masks = NULL: the backward gates with its own bf16 recompute.  Near-zero activations then gate differently
    from the fp32 reference (a fraction f of flipped gates moves the gradient by ~sqrt(f) per layer), so only the looser
    bound cos >= 0.99 holds -- which is why the autograd path always passes the forward's gates."""
    from nerfw import ops
    model, emb = cuda_model
    gen = torch.Generator(device="cuda").manual_seed(99)
    b, n = 300, 64
    o = torch.randn(b, 3, device="cuda", generator=gen)
    d = torch.nn.functional.normalize(torch.randn(b, 3, device="cuda", generator=gen), dim=-1)
    z = torch.sort(torch.rand(b, n, device="cuda", generator=gen) * 4 + 2, dim=-1).values
    d_raw = torch.randn(b * n, 4, device="cuda", generator=gen)
    names, tensors = model.kernel_params()
    params = {k: t.detach() for k, t in zip(names, tensors)}
    packed = model.packed_weights(names, tensors)
    g_ref = {k: torch.zeros_like(t) for k, t in params.items()}
    g_tc = {k: torch.zeros_like(t) for k, t in params.items()}
    ops.mlp_bwd(params, g_ref, o, d, z, None, d_raw, None)
    ops.mlp_bwd_tc(params, g_tc, packed, o, d, z, None, d_raw, None, None)
    for k in params:
        if k.startswith("appearance"):
            continue
        ref, got = g_ref[k].double(), g_tc[k].double()
        cos = float((ref * got).sum() / (ref.norm() * got.norm() + 1e-30))
        assert cos >= 0.99, (k, cos)


def test_training_forward_keeps_split_direction_layer(cuda_model, golden):
    """With ReLU masks requested (training) the bf16x3 forward also splits the direction layer: fp32-level rgb."""
    from nerfw import ops
    model, emb = cuda_model
    g = golden("mlp_64")
    x, d = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["d"]).cuda()
    names, tensors = model.kernel_params()
    params = {k: t.detach() for k, t in zip(names, tensors)}
    packed = model.packed_weights(names, tensors)
    raw, masks = ops.mlp_fwd(params, packed, x, d, None, emb.unsqueeze(0).contiguous(), 1, want_masks=True)
    e_rgb, e_sig = maxabs(raw[:, :3], g["rgb"]), maxabs(raw[:, 3:], g["sigma"])
    record("mlp_fwd_golden_bf16x3_training", rgb=e_rgb, sigma=e_sig)
    assert e_rgb <= 2e-5 and e_sig <= 2e-5, (e_rgb, e_sig)
    assert masks.numel() == 9 * 128 * 8 * 4


@pytest.mark.parametrize("mode", ["bf16x3", "fp16", "bf16"])
def test_sigma_only_flag(cuda_model, oracle, mode):
    """NERFW_MLP_SIGMA_ONLY: same sigma bits as the full forward, rgb = 0; and the hierarchical render is bit-identical
    whether the coarse pass evaluates colour or not (inference default: it does not)."""
    import nerfw
    from nerfw import ops
    model, emb = cuda_model
    names, tensors = model.kernel_params()
    params = {k: t.detach() for k, t in zip(names, tensors)}
    packed = model.packed_weights(names, tensors)
    mode_id = nerfw.models.resolve_mode(mode)
    g = torch.Generator(device="cuda").manual_seed(4)
    for b, n in ((37, 64), (300, 129)):
        o = torch.randn(b, 3, device="cuda", generator=g)
        d = torch.nn.functional.normalize(torch.randn(b, 3, device="cuda", generator=g), dim=-1)
        z = torch.sort(torch.rand(b, n, device="cuda", generator=g) * 4 + 2, dim=-1).values
        full = ops.mlp_fwd(params, packed, o, d, z, emb.reshape(1, -1), mode_id)
        sig = ops.mlp_fwd(params, packed, o, d, z, emb.reshape(1, -1), mode_id, sigma_only=True)
        assert torch.equal(sig[:, 3], full[:, 3]) and float(sig[:, :3].abs().max()) == 0.0
    h, w, focal, c2w = oracle.golden_camera()
    ro, rd = nerfw.get_rays(h, w, focal, c2w.cuda())
    oc, dc = ro[40:60, 40:60].reshape(-1, 3), rd[40:60, 40:60].reshape(-1, 3)
    u = torch.rand(400, 128, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        a = nerfw.volume_render(model, oc, dc, 2.0, 6.0, 64, 128, appearance_embedding=emb, perturb=False, mlp_dtype=mode, u_rand=u,
                                reuse_coarse=False)
        bb = nerfw.volume_render(model, oc, dc, 2.0, 6.0, 64, 128, appearance_embedding=emb, perturb=False, mlp_dtype=mode, u_rand=u,
                                 coarse_rgb=True, reuse_coarse=False)
    assert "rgb_coarse" not in a[2] and "rgb_coarse" in bb[2]
    assert torch.equal(a[0], bb[0]) and torch.equal(a[1], bb[1]) and torch.equal(a[2]["z_vals"], bb[2]["z_vals"])
    assert torch.equal(a[2]["depth_coarse"], bb[2]["depth_coarse"])


def test_fp16_mode_saturates_instead_of_overflowing(state_dict):
    """fp16 operands saturate at 65 504 (F2FP.SATFINITE): a network whose activations leave the fp16 range still gives
    finite outputs in the fp16 / mixed modes (and the bf16x3 mode, which has the range of fp32, is unaffected)."""
    import nerfw
    from config import Config
    sd, emb = state_dict
    sd = {k: v.clone() for k, v in sd.items()}
    sd["pts_linears.0.weight"] *= 3e4          # first-layer activations ~1e5
    sd["pts_linears.1.weight"] *= 1e-4         # keep the rest of the trunk in range
    m = nerfw.NeRF(Config())
    m.load_state_dict(sd)
    m = m.cuda()
    g = torch.Generator(device="cuda").manual_seed(6)
    x = (torch.rand(777, 3, device="cuda", generator=g) - 0.5) * 6
    d = torch.nn.functional.normalize(torch.randn(777, 3, device="cuda", generator=g), dim=-1)
    with torch.no_grad():
        for mode in ("fp16", "bf16x3"):
            m.mlp_mode = mode
            rgb, sigma = m(x, d, emb.cuda())
            assert bool(torch.isfinite(rgb).all()) and bool(torch.isfinite(sigma).all()), mode


def test_appearance_offsets_are_cached_per_embedding_and_weights(cuda_model):
    """The per-embedding rgb-logit offsets are computed once for launches that share weights and embedding rows
    (NERFW_MLP_APP_CACHED): a second call launches one kernel fewer and returns the same bits; changing the embedding's
    values (in place: version bump; or another tensor) or the weights recomputes them."""
    from nerfw import ops
    model, emb = cuda_model
    x = torch.rand(500, 3, device="cuda") * 4 - 2
    d = torch.nn.functional.normalize(torch.randn(500, 3, device="cuda"), dim=-1)
    e = emb.clone()
    with torch.no_grad():
        model.invalidate_packed()
        n0 = ops.launch_count()
        a = model.run_mlp(x, d, None, e.unsqueeze(0), "bf16x3")
        n1 = ops.launch_count()
        b = model.run_mlp(x, d, None, e.unsqueeze(0), "bf16x3")
        n2 = ops.launch_count()
        assert torch.equal(a, b) and (n2 - n1) == (n1 - n0) - 3      # no re-pack (2 kernels), no offset kernel
        e.add_(0.5)                                                    # same storage, new values
        c = model.run_mlp(x, d, None, e.unsqueeze(0), "bf16x3")
        assert float((c[:, :3] - a[:, :3]).abs().max()) > 1e-3 and torch.equal(c[:, 3], a[:, 3])
        fresh = model.run_mlp(x, d, None, e.clone().unsqueeze(0), "bf16x3")
        assert torch.equal(fresh, c)
        saved = model.appearance_projection.bias.clone()
        model.appearance_projection.bias.add_(0.25)                   # weights change -> new packed image -> offsets recomputed
        w = model.run_mlp(x, d, None, e.unsqueeze(0), "bf16x3")
        model.appearance_projection.bias.copy_(saved)
        back = model.run_mlp(x, d, None, e.unsqueeze(0), "bf16x3")
    assert float((w[:, :3] - c[:, :3]).abs().max()) > 1e-4 and torch.equal(back, c)
