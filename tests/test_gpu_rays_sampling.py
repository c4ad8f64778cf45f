"""K1/K2/K3/K8 parity on the GPU: rays, stratified depths and resampling indices are BIT-exact against the reference's
CPU PyTorch results (golden vectors) and the oracle on fresh seeded inputs."""
import hashlib

import numpy as np
import pytest
import torch

from gpu_util import maxabs, record

pytestmark = pytest.mark.gpu


def sha(t):
    return hashlib.sha256(t.contiguous().cpu().numpy().tobytes()).hexdigest()


def test_get_rays_bit_exact_golden(golden, manifest):
    import nerfw
    g = golden("rays_100")
    c2w = torch.from_numpy(g["c2w"])
    o, d = nerfw.get_rays(100, 100, float(g["focal"]), c2w)          # CPU c2w -> CPU outputs like the reference
    assert not d.is_cuda and d.shape == (100, 100, 3)
    assert torch.equal(d, torch.from_numpy(g["rays_d"]))
    assert torch.equal(o[3, 7], torch.from_numpy(g["origin"])) and o.stride()[:2] == (0, 0)
    o, d = nerfw.get_rays(100, 100, float(g["focal"]), c2w.cuda())   # CUDA c2w works too (the reference raises, F5)
    assert d.is_cuda and o.is_cuda and torch.equal(d.cpu(), torch.from_numpy(g["rays_d"]))
    for name, case in manifest["cases"]["rays_rot"].items():        # includes a full 800x800 view
        _, d = nerfw.get_rays(case["h"], case["w"], case["focal"], torch.tensor(case["c2w"]).cuda())
        assert sha(d) == case["sha256_d"], name


def test_get_rays_vs_oracle_random_views(oracle):
    import nerfw
    g = torch.Generator().manual_seed(42)
    for h, w in [(1, 1), (3, 5), (31, 17), (240, 320)]:
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
        m = torch.eye(4)
        m[:3, :3] = q
        m[:3, 3] = torch.randn(3, generator=g) * 3
        f = 0.5 * w / np.tan(0.5 * (0.3 + float(torch.rand(1, generator=g))))
        _, want = oracle.rays_for_view(h, w, f, m)
        _, got = nerfw.get_rays(h, w, f, m[:3])   # (3,4) matrices are accepted too
        assert torch.equal(got, want), (h, w)


def test_normalize_bit_exact():
    from nerfw import ops
    g = torch.Generator().manual_seed(1)
    d = torch.randn(5000, 3, generator=g) * torch.logspace(-3, 3, 5000).unsqueeze(-1)
    d[7] = 0.0   # zero vector -> zeros (F.normalize eps clamp)
    want = torch.nn.functional.normalize(d, dim=-1)
    got = ops.normalize_dirs(d.cuda()).cpu()
    assert torch.equal(got, want)


def test_stratified_bit_exact(golden, oracle):
    import nerfw
    g = golden("stratified")
    o, d = torch.from_numpy(g["o"]).cuda(), torch.from_numpy(g["d"]).cuda()
    z, pts = nerfw.sample_stratified(o, d, 2.0, 6.0, 64, perturb=False)
    assert z.shape == (4, 64) and pts.shape == (4, 64, 3)
    assert torch.equal(z[0].cpu(), torch.from_numpy(g["z_det"]))
    z, pts = nerfw.sample_stratified(o, d, 2.0, 6.0, 64, perturb=True, t_rand=torch.from_numpy(g["t_rand"]))
    assert torch.equal(z.cpu(), torch.from_numpy(g["z_perturb"]))
    assert torch.equal(pts.cpu(), torch.from_numpy(g["pts_perturb"]))
    # other sample counts / leading shapes / near-far, against the oracle
    gen = torch.Generator().manual_seed(5)
    for n, lead in [(2, (3,)), (7, (2, 5)), (192, (33,)), (256, (1,))]:
        ro = torch.randn(*lead, 3, generator=gen)
        rd = torch.nn.functional.normalize(torch.randn(*lead, 3, generator=gen), dim=-1)
        tr = torch.rand(*lead, n, generator=gen)
        zw, pw = oracle.stratified_depths(ro, rd, 0.5, 9.25, n, perturb=True, t_rand=tr)
        zg, pg = nerfw.sample_stratified(ro.cuda(), rd.cuda(), 0.5, 9.25, n, perturb=True, t_rand=tr)
        assert torch.equal(zg.cpu(), zw) and torch.equal(pg.cpu(), pw), n
    # empty batch
    z, pts = nerfw.sample_stratified(torch.zeros(0, 3).cuda(), torch.zeros(0, 3).cuda(), 2.0, 6.0, 8, perturb=False)
    assert z.shape == (0, 8) and pts.shape == (0, 8, 3)


def test_posenc_matches_reference_formula(oracle, golden):
    import nerfw
    g = golden("mlp_64")
    x = torch.from_numpy(g["x"])
    pe = nerfw.PositionalEncoding(10)(x.cuda())
    assert pe.shape == (64, 63)
    err = maxabs(pe, g["pe"])
    record("posenc_L10", maxabs=err)
    assert err <= 1e-6   # sincosf vs ATen sin/cos: <= 2 ulp near |v| <= 1
    pe4 = nerfw.PositionalEncoding(4, include_input=False)(x.cuda().reshape(8, 8, 3))
    want = oracle.encode(x, 4)[:, 3:].reshape(8, 8, 24)
    assert pe4.shape == (8, 8, 24) and maxabs(pe4, want) <= 1e-6


def test_sample_pdf_golden_bit_exact(golden):
    from nerfw import ops
    g = golden("resample")
    for tag in ("safe", "gen"):
        z = torch.from_numpy(g[f"z_{tag}"]).cuda()
        w = torch.from_numpy(g[f"w_{tag}"]).cuda()
        u = torch.from_numpy(g[f"u_rand_{tag}"]).cuda()
        out, aux = ops.sample_pdf(z, w, 128, u, want_aux=True)
        inds_want = torch.from_numpy(g[f"inds_{tag}"])
        mism = float((aux["inds"].cpu() != inds_want).float().mean())
        record(f"sample_pdf_inds_{tag}", mismatch_rate=mism)
        assert mism == 0.0, f"sample indices differ from the reference ({tag})"
        assert torch.equal(aux["z_fine"].cpu(), torch.from_numpy(g[f"zfine_{tag}"]))
        assert torch.equal(out.cpu(), torch.from_numpy(g[f"out_{tag}"]))


@pytest.mark.parametrize("n,ni,b", [(64, 128, 777), (256, 512, 129), (64, 96, 50), (8, 8, 33), (1, 5, 4), (40, 100, 64)])
def test_sample_pdf_vs_oracle(oracle, n, ni, b):
    import nerfw
    from nerfw import ops
    gen = torch.Generator().manual_seed(n * 1000 + ni)
    z = torch.sort(torch.rand(b, n, generator=gen) * 4 + 2, dim=-1).values
    w = torch.rand(b, n, generator=gen) ** 4
    w[: b // 8] = 0.0                       # all-zero weights -> uniform pdf
    w[b // 8: b // 4, : n // 2] = 0.0       # empty leading half
    u = torch.rand(b, ni, generator=gen)
    o = torch.randn(b, 3, generator=gen)
    d = torch.nn.functional.normalize(torch.randn(b, 3, generator=gen), dim=-1)
    zw, pw, aux = oracle.resample_pdf(o, d, z, w, ni, u_rand=u, return_aux=True)
    got, gaux = ops.sample_pdf(z.cuda(), w.cuda(), ni, u.cuda(), want_aux=True)
    mism = float((gaux["inds"].cpu() != aux["inds"]).float().mean())
    record(f"sample_pdf_inds_{n}_{ni}", mismatch_rate=mism)
    if n % 8 == 0:
        assert mism == 0.0               # ATen's sum order is reproduced exactly for N % 8 == 0
        assert torch.equal(gaux["cdf"].cpu(), aux["cdf"])
        assert torch.equal(got.cpu(), zw)
    else:
        assert mism <= 1e-3
        assert maxabs(got, zw) <= 1e-4
    zz, pts = nerfw.sample_importance(o.cuda(), d.cuda(), z.cuda(), w.cuda(), ni, u_rand=u)
    assert zz.shape == (b, n + ni) and pts.shape == (b, n + ni, 3)
    assert bool((zz[:, 1:] >= zz[:, :-1]).all())
    if n % 8 == 0:
        assert torch.equal(pts.cpu(), pw)


def test_sample_pdf_merge_path_equals_general_path(oracle, monkeypatch):
    """The merge path (sorted u: no searches) and the general search + sort path return the same bits; rays whose u is
    not sorted (rand outside [0,1)), whose depths are unsorted or whose weights are negative fall back inside the kernel."""
    from nerfw import ops
    gen = torch.Generator(device="cuda").manual_seed(21)
    b, n, ni = 20000, 64, 128
    z = torch.sort(torch.rand(b, n, device="cuda", generator=gen) * 4 + 2, dim=-1).values
    w = torch.rand(b, n, device="cuda", generator=gen) ** 8
    w[:500] = 0.0
    w[500:1000, 5:] = 0.0                    # one occupied bin: most cdf entries equal
    u = torch.rand(b, ni, device="cuda", generator=gen)
    u[1000:1500] = u[1000:1500] * 3.0        # unsorted u -> in-kernel fallback
    u[1500:1600] = 0.0
    u[1600:1700] = 1.0 - 2.0 ** -24
    z[1700:1800] = z[1700:1800].flip(-1)     # unsorted depths -> fallback
    w[1800:1900, ::7] = -0.5                 # non-monotone cdf -> fallback
    got, aux = ops.sample_pdf(z, w, ni, u, want_aux=True)
    want, aux_w = ops.sample_pdf(z, w, ni, u, want_aux=True, general_path=True)   # nerfw_sample_pdf_general
    assert torch.equal(got, want)
    assert torch.equal(aux["inds"], aux_w["inds"]) and torch.equal(aux["z_fine"], aux_w["z_fine"])
    # and both equal the oracle where its searchsorted is well defined (sorted cdf)
    sel = torch.cat([torch.arange(0, 1700), torch.arange(1900, 4000)])
    o = torch.zeros(len(sel), 3)
    zw, _, oaux = oracle.resample_pdf(o, o, z[sel].cpu(), w[sel].cpu(), ni, u_rand=u[sel].cpu(), return_aux=True)
    assert torch.equal(aux["inds"][sel].cpu(), oaux["inds"])
    assert torch.equal(got[sel].cpu(), zw)


def test_sample_pdf_full_size_properties():
    """800x800 rays x (64 -> +128): sortedness, range and multiset-preservation of the coarse depths."""
    from nerfw import ops
    b, n, ni = 640000, 64, 128
    gen = torch.Generator(device="cuda").manual_seed(0)
    z = torch.sort(torch.rand(b, n, device="cuda", generator=gen) * 4 + 2, dim=-1).values
    w = torch.rand(b, n, device="cuda", generator=gen) ** 8
    u = torch.rand(b, ni, device="cuda", generator=gen)
    out = ops.sample_pdf(z, w, ni, u)
    assert bool((out[:, 1:] >= out[:, :-1]).all())
    assert float(out.min()) >= float(z.min()) - 1e-5 and float(out.max()) <= float(z.max()) + 1e-5
    # every coarse depth is still present: sum over the merged list minus the fine samples equals the coarse sum
    _, aux = ops.sample_pdf(z[:4096], w[:4096], ni, u[:4096], want_aux=True)
    merged = torch.sort(torch.cat([z[:4096], aux["z_fine"]], dim=-1), dim=-1).values
    assert torch.equal(merged, out[:4096])
