"""K6/K7 compositing forward/backward against the oracle (src/render.py:56-80) and its autograd."""
import pytest
import torch

from gpu_util import maxabs, record

pytestmark = pytest.mark.gpu


def make_inputs(b, n, seed, dense=False):
    g = torch.Generator().manual_seed(seed)
    sigma = torch.rand(b, n, 1, generator=g) * (60.0 if dense else 3.0)
    sigma[torch.rand(b, n, 1, generator=g) < 0.3] = 0.0
    rgb = torch.rand(b, n, 3, generator=g)
    z = torch.sort(torch.rand(b, n, generator=g) * 4 + 2, dim=-1).values
    return sigma, rgb, z


@pytest.mark.parametrize("b,n,dense", [(1, 2, False), (5, 2, False), (300, 64, False), (300, 64, True), (64, 192, True),
                                       (33, 100, False), (9, 768, False), (3, 1024, True)])
def test_composite_forward(oracle, b, n, dense):
    from nerfw import ops
    sigma, rgb, z = make_inputs(b, n, b * 7 + n, dense)
    want_rgb, want_depth, want_w = oracle.composite(sigma, rgb, z)
    raw = torch.cat([rgb, sigma], dim=-1).reshape(b * n, 4).cuda()
    got_rgb, got_depth, got_acc, got_w = ops.composite_fwd(raw, z.cuda())
    e = dict(rgb=maxabs(got_rgb, want_rgb), depth=maxabs(got_depth, want_depth), w=maxabs(got_w, want_w[..., 0]),
             acc=maxabs(got_acc, want_w.sum(1)))
    record(f"composite_fwd_{b}x{n}{'_dense' if dense else ''}", **e)
    assert e["rgb"] <= 2e-5 and e["w"] <= 2e-6 and e["acc"] <= 2e-5 and e["depth"] <= 5e-5, e


@pytest.mark.parametrize("b,n,dense", [(4, 2, False), (40, 64, False), (40, 64, True), (16, 192, False), (5, 768, False), (2, 1100, False)])
def test_composite_backward_vs_autograd(oracle, b, n, dense):
    from nerfw import ops
    sigma, rgb, z = make_inputs(b, n, 100 + b + n, dense)
    g = torch.Generator().manual_seed(9)
    d_rgb = torch.randn(b, 3, generator=g)
    d_depth = torch.randn(b, 1, generator=g)
    d_acc = torch.randn(b, 1, generator=g)
    d_w = torch.randn(b, n, generator=g)
    s64 = sigma.double().requires_grad_(True)
    c64 = rgb.double().requires_grad_(True)
    o_rgb, o_depth, o_w = oracle.composite(s64, c64, z.double())
    loss = (o_rgb * d_rgb.double()).sum() + (o_depth * d_depth.double()).sum() + (o_w.sum(1) * d_acc.double()).sum() \
        + (o_w[..., 0] * d_w.double()).sum()
    loss.backward()
    raw = torch.cat([rgb, sigma], dim=-1).reshape(b * n, 4).cuda()
    d_raw = ops.composite_bwd(raw, z.cuda(), d_rgb.cuda(), d_depth.cuda(), d_acc.cuda(), d_w.cuda()).reshape(b, n, 4).cpu()
    scale_s = float(s64.grad.abs().max()) + 1e-12
    scale_c = float(c64.grad.abs().max()) + 1e-12
    es = maxabs(d_raw[..., 3:], s64.grad) / scale_s
    ec = maxabs(d_raw[..., :3], c64.grad) / scale_c
    record(f"composite_bwd_{b}x{n}{'_dense' if dense else ''}", dsigma_rel=es, drgb_rel=ec)
    assert es <= 2e-4 and ec <= 1e-5, (es, ec)
    # optional cotangents may be absent
    d_raw2 = ops.composite_bwd(raw, z.cuda(), d_rgb.cuda(), None, None, None).reshape(b, n, 4).cpu()
    s2 = sigma.double().requires_grad_(True)
    o_rgb2, _, _ = oracle.composite(s2, rgb.double(), z.double())
    (o_rgb2 * d_rgb.double()).sum().backward()
    assert maxabs(d_raw2[..., 3:], s2.grad) / (float(s2.grad.abs().max()) + 1e-12) <= 2e-4


def test_single_sample_rays():
    """N = 1 (the reference itself breaks there: its delta tensor comes out empty): w = 1 - exp(-sigma * 1e-3)."""
    from nerfw import ops
    raw = torch.tensor([[0.2, 0.4, 0.6, 50.0], [1.0, 1.0, 1.0, 0.0]], device="cuda")
    z = torch.tensor([[3.0], [5.0]], device="cuda")
    rgb, depth, acc, w = ops.composite_fwd(raw, z)
    a = 1.0 - torch.exp(torch.tensor(-50.0 * 1e-3))
    assert abs(float(w[0, 0]) - float(a)) <= 1e-7 and float(w[1, 0]) == 0.0
    assert abs(float(depth[0, 0]) - 3.0) <= 1e-5 and float(depth[1, 0]) == 0.0


def test_composite_full_size_properties():
    """640k rays x 192 samples: weights in [0,1], acc <= 1, depth within [near, far], linear in rgb."""
    from nerfw import ops
    b, n = 640000, 192
    g = torch.Generator(device="cuda").manual_seed(1)
    raw = torch.rand(b * n, 4, device="cuda", generator=g)
    raw[:, 3] *= 8.0
    z = torch.sort(torch.rand(b, n, device="cuda", generator=g) * 4 + 2, dim=-1).values
    rgb, depth, acc, w = ops.composite_fwd(raw, z)
    assert float(w.min()) >= 0.0 and float(acc.max()) <= 1.0 + 1e-4
    assert float(depth.min()) >= 2.0 - 1e-3 and float(depth.max()) <= 6.0 + 1e-3
    raw2 = raw.clone()
    raw2[:, :3] *= 0.5
    rgb2, depth2, _, _ = ops.composite_fwd(raw2, z)
    assert maxabs(rgb2, rgb * 0.5) <= 1e-6 and torch.equal(depth2, depth)
