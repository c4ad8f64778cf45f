"""Randomised shapes (hypothesis) for the HBM-bound kernels against the oracle: ragged ray counts, odd sample counts."""
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from gpu_util import maxabs

pytestmark = pytest.mark.gpu
COMMON = dict(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])


@settings(**COMMON)
@given(b=st.integers(1, 300), n=st.integers(2, 300), seed=st.integers(0, 2**16), scale=st.sampled_from([0.5, 5.0, 80.0]))
def test_composite_random_shapes(oracle, b, n, seed, scale):
    from nerfw import ops
    g = torch.Generator().manual_seed(seed)
    sigma = torch.rand(b, n, 1, generator=g) * scale
    rgb = torch.rand(b, n, 3, generator=g)
    z = torch.sort(torch.rand(b, n, generator=g) * 4 + 2, dim=-1).values
    want_rgb, want_depth, want_w = oracle.composite(sigma, rgb, z)
    raw = torch.cat([rgb, sigma], dim=-1).reshape(b * n, 4).cuda()
    got_rgb, got_depth, got_acc, got_w = ops.composite_fwd(raw, z.cuda())
    assert maxabs(got_rgb, want_rgb) <= 3e-5 and maxabs(got_w, want_w[..., 0]) <= 3e-6
    assert maxabs(got_acc, want_w.sum(1)) <= 3e-5 and maxabs(got_depth, want_depth) <= 1e-4


@settings(**COMMON)
@given(b=st.integers(1, 200), n8=st.integers(1, 24), ni=st.integers(1, 300), seed=st.integers(0, 2**16),
       power=st.sampled_from([1, 4, 16]))
def test_sample_pdf_random_shapes_bit_exact(oracle, b, n8, ni, seed, power):
    """N a multiple of 8 (where ATen's summation order is reproduced): indices, cdf and merged depths bit-exact."""
    from nerfw import ops
    n = 8 * n8
    g = torch.Generator().manual_seed(seed)
    z = torch.sort(torch.rand(b, n, generator=g) * 4 + 2, dim=-1).values
    w = torch.rand(b, n, generator=g) ** power
    u = torch.rand(b, ni, generator=g)
    o = torch.zeros(b, 3)
    d = torch.ones(b, 3)
    zw, _, aux = oracle.resample_pdf(o, d, z, w, ni, u_rand=u, return_aux=True)
    got, gaux = ops.sample_pdf(z.cuda(), w.cuda(), ni, u.cuda(), want_aux=True)
    assert torch.equal(gaux["cdf"].cpu(), aux["cdf"])
    assert torch.equal(gaux["inds"].cpu(), aux["inds"])
    assert torch.equal(got.cpu(), zw)


@settings(**COMMON)
@given(b=st.integers(1, 64), n=st.integers(1, 200), seed=st.integers(0, 2**16), perturb=st.booleans())
def test_stratified_random_shapes_bit_exact(oracle, b, n, seed, perturb):
    import nerfw
    g = torch.Generator().manual_seed(seed)
    ro = torch.randn(b, 3, generator=g)
    rd = torch.nn.functional.normalize(torch.randn(b, 3, generator=g), dim=-1)
    tr = torch.rand(b, n, generator=g)
    near, far = 0.25, 7.5
    zw, pw = oracle.stratified_depths(ro, rd, near, far, n, perturb=perturb, t_rand=tr)
    zg, pg = nerfw.sample_stratified(ro.cuda(), rd.cuda(), near, far, n, perturb=perturb, t_rand=tr)
    assert torch.equal(zg.cpu(), zw.expand(b, n)) and torch.equal(pg.cpu(), pw)
