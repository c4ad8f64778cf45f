"""Training step (src/train.py:54-92) on the GPU: fused Adam equals torch.optim.Adam, the Trainer's step equals a
plain torch loop over the same kernels, and the loss goes down."""
import pytest
import torch

from gpu_util import maxabs, record

pytestmark = pytest.mark.gpu


def test_fused_adam_matches_torch():
    from nerfw import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    p = torch.randn(100003, device="cuda", generator=g)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=5e-4)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 6):
        grad = torch.randn(p.shape, device="cuda", generator=g) * 0.1
        ref.grad = grad.clone()
        opt.step()
        ops.adam_step(p, grad, m, v, step, 5e-4)
        assert maxabs(p, ref) <= 1e-6, step   # fp32 rounding of the bias-corrected step size
    # grad_scale = 1/world folds the data-parallel averaging into the update
    p2, m2, v2 = p.clone(), m.clone(), v.clone()
    grad = torch.randn(p.shape, device="cuda", generator=g)
    ops.adam_step(p, grad, m, v, 6, 5e-4)
    ops.adam_step(p2, grad * 4, m2, v2, 6, 5e-4, grad_scale=0.25)
    assert maxabs(p, p2) <= 1e-7


def test_mse_kernel():
    from nerfw import ops
    a = torch.rand(4096, 3, device="cuda")
    b = torch.rand(4096, 3, device="cuda")
    loss, d = ops.mse(a, b)
    ar = a.clone().requires_grad_(True)
    want = torch.nn.functional.mse_loss(ar, b)
    want.backward()
    assert abs(float(loss) - float(want)) <= 1e-6 and maxabs(d, ar.grad) <= 1e-9


def test_trainer_step_matches_torch_loop(state_dict, oracle):
    """Two steps of Trainer.step == volume_render + mse_loss + backward + torch.optim.Adam on a twin model."""
    import nerfw
    from config import Config
    from nerfw.train import Trainer
    sd, _ = state_dict
    h, w, focal, c2w = oracle.golden_camera()
    o, d = nerfw.get_rays(h, w, focal, c2w.cuda())
    sel = torch.arange(0, 10000, 79, device="cuda")[:96]
    o, d = o.reshape(-1, 3)[sel].contiguous(), d.reshape(-1, 3)[sel].contiguous()
    tgt = torch.rand(96, 3, device="cuda")
    models, tables = [], []
    for _ in range(2):
        m = nerfw.NeRF(Config())
        m.load_state_dict(sd)
        models.append(m.cuda())
        torch.manual_seed(4)
        tables.append(torch.nn.Parameter(torch.randn(5, 32, device="cuda")))
    tr = Trainer(models[0], tables[0], lr=5e-4, mlp_dtype="fp32", coarse_loss=False)
    opt = torch.optim.Adam(list(models[1].parameters()) + [tables[1]], lr=5e-4)
    losses = []
    for step in range(2):
        torch.manual_seed(100 + step)   # same device RNG state -> same stratified jitter for both paths
        la = tr.step(o, d, tgt, 2, 2.0, 6.0, 64, 0, perturb=True)
        torch.manual_seed(100 + step)
        opt.zero_grad()
        rgb, _, _ = nerfw.volume_render(models[1], o, d, 2.0, 6.0, 64, 0, appearance_embedding=tables[1][2],
                                        perturb=True, mlp_dtype="fp32")
        lb = torch.nn.functional.mse_loss(rgb, tgt)
        lb.backward()
        opt.step()
        losses.append((float(la), float(lb)))
        assert abs(float(la) - float(lb)) <= 1e-6
    # gradients of the last step agree up to fp32 summation-order noise (both backward passes accumulate with atomics)
    # carried through one Adam update
    for (k, pa), (_, pb) in zip(models[0].named_parameters(), models[1].named_parameters()):
        rel = maxabs(pa.grad, pb.grad) / (float(pb.grad.abs().max()) + 1e-20)
        assert rel <= 2e-3, (k, rel)
    # parameters: Adam normalises every element to a step of ~lr, so an element whose tiny gradient changes sign with the
    # summation order may move the other way: bound the worst case by 2 steps x 2 lr and require 99.9 % within 2e-6
    diffs = torch.cat([(pa - pb).abs().reshape(-1) for (_, pa), (_, pb) in
                       zip(models[0].named_parameters(), models[1].named_parameters())] + [(tables[0] - tables[1]).abs().reshape(-1)])
    worst = float(diffs.max())
    frac_close = float((diffs <= 2e-6).float().mean())
    record("trainer_vs_torch_loop", worst_param_abs=worst, frac_within_2e6=frac_close)
    assert worst <= 4 * 5e-4 and frac_close >= 0.999, (worst, frac_close)
    assert float(tables[0][0].sub(tables[1][0]).abs().max()) == 0.0   # untouched rows stay equal


@pytest.mark.parametrize("mode", ["fp32", "bf16x3"])
def test_loss_decreases_coarse_and_fine(state_dict, oracle, mode):
    import nerfw
    from config import Config
    from nerfw.train import Trainer
    sd, _ = state_dict
    m = nerfw.NeRF(Config())
    m.load_state_dict(sd)
    m = m.cuda()
    table = torch.nn.Parameter(torch.randn(3, 32))      # CPU-resident like dataset.appearance_embeddings (src/dataset.py:81)
    tr = Trainer(m, table, lr=5e-4, mlp_dtype=mode)
    h, w, focal, c2w = oracle.golden_camera()
    o, d = nerfw.get_rays(h, w, focal, c2w.cuda())
    sel = torch.arange(0, 10000, 39, device="cuda")[:256]
    o, d = o.reshape(-1, 3)[sel].contiguous(), d.reshape(-1, 3)[sel].contiguous()
    tgt = torch.full((256, 3), 0.8, device="cuda")
    losses = [float(tr.step(o, d, tgt, 1, 2.0, 6.0, 64, 128)) for _ in range(12)]
    record(f"train_loss_{mode}", first=losses[0], last=losses[-1])
    assert losses[-1] < losses[0] * 0.9, losses
    assert all(torch.isfinite(p).all() for p in m.parameters())
