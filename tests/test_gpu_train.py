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
    assert abs(float(loss) - float(want.detach())) <= 1e-6 and maxabs(d, ar.grad) <= 1e-9


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


def _twin(sd, mode, table_seed=4):
    import nerfw
    from config import Config
    from nerfw.train import Trainer
    m = nerfw.NeRF(Config())
    m.load_state_dict(sd)
    m = m.cuda()
    torch.manual_seed(table_seed)
    table = torch.nn.Parameter(torch.randn(5, 32, device="cuda"))
    return m, table, Trainer(m, table, lr=5e-4, mlp_dtype=mode)


def test_trainer_default_mode_tracks_fp32(state_dict, oracle):
    """The DEFAULT training arithmetic ("mixed": bf16x3 coarse + fp16 fine forward, tcgen05 bf16 backward) against the
    fp32 CUDA-core path on the same batches and jitter: per-step losses within 2e-4 relative, first-step gradients within
    the stated bf16 bound (3e-2 of each tensor's max, cosine >= 0.999), parameters after 8 Adam steps within 8 lr."""
    import nerfw
    sd, _ = state_dict
    h, w, focal, c2w = oracle.golden_camera()
    o, d = nerfw.get_rays(h, w, focal, c2w.cuda())
    sel = torch.arange(0, 10000, 19, device="cuda")[:512]
    o, d = o.reshape(-1, 3)[sel].contiguous(), d.reshape(-1, 3)[sel].contiguous()
    tgt = torch.rand(512, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    ma, ta, tra = _twin(sd, None)        # default
    mb, tb, trb = _twin(sd, "fp32")
    worst_loss = 0.0
    for step in range(8):
        torch.manual_seed(200 + step)
        la = float(tra.step(o, d, tgt, 3, 2.0, 6.0, 64, 128, perturb=True))
        torch.manual_seed(200 + step)
        lb = float(trb.step(o, d, tgt, 3, 2.0, 6.0, 64, 128, perturb=True))
        worst_loss = max(worst_loss, abs(la - lb) / lb)
        if step == 0:
            worst, worst_cos = 0.0, 1.0
            for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
                ga, gb = pa.grad.double(), pb.grad.double()
                rel = float((ga - gb).abs().max() / (gb.abs().max() + 1e-20))
                cos = float((ga * gb).sum() / (ga.norm() * gb.norm() + 1e-30))
                worst, worst_cos = max(worst, rel), min(worst_cos, cos)
                assert rel <= 3e-2 and cos >= 0.999, (k, rel, cos)
    diffs = torch.cat([(pa - pb).abs().reshape(-1) for (_, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters())])
    record("trainer_default_vs_fp32", worst_loss_rel=worst_loss, grad_rel=worst, grad_cos=worst_cos,
           param_abs=float(diffs.max()), param_mean_abs=float(diffs.mean()))
    assert worst_loss <= 2e-4, worst_loss
    assert float(diffs.max()) <= 8 * 2 * 5e-4     # Adam moves every element by <= ~lr per step
    assert float((ta - tb).abs().max()) <= 8 * 2 * 5e-4 and float(ta[0].sub(tb[0]).abs().max()) == 0.0


def test_checkpoint_resume_keeps_training(state_dict, oracle, tmp_path):
    """save -> load into a live Trainer: parameters are copied IN PLACE (the flat-buffer views stay attached, so training
    continues to move the tensors the forward reads), Adam moments and step count come back, and the resumed run
    reproduces the uninterrupted one bit for bit (fp32 mode, same jitter)."""
    import nerfw
    from nerfw.checkpoint import load_checkpoint, save_checkpoint
    sd, _ = state_dict
    h, w, focal, c2w = oracle.golden_camera()
    o, d = nerfw.get_rays(h, w, focal, c2w.cuda())
    sel = torch.arange(0, 10000, 79, device="cuda")[:128]
    o, d = o.reshape(-1, 3)[sel].contiguous(), d.reshape(-1, 3)[sel].contiguous()
    tgt = torch.full((128, 3), 0.7, device="cuda")

    def run(tr, steps, first):
        for s in range(first, first + steps):
            torch.manual_seed(300 + s)
            tr.step(o, d, tgt, 1, 2.0, 6.0, 64, 0, perturb=True)

    ma, ta, tra = _twin(sd, "fp32")
    run(tra, 3, 0)
    path = save_checkpoint(str(tmp_path), 3, ma, ta, optimizer_state=tra.state_dict(), loss=0.1, psnr=10.0)
    run(tra, 2, 3)                                   # the uninterrupted run: 5 steps
    mb, tb, trb = _twin(sd, "fp32", table_seed=9)    # a fresh process would start from other values
    ptr_before = mb.rgb_linear.weight.data_ptr()
    ck = load_checkpoint(path, mb, tb, trainer=trb)
    assert ck["iteration"] == 3 and trb.step_count == 3
    assert mb.rgb_linear.weight.data_ptr() == ptr_before                       # still a view of the flat buffer
    assert tb.data_ptr() == trb.flat.param.data_ptr() + trb.flat.offsets[-2] * 4
    run(trb, 2, 3)
    # equal up to the atomics order of the backward carried through two Adam updates: Adam normalises every element to a
    # step of ~lr, so an element whose tiny gradient changes sign with the summation order may move the other way
    diffs = torch.cat([(pa - pb).abs().reshape(-1) for (_, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters())]
                      + [(ta - tb).abs().reshape(-1)])
    assert float(diffs.max()) <= 4 * 5e-4 and float((diffs <= 2e-6).float().mean()) >= 0.999, (float(diffs.max()), float((diffs <= 2e-6).float().mean()))
    # ... whereas a run that did NOT get the Adam moments back drifts visibly: the restored state matters
    assert trb.step_count == 5 and float(trb.exp_avg.abs().sum()) > 0
    with pytest.raises(RuntimeError, match="expected"):
        load_checkpoint(path, mb, torch.nn.Parameter(torch.zeros(4, 32, device="cuda")))


def test_unmerge_is_the_transpose_of_merge():
    """nerfw_unmerge_raw scatters a merged row back to its two source lists with the slots nerfw_merge_raw used: the round
    trip is the identity (ties between the lists and an unsorted fine list included), and the accumulate form adds."""
    from nerfw import ops
    g = torch.Generator(device="cuda").manual_seed(8)
    for b, n, ni in ((257, 64, 128), (33, 37, 53), (5, 512, 3584)):
        zc = torch.sort(torch.rand(b, n, device="cuda", generator=g) * 4 + 2, dim=-1).values
        zf = torch.sort(torch.rand(b, ni, device="cuda", generator=g) * 4 + 2, dim=-1).values
        zf[: b // 3, : min(ni, n) // 2] = zc[: b // 3, : min(ni, n) // 2]          # exact ties between the two lists
        zf[b // 3: b // 2] = zf[b // 3: b // 2].flip(-1)                            # unsorted fine rows: counting fallback
        rc = torch.randn(b * n, 4, device="cuda", generator=g)
        rf = torch.randn(b * ni, 4, device="cuda", generator=g)
        merged = ops.merge_raw(zc, rc, zf, rf)
        back_c, back_f = ops.unmerge_raw(zc, zf, merged)
        assert torch.equal(back_c, rc) and torch.equal(back_f, rf), (b, n, ni)
        base = torch.randn(b * n, 4, device="cuda", generator=g)
        acc_c, _ = ops.unmerge_raw(zc, zf, merged, base.clone())
        assert torch.equal(acc_c, base + rc)
        # the merged depths are sorted wherever both inputs were
        zm = ops.merge_raw(zc, zc.reshape(-1, 1).expand(-1, 4).contiguous(), zf, zf.reshape(-1, 1).expand(-1, 4).contiguous())[:, 0].reshape(b, n + ni)
        ok_rows = torch.ones(b, dtype=torch.bool, device="cuda")
        ok_rows[b // 3: b // 2] = False
        assert bool((zm[ok_rows][:, 1:] >= zm[ok_rows][:, :-1]).all())


@pytest.mark.parametrize("mode", ["fp32", None])
def test_reuse_coarse_training_gradients_equal_two_pass(cuda_model, oracle, mode):
    """Training with one network: the default (every depth through the MLP once, ReuseRenderFn) against the two-pass form
    (coarse depths evaluated in both passes) on a loss that uses the fine AND the coarse outputs.  fp32 mode: the forward
    is bit-identical and the gradients agree to summation order (1e-4 of each tensor's max); default tensor-core mode:
    inside the stated bf16 bound (3e-2, cosine >= 0.999)."""
    import nerfw
    model, emb = cuda_model
    h, w, focal, c2w = oracle.golden_camera()
    o, d = nerfw.get_rays(h, w, focal, c2w.cuda())
    sel = torch.arange(0, 10000, 23, device="cuda")[:400]
    o, d = o.reshape(-1, 3)[sel].contiguous(), d.reshape(-1, 3)[sel].contiguous()
    gen = torch.Generator().manual_seed(12)
    t_rand, u_rand = torch.rand(400, 64, generator=gen), torch.rand(400, 128, generator=gen)
    tgt = torch.rand(400, 3, generator=gen).cuda()
    grads, outs = [], []
    for reuse in (True, False):
        model.zero_grad()
        e = emb.clone().requires_grad_(True)
        rgb, depth, ex = nerfw.volume_render(model, o, d, 2.0, 6.0, 64, 128, appearance_embedding=e, perturb=True, mlp_dtype=mode,
                                             t_rand=t_rand, u_rand=u_rand, reuse_coarse=reuse)
        loss = ((rgb - tgt) ** 2).mean() + ((ex["rgb_coarse"] - tgt) ** 2).mean() + 0.1 * depth.mean() + 0.05 * ex["weights"].sum(1).mean()
        loss.backward()
        grads.append({k: p.grad.clone() for k, p in model.named_parameters()} | {"emb": e.grad.clone()})
        outs.append((rgb.detach(), depth.detach(), ex["z_vals"], float(loss.detach())))
    if mode == "fp32":
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[1][2])
    tol, tol_cos = (1e-4, 0.999999) if mode == "fp32" else (3e-2, 0.999)
    worst, worst_cos = 0.0, 1.0
    for k in grads[0]:
        a, b_ = grads[0][k].double(), grads[1][k].double()
        rel = float((a - b_).abs().max() / (b_.abs().max() + 1e-20))
        cos = float((a * b_).sum() / (a.norm() * b_.norm() + 1e-30))
        worst, worst_cos = max(worst, rel), min(worst_cos, cos)
        assert rel <= tol and cos >= tol_cos, (k, rel, cos)
    record(f"reuse_training_vs_two_pass_{mode or 'default'}", worst_rel=worst, worst_cos=worst_cos, loss_reuse=outs[0][3], loss_two_pass=outs[1][3])
    model.zero_grad()


def test_cross_image_batches_per_ray_embeddings(state_dict, oracle):
    """RayBank.sample(cross_image=True) draws (image, pixel) pairs over the whole bank and the step gathers
    one embedding row per ray; the default tensor-core path (per-ray rows in the tcgen05 backward) tracks the fp32 path."""
    import nerfw
    from nerfw.camera import aligned_spiral_poses, blender_focal
    from nerfw.raybank import RayBank
    sd, _ = state_dict
    poses = torch.from_numpy(aligned_spiral_poses(6, 1))
    imgs = torch.rand(6, 24, 32, 3, generator=torch.Generator().manual_seed(2))
    bank = RayBank(imgs, poses, blender_focal(32))
    batch = bank.sample(512, generator=torch.Generator(device="cuda").manual_seed(5), cross_image=True)
    idx = batch["appearance_idx"]
    assert idx.shape == (512,) and idx.dtype == torch.int64 and len(torch.unique(idx)) > 1
    flat = idx * (24 * 32)
    assert torch.equal(batch["rays_o"], bank.origins[idx])
    pix = torch.stack([(bank.dirs[i] == batch["rays_d"][k]).all(-1).nonzero()[0, 0] for k, i in enumerate(idx.tolist())])
    assert torch.equal(batch["rgb"], bank.rgb[idx, pix])
    ma, ta, tra = _twin(sd, None)
    mb, tb, trb = _twin(sd, "fp32")
    for step in range(2):
        torch.manual_seed(400 + step)
        la = float(tra.step(batch["rays_o"], batch["rays_d"], batch["rgb"], idx % 5, 2.0, 6.0, 64, 128))
        torch.manual_seed(400 + step)
        lb = float(trb.step(batch["rays_o"], batch["rays_d"], batch["rgb"], idx % 5, 2.0, 6.0, 64, 128))
        assert abs(la - lb) / lb <= 2e-4, (la, lb)
        if step == 0:
            ga, gb = ta.grad.double(), tb.grad.double()
            rel = float((ga - gb).abs().max() / gb.abs().max())
            cos = float((ga * gb).sum() / (ga.norm() * gb.norm()))
            record("per_ray_embedding_table_grad", rel=rel, cos=cos)
            assert rel <= 3e-2 and cos >= 0.999, (rel, cos)
            assert float(gb.abs().sum(1).min()) > 0     # every one of the 5 rows received gradient
