"""SURVEY.md section 8(d) "secondary baseline": the reference algorithm (the oracle's torch restatement) run in PyTorch
eager mode ON the B200 -- fp32 nn.Linear through cuBLAS, TF32 off -- next to the sm_100a path, same rays and weights.
Records both times (gpurun_out/parity_errors.jsonl) and checks the outputs against each other."""
import pytest
import torch

from gpu_util import maxabs, record

pytestmark = pytest.mark.gpu


def _time(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def test_torch_eager_on_gpu_vs_sm100a_path(cuda_model, oracle, state_dict):
    import nerfw
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model, emb = cuda_model
    sd, _ = state_dict
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    g = torch.Generator(device="cuda").manual_seed(3)
    b = 4096
    o = torch.tensor([0.0, 0.0, 4.0], device="cuda").expand(b, 3).contiguous()
    d = torch.nn.functional.normalize(torch.randn(b, 3, device="cuda", generator=g) * 0.25 + torch.tensor([0.0, 0.0, -1.0], device="cuda"), dim=-1)
    u = torch.rand(b, 128, device="cuda", generator=g)
    with torch.no_grad():
        ms_eager, ref = _time(lambda: oracle.render_hier(sd_gpu, sd_gpu, o, d, 2.0, 6.0, 64, 128, emb=emb, perturb=False, u_rand=u))
        ms_ours, got = _time(lambda: nerfw.volume_render(model, o, d, 2.0, 6.0, 64, 128, appearance_embedding=emb,
                                                         perturb=False, u_rand=u), reps=10)
    e = dict(rgb=maxabs(got[0], ref[0]), depth=maxabs(got[1], ref[1]), acc=maxabs(got[2]["acc"], ref[2]["acc"]))
    record("eager_gpu_baseline_4096x(64+192)", eager_ms=ms_eager, ours_ms=ms_ours, speedup=ms_eager / ms_ours, **e)
    print(f"torch eager on the GPU: {ms_eager:.2f} ms, sm_100a path: {ms_ours:.3f} ms ({ms_eager / ms_ours:.1f}x) for 4096 rays; "
          f"max abs diff rgb {e['rgb']:.1e} depth {e['depth']:.1e} acc {e['acc']:.1e}")
    assert e["rgb"] <= 1e-3 and e["depth"] <= 1e-3 and e["acc"] <= 1e-3, e
    assert ms_ours < ms_eager
