"""Host logic of the multi-GPU path on CPU: partitioning, the gather of per-rank results and the flat gradient
all-reduce, with world_size 2 over gloo (the GPU box runs the same code over NCCL)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_bounds_cover_everything():
    from nerfw.parallel import frames_for_rank, shard_bounds
    for n in (0, 1, 7, 640000, 640001):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in cuts]
            assert max(sizes) - min(sizes) <= 1
    assert frames_for_rank(120, 3, 8) == list(range(3, 120, 8))
    assert sorted(sum((frames_for_rank(120, r, 8) for r in range(8)), [])) == list(range(120))
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_rays, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nerfw.parallel import FlatParams, render_sharded

    torch.manual_seed(0)
    o = torch.randn(n_rays, 3)
    d = torch.randn(n_rays, 3)

    def fake_render(oo, dd):   # any per-ray function: sharded + gathered must equal the unsharded call
        rgb = torch.sin(oo) * torch.cos(dd)
        depth = (oo * dd).sum(-1, keepdim=True)
        return rgb, depth, {"acc": (oo - dd).norm(dim=-1, keepdim=True)}

    want = fake_render(o, d)
    got = render_sharded(fake_render, o, d)
    ok = torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]) and torch.equal(got[2], want[2]["acc"])
    only0 = render_sharded(fake_render, o, d, dst=0)
    ok = ok and ((only0 is None) == (rank != 0))

    # flat gradient all-reduce == gradient of the concatenated batch
    lin = torch.nn.Linear(5, 3)
    emb = torch.nn.Parameter(torch.zeros(4, 2))
    with torch.no_grad():
        for p in lin.parameters():
            p.copy_(torch.arange(p.numel(), dtype=torch.float32).reshape(p.shape) / 10)
    flat = FlatParams(list(lin.parameters()) + [emb])
    x = torch.arange(20, dtype=torch.float32).reshape(4, 5) / 7
    flat.zero_grad()
    xs = x[rank * 2:(rank + 1) * 2]
    (lin(xs).pow(2).sum() + emb[rank].sum()).backward()
    assert lin.weight.grad.data_ptr() == flat.grad.data_ptr()          # grads accumulate straight into the flat buffer
    world_n = flat.all_reduce()
    ref = torch.nn.Linear(5, 3)
    ref.load_state_dict(lin.state_dict())
    ref(x).pow(2).sum().backward()
    ok = ok and world_n == world and torch.allclose(lin.weight.grad, ref.weight.grad, atol=1e-5) \
        and torch.allclose(lin.bias.grad, ref.bias.grad, atol=1e-5)
    ok = ok and torch.equal(emb.grad[:2], torch.ones(2, 2)) and float(emb.grad[2:].abs().sum()) == 0.0
    out[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rays", [10, 7])
def test_sharded_render_and_flat_allreduce_gloo(n_rays):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_rays, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(out) == {0: True, 1: True}
