"""Multi-GPU check under torchrun (NCCL): a ray batch rendered sharded over the ranks and gathered equals the
single-GPU render bit for bit, and one data-parallel training step leaves identical weights on every rank.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_sharded.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import nerfw  # noqa: E402
import nerfw_oracle as orc  # noqa: E402
from config import Config  # noqa: E402
from nerfw.parallel import render_sharded  # noqa: E402
from nerfw.train import Trainer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sd = orc.make_state_dict(0)
emb = torch.randn(32).cuda()
m = nerfw.NeRF(Config())
m.load_state_dict(sd)
m = m.cuda()
h, w, focal, c2w = orc.golden_camera()
o, d = nerfw.get_rays(h, w, focal, c2w.cuda())
o, d = o.reshape(-1, 3).contiguous(), d.reshape(-1, 3).contiguous()
torch.manual_seed(5)
u_all = torch.rand(10000, 128, device="cuda")


def render(oo, dd, u):
    with torch.no_grad():
        return nerfw.volume_render(m, oo, dd, 2.0, 6.0, 64, 128, appearance_embedding=emb, perturb=False, u_rand=u)


from nerfw.parallel import shard_bounds  # noqa: E402
s, e = shard_bounds(10000, rank, world)
got = render_sharded(lambda oo, dd: render(oo, dd, u_all[s:e]), o, d)
want = render(o, d, u_all)
ok_render = torch.equal(got[0], want[0]) and torch.equal(got[1], want[1]) and torch.equal(got[2], want[2]["acc"])

table = torch.nn.Parameter(torch.randn(4, 32, device="cuda"))
dist.broadcast(table.data, 0)
tr = Trainer(m, table, lr=5e-4, mlp_dtype="fp32")
tgt = torch.rand(10000, 3, device="cuda")
dist.broadcast(tgt, 0)
g = torch.Generator(device="cuda").manual_seed(100 + rank)
loss = tr.step(o[:4096], d[:4096], tgt[:4096], 1, 2.0, 6.0, 64, 0, perturb=True, shard=True, generator=g)
flat = tr.flat.param.clone()
ref = flat.clone()
dist.broadcast(ref, 0)
ok_train = bool(torch.equal(flat, ref))
res = torch.tensor([int(ok_render), int(ok_train)], device="cuda")
dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world={world} sharded_render_bitwise_equal={bool(res[0])} dp_weights_identical_across_ranks={bool(res[1])} loss={float(loss):.6f}")
dist.destroy_process_group()
sys.exit(0 if bool(res.min()) else 1)
