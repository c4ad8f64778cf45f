import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "depth-aware-shader-effects-for-nerf_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, nerfw, nerfw_oracle as orc
from config import Config
from nerfw import ops
sd = orc.make_state_dict(0); emb = torch.randn(32)
m = nerfw.NeRF(Config()); m.load_state_dict(sd); m = m.cuda()
for (b, n, with_emb) in [(700, 64, True), (3, 64, False)]:
    gen = torch.Generator(device="cuda").manual_seed(b * 13 + n)
    o = torch.randn(b, 3, device="cuda", generator=gen)
    d = torch.nn.functional.normalize(torch.randn(b, 3, device="cuda", generator=gen), dim=-1)
    z = torch.sort(torch.rand(b, n, device="cuda", generator=gen) * 4 + 2, dim=-1).values
    d_raw = torch.randn(b * n, 4, device="cuda", generator=gen)
    e = emb.cuda().unsqueeze(0).contiguous() if with_emb else None
    names, tensors = m.kernel_params()
    params = {k: t.detach() for k, t in zip(names, tensors)}
    packed = m.packed_weights(names, tensors)
    g_ref = {k: torch.zeros_like(t) for k, t in params.items()}
    g_tc = {k: torch.zeros_like(t) for k, t in params.items()}
    de_ref = torch.zeros(1, 32, device="cuda") if with_emb else None
    de_tc = torch.zeros(1, 32, device="cuda") if with_emb else None
    ops.mlp_bwd(params, g_ref, o, d, z, e, d_raw, de_ref)
    _, masks = ops.mlp_fwd(params, packed, o, d, z, e, 1, want_masks=True)   # gates of the bf16x3 forward
    ops.mlp_bwd_tc(params, g_tc, packed, o, d, z, e, d_raw, de_tc, masks)
    torch.cuda.synchronize()
    print(f"--- b={b} n={n} emb={with_emb}")
    for k in params:
        ref, got = g_ref[k].double(), g_tc[k].double()
        rel = float((ref - got).abs().max() / (ref.abs().max() + 1e-20))
        cos = float((ref * got).sum() / (ref.norm() * got.norm() + 1e-30))
        ratio = float(got.norm() / (ref.norm() + 1e-30))
        print(f"{k:34s} rel={rel:.4f} cos={cos:.6f} norm_ratio={ratio:.4f}")
        for key, parts in (("pts_linears.4.weight", (("h", slice(0, 256)), ("enc", slice(256, 319)))),
                           ("dir_linear.weight", (("h", slice(0, 256)), ("enc", slice(256, 283))))):
            if k == key:
                for nm, sl in parts:
                    r, g = ref[:, sl], got[:, sl]
                    print(f"    part {nm}: rel={float((r-g).abs().max()/(r.abs().max()+1e-20)):.4f} cos={float((r*g).sum()/(r.norm()*g.norm()+1e-30)):.6f}")
    if with_emb:
        print("d_emb rel", float((de_ref - de_tc).abs().max() / (de_ref.abs().max() + 1e-20)))
