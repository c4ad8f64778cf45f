"""TEST INFRASTRUCTURE (CPU, oracle only): which operand precision does each pass of the hierarchical render need?

Emulates the operand rounding of the tensor-core modes inside the oracle's MLP (products in float64 of the rounded
operands, so only the operand format differs) and prints the max / mean abs error of rgb, depth and acc against the
fp32 oracle for every (coarse pass, fine pass) combination, on a 48 x 48 crop of the golden view, random init and two
dense variants.  Evidence for the `mixed` default: the error of the cheap modes comes from the
coarse pass (it places the fine samples); a single fp16 MMA per product in the fine pass matches the bf16x3 split.

    python oracle/emulate_mixed_precision.py            # ~10 min on 16 cores
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import nerfw_oracle as orc  # noqa: E402

F = torch.nn.functional


def rnd(t, dt):
    return t.to(dt).to(torch.float32)


def split(t, dt):
    hi = rnd(t, dt)
    return hi, rnd(t - hi, dt)


def make_linear(scheme):
    def lin(x, w, b):
        xd, wd = x.double(), w.double()
        if scheme == "fp32":
            y = xd @ wd.T
        elif scheme in ("bf16", "fp16"):
            dt = torch.bfloat16 if scheme == "bf16" else torch.float16
            y = rnd(x, dt).double() @ rnd(w, dt).double().T
        elif scheme == "bf16x3":
            xh, xl = split(x, torch.bfloat16)
            wh, wl = split(w, torch.bfloat16)
            y = xh.double() @ wh.double().T + xl.double() @ wh.double().T + xh.double() @ wl.double().T
        elif scheme == "fp16_x1_w2":      # fp16 activations x split fp16 weights (2 MMAs)
            wh, wl = split(w, torch.float16)
            y = rnd(x, torch.float16).double() @ (wh.double() + wl.double()).T
        elif scheme == "fp16_x2_w1":      # split fp16 activations x fp16 weights (2 MMAs)
            xh, xl = split(x, torch.float16)
            y = (xh.double() + xl.double()) @ rnd(w, torch.float16).double().T
        elif scheme.startswith("fp16_fp8c"):
            # main term in fp16, both first-order correction terms in fp8 (kind::f8f6f4 runs at twice the fp16 rate):
            #   x w ~ x16 w16 + e4m3(lo_x 2^sa) e5m2(w 2^-sa) + e5m2(x 2^-sw) e4m3(lo_w 2^sw),  lo = value - fp16(value)
            sa, sw = 6, 10
            if ":" in scheme:
                sa, sw = (int(v) for v in scheme.split(":")[1:3])
            f8 = lambda t, dt, lim: t.clamp(-lim, lim).to(dt).to(torch.float32)
            x16, w16 = rnd(x, torch.float16), rnd(w, torch.float16)
            lox, low = x - x16, w - w16
            c1 = f8(lox * 2.0 ** sa, torch.float8_e4m3fn, 448.0).double() @ f8(w * 2.0 ** -sa, torch.float8_e5m2, 57344.0).double().T
            c2 = f8(x * 2.0 ** -sw, torch.float8_e5m2, 57344.0).double() @ f8(low * 2.0 ** sw, torch.float8_e4m3fn, 448.0).double().T
            y = x16.double() @ w16.double().T + c1 + c2
        elif scheme in ("fp16_i8c", "bf16_i8c"):
            # main term in fp16 (bf16), both first-order correction terms as ONE int8 x int8 -> int32 accumulation
            # (kind::i8 issues at twice the fp16 rate: 2 instead of 3 MMA units per product).  Per-row power-of-two scales
            # sx_i (activations) and sw_j (weight rows); the residuals are quantised with the scales sx_i 2^-q, sw_j 2^-q
            # (q = 11 for fp16, 8 for bf16: |lo| <= 2^-(q+1) |value|), so both products carry the scale sx_i sw_j 2^-q and can
            # share an integer accumulator:  x w ~ x16 w16 + sx sw 2^-q (Q(lo_x) Q(w)^T + Q(x) Q(lo_w)^T).
            dt, q = (torch.float16, 11) if scheme == "fp16_i8c" else (torch.bfloat16, 8)
            xm, wm = rnd(x, dt), rnd(w, dt)
            lox, low = x - xm, w - wm
            p2 = lambda t: torch.exp2(torch.ceil(torch.log2(t.abs().amax(-1, keepdim=True).clamp_min(1e-30) / 127.0)))
            sx, sw = p2(x), p2(w)
            qi = lambda t, sc: torch.round(t / sc).clamp(-127, 127)
            corr = qi(lox, sx * 2.0 ** -q).double() @ qi(w, sw).double().T + qi(x, sx).double() @ qi(low, sw * 2.0 ** -q).double().T
            y = xm.double() @ wm.double().T + corr * (sx.double() * sw.double().T * 2.0 ** -q)
        else:
            raise ValueError(scheme)
        return (y + b.double()).float()
    return lin


def mlp(sd, x, d, emb, lin):
    enc_x, enc_d = orc.encode(x, orc.POS_LEVELS), orc.encode(d, orc.DIR_LEVELS)
    h = enc_x
    for i in range(8):
        if i == 4:
            h = torch.cat([h, enc_x], -1)
        h = F.relu(lin(h, sd[f"pts_linears.{i}.weight"], sd[f"pts_linears.{i}.bias"]))
    sigma = F.relu(F.linear(h, sd["density_head.weight"], sd["density_head.bias"]))
    hd = F.relu(lin(torch.cat([h, enc_d], -1), sd["dir_linear.weight"], sd["dir_linear.bias"]))
    hd = hd + F.linear(emb.reshape(1, -1), sd["appearance_projection.weight"], sd["appearance_projection.bias"])
    return torch.sigmoid(F.linear(hd, sd["rgb_linear.weight"], sd["rgb_linear.bias"])), sigma


def render(sd, o, d, emb, coarse, fine, u):
    b = o.shape[0]
    z, _ = orc.stratified_depths(o, d, 2.0, 6.0, 64, perturb=False)

    def shade(z, lin):
        n = z.shape[1]
        pts = o[:, None] + d[:, None] * z[..., None]
        dirs = d[:, None].expand(-1, n, -1).reshape(-1, 3)
        rgb, sig = mlp(sd, pts.reshape(-1, 3), dirs, emb, lin)
        return orc.composite(sig.reshape(b, n, 1), rgb.reshape(b, n, 3), z)
    _, _, wc = shade(z, make_linear(coarse))
    z_all, _ = orc.resample_pdf(o, d, z, wc.squeeze(-1), 128, u_rand=u)
    rf, df, wf = shade(z_all, make_linear(fine))
    return rf, df, wf.sum(1)


def render_reuse(sd, o, d, emb, coarse, fine, u):
    """The reuse_coarse path of nerfw.render: the coarse pass's (rgb, sigma) records are kept for the 64 coarse depths,
    only the 128 new depths go through the fine-pass arithmetic, and the merged row is composited."""
    b = o.shape[0]
    z, _ = orc.stratified_depths(o, d, 2.0, 6.0, 64, perturb=False)

    def raw(z, lin):
        n = z.shape[1]
        pts = o[:, None] + d[:, None] * z[..., None]
        dirs = d[:, None].expand(-1, n, -1).reshape(-1, 3)
        rgb, sig = mlp(sd, pts.reshape(-1, 3), dirs, emb, lin)
        return rgb.reshape(b, n, 3), sig.reshape(b, n, 1)
    rgb_c, sig_c = raw(z, make_linear(coarse))
    _, _, wc = orc.composite(sig_c, rgb_c, z)
    _, _, aux = orc.resample_pdf(o, d, z, wc.squeeze(-1), 128, u_rand=u, return_aux=True)
    z_new = aux["z_fine"]
    rgb_f, sig_f = raw(z_new, make_linear(fine))
    z_cat = torch.cat([z, z_new], -1)
    z_all, order = torch.sort(z_cat, dim=-1, stable=True)
    rgb_all = torch.gather(torch.cat([rgb_c, rgb_f], 1), 1, order[..., None].expand(-1, -1, 3))
    sig_all = torch.gather(torch.cat([sig_c, sig_f], 1), 1, order[..., None])
    rf, df, wf = orc.composite(sig_all, rgb_all, z_all)
    return rf, df, wf.sum(1)


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    h = w = 48
    _, _, focal, c2w = orc.golden_camera(h, w)
    ro, rd = orc.rays_for_view(h, w, focal, c2w)
    o, d = ro.reshape(-1, 3).contiguous(), F.normalize(rd.reshape(-1, 3), dim=-1)
    g = torch.Generator().manual_seed(5)
    u = torch.rand(o.shape[0], 128, generator=g)
    emb = torch.randn(32, generator=g)
    combos = [("bf16x3", "bf16x3"), ("bf16x3", "fp16"), ("bf16x3", "bf16"), ("fp16", "bf16x3"), ("bf16", "bf16x3"),
              ("bf16", "bf16"), ("fp16", "fp16"), ("fp16_x1_w2", "fp16_x1_w2"), ("fp16_x2_w1", "fp16_x2_w1"),
              ("fp16_x1_w2", "fp16"), ("fp16_x2_w1", "fp16")]
    if len(sys.argv) > 1:   # e.g. "bf16x3:fp16,fp16:fp16"
        combos = [tuple(c.split(":")) for c in sys.argv[1].split(",")]
    for variant in ("random-init", "dense x200", "dense x30"):
        sd = orc.make_state_dict(0)
        if variant != "random-init":
            k = float(variant.split("x")[1])
            sd["density_head.weight"] = sd["density_head.weight"] * k
            sd["density_head.bias"] = sd["density_head.bias"] + 1.0
        ref = render(sd, o, d, emb, "fp32", "fp32", u)
        print(f"{variant}: mean acc {float(ref[2].mean()):.3f}")
        for coarse, fine in combos:
            got = render(sd, o, d, emb, coarse, fine, u)
            e = [float((a - b).abs().max()) for a, b in zip(got, ref)]
            m = [float((a - b).abs().mean()) for a, b in zip(got, ref)]
            print(f"  coarse {coarse:11s} fine {fine:11s} max rgb {e[0]:.2e} depth {e[1]:.2e} acc {e[2]:.2e} | "
                  f"mean rgb {m[0]:.1e} depth {m[1]:.1e} acc {m[2]:.1e}", flush=True)
            got = render_reuse(sd, o, d, emb, coarse, fine, u)
            e = [float((a - b).abs().max()) for a, b in zip(got, ref)]
            print(f"     reuse_coarse (64 coarse records kept + 128 new)  max rgb {e[0]:.2e} depth {e[1]:.2e} acc {e[2]:.2e}", flush=True)


if __name__ == "__main__":
    main()
