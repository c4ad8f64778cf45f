"""CPU oracle for the NeRF-W ray-marching hot path.  TEST INFRASTRUCTURE ONLY.

This file is a restatement, in plain torch-CPU tensor arithmetic, of the algorithm the
reference implements in src/ray_utils.py, src/render.py and src/models.py.  It exists so the
CUDA path can be checked against something that runs without the reference tree (which is not
present on the GPU box).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import it; the product package never does.

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned
against the reference itself: oracle/make_golden.py imports /root/reference, runs both on the
same seeded inputs, requires bit-equality, and writes tests/golden/*.npz + manifest.json which
tests/test_oracle_golden.py re-checks everywhere (CPU box and GPU box).

Every function names the reference lines it follows.  Arithmetic is kept op-for-op (same torch
kernels in the same order) because several outputs are compared bit-exactly and because the
oracle doubles as the CPU baseline whose cost must equal the reference's.
"""
from __future__ import annotations

import torch

POS_LEVELS = 10   # config.py:32
DIR_LEVELS = 4    # config.py:33
LAST_DELTA = 1e-3  # src/render.py:58


# --------------------------------------------------------------------------- rays
def rays_for_view(height, width, focal, c2w):
    """Per-pixel world-space rays.  Follows src/ray_utils.py:4-50.

    Pixel (row i, col j) -> camera dir ((j - W/2)/f, -(i - H/2)/f, -1)  (:26-28, no half-pixel
    offset), rotated by c2w[:3,:3] as a broadcast multiply + sum over the last axis (:40-42),
    divided by its L2 norm (:45).  Origins are c2w[:3,3] expanded, stride 0 (:48).
    """
    rows = torch.arange(height, dtype=torch.float32)
    cols = torch.arange(width, dtype=torch.float32)
    ii, jj = torch.meshgrid(rows, cols, indexing="ij")
    cam = torch.stack(
        [(jj - width * 0.5) / focal, -(ii - height * 0.5) / focal, -torch.ones_like(jj)], dim=-1
    )
    rot = c2w[..., :3, :3] if c2w.shape[-1] == 4 else c2w
    world = (cam.unsqueeze(-2) * rot).sum(dim=-1)
    world = world / torch.norm(world, dim=-1, keepdim=True)
    origins = c2w[..., :3, 3].expand(world.shape)
    return origins, world


# --------------------------------------------------------------------------- sampling
def depth_table(near, far, n_samples, device="cpu"):
    """The N-entry z table of src/ray_utils.py:69-70 (linspace, then mul, then add)."""
    t = torch.linspace(0.0, 1.0, n_samples, device=device)
    return near + t * (far - near)


def stratified_depths(rays_o, rays_d, near, far, n_samples, perturb=True, t_rand=None):
    """Stratified depths and points.  Follows src/ray_utils.py:52-88.

    perturb: bins are [z0, mids..] .. [mids.., z_last]; z = lower + (upper-lower)*U(0,1)  (:77-81).
    `t_rand` lets a test inject the uniforms; by default they are drawn exactly as :80 does
    (one torch.rand of the (…,N) shape) so the global RNG stream advances identically.
    """
    z = depth_table(near, far, n_samples, rays_o.device)
    z = z.expand(list(rays_o.shape[:-1]) + [n_samples])
    if perturb:
        mid = 0.5 * (z[..., 1:] + z[..., :-1])
        hi = torch.cat([mid, z[..., -1:]], dim=-1)
        lo = torch.cat([z[..., :1], mid], dim=-1)
        if t_rand is None:
            t_rand = torch.rand(z.shape, device=rays_o.device)
        z = lo + (hi - lo) * t_rand
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    return z, pts


def resample_pdf(rays_o, rays_d, z_vals, weights, n_importance, u_rand=None, return_aux=False):
    """Inverse-CDF resampling.  Follows src/ray_utils.py:90-149 with the F2 patch.

    pdf = (w + 1e-5)/sum (:106-108); cdf = [0, cumsum(pdf)] (N+1 entries, :111-112);
    u_k = k/NI + U(0,1)/NI (:115-119); idx = lower_bound(cdf, u) (:122);
    below = max(idx-1, 0), above = min(idx, N) (:123-124); the cdf is gathered at (below, above)
    (:127-129); z is gathered at the same indices (:131-133) -- the reference raises there when
    an index equals N (z has N entries, cdf N+1).  PATCH: the z gather index is
    clamped to N-1.  This changes nothing where the reference runs without raising.
    Then t = (u - cdf_b)/max-guarded denom (:136-138), z_f = z_b + t (z_a - z_b) (:139),
    concatenate, sort ascending (:142-144), pts = o + d z (:147).
    """
    n = z_vals.shape[-1]
    pdf = weights + 1e-5
    pdf = pdf / pdf.sum(dim=-1, keepdim=True)
    cdf = torch.cumsum(pdf, dim=-1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)

    u = torch.linspace(0.0, 1.0, n_importance + 1, device=rays_o.device)[:-1]
    u = u.expand(list(cdf.shape[:-1]) + [n_importance])
    if u_rand is None:
        u_rand = torch.rand(u.shape, device=u.device)
    u = u + u_rand / n_importance

    idx = torch.searchsorted(cdf, u)
    below = torch.clamp_min(idx - 1, 0)
    above = torch.clamp_max(idx, n)
    pair = torch.stack([below, above], dim=-1)
    cdf_pair = torch.gather(
        cdf.unsqueeze(-2).expand(*cdf.shape[:-1], n_importance, n + 1), dim=-1, index=pair
    )
    z_pair = torch.gather(
        z_vals.unsqueeze(-2).expand(*z_vals.shape[:-1], n_importance, n),
        dim=-1,
        index=pair.clamp_max(n - 1),  # F2 patch
    )
    denom = cdf_pair[..., 1] - cdf_pair[..., 0]
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_pair[..., 0]) / denom
    z_fine = z_pair[..., 0] + t * (z_pair[..., 1] - z_pair[..., 0])

    merged = torch.cat([z_vals, z_fine], dim=-1)
    _, order = torch.sort(merged, dim=-1)
    merged = torch.gather(merged, dim=-1, index=order)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * merged[..., :, None]
    if return_aux:
        return merged, pts, {"inds": idx, "z_fine": z_fine, "cdf": cdf, "u": u}
    return merged, pts


# --------------------------------------------------------------------------- model
def encode(x, levels):
    """[x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...] -- src/models.py:35-44 (no pi factor)."""
    parts = [x]
    for lv in range(levels):
        f = 2 ** lv
        parts.append(torch.sin(f * x))
        parts.append(torch.cos(f * x))
    return torch.cat(parts, dim=-1)


def mlp_forward(sd, x, d, emb=None, skips=(4,), n_layers=8):
    """NeRF-W MLP on a state_dict `sd` (reference key names).  Follows src/models.py:105-162.

    trunk: h = enc_x; before layer i in skips, h = [h, enc_x] (:130-131); h = relu(W h + b).
    sigma = relu(density_head(h)) (:137-138); h_dir = relu(dir_linear([h, enc_d])) (:141-143);
    h_dir += appearance_projection(emb) *after* the relu (:146-156); rgb = sigmoid(rgb_linear) (:159-160).
    """
    F = torch.nn.functional
    enc_x = encode(x, POS_LEVELS)
    enc_d = encode(d, DIR_LEVELS)
    h = enc_x
    for i in range(n_layers):
        if i in skips:
            h = torch.cat([h, enc_x], dim=-1)
        h = F.relu(F.linear(h, sd[f"pts_linears.{i}.weight"], sd[f"pts_linears.{i}.bias"]))
    sigma = F.relu(F.linear(h, sd["density_head.weight"], sd["density_head.bias"]))
    hd = torch.cat([h, enc_d], dim=-1)
    hd = F.relu(F.linear(hd, sd["dir_linear.weight"], sd["dir_linear.bias"]))
    if emb is not None and "appearance_projection.weight" in sd:
        if emb.dim() == 1:
            emb = emb.unsqueeze(0)
        if emb.shape[0] == 1 and hd.shape[0] > 1:
            emb = emb.expand(hd.shape[0], -1)
        hd = hd + F.linear(emb, sd["appearance_projection.weight"], sd["appearance_projection.bias"])
    rgb = torch.sigmoid(F.linear(hd, sd["rgb_linear.weight"], sd["rgb_linear.bias"]))
    return rgb, sigma


# --------------------------------------------------------------------------- compositing
def composite(sigma, rgb, z_vals):
    """Front-to-back alpha compositing.  Follows src/render.py:56-80.

    sigma (B,N,1), rgb (B,N,3), z (B,N).  delta_i = z_{i+1}-z_i, last = 1e-3 (:56-58);
    alpha = 1-exp(-sigma*delta) (:67); T = exclusive cumprod of (1-alpha+1e-10) (:70-73);
    w = alpha*T (:76); rgb_map = sum w*rgb (:79); depth = sum w*z / (sum w + 1e-10) (:80).
    Returns rgb_map (B,3), depth (B,1), weights (B,N,1).
    """
    delta = z_vals[..., 1:] - z_vals[..., :-1]
    delta = torch.cat([delta, torch.ones_like(delta[..., :1]) * LAST_DELTA], dim=-1).unsqueeze(-1)
    alpha = 1.0 - torch.exp(-sigma * delta)
    trans = torch.cumprod(
        torch.cat([torch.ones_like(alpha[:, :1, :]), 1.0 - alpha + 1e-10], dim=1), dim=1
    )[:, :-1, :]
    w = alpha * trans
    rgb_map = torch.sum(w * rgb, dim=1)
    depth = torch.sum(w * z_vals.unsqueeze(-1), dim=1) / (torch.sum(w, dim=1) + 1e-10)
    return rgb_map, depth, w


def _expand_embedding(emb, batch, n):
    """src/render.py:33-46: (D,)|(1,D)|(B,D) -> (B*n, D)."""
    if emb is None:
        return None
    if emb.dim() == 1:
        emb = emb.unsqueeze(0)
    if emb.shape[0] == 1 and batch > 1:
        emb = emb.expand(batch, -1)
    return emb.unsqueeze(1).expand(-1, n, -1).reshape(-1, emb.shape[-1])


def shade(sd, rays_o, rays_d_unit, z_vals, emb):
    """Model + compositing on given depths: the body of src/render.py:29-80."""
    b, n = z_vals.shape
    pts = rays_o[..., None, :] + rays_d_unit[..., None, :] * z_vals[..., :, None]
    dirs = rays_d_unit.unsqueeze(1).expand(-1, n, -1).reshape(-1, 3)
    rgb, sigma = mlp_forward(sd, pts.reshape(-1, 3), dirs, _expand_embedding(emb, b, n))
    return composite(sigma.reshape(b, n, 1), rgb.reshape(b, n, 3), z_vals)


def render_coarse(sd, rays_o, rays_d, near, far, n_samples, emb=None, perturb=True, t_rand=None):
    """The reference's volume_render as it actually executes (coarse only).

    src/render.py:5-97: flatten, F.normalize(d) (:19), stratified depths (:22), model (:49),
    compositing (:56-80), reshape (:89-90).  Returns rgb (…,3), depth (…,1), extras.
    """
    shape = rays_o.shape
    o = rays_o.reshape(-1, 3)
    d = torch.nn.functional.normalize(rays_d.reshape(-1, 3), dim=-1)
    z, _ = stratified_depths(o, d, near, far, n_samples, perturb=perturb, t_rand=t_rand)
    rgb_map, depth, w = shade(sd, o, d, z, emb)
    return (
        rgb_map.reshape(*shape[:-1], 3),
        depth.reshape(*shape[:-1], 1),
        {"weights": w, "z_vals": z, "acc": w.sum(dim=1)},
    )


def render_hier(sd_coarse, sd_fine, rays_o, rays_d, near, far, n_samples, n_importance,
                emb=None, perturb=True, t_rand=None, u_rand=None):
    """Composed coarse+fine oracle: coarse pass, patched resample on the detached
    coarse weights, second shading pass on the merged N+NI depths.  RNG order: rand(B,N) if
    perturb, then rand(B,NI)."""
    shape = rays_o.shape
    o = rays_o.reshape(-1, 3)
    d = torch.nn.functional.normalize(rays_d.reshape(-1, 3), dim=-1)
    z, _ = stratified_depths(o, d, near, far, n_samples, perturb=perturb, t_rand=t_rand)
    rgb_c, depth_c, w_c = shade(sd_coarse, o, d, z, emb)
    z_all, _, aux = resample_pdf(o, d, z, w_c.squeeze(-1).detach(), n_importance, u_rand=u_rand, return_aux=True)
    rgb_f, depth_f, w_f = shade(sd_fine, o, d, z_all, emb)
    extras = {
        "inds": aux["inds"], "z_fine": aux["z_fine"],
        "weights": w_f, "z_vals": z_all, "acc": w_f.sum(dim=1),
        "rgb_coarse": rgb_c.reshape(*shape[:-1], 3), "depth_coarse": depth_c.reshape(*shape[:-1], 1),
        "weights_coarse": w_c, "z_vals_coarse": z,
    }
    return rgb_f.reshape(*shape[:-1], 3), depth_f.reshape(*shape[:-1], 1), extras


# --------------------------------------------------------------------------- fixtures
def make_state_dict(seed=0, appearance=True):
    """Random-init weights with the reference's module construction order (src/models.py:80-103),
    so `torch.manual_seed(seed)` yields the same tensors as `NeRF(Config())` in the reference."""
    nn = torch.nn
    torch.manual_seed(seed)
    mods = []
    mods.append(("pts_linears.0", nn.Linear(63, 256)))
    for i in range(1, 8):
        mods.append((f"pts_linears.{i}", nn.Linear(256 + (63 if i == 4 else 0), 256)))
    mods.append(("density_head", nn.Linear(256, 1)))
    mods.append(("dir_linear", nn.Linear(256 + 27, 128)))
    if appearance:
        mods.append(("appearance_projection", nn.Linear(32, 128)))
    mods.append(("rgb_linear", nn.Linear(128, 3)))
    sd = {}
    for name, m in mods:
        sd[name + ".weight"] = m.weight.detach().clone()
        sd[name + ".bias"] = m.bias.detach().clone()
    return sd


def golden_camera(h=100, w=100):
    """c2w = I with z translation 4; focal from the Blender fov."""
    import numpy as np
    c2w = torch.eye(4)
    c2w[2, 3] = 4.0
    focal = 0.5 * w / np.tan(0.5 * 0.6911112070083618)
    return h, w, focal, c2w
