"""CPU restatement of PoseGen's A-NeRF render path (the ORACLE).

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module;
nothing under ``posegen_b200/`` does.  It is a plain PyTorch (fp32 or fp64, any
device) re-statement of the reference algorithm, op for op, so that (a) it can be
pinned against the unmodified reference (``oracle/make_golden.py`` ->
``tests/golden/*.npz``; tests/test_oracle_golden.py) and (b) its CPU timing is
representative of the reference's own PyTorch CPU path.

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md §4), so the pin is the reference itself executed in the build container
under ``oracle/ref_shim.py``; the fixtures it produced are committed under
``tests/golden/`` together with the generating script.

Every function cites the reference lines it follows (paths relative to the
reference repo root).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

N_JOINTS = 24


# ----------------------------------------------------------------------------
# core/utils/ray_utils.py:292-344  get_near_far_in_cylinder
# ----------------------------------------------------------------------------
def near_far_in_cylinder(rays_o, rays_d, cyl, near, far):
    g = [0, 2]
    r_near = (rays_o + rays_d * near)[..., g]
    r_far = (rays_o + rays_d * far)[..., g]
    radius = cyl[..., 2:3]
    center = cyl[..., :2]
    nc = center - r_near
    nf = r_far - r_near
    nf_norm = torch.norm(nf, dim=-1, p=2)
    scale = torch.norm(rays_d[..., g], dim=-1, p=2)[..., None]
    cross = nc[..., 0] * nf[..., 1] - nc[..., 1] * nf[..., 0]
    dist = (torch.abs(cross) / nf_norm)[..., None]
    Q = (radius.pow(2) - dist.pow(2)).pow(0.5)
    K = ((nc * nf).sum(-1) / nf_norm)[..., None]
    mask = (Q < K).to(rays_o.dtype)
    new_near = near + mask * (K - Q) / scale
    new_far = near + (K + Q) / scale
    if torch.isnan(new_near).any():
        # rays that miss the cylinder (Q = NaN) take the mean near/far of the chunk
        # (ray_utils.py:328-342; numpy nanmean is a float32 pairwise mean)
        miss = torch.isnan(Q)[..., 0]
        avg_near = _nanmean32(new_near)
        new_near[miss] = avg_near if avg_near == avg_near else near[miss]
        avg_far = _nanmean32(new_far)
        new_far[miss] = avg_far if avg_far == avg_far else far[miss]
    return new_near, new_far


def _nanmean32(x):
    import numpy as np
    return float(np.nanmean(x.detach().cpu().numpy()))


# ----------------------------------------------------------------------------
# core/utils/ray_utils.py:204-251  sample_from_lineseg (lindisp False).  perturb > 0 (training): the
# stratified jitter takes its uniform numbers from `t_rand` [N,n_samples] (the reference draws torch.rand).
# ----------------------------------------------------------------------------
def coarse_z_vals(near, far, n_samples, t_rand=None, lindisp=False):
    t = torch.linspace(0., 1., steps=n_samples, dtype=near.dtype, device=near.device)
    t = t.expand(near.shape[0], n_samples)
    if not lindisp:
        z_vals = near * (1. - t) + far * t
    else:                                   # linear in inverse depth (ray_utils.py:224-227)
        z_vals = 1. / (1. / near * (1. - t) + 1. / far * t)
    if t_rand is not None:
        mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
        upper = torch.cat([mids, z_vals[..., -1:]], -1)
        lower = torch.cat([z_vals[..., :1], mids], -1)
        z_vals = lower + (upper - lower) * t_rand
    return z_vals


# ----------------------------------------------------------------------------
# core/encoders.py:8-37  transform_batch_pts / transform_batch_rays
# ----------------------------------------------------------------------------
def to_joint_frames(pts, skts):
    """pts [N,S,3], skts [N,24,4,4] -> pts_t [N,S,24,3] via the homogeneous matmul."""
    n, s = pts.shape[:2]
    homo = torch.cat([pts, torch.ones(n, s, 1, dtype=pts.dtype, device=pts.device)], -1)
    homo = homo.view(n, 1, s, 4).expand(-1, skts.shape[1], -1, -1).transpose(3, 2).contiguous()
    mm = (skts @ homo).permute(0, 3, 1, 2).contiguous()
    return mm[..., :3]


def dirs_to_joint_frames(rays_d, skts):
    """rays_d [N,3] -> [N,1,24,3], rotation part only."""
    n = rays_d.shape[0]
    rot = skts[..., :3, :3]
    d = rays_d.view(n, 1, 1, 3).expand(-1, skts.shape[1], -1, -1).transpose(3, 2).contiguous()
    return (rot @ d).permute(0, 3, 1, 2).contiguous()


# ----------------------------------------------------------------------------
# core/cutoff_embedder.py:111-174  CutoffEmbedder._embed
# ----------------------------------------------------------------------------
def cutoff_embed(x, dists, n_freqs, tau, cutoff_dist, dist_inputs):
    """x [N,S,C]; dists [N,S,24].  Returns [N,S,(1+2L)*C]; channel order k*C + c.
    include_input and cutoff_inputs are both on for surreal.txt."""
    freqs = 2. ** torch.linspace(0., n_freqs - 1, steps=n_freqs, dtype=x.dtype, device=x.device)
    if dist_inputs:
        expand = x.shape[-1] // dists.shape[-1]
        d = dists[..., None].expand(*dists.shape, expand).flatten(start_dim=-2)
        cut = cutoff_dist[:, None].expand(-1, expand).flatten(start_dim=-2)
    else:
        d = x
        cut = cutoff_dist
    x_freq = freqs.view(1, -1, 1) * x[..., None, :]
    v = (tau * (d - cut))[..., None, :]
    w = 1. - torch.sigmoid(v)
    emb = torch.stack([torch.sin(x_freq), torch.cos(x_freq)], dim=-2).flatten(start_dim=-3, end_dim=-2)
    emb = torch.cat([x[..., None, :], emb], dim=-2) * w
    return emb.flatten(start_dim=-2), w


# ----------------------------------------------------------------------------
# core/raycasters.py:476-555  encode_inputs  (reldist / reldir / relray, surreal.txt)
# ----------------------------------------------------------------------------
def encode(pts, rays_d, skts, emb):
    """Returns the [N,S,1080] network input: [v_emb(360) | r(72) | d_emb(648)]."""
    pts_t = to_joint_frames(pts, skts)
    rays_t = dirs_to_joint_frames(rays_d, skts)
    v = torch.norm(pts_t, dim=-1, p=2)                                   # encoders.py:110-122
    r = F.normalize(pts_t, dim=-1, p=2).flatten(start_dim=2)             # encoders.py:181-193
    d = F.normalize(rays_t, dim=-1, p=2).flatten(start_dim=2).expand(*pts_t.shape[:2], -1)
    v_emb, _ = cutoff_embed(v, v, emb["multires"], emb["tau_v"], emb["cutoff_v"], dist_inputs=False)
    d_emb, _ = cutoff_embed(d, v, emb["multires_views"], emb["tau_d"], emb["cutoff_d"], dist_inputs=True)
    return torch.cat([v_emb, r, d_emb], dim=-1)


# ----------------------------------------------------------------------------
# core/networks/nerf.py:94-148  NeRF.forward (use_viewdirs, skips=[4]); frame_code [rows,16] = the Optcodes
# frame code of each row's camera (nerf.py:104-131: input_views = cat([input_views, framecodes])), None without
# ----------------------------------------------------------------------------
def frame_codes(net, cams, n_rows_per_ray, n_rays, training=False):
    """core/networks/embedding.py:19-31: codes[cam] per ray, the mean code in eval when every index is < 0 (cams None
    counts as -1), repeated for the ray's samples -> [n_rays * n_rows_per_ray, 16]; None for a net without codes."""
    if "framecodes.codes.weight" not in net:
        return None
    codes = net["framecodes.codes.weight"]
    if cams is None:
        cams = torch.full((n_rays,), -1, dtype=torch.long)
    if not training and int(cams.max()) < 0:
        c = codes.mean(0, keepdim=True).expand(n_rays, -1)
    else:
        c = codes[cams.long()]
    return c[:, None, :].expand(-1, n_rows_per_ray, -1).reshape(-1, codes.shape[1])


def nerf_forward(x, net, chunk=1024 * 64, frame_code=None):
    outs = []
    for i in range(0, x.shape[0], chunk):
        xi = x[i:i + chunk]
        x_p, x_v = xi[:, :432], xi[:, 432:]
        h = x_p
        for l in range(8):
            h = F.relu(F.linear(h, net[f"pts_linears.{l}.weight"], net[f"pts_linears.{l}.bias"]))
            if l == 4:
                h = torch.cat([x_p, h], -1)
        alpha = F.linear(h, net["alpha_linear.weight"], net["alpha_linear.bias"])
        feat = F.linear(h, net["feature_linear.weight"], net["feature_linear.bias"])
        vin = [feat, x_v] if frame_code is None else [feat, x_v, frame_code[i:i + chunk]]
        g = F.relu(F.linear(torch.cat(vin, -1), net["views_linears.0.weight"],
                            net["views_linears.0.bias"]))
        rgb = F.linear(g, net["rgb_linear.weight"], net["rgb_linear.bias"])
        outs.append(torch.cat([rgb, alpha], -1))
    return torch.cat(outs, 0)


# ----------------------------------------------------------------------------
# core/networks/nerf.py:150-205  raw2outputs (act=relu, B=density_scale).  `noise` [N,S] (training,
# raw_noise_std > 0) = randn * raw_noise_std * B, added to raw_sigma / B before the ReLU (nerf.py:165,176-186).
# ----------------------------------------------------------------------------
def raw2outputs(raw, z_vals, rays_d, density_scale=1.0, rgb_eps=0.001, noise=None):
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    big = torch.full_like(dists[..., :1], 1e10)
    dists = torch.cat([dists, big], -1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :3]) * (1 + 2 * rgb_eps) - rgb_eps
    alpha = 1. - torch.exp(-F.relu(raw[..., 3] / density_scale + (0. if noise is None else noise)) * dists)
    ones = torch.ones((alpha.shape[0], 1), dtype=alpha.dtype, device=alpha.device)
    weights = alpha * torch.cumprod(torch.cat([ones, 1. - alpha + 1e-10], -1), -1)[:, :-1]
    rgb_map = torch.sum(weights[..., None] * rgb, -2)
    depth_map = torch.sum(weights * z_vals, -1)
    acc = torch.sum(weights, -1)
    disp_map = 1. / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / (acc + 1e-10))
    disp_map = disp_map * (~torch.isclose(acc, torch.zeros_like(acc))).to(disp_map.dtype)
    acc_map = torch.minimum(acc, torch.ones_like(acc))
    return {"rgb_map": rgb_map, "disp_map": disp_map, "acc_map": acc_map,
            "weights": weights, "alpha": alpha}


# ----------------------------------------------------------------------------
# core/utils/ray_utils.py:157-201  sample_pdf (det=True; det=False when `u` [N,n_samples] is given: the
# reference then draws torch.rand)
# ----------------------------------------------------------------------------
def sample_pdf_det(bins, weights, n_samples, u=None):
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if u is None:
        u = torch.linspace(0., 1., steps=n_samples, dtype=cdf.dtype, device=cdf.device)
        u = u.expand(list(cdf.shape[:-1]) + [n_samples])
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_b, bin_a = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    return bin_b + t * (bin_a - bin_b), inds, cdf


# ----------------------------------------------------------------------------
# core/utils/ray_utils.py:255-289  isample_from_lineseg (is_only False)
# ----------------------------------------------------------------------------
def importance_z_vals(z_vals, weights, n_importance, u=None):
    mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
    z_samples, inds, cdf = sample_pdf_det(mids, weights[..., 1:-1], n_importance, u=u)
    z_all, sorted_idxs = torch.sort(torch.cat([z_vals, z_samples], -1), -1)
    return z_all, z_samples, sorted_idxs, inds, cdf


# ----------------------------------------------------------------------------
# core/raycasters.py:361-474  render_rays  (eval path: perturb 0, no noise)
# ----------------------------------------------------------------------------
def render_rays(ray_batch, skts, cyls, nets, emb, n_samples=64, n_importance=16,
                density_scale=1.0, taps=None, cams=None, lindisp=False):
    """ray_batch [N,11]; skts [N,24,4,4]; cyls [N,5]; nets = (coarse, fine) state dicts.
    Returns the reference's output dict (core/raycasters.py:711-724).  `taps`, if a
    dict, receives intermediate tensors for stage-level parity tests."""
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    near, far = near_far_in_cylinder(rays_o, rays_d, cyls, near, far)
    z_vals = coarse_z_vals(near, far, n_samples, lindisp=lindisp)
    pts = rays_o[:, None, :] + rays_d[:, None, :] * z_vals[:, :, None]
    enc = encode(pts, rays_d, skts, emb)
    raw = nerf_forward(enc.reshape(-1, enc.shape[-1]), nets[0],
                       frame_code=frame_codes(nets[0], cams, enc.shape[1], enc.shape[0])).reshape(*enc.shape[:2], 4)
    ret0 = raw2outputs(raw, z_vals, rays_d, density_scale)
    if taps is not None:
        taps.update(near=near, far=far, z_coarse=z_vals, enc_coarse=enc, raw_coarse=raw,
                    weights_coarse=ret0["weights"])

    z_all, z_samples, sorted_idxs, inds, cdf = importance_z_vals(z_vals, ret0["weights"], n_importance)
    pts_is = rays_o[:, None, :] + rays_d[:, None, :] * z_samples[:, :, None]
    enc_is = encode(pts_is, rays_d, skts, emb)
    # core/raycasters.py:679-709,796-812: concat coarse + new encodings, gather in z order
    merged = torch.cat([enc, enc_is], dim=1)
    merged = torch.gather(merged, 1, sorted_idxs[..., None].expand(-1, -1, merged.shape[-1]))
    raw_f = nerf_forward(merged.reshape(-1, merged.shape[-1]), nets[1],
                         frame_code=frame_codes(nets[1], cams, merged.shape[1], merged.shape[0])).reshape(*merged.shape[:2], 4)
    ret = raw2outputs(raw_f, z_all, rays_d, density_scale)
    if taps is not None:
        taps.update(z_samples=z_samples, z_fine=z_all, sorted_idxs=sorted_idxs, pdf_inds=inds,
                    cdf=cdf, raw_fine=raw_f, weights_fine=ret["weights"])
    return {"rgb_map": ret["rgb_map"], "disp_map": ret["disp_map"], "acc_map": ret["acc_map"],
            "alpha": ret["alpha"], "rgb0": ret0["rgb_map"], "disp0": ret0["disp_map"],
            "acc0": ret0["acc_map"], "alpha0": ret0["alpha"]}


# ----------------------------------------------------------------------------
# core/trainer.py:64-81  batchify_rays (chunk loop; chunk matters for the NaN fill)
# ----------------------------------------------------------------------------
@torch.no_grad()
def render(ray_batch, skts, cyls, nets, emb, chunk=4096, **kw):
    """skts [24,4,4] or [N,24,4,4]; cyls [5] or [N,5] (expanded per ray like run_nerf.py:63-90)."""
    n = ray_batch.shape[0]
    if skts.dim() == 3:
        skts = skts[None].expand(n, -1, -1, -1)
    if cyls.dim() == 1:
        cyls = cyls[None].expand(n, -1)
    outs = {}
    cams = kw.pop("cams", None)
    for i in range(0, n, chunk):
        ret = render_rays(ray_batch[i:i + chunk], skts[i:i + chunk], cyls[i:i + chunk], nets, emb,
                          cams=None if cams is None else cams[i:i + chunk], **kw)
        for k, v in ret.items():
            outs.setdefault(k, []).append(v)
    return {k: torch.cat(v, 0) for k, v in outs.items()}


def default_embed_params(dtype=torch.float32, device="cpu", tau=20.0, cutoff=0.5):
    """surreal.txt embedder scalars (core/raycasters.py:30-79, cutoff_embedder.py:94-95)."""
    c = torch.full((N_JOINTS,), cutoff, dtype=dtype, device=device)
    return {"multires": 7, "multires_views": 4, "tau_v": torch.tensor(tau, dtype=dtype, device=device),
            "tau_d": torch.tensor(tau, dtype=dtype, device=device), "cutoff_v": c, "cutoff_d": c.clone()}


def nets_from_ckpt(ckpt, dtype=torch.float32, device="cpu"):
    def conv(sd):
        return {k: torch.as_tensor(v).to(dtype=dtype, device=device) for k, v in sd.items()}
    return conv(ckpt["network_fn_state_dict"]), conv(ckpt["network_fine_state_dict"])


def embed_params_from_ckpt(ckpt, dtype=torch.float32, device="cpu"):
    e, d = ckpt["embed_state_dict"], ckpt["embeddirs_state_dict"]
    t = lambda x: torch.as_tensor(x).to(dtype=dtype, device=device)  # noqa: E731
    return {"multires": 7, "multires_views": 4, "tau_v": t(e["tau"]), "tau_d": t(d["tau"]),
            "cutoff_v": t(e["cutoff_dist"]), "cutoff_d": t(d["cutoff_dist"])}
