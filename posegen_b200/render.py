"""Callers of the hot path.

`render` / `batchify_rays` restate core/trainer.py:64-147 so that code written against the
reference driver keeps working; `render_path` is the per-image driver of run_nerf.py:27-147
re-designed for the GPU: rays of the cylinder bbox are generated on the device, each image is
one kernel launch over all of its rays (no 4096-ray Python chunk loop, no per-chunk H2D of
replicated pose tensors), and the white-background composite + scatter happens on the device.
`render_pose_batch` is the batched, host-free form for the PoseGen generation loop (one camera, B generated
poses): device FK -> device cylinders -> device bboxes -> ONE ray batch and ONE fused render launch per group
of poses (the `pose_idx` form of the C ABI) -> device frame composition; the only host round trip is one
read-back of the B integer bboxes (16 B per pose) that sizes the launches.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import synthetic as syn


def batchify_rays(rays_flat, chunk=1024 * 32, ray_caster=None, **kwargs):
    """core/trainer.py:64-81.  Chunks are kept (the NaN near/far fill is per chunk)."""
    all_ret = {}
    dev = torch.device("cuda", torch.cuda.current_device())
    for i in range(0, rays_flat.shape[0], chunk):
        batch_kwargs = {k: (v[i:i + chunk].to(dev) if torch.is_tensor(v) else v) for k, v in kwargs.items()}
        ret = ray_caster(rays_flat[i:i + chunk].to(dev), **batch_kwargs)
        for k, v in ret.items():
            if v is not None:
                all_ret.setdefault(k, []).append(v)
    return {k: torch.cat(v, 0) for k, v in all_ret.items()}


def render(H, W, focal, chunk=1024 * 32, rays=None, c2w=None, near=0., far=1., center=None,
           use_viewdirs=False, c2w_staticcam=None, **kwargs):
    """core/trainer.py:84-147 for the `rays=(rays_o, rays_d)` form every PoseGen caller uses."""
    if rays is None:
        raise NotImplementedError("render(c2w=...) without rays drops into pdb in the reference (core/trainer.py:111-113)")
    rays_o, rays_d = rays
    sh = rays_d.shape
    rays_o = torch.reshape(rays_o, [-1, 3]).float()
    rays_d = torch.reshape(rays_d, [-1, 3]).float()
    cols = [rays_o, rays_d, near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])]
    if use_viewdirs:
        cols.append(rays_d / torch.norm(rays_d, dim=-1, keepdim=True))
    else:
        raise NotImplementedError("use_viewdirs=False is not part of the surreal.txt path")
    all_ret = batchify_rays(torch.cat(cols, -1), chunk, **kwargs)
    for k in all_ret:
        all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))
    return all_ret


@torch.no_grad()
def render_path(render_poses, hwf, chunk, render_kwargs, kp=None, skts=None, cyls=None, bones=None,
                white_bkgd=False, ret_acc=False, ext_scale=0.001, base_bg=1.0, full_frame=False,
                to_numpy=True, precision=None):
    """run_nerf.py:27-147 for the PoseGen generation loop (run_gan.py:2299-2337).

    render_poses [B,4,4] c2w, kp [P,24,3], skts [P,24,4,4]; pose i%P is rendered from camera i.
    Returns (rgbs [B,H,W,3], disps [B,H,W,1], accs, valid_idxs, bboxes) like the reference
    (numpy when to_numpy, else CUDA tensors).  `chunk` is honoured for the NaN-fill semantics.
    """
    H, W, focal = hwf
    ray_caster = render_kwargs["ray_caster"]
    dev = next(ray_caster.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("render_path needs the RayCaster on a CUDA device")
    eng = ray_caster.engine(dev)
    kp_np = kp.detach().cpu().numpy() if torch.is_tensor(kp) else np.asarray(kp)
    if cyls is None:
        cyls_np = syn.bounding_cylinder(kp_np, ext_scale=ext_scale)
    else:
        cyls_np = cyls.detach().cpu().numpy() if torch.is_tensor(cyls) else np.asarray(cyls)
    skts_dev = torch.as_tensor(skts, dtype=torch.float32).to(dev)
    cyls_dev = torch.as_tensor(cyls_np, dtype=torch.float32).to(dev)
    bg = base_bg if white_bkgd else 0.0
    rgbs, disps, accs, valid_idxs, bboxes = [], [], [], [], []
    poses = render_poses.detach().cpu().numpy() if torch.is_tensor(render_poses) else np.asarray(render_poses)
    for i, c2w in enumerate(poses):
        p = i % kp_np.shape[0]
        f = float(focal) if np.isscalar(focal) else float(focal[i])
        if full_frame:
            tl, br = np.array([0, 0]), np.array([W, H])
        else:
            tl, br = syn.cylinder_bbox_2d(cyls_np[p], H, W, f, c2w)
        x0, y0, x1, y1 = int(tl[0]), int(tl[1]), int(br[0]), int(br[1])
        bboxes.append((tl, br))
        valid_idxs.append(syn.bbox_pixel_indices(tl, br, W))
        n = max(x1 - x0, 0) * max(y1 - y0, 0)
        if n > 0:
            rb = eng.generate_rays(H, W, f, c2w, x0, y0, x1, y1)
            ret = eng.render(rb, skts_dev[p], cyls_dev[p], nanfill_chunk=chunk,
                             precision=precision or ray_caster.precision, return_alpha=False)
            img = eng.compose_frame(H, W, x0, y0, x1, y1, ret["rgb_map"], ret["acc_map"], bg)
            disp = torch.zeros(H * W, device=dev)
            acc = torch.zeros(H * W, device=dev)
            # flat pixel indices of the bbox formed on the device (a pageable H2D copy of valid_idxs would stall the
            # launch queue behind the render)
            idx = ((torch.arange(y0, y1, device=dev) * W)[:, None] + torch.arange(x0, x1, device=dev)[None, :]).reshape(-1)
            disp[idx] = ret["disp_map"]
            acc[idx] = ret["acc_map"]
        else:
            img = torch.full((H, W, 3), bg, device=dev)
            disp = torch.zeros(H * W, device=dev)
            acc = torch.zeros(H * W, device=dev)
        rgbs.append(img)
        disps.append(disp.view(H, W, 1))
        accs.append(acc.view(H, W, 1))
    rgbs, disps, accs = torch.stack(rgbs), torch.nan_to_num(torch.stack(disps), nan=0.0), torch.stack(accs)
    torch.cuda.current_stream(dev).synchronize()
    eng.check_status()                 # a tripped device-side watchdog must not hand back garbage frames silently
    if to_numpy:
        return rgbs.cpu().numpy(), disps.cpu().numpy(), (accs.cpu().numpy() if ret_acc else []), valid_idxs, bboxes
    return rgbs, disps, (accs if ret_acc else []), valid_idxs, bboxes


@torch.no_grad()
def render_pose_batch(ray_caster, bones, rest_pose, c2w, hwf, crop=None, bg=1.0, chunk=4096, poses_per_launch=16,
                      ext_scale=0.001, precision=None, hmr_res=224):
    """bones [B,24,3] axis-angle (CUDA) -> (frames [B,H,W,3], HMR inputs [B,3,R,R] or None, number of rays rendered).

    The generation loop of run_gan.py:2299-2347 (render every generated pose from the fixed camera, hand the crop to
    HMR) without per-image host work: `pgn_pose_to_skts` -> `pgn_cylinder_bboxes` -> [one 16 B/pose read-back] ->
    per group of `poses_per_launch` poses: `pgn_generate_rays_batch` -> `pgn_render_forward` (pose_idx form, explicit
    per-image chunk table so the near/far NaN fill keeps the reference's per-image 4096-ray chunks) ->
    `pgn_compose_frames_batch`; then `pgn_frame_to_hmr_input` per frame when `crop` is given."""
    H, W, focal = hwf
    dev = bones.device
    eng = ray_caster.engine(dev)
    B = bones.shape[0]
    skts, kps, cyls = eng.pose_to_skts(bones.float().contiguous(), rest_pose, ext_scale=ext_scale)
    bboxes = eng.cylinder_bboxes(cyls, c2w, H, W, float(focal))
    bb = bboxes.cpu().numpy().astype(np.int64)                       # the loop's only device->host read
    areas = np.clip(bb[:, 2] - bb[:, 0], 0, None) * np.clip(bb[:, 3] - bb[:, 1], 0, None)
    # launch tables for every group, built and uploaded once, before the first render launch
    groups, flat = [], []
    for s in range(0, B, poses_per_launch):
        a = areas[s:s + poses_per_launch]
        off = np.concatenate([[0], np.cumsum(a)]).astype(np.int64)
        cs = [np.arange(off[i], off[i + 1], chunk, dtype=np.int64) for i in range(len(a))]
        cs = np.concatenate(cs + [off[-1:]]) if off[-1] > 0 else np.zeros(1, np.int64)
        groups.append((s, len(a), int(off[-1]), int(a.max()) if len(a) else 0, len(off), len(cs)))
        flat += [off, cs]
    table = torch.from_numpy(np.concatenate(flat)).pin_memory().to(dev, non_blocking=True)
    frames = torch.empty((B, H, W, 3), dtype=torch.float32, device=dev)
    pos = 0
    for s, nb, n_total, max_rays, n_off, n_cs in groups:
        off_d, cs_d = table[pos:pos + n_off], table[pos + n_off:pos + n_off + n_cs]
        pos += n_off + n_cs
        if n_total == 0:
            frames[s:s + nb] = bg
            continue
        rb, pidx = eng.generate_rays_batch(H, W, float(focal), c2w, bboxes[s:s + nb], off_d, n_total, max_rays)
        ret = eng.render(rb, skts[s:s + nb], cyls[s:s + nb], pose_idx=pidx, chunk_starts=cs_d,
                         precision=precision or ray_caster.precision, return_alpha=False)
        frames[s:s + nb] = eng.compose_frames_batch(H, W, bboxes[s:s + nb], off_d, ret["rgb_map"], ret["acc_map"], bg)
    hmr = None
    if crop is not None:
        hmr = torch.stack([eng.frame_to_hmr_input(f, crop=crop, out_res=hmr_res) for f in frames])
    eng.poll_status()
    return frames, hmr, int(areas.sum())
