"""Differentiable forward kinematics in torch (host-side glue of BASELINE.json configs[4]).

`smpl_skts(bones, rest_pose)` maps axis-angle joint rotations to the world->joint-local transforms the renderer
consumes, with autograd: it restates `get_smpl_l2ws_torch` (core/utils/skeleton_utils.py:379-463) followed by the
rigid inverse the reference obtains with `torch.inverse` (core/pose_opt.py:435, run_gan.py:447-449).  Together
with the renderer's dL/d skts (`pgn_encode_backward`) this closes the chain pose generator -> bones -> skts ->
render -> loss.  The forward-only device kernel for the generation loop is `pgn_pose_to_skts`.
"""
from __future__ import annotations

import torch

from .synthetic import SMPL_PARENTS


def axis_angle_to_matrix(rotvec: torch.Tensor) -> torch.Tensor:
    """Rodrigues formula, [...,3] -> [...,3,3] (matches scipy Rotation.from_rotvec / pytorch3d)."""
    theta = rotvec.norm(dim=-1, keepdim=True)
    small = theta < 1e-8
    k = rotvec / torch.where(small, torch.ones_like(theta), theta)
    kx, ky, kz = k[..., 0], k[..., 1], k[..., 2]
    zero = torch.zeros_like(kx)
    K = torch.stack([zero, -kz, ky, kz, zero, -kx, -ky, kx, zero], -1).reshape(*rotvec.shape[:-1], 3, 3)
    s, c = torch.sin(theta)[..., None], torch.cos(theta)[..., None]
    eye = torch.eye(3, dtype=rotvec.dtype, device=rotvec.device).expand_as(K)
    return eye + s * K + (1.0 - c) * (K @ K)


def smpl_l2ws(bones: torch.Tensor, rest_pose: torch.Tensor) -> torch.Tensor:
    """bones [B,24,3] axis-angle, rest_pose [24,3] (scaled) -> local-to-world transforms [B,24,4,4]."""
    B = bones.shape[0]
    R = axis_angle_to_matrix(bones)
    rest = rest_pose.to(bones.dtype)
    bottom = torch.tensor([0., 0., 0., 1.], dtype=bones.dtype, device=bones.device).expand(B, 1, 4)
    out = []
    for i, p in enumerate(SMPL_PARENTS):
        off = rest[i] if i == 0 else rest[i] - rest[p]
        rel = torch.cat([torch.cat([R[:, i], off.expand(B, 3)[:, :, None]], -1), bottom], 1)
        out.append(rel if i == 0 else out[p] @ rel)
    return torch.stack(out, 1)


def smpl_skts(bones: torch.Tensor, rest_pose: torch.Tensor):
    """-> (skts [B,24,4,4] world->joint-local, kps [B,24,3]); closed-form rigid inverse [R^T | -R^T t]."""
    l2w = smpl_l2ws(bones, rest_pose)
    Rt = l2w[..., :3, :3].transpose(-1, -2)
    t = -(Rt @ l2w[..., :3, 3:4])
    top = torch.cat([Rt, t], -1)
    bottom = l2w[..., 3:4, :].detach() * 0 + torch.tensor([0., 0., 0., 1.], dtype=l2w.dtype, device=l2w.device)
    return torch.cat([top, bottom], -2), l2w[..., :3, 3]
