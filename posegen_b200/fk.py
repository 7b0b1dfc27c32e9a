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
    # R = I + a [r]x + b [r]x^2 with a = sin(t)/t, b = (1 - cos t)/t^2 on the UN-normalised skew matrix and a Taylor
    # branch near t = 0 (pytorch3d's axis_angle_to_matrix, which the reference uses, skeleton_utils.py:498-499, does
    # the same): the gradient at a zero rotation is the skew generators, not zero
    t2 = (rotvec * rotvec).sum(-1, keepdim=True)
    small = t2 < 1e-8
    t2s = torch.where(small, torch.ones_like(t2), t2)
    t = torch.sqrt(t2s)
    a = torch.where(small, 1.0 - t2 / 6.0, torch.sin(t) / t)[..., None]
    b = torch.where(small, 0.5 - t2 / 24.0, (1.0 - torch.cos(t)) / t2s)[..., None]
    rx, ry, rz = rotvec[..., 0], rotvec[..., 1], rotvec[..., 2]
    zero = torch.zeros_like(rx)
    K = torch.stack([zero, -rz, ry, rz, zero, -rx, -ry, rx, zero], -1).reshape(*rotvec.shape[:-1], 3, 3)
    eye = torch.eye(3, dtype=rotvec.dtype, device=rotvec.device).expand_as(K)
    return eye + a * K + b * (K @ K)


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


class _DeviceFkFn(torch.autograd.Function):
    """bones -> (skts, kps, cyls) through the device kernels: forward `pgn_pose_to_skts`, backward `pgn_pose_fk_backward`
    (one thread per pose, fp64 chain; replaces 23 chained autograd matmuls per pose)."""

    @staticmethod
    def forward(ctx, eng, bones, rest_pose, ext_scale):
        skts, kps, cyls = eng.pose_to_skts(bones.detach().float().contiguous(), rest_pose, ext_scale=ext_scale)
        ctx.eng, ctx.rest = eng, rest_pose
        ctx.save_for_backward(bones)
        ctx.mark_non_differentiable(cyls)         # bbox / near-far geometry is detached like the reference's numpy helpers
        return skts, kps, cyls

    @staticmethod
    def backward(ctx, g_skts, g_kps, _g_cyls):
        (bones,) = ctx.saved_tensors
        if g_skts is None and g_kps is None:
            return None, None, None, None
        if g_skts is None:
            g_skts = torch.zeros((bones.shape[0], 24, 4, 4), dtype=torch.float32, device=bones.device)
        g = ctx.eng.pose_fk_backward(bones.detach().float().contiguous(), ctx.rest, g_skts, g_kps)
        return None, g.to(bones.dtype), None, None


def device_smpl_skts(eng, bones: torch.Tensor, rest_pose, ext_scale: float = 0.001):
    """Differentiable device FK: bones [B,24,3] (CUDA, may require grad) -> (skts [B,24,4,4], kps [B,24,3], cyls [B,5])."""
    return _DeviceFkFn.apply(eng, bones, rest_pose, ext_scale)
