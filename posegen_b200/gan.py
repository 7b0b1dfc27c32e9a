"""Differentiable frame render for the PoseGen GAN step (BASELINE.json configs[4]).

The reference's generator step renders `rpi` generated poses to PNG files, reads them back, crops / resizes them
with skimage and runs HMR under `no_grad` (run_gan.py:2040-2091, 2299-2347): no gradient reaches the generator
through the image.  configs[4] asks for that chain with the render differentiable in the pose:

    bones --fk.smpl_skts--> skts --render_frame--> image --hmr_input--> HMR --> loss
                 (torch)          (this file)             (this file)

`render_frame` is one fused-kernel launch over all bbox rays of the frame (`pgn_render_forward_masks`).  The NeRF is
frozen in the GAN step (run_gan.py:159-160), so the backward needs no weight gradients and therefore no activations:
the forward keeps only the fine pass's ReLU masks as bits (272 B per sample, 4.7 GB for a 512x512 frame instead of the
134 GB of a full activation dump; the image reads `rgb_map` / `acc_map`, so nothing flows into the coarse network, and
the importance samples are detached, core/utils/ray_utils.py:286) plus the per-sample network outputs.  The backward
walks the rays with a non-zero upstream gradient (the HMR crop) in chunks: `pgn_composite_backward` ->
`pgn_view_delta_from_mask` -> `pgn_mlp_delta_chain` (the trunk's delta chain on tcgen05) -> three input-gradient
GEMMs -> `pgn_encode_backward_bf16` -> dL/d skts.  Nothing is re-rendered.

`hmr_input` is `pgn_frame_to_hmr_input` (uint8 quantisation as a straight-through estimator, crop, normalise,
anti-aliased resize) with the adjoint of the separable resize as its backward; the 1-D operator is read off the
CUDA kernel itself (one probe launch per crop row, cached), so forward and backward cannot drift apart.
"""
from __future__ import annotations

import functools
from typing import Dict, Tuple

import numpy as np
import torch

from . import synthetic as syn
from .train import T, mlp_backward


class _FrameRenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rc, ray_batch, skts, cyl, chunk):
        eng = rc.engine(ray_batch.device)
        ret, masks = eng.render_masks(ray_batch, skts, cyl, nanfill_chunk=0)
        ctx.rc, ctx.eng, ctx.chunk, ctx.masks = rc, eng, int(chunk), masks
        ctx.save_for_backward(ray_batch, skts, cyl, ret["raw"], ret["z_fine"])
        return ret["rgb_map"], ret["acc_map"]

    @staticmethod
    def backward(ctx, g_rgb, g_acc):
        rb_all, sk, cy, raw_all, z_all = ctx.saved_tensors
        rc, eng = ctx.rc, ctx.eng
        trunk_mask, view_mask = ctx.masks
        ctx.masks = None
        dev = rb_all.device
        n = rb_all.shape[0]
        g_rgb = torch.zeros((n, 3), device=dev) if g_rgb is None else g_rgb.float()
        g_acc = torch.zeros((n,), device=dev) if g_acc is None else g_acc.float()
        live = ((g_rgb.abs().sum(-1) + g_acc.abs()) > 0).nonzero().squeeze(-1)      # rays the loss reads (e.g. the HMR crop)
        d_skts = torch.zeros((24, 4, 4), dtype=torch.float32, device=dev)
        pd = dict(rc.network_fine.named_parameters())
        rows_all = trunk_mask.shape[2]
        if rows_all % T:
            raise RuntimeError("mask dump rows are not whole rays")
        trunk_rays = trunk_mask.view(64, rows_all // T, T)                          # per (layer, word) plane and ray: 80 samples
        view_rays = view_mask.view(4, rows_all // T, T)
        for i in range(0, live.numel(), ctx.chunk):
            idx = live[i:i + ctx.chunk].contiguous()
            k = idx.numel()
            rb = rb_all.index_select(0, idx)
            z = eng.gather_ray_rows(z_all, idx)                                      # the rays' samples in the fine pass
            d_raw = eng.composite_backward(rb, sk, cy, eng.gather_ray_rows(raw_all, idx), z, g_rgb.index_select(0, idx).contiguous(),
                                           g_acc.index_select(0, idx).contiguous())
            mask_dump = (eng.gather_ray_rows(trunk_rays, idx, n_planes=64).view(8, 8, k * T),
                         eng.gather_ray_rows(view_rays, idx, n_planes=4).view(4, k * T))
            gd = mlp_backward(pd, None, None, d_raw.reshape(-1, 4), eng.mlp_delta, want_input_grad=True, want_weight_grad=False,
                              chain=functools.partial(eng.mlp_delta_chain_net, 1), mask_dump=mask_dump, view_delta=eng.view_delta_from_mask,
                              input_grads=functools.partial(eng.mlp_input_grads, 1))
            d = eng.encode_backward_bf16(rb, sk, cy, z, gd["_g_xp"], gd["_g_d"], tile_blocked=gd.get("_g_tb", False))
            d_skts += d.sum(0)
            del gd, d, mask_dump
        return None, None, d_skts, None, None


def render_frame(rc, ray_batch: torch.Tensor, skts: torch.Tensor, cyl: torch.Tensor, chunk: int = 16384) -> Tuple[torch.Tensor, torch.Tensor]:
    """(rgb_map [n,3], acc_map [n]) of all rays of one frame, differentiable w.r.t. skts [24,4,4] (one pose)."""
    if skts.shape != (24, 4, 4):
        raise ValueError("render_frame renders one pose: skts must be [24,4,4]")
    if torch.is_grad_enabled() and any(p.requires_grad for net in (rc.network, rc.network_fine) for p in net.parameters()):
        raise RuntimeError("render_frame back-propagates to the pose only and keeps no activations: freeze the NeRF "
                           "(run_gan.py:159-160) or use RayCaster.forward in train mode for weight gradients")
    return _FrameRenderFn.apply(rc, ray_batch.float().contiguous(), skts.float(), cyl.float().contiguous(), chunk)


def compose_white(rgb_map, acc_map, H, W, x0, y0, x1, y1, bg: float = 1.0) -> torch.Tensor:
    """run_nerf.py:100-133 as differentiable torch indexing: image[bbox] = rgb + (1 - acc) * bg, `bg` elsewhere."""
    img = torch.full((H, W, 3), float(bg), dtype=rgb_map.dtype, device=rgb_map.device)
    patch = rgb_map + (1.0 - acc_map[:, None]) * bg
    img[y0:y1, x0:x1] = patch.view(y1 - y0, x1 - x0, 3)
    return img


_RESIZE_OPS: Dict[Tuple[int, int, int], torch.Tensor] = {}


def resize_operator(eng, n_in: int, n_out: int) -> torch.Tensor:
    """The 1-D operator A [n_out, n_in] of pgn_frame_to_hmr_input's separable anti-aliased resize, read off the kernel:
    an image that is 1 on row k and 0 elsewhere (mean 0, std 1, no quantisation) comes out as A[:, k] x (A 1),
    and A 1 = 1 (mirror boundaries keep constants)."""
    key = (eng.device.index or 0, n_in, n_out)
    if key not in _RESIZE_OPS:
        A = torch.empty((n_out, n_in), dtype=torch.float32, device=eng.device)
        img = torch.zeros((n_in, n_in, 3), dtype=torch.float32, device=eng.device)
        for k in range(n_in):
            img[k] = 1.0
            out = eng.frame_to_hmr_input(img, crop=(0, 0, n_in, n_in), out_res=n_out, mean=(0., 0., 0.), std=(1., 1., 1.),
                                         quantize_u8=False)
            A[:, k] = out[0, :, n_out // 2]
            img[k] = 0.0
        _RESIZE_OPS[key] = A
    return _RESIZE_OPS[key]


class _HmrInputFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, image, crop, out_res, mean, std, quantize_u8):
        x0, y0, x1, y1 = crop
        if x1 - x0 != y1 - y0:
            raise ValueError("hmr_input: square crops only (run_gan.py:2059 crops [100:412, 100:412])")
        ctx.eng, ctx.crop, ctx.out_res, ctx.std, ctx.shape = eng, crop, out_res, std, image.shape
        return eng.frame_to_hmr_input(image.detach(), crop, out_res, mean, std, quantize_u8)

    @staticmethod
    def backward(ctx, g_out):
        x0, y0, x1, y1 = ctx.crop
        A = resize_operator(ctx.eng, x1 - x0, ctx.out_res)
        inv_std = torch.tensor([1.0 / s for s in ctx.std], dtype=torch.float32, device=g_out.device)
        g_crop = torch.matmul(torch.matmul(A.t(), g_out.float()), A) * inv_std[:, None, None]        # [3, n_in, n_in]
        g_img = torch.zeros(ctx.shape, dtype=torch.float32, device=g_out.device)
        g_img[y0:y1, x0:x1] = g_crop.permute(1, 2, 0)
        return None, g_img, None, None, None, None, None


def hmr_input(eng, image: torch.Tensor, crop=(100, 100, 412, 412), out_res: int = 224, mean=(0.485, 0.456, 0.406),
              std=(0.485, 0.456, 0.406), quantize_u8: bool = True) -> torch.Tensor:
    """Frame [H,W,3] in [0,1] -> HMR input [3,R,R] (run_gan.py:2057-2071, 2343), differentiable w.r.t. the frame
    (the uint8 quantisation of the PNG round trip is a straight-through estimator)."""
    return _HmrInputFn.apply(eng, image, tuple(crop), int(out_res), tuple(mean), tuple(std), bool(quantize_u8))


def render_pose_images(rc, bones: torch.Tensor, rest_pose: torch.Tensor, c2w: np.ndarray, H: int = 512, W: int = 512,
                       focal: float = 1000.0, ext_scale: float = 0.001, chunk: int = 16384, bg: float = 1.0):
    """bones [B,24,3] axis-angle (requires grad) -> frames [B,H,W,3] with autograd back to `bones`.

    FK forward and backward on the device (`fk.device_smpl_skts`: `pgn_pose_to_skts` / `pgn_pose_fk_backward`), cylinders
    and bboxes on the device (`pgn_cylinder_bboxes`, detached like the reference's numpy `kp_to_valid_rays`; one read-back
    of the B integer bboxes sizes the launches), rays on the device."""
    from . import fk
    dev = bones.device
    eng = rc.engine(dev)
    skts, kps, cyls = fk.device_smpl_skts(eng, bones, rest_pose, ext_scale)
    bb = eng.cylinder_bboxes(cyls, c2w, H, W, float(focal)).cpu().numpy()
    frames = []
    for b in range(bones.shape[0]):
        x0, y0, x1, y1 = (int(v) for v in bb[b])
        rb = eng.generate_rays(H, W, float(focal), c2w, x0, y0, x1, y1)
        rgb, acc = render_frame(rc, rb, skts[b], cyls[b], chunk)
        frames.append(compose_white(rgb, acc, H, W, x0, y0, x1, y1, bg))
    return torch.stack(frames), kps
