"""Differentiable frame render for the PoseGen GAN step (BASELINE.json configs[4]).

The reference's generator step renders `rpi` generated poses to PNG files, reads them back, crops / resizes them
with skimage and runs HMR under `no_grad` (run_gan.py:2040-2091, 2299-2347): no gradient reaches the generator
through the image.  configs[4] asks for that chain with the render differentiable in the pose:

    bones --fk.smpl_skts--> skts --render_frame--> image --hmr_input--> HMR --> loss
                 (torch)          (this file)             (this file)

`render_frame` is one fused-kernel launch over all bbox rays of the frame in the forward (no activation dump: a
512x512 frame would need 134 GB of it).  Its backward walks the frame in ray chunks: each chunk is re-rendered with
`pgn_render_forward_train` (fine-pass dump only: the image reads `rgb_map` / `acc_map`, so nothing flows into the
coarse network, and the importance samples are detached, core/utils/ray_utils.py:286), then
`pgn_composite_backward` -> the input-gradient GEMM chain of `train.mlp_backward` (the NeRF is frozen in the GAN
step, run_gan.py:159-160: no weight gradients) -> `pgn_encode_backward` -> dL/d skts.  Rays whose upstream gradient
is zero (outside the HMR crop) are skipped.

`hmr_input` is `pgn_frame_to_hmr_input` (uint8 quantisation as a straight-through estimator, crop, normalise,
anti-aliased resize) with the adjoint of the separable resize as its backward; the 1-D operator is read off the
CUDA kernel itself (one probe launch per crop row, cached), so forward and backward cannot drift apart.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch

from . import synthetic as syn
from . import train
from .train import mlp_backward


class _FrameRenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rc, ray_batch, skts, cyl, chunk, live, dump_budget):
        eng = rc.engine(ray_batch.device)
        ctx.rc, ctx.eng, ctx.chunk = rc, eng, int(chunk)
        ctx.kept = None
        if live is None:
            ret = eng.render(ray_batch, skts, cyl, nanfill_chunk=0, precision="bf16", return_alpha=False)
            ctx.save_for_backward(ray_batch, skts, cyl)
            return ret["rgb_map"], ret["acc_map"]
        # the caller names the rays the loss can reach: the others take the plain kernel, the live ones the dumping
        # kernel, and their fine-pass dumps are kept for the backward while they fit the budget (no recompute)
        n, dev = ray_batch.shape[0], ray_batch.device
        rgb, acc = torch.empty((n, 3), device=dev), torch.empty((n,), device=dev)
        idx_dead = (~live).nonzero().squeeze(-1)
        idx_live = live.nonzero().squeeze(-1)
        if idx_dead.numel():
            ret = eng.render(ray_batch.index_select(0, idx_dead), skts, cyl, nanfill_chunk=0, precision="bf16", return_alpha=False)
            rgb[idx_dead], acc[idx_dead] = ret["rgb_map"], ret["acc_map"]
        kept, used = [], 0
        for i in range(0, idx_live.numel(), ctx.chunk):
            idx = idx_live[i:i + ctx.chunk]
            rb = ray_batch.index_select(0, idx)
            ret, acts = eng.render_train(rb, skts, cyl, nanfill_chunk=0, dump_coarse=False)
            rgb[idx], acc[idx] = ret["rgb_map"], ret["acc_map"]
            nbytes = acts["f"].numel() * 2
            if used + nbytes <= dump_budget:
                kept.append((idx, rb, acts["f"], ret["raw"], ret["z_fine"]))
                used += nbytes
            else:
                kept.append((idx, rb, None, None, None))
        ctx.kept = kept
        ctx.save_for_backward(ray_batch, skts, cyl)
        return rgb, acc

    @staticmethod
    def backward(ctx, g_rgb, g_acc):
        rb_all, sk, cy = ctx.saved_tensors
        rc, eng = ctx.rc, ctx.eng
        dev = rb_all.device
        n = rb_all.shape[0]
        g_rgb = torch.zeros((n, 3), device=dev) if g_rgb is None else g_rgb.float()
        g_acc = torch.zeros((n,), device=dev) if g_acc is None else g_acc.float()
        if ctx.kept is None:
            live = ((g_rgb.abs().sum(-1) + g_acc.abs()) > 0).nonzero().squeeze(-1)
            work = [(live[i:i + ctx.chunk], None, None, None, None) for i in range(0, live.numel(), ctx.chunk)]
        else:
            work, ctx.kept = ctx.kept, None
        d_skts = torch.zeros((24, 4, 4), dtype=torch.float32, device=dev)
        pd = dict(rc.network_fine.named_parameters())
        while work:
            idx, rb, acts_f, raw, z = work.pop(0)                      # popped: a chunk's dump is freed as soon as it is used
            if rb is None:
                rb = rb_all.index_select(0, idx)
            if acts_f is None:                                           # recompute this chunk with the fine-pass dump
                ret, acts = eng.render_train(rb, sk, cy, nanfill_chunk=0, dump_coarse=False)
                acts_f, raw, z = acts["f"], ret["raw"], ret["z_fine"]
            d_raw = eng.composite_backward(rb, sk, cy, raw, z, g_rgb.index_select(0, idx).contiguous(),
                                           g_acc.index_select(0, idx).contiguous())
            gd = mlp_backward(pd, None, acts_f, d_raw.reshape(-1, 4), eng.mlp_delta, want_input_grad=True,
                              want_weight_grad=False, chain=eng.mlp_delta_chain if train.USE_DELTA_CHAIN else None)
            d = eng.encode_backward_bf16(rb, sk, cy, z, gd["_g_xp"], gd["_g_d"])
            d_skts += d.sum(0)
            del acts_f, gd, d, raw, z
        return None, None, d_skts, None, None, None, None


def render_frame(rc, ray_batch: torch.Tensor, skts: torch.Tensor, cyl: torch.Tensor, chunk: int = 16384,
                 live: torch.Tensor | None = None, dump_budget_bytes: int = 48 << 30) -> Tuple[torch.Tensor, torch.Tensor]:
    """(rgb_map [n,3], acc_map [n]) of all rays of one frame, differentiable w.r.t. skts [24,4,4] (one pose).

    live (optional bool [n]): a promise that the loss only reads the rays marked True (e.g. those inside the HMR crop);
    the others are rendered as constants.  With it the live rays go through the dumping kernel in the forward and
    their fine-pass dumps (369 KB per ray) are kept for the backward while they fit dump_budget_bytes, which removes
    the recompute; without it the backward finds the rays with a non-zero upstream gradient and re-renders them."""
    if skts.shape != (24, 4, 4):
        raise ValueError("render_frame renders one pose: skts must be [24,4,4]")
    return _FrameRenderFn.apply(rc, ray_batch.float().contiguous(), skts.float(), cyl.float().contiguous(), chunk, live,
                                int(dump_budget_bytes))


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
                       focal: float = 1000.0, ext_scale: float = 0.001, chunk: int = 16384, bg: float = 1.0,
                       live_crop=None, dump_budget_bytes: int = 48 << 30):
    """bones [B,24,3] axis-angle (requires grad) -> frames [B,H,W,3] with autograd back to `bones`.

    FK with autograd (`fk.smpl_skts`), bbox from the detached key points (`kp_to_valid_rays`), rays on the device.
    live_crop = (x0, y0, x1, y1): only pixels inside it will be read by the loss (the HMR crop, run_gan.py:2059):
    `render_frame(live=...)`."""
    from . import fk
    dev = bones.device
    eng = rc.engine(dev)
    skts, kps = fk.smpl_skts(bones.double(), torch.as_tensor(rest_pose, dtype=torch.float64, device=dev))
    skts = skts.float()
    kps_np = kps.detach().cpu().numpy()
    frames = []
    for b in range(bones.shape[0]):
        cyl_np = syn.bounding_cylinder(kps_np[b], ext_scale=ext_scale)
        tl, br = syn.cylinder_bbox_2d(cyl_np, H, W, focal, c2w)
        x0, y0, x1, y1 = int(tl[0]), int(tl[1]), int(br[0]), int(br[1])
        rb = eng.generate_rays(H, W, float(focal), c2w, x0, y0, x1, y1)
        cyl = torch.as_tensor(cyl_np, dtype=torch.float32, device=dev)
        live = None
        if live_crop is not None:
            cx0, cy0, cx1, cy1 = live_crop
            ys = torch.arange(y0, y1, device=dev)[:, None]
            xs = torch.arange(x0, x1, device=dev)[None, :]
            live = ((ys >= cy0) & (ys < cy1) & (xs >= cx0) & (xs < cx1)).reshape(-1)
        rgb, acc = render_frame(rc, rb, skts[b], cyl, chunk, live=live, dump_budget_bytes=dump_budget_bytes // max(1, bones.shape[0]))
        frames.append(compose_white(rgb, acc, H, W, x0, y0, x1, y1, bg))
    return torch.stack(frames), kps
