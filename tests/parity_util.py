"""Shared helpers of the parity tests: golden fixtures, synthetic cases, error metrics."""
from __future__ import annotations

import hashlib
import os

import numpy as np
import torch

from oracle import render_oracle as orc
from posegen_b200 import synthetic as syn

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
IMAGE_KEYS = ("rgb_map", "acc_map", "disp_map", "rgb0", "acc0", "disp0")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False))


def case_from_golden(g):
    """Rebuild the exact inputs a golden fixture was generated from (frame, ckpt, cyl)."""
    frame = syn.synthetic_frame(int(g["meta_pose_seed"]), int(g["meta_res"]), int(g["meta_res"]))
    gain = float(g["meta_alpha_gain"])
    calibrated = bool(g["meta_calibrated"])
    ckpt = syn.synthetic_raycaster_state(int(g["meta_weight_seed"]), alpha_gain=None if (calibrated or gain == 0) else gain,
                                         n_framecodes=int(g["meta_n_framecodes"]) if "meta_n_framecodes" in g else 0)
    if calibrated:
        for key, raw_key in (("network_fn_state_dict", "raw_coarse"), ("network_fine_state_dict", "raw_fine")):
            ckpt[key]["alpha_linear.bias"] = np.zeros_like(ckpt[key]["alpha_linear.bias"])
            syn.calibrate_alpha_head(ckpt[key], float(g[f"sigma_far_max_{raw_key}"]))
    rb = syn.ray_batch(frame.rays_o, frame.rays_d)
    cyl = np.asarray(g["in_cyl"], dtype=np.float32)
    return frame, ckpt, rb, cyl


def oracle_render(rb, skts, cyl, ckpt, chunk=4096, dtype=torch.float32, device="cpu", taps=None, cams=None, lindisp=False):
    rbt = torch.as_tensor(rb).to(device=device, dtype=dtype)
    nets = orc.nets_from_ckpt(ckpt, dtype, device)
    emb = orc.embed_params_from_ckpt(ckpt, dtype, device)
    sk = torch.as_tensor(skts).to(device=device, dtype=dtype)
    cy = torch.as_tensor(cyl).to(device=device, dtype=dtype)
    with torch.no_grad():
        if taps is not None and rbt.shape[0] <= chunk:
            n = rbt.shape[0]
            out = orc.render_rays(rbt, sk[None].expand(n, -1, -1, -1), cy[None].expand(n, -1), nets, emb, taps=taps, cams=cams, lindisp=lindisp)
        else:
            out = orc.render(rbt, sk, cy, nets, emb, chunk=chunk, cams=cams, lindisp=lindisp)
    return {k: v.float().cpu().numpy() for k, v in out.items()}


def max_abs(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max()) if a.size else 0.0


def psnr(a, b, peak=1.0):
    mse = float(np.mean((np.asarray(a, np.float64) - np.asarray(b, np.float64)) ** 2))
    return 99.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)


def gpu_render(engine, rb, skts, cyl, ckpt, precision, chunk=4096, taps=False, cams=None, lindisp=False):
    """Render through the C ABI in batchify-sized calls (the reference's chunk semantics)."""
    dev = engine.device
    engine.load_checkpoint(ckpt)
    rbt = torch.as_tensor(rb, device=dev)
    sk = torch.as_tensor(skts, device=dev)
    cy = torch.as_tensor(cyl, device=dev)
    ret = engine.render(rbt, sk, cy, nanfill_chunk=chunk, precision=precision, return_alpha=True, taps=taps, cams=cams, lindisp=lindisp)
    torch.cuda.synchronize()
    engine.check_status()
    return {k: v.cpu().numpy() for k, v in ret.items()}
