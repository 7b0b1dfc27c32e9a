"""GPU bring-up / parity report (run on the B200 box):

    python tests/gpu_check.py <step> [<step> ...]     steps: probe stages mlp render time

Prints error metrics of every stage against the oracle / golden fixtures and appends them
to gpurun_out/gpu_check.jsonl.  Each step is independent so a fault in one does not hide
the others (run them as separate processes under `timeout`).
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import render_oracle as orc            # noqa: E402  (checker only)
from posegen_b200 import synthetic as syn          # noqa: E402
from posegen_b200.engine import Engine             # noqa: E402
from tests import parity_util as pu                # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def log(step, **kv):
    rec = {"step": step, **kv}
    print(json.dumps(rec), flush=True)
    with open(os.path.join(OUT, "gpu_check.jsonl"), "a") as f:
        f.write(json.dumps(rec) + "\n")


def step_probe(eng):
    torch.manual_seed(0)
    for (K, N) in [(16, 256), (64, 256), (256, 256), (144, 128), (432, 256)]:
        A = torch.randn(128, K, device="cuda")
        B = torch.randn(N, K, device="cuda")
        ref = A.bfloat16().float() @ B.bfloat16().float().t()
        for variant in (0, 1):
            D = eng.debug_umma_gemm(A, B, variant)
            torch.cuda.synchronize()
            try:
                eng.check_status()
                st = "ok"
            except Exception as e:  # noqa: BLE001
                st = str(e)
            log("probe", K=K, N=N, variant=variant, max_err=float((D - ref).abs().max()), ref_max=float(ref.abs().max()), status=st)


def _case(name):
    g = pu.load_golden(name)
    frame, ckpt, rb, cyl = pu.case_from_golden(g)
    return g, frame, ckpt, rb, cyl


def step_stages(eng):
    g, frame, ckpt, rb, cyl = _case("a_32_boost_taps")
    eng.load_checkpoint(ckpt)
    dev = eng.device
    rbt, sk, cy = torch.as_tensor(rb, device=dev), torch.as_tensor(frame.pose.skts, device=dev), torch.as_tensor(cyl, device=dev)
    nf = eng.near_far(rbt, sk, cy, nanfill_chunk=4096).cpu().numpy()
    log("near_far", max_err_near=pu.max_abs(nf[:, 0], g["near"][:, 0]), max_err_far=pu.max_abs(nf[:, 1], g["far"][:, 0]),
        bit_exact=bool(np.array_equal(nf[:, 0], g["near"][:, 0]) and np.array_equal(nf[:, 1], g["far"][:, 0])))
    ge, fe, ce, rbe, cyle = _case("e_32_nanfill")
    nfe = eng.near_far(torch.as_tensor(rbe, device=dev), torch.as_tensor(fe.pose.skts, device=dev),
                       torch.as_tensor(cyle, device=dev), nanfill_chunk=4096).cpu().numpy()
    log("near_far_nanfill", max_err_near=pu.max_abs(nfe[:, 0], ge["near"][:, 0]), max_err_far=pu.max_abs(nfe[:, 1], ge["far"][:, 0]),
        n_nan=int(np.isnan(nfe).sum()))
    # encode vs golden (2 rays x 64 coarse samples) and vs oracle on 32 rays
    taps = {}
    ref = pu.oracle_render(rb, frame.pose.skts, cyl, ckpt, taps=taps)
    zc = taps["z_coarse"].to(dev)
    enc = eng.encode(rbt[:32], sk, cy, zc[:32].contiguous())
    log("encode", max_err_vs_golden=pu.max_abs(enc[:2].cpu().numpy(), g["enc_coarse_head"]),
        max_err_vs_oracle=pu.max_abs(enc.cpu().numpy(), taps["enc_coarse"][:32].numpy()))
    # composite on the oracle's raw
    comp = eng.composite(rbt, sk, cy, taps["raw_coarse"].to(dev), zc)
    log("composite", **{k: pu.max_abs(comp[k2].cpu().numpy(), v) for k, k2, v in [
        ("rgb0", "rgb_map", ref["rgb0"]), ("acc0", "acc_map", ref["acc0"]), ("disp0", "disp_map", ref["disp0"]),
        ("alpha0", "alpha", ref["alpha0"]), ("weights0", "weights", taps["weights_coarse"].numpy())]})
    compf = eng.composite(rbt, sk, cy, taps["raw_fine"].to(dev), taps["z_fine"].to(dev))
    log("composite_fine", rgb=pu.max_abs(compf["rgb_map"].cpu().numpy(), ref["rgb_map"]),
        acc=pu.max_abs(compf["acc_map"].cpu().numpy(), ref["acc_map"]), alpha=pu.max_abs(compf["alpha"].cpu().numpy(), ref["alpha"]))
    # sample_pdf on the golden weights
    sp = eng.sample_pdf(zc, torch.as_tensor(g["weights_coarse"], device=dev))
    inds = sp["pdf_inds"].cpu().numpy()
    gi = g["pdf_inds"].astype(np.int32)
    log("sample_pdf", mismatch_cols_0_14=int((inds[:, :15] != gi[:, :15]).sum()), mismatch_col_15=int((inds[:, 15] != gi[:, 15]).sum()),
        n=int(inds.shape[0]), z_samples_max_err=pu.max_abs(sp["z_samples"].cpu().numpy(), g["z_samples"]),
        z_sorted_max_err=pu.max_abs(sp["z_sorted"].cpu().numpy(), taps["z_fine"].numpy()),
        sorted=bool((np.diff(sp["z_sorted"].cpu().numpy(), axis=1) >= 0).all()))


def step_mlp(eng, precisions=("fp32", "bf16")):
    g, frame, ckpt, rb, cyl = _case("a_32_boost_taps")
    eng.load_checkpoint(ckpt)
    dev = eng.device
    taps = {}
    pu.oracle_render(rb[:16], frame.pose.skts, cyl, ckpt, taps=taps)
    enc = taps["enc_coarse"].reshape(-1, 1080)[:700].contiguous()
    nets = orc.nets_from_ckpt(ckpt)
    for net_id in (0, 1):
        ref = orc.nerf_forward(enc, nets[net_id]).numpy()
        for prec in precisions:
            t0 = time.time()
            raw = eng.mlp(net_id, enc.to(dev), prec)
            torch.cuda.synchronize()
            try:
                eng.check_status()
                st = "ok"
            except Exception as e:  # noqa: BLE001
                st = str(e)
            raw = raw.cpu().numpy()
            log("mlp", net=net_id, precision=prec, rgb_raw_err=pu.max_abs(raw[:, :3], ref[:, :3]),
                sigma_raw_err=pu.max_abs(raw[:, 3], ref[:, 3]), ref_sigma_absmax=float(np.abs(ref[:, 3]).max()),
                ref_rgb_absmax=float(np.abs(ref[:, :3]).max()), status=st, secs=round(time.time() - t0, 3))


def step_render(eng, precisions=("fp32", "bf16")):
    for name in ["a_32_boost_taps", "e_32_nanfill", "b_64_boost", "c_64_plain", "d_64_calibrated"]:
        g, frame, ckpt, rb, cyl = _case(name)
        for prec in precisions:
            try:
                out = pu.gpu_render(eng, rb, frame.pose.skts, cyl, ckpt, prec, chunk=int(g["meta_chunk"]), taps=True)
            except Exception as e:  # noqa: BLE001
                log("render", case=name, precision=prec, error=str(e))
                continue
            rec = {k: pu.max_abs(out[k], g[k]) for k in pu.IMAGE_KEYS}
            rec["psnr_rgb"] = pu.psnr(out["rgb_map"], g["rgb_map"])
            rec["psnr_acc"] = pu.psnr(out["acc_map"], g["acc_map"])
            if "alpha" in g:
                rec["alpha"] = pu.max_abs(out["alpha"], g["alpha"])
                rec["alpha0"] = pu.max_abs(out["alpha0"], g["alpha0"])
                gi = g["pdf_inds"].astype(np.int32)
                rec["pdf_mismatch_0_14"] = int((out["pdf_inds"][:, :15] != gi[:, :15]).sum())
                rec["pdf_mismatch_15"] = int((out["pdf_inds"][:, 15] != gi[:, 15]).sum())
                rec["z_samples"] = pu.max_abs(out["z_samples"], g["z_samples"])
                rec["raw0_head"] = pu.max_abs(out["raw0"][:64], g["raw_coarse_head"])
                rec["raw_head"] = pu.max_abs(out["raw"][:64], g["raw_fine_head"])
            log("render", case=name, precision=prec, n_rays=int(rb.shape[0]), acc_mean=float(g["acc_map"].mean()), **rec)


def step_time(eng, precisions=("bf16", "fp32")):
    ckpt = syn.synthetic_raycaster_state(0, alpha_gain=400.)
    eng.load_checkpoint(ckpt)
    dev = eng.device
    for res in (128, 512):
        frame = syn.synthetic_frame(5, res, res)
        rb = torch.as_tensor(syn.ray_batch(frame.rays_o, frame.rays_d), device=dev)
        sk, cy = torch.as_tensor(frame.pose.skts, device=dev), torch.as_tensor(frame.pose.cyl, device=dev)
        for prec in precisions:
            if prec == "fp32" and res > 128:
                continue
            for it in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.render(rb, sk, cy, nanfill_chunk=4096, precision=prec, return_alpha=False)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
            eng.check_status()
            n = rb.shape[0]
            log("time", res=res, precision=prec, n_rays=n, ms=round(ms, 3), rays_per_s=round(n / ms * 1e3),
                tflops=round(n * 248205312 / ms * 1e3 / 1e12, 2))


def step_phases(eng, precisions=("bf16",)):
    ckpt = syn.synthetic_raycaster_state(0, alpha_gain=400.)
    eng.load_checkpoint(ckpt)
    dev = eng.device
    frame = syn.synthetic_frame(5, 512, 512)
    rb = torch.as_tensor(syn.ray_batch(frame.rays_o, frame.rays_d), device=dev)
    sk, cy = torch.as_tensor(frame.pose.skts, device=dev), torch.as_tensor(frame.pose.cyl, device=dev)
    eng.phase_timers(True)
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.render(rb, sk, cy, nanfill_chunk=4096, precision="bf16", return_alpha=False)
        e1.record()
        torch.cuda.synchronize()
    t = eng.phase_timers(False, read=True)
    n = rb.shape[0]
    tiles_per_slot = (n / 8) * 9 / 148 / 2          # timers cover pipeline slot 0 of every CTA
    log("phases", ms=round(e0.elapsed_time(e1), 3), n_rays=n, tiles_per_slot=round(tiles_per_slot, 1),
        **{k: round(v / tiles_per_slot) for k, v in t.items()})


STEPS = {"phases": step_phases, "probe": step_probe, "stages": step_stages, "mlp": step_mlp, "render": step_render, "time": step_time}

if __name__ == "__main__":
    eng = Engine()
    args = sys.argv[1:] or list(STEPS)
    for s in args:
        if ":" in s:
            name, precs = s.split(":")
            STEPS[name](eng, tuple(precs.split(",")))
        else:
            STEPS[s](eng)
