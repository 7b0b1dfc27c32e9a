#!/usr/bin/env python
"""PoseGen generation loop (BASELINE.json configs[2] / SURVEY.md §8d config 3): a batch of synthetic poses
(axis-angle bones, seeds 0..P-1) -> device FK (`pgn_pose_to_skts`) -> cylinder bbox -> device rays -> fused render ->
white-background frame -> HMR input (crop / resize 224 / normalise), sharded by image over the GPUs
(`pose_idx % world_size`), weights replicated, one final all_gather of the frames.  This is the reference's
`run_render` + the image hand-off of `train_gan` (run_gan.py:2299-2347, 2057-2071) without the PNG round trip.
Time = max over ranks (CUDA events), gather included.  Prints one JSON line on rank 0.

    python tools/generation_loop_bench.py [--poses 256]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/generation_loop_bench.py --poses 256
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from posegen_b200 import dist as pdist, synthetic as syn                         # noqa: E402
from posegen_b200.raycaster import raycaster_from_checkpoint                     # noqa: E402
from posegen_b200.render import render_pose_batch                                # noqa: E402


def run(rank, world, local, poses=256, res=512):
    """One leg (process group already initialised for world > 1): returns the result dict on every rank."""
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    rc = raycaster_from_checkpoint(syn.synthetic_raycaster_state(0, alpha_gain=400.), device=dev, precision="bf16")
    rc.eval()
    eng = rc.engine(dev)
    mine = pdist.shard_indices(poses, rank, world)
    bones = np.stack([syn.synthetic_pose(s).bones for s in mine]).astype(np.float32)
    rest = (syn.SMPL_REST_POSE * syn.BODY_SCALE).astype(np.float32)
    c2w = syn.run_gan_c2w()
    focal = 1000.0 * res / 512
    crop = tuple(int(round(v * res / 512)) for v in (100, 100, 412, 412))

    def go(bones_np):
        b = torch.as_tensor(bones_np, device=dev)
        rgbs, hmr, n_rays = render_pose_batch(rc, b, rest, c2w, (res, res, focal), crop=crop)
        return rgbs, hmr, n_rays

    go(bones[:2])                                                      # warm-up
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rgbs, hmr, n_rays = go(bones)
    frames_u8 = (rgbs.clamp(0, 1) * 255).to(torch.uint8)               # what the reference writes to PNG (run_nerf.py to8b)
    all_frames = pdist.gather_frames(frames_u8, poses, rank, world)
    e1.record()
    torch.cuda.synchronize()
    eng.check_status()
    ms = pdist.max_over_ranks(e0.elapsed_time(e1), dev)
    tot = torch.tensor([float(n_rays)], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(tot)
    assert all_frames.shape[0] == poses
    # the same kernel back to back on one resident pose, no driver work: what the loop is compared with (the headline
    # bench line pauses between steps to flush L2, so its clocks sit higher than a seconds-long loop at the power cap)
    f0 = syn.synthetic_frame(int(mine[0]) if len(mine) else 0, res, res)
    rb0 = torch.as_tensor(syn.ray_batch(f0.rays_o, f0.rays_d), device=dev)
    sk0, cy0 = torch.as_tensor(f0.pose.skts, device=dev), torch.as_tensor(f0.pose.cyl, device=dev)
    reps = 24
    for _ in range(2):
        eng.render(rb0, sk0, cy0, nanfill_chunk=4096, return_alpha=False)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(reps):
        eng.render(rb0, sk0, cy0, nanfill_chunk=4096, return_alpha=False)
    k1.record()
    torch.cuda.synchronize()
    kernel_rps = rb0.shape[0] * reps / (k0.elapsed_time(k1) * 1e-3)
    loop_rps_rank = n_rays / (ms * 1e-3)
    return {"metric": "frames_per_sec_512", "value": poses / ms * 1e3, "unit": "frames/s", "n_gpus": world,
            "poses": poses, "seconds": ms * 1e-3, "rays_per_sec": float(tot[0]) / ms * 1e3,
            "kernel_only_sustained_rays_per_sec_rank0": kernel_rps, "loop_over_kernel_only_rank0": loop_rps_rank / kernel_rps,
            "hmr_inputs": list(hmr.shape), "finite": bool(torch.isfinite(hmr).all()), "gpu_launches_rank0": int(eng.launch_count - l0),
            "config": f"{poses} synthetic poses x {res}x{res} bbox renders sharded pose_idx % {world}; device FK + cylinder + bbox, "
                      "device rays, ONE fused bf16 render launch per pose batch (pose_idx form), white-bg frames, HMR input 224; "
                      "one final all_gather of the uint8 frames"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--poses", type=int, default=256)
    ap.add_argument("--res", type=int, default=512)
    a = ap.parse_args()
    rank, world, local = pdist.env_rank_world()
    pdist.init_process_group("nccl" if world > 1 else None)
    line = run(rank, world, local, a.poses, a.res)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
