#!/usr/bin/env python
"""posegen_b200 benchmark (driver contract: prints ONE JSON line on rank 0).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA library)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: CPU restatement of the
                                                             # reference's PyTorch path on the host cores

Workload (BASELINE.json configs[1]; configs[2] for N>1): one "step" = one synthetic SMPL pose
rendered at 512x512, coarse+fine (64+16 samples, two 8x256 A-NeRF MLPs), rays restricted to the
cylinder bbox exactly like the reference's kp_to_valid_rays (run_gan camera, focal 1000).
N>1: every rank renders its own pose per step (images shard across GPUs, weights replicated,
no data-path collective); the frames of all K steps are exchanged by ONE final all_gather, which is
timed (CUDA events) and added to the rank's step time (BASELINE.json configs[2]: "no communication
except a final gather").
Metric: rays/s over all ranks (max-over-ranks device time, CUDA events).

The same invocation also measures the other three BASELINE.json configurations and attaches them to the JSON line as
`aux` (each with its own warm-up, CUDA-event timing and max over ranks; `--aux none` skips them):
  aux.genloop_256  configs[2]: 256 synthetic poses -> device FK -> bbox -> fused render -> frame -> HMR input, poses
                   sharded pose_idx % N, one final all_gather (tools/generation_loop_bench.py)
  aux.train_step   configs[3]: 3072-ray training step per GPU, forward + backward + NCCL all-reduce + Adam, CUDA graph;
                   the all-reduce is also timed alone (tools/train_step_bench.py)
  aux.gan_step     configs[4]: generator -> device FK -> differentiable 512x512 render -> HMR stand-in -> MPJPE ->
                   backward to the generator, all-reduce of the generator gradients (tools/gan_step_bench.py)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_RAY = 248_205_312          # 144 samples x 1,723,648 FLOP (un-padded nn.Linear MACs x 2)
N_POSES = 5                          # distinct synthetic poses cycled through the steps (rank r starts at pose r:
                                     # every rank renders the same pool, different images at any one time)


NCU_CAPTURE = "r2_bf16_render_512.json"        # ncu --set full capture of the shipped render kernel (profiles/)


def ncu_traffic():
    """DRAM bytes of one launch of the dominant kernel from the committed ncu capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", NCU_CAPTURE)
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "r1_bf16_render_512.json")
    try:
        prof = json.load(open(path))
        m = prof["kernels"][0]["metrics"]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd, wr = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
        return float(rd["value"]) * scale[rd["unit"]] + float(wr["value"]) * scale[wr["unit"]]
    except Exception:  # noqa: BLE001
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_sustained": p["bf16_tflops_sustained"], "bf16_burst": p["bf16_tflops"], "hbm": p["hbm_gbs"], "src": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]),
                "power_w_max": max(float(s[2]) for s in self.samples), "samples": len(sm), "reasons": reasons}


def make_jobs(res, rank, world):
    from posegen_b200 import synthetic as syn
    jobs = []
    for i in range(N_POSES):
        frame = syn.synthetic_frame(100 + (rank + i) % N_POSES, res, res)
        jobs.append((frame, syn.ray_batch(frame.rays_o, frame.rays_d)))
    return jobs


# ------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's CPU path (restated op-for-op in oracle/render_oracle.py; the reference itself is
    Python and cannot travel to the GPU box) on all host threads, on a bounded sample of the workload."""
    if rank != 0:
        return
    from oracle import render_oracle as orc
    from posegen_b200 import synthetic as syn
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    ckpt = syn.synthetic_raycaster_state(0, alpha_gain=400.)
    nets, emb = orc.nets_from_ckpt(ckpt), orc.embed_params_from_ckpt(ckpt)
    # the SAME pose pool as our arm (step i renders pose i % N_POSES), each step a bounded strided sample of that frame
    jobs = make_jobs(args.res, 0, 1)
    samples = []
    for frame, rb in jobs:
        n_sample = min(args.cpu_rays, rb.shape[0])
        sel = np.linspace(0, rb.shape[0] - 1, n_sample).astype(np.int64)
        samples.append((torch.from_numpy(rb[sel]), torch.from_numpy(frame.pose.skts), torch.from_numpy(frame.pose.cyl)))
    sec_total, rays_total, frame_rays = 0.0, 0, []
    for it in range(args.warmup + args.steps):
        rbt, sk, cy = samples[it % N_POSES]
        t0 = time.perf_counter()
        orc.render(rbt, sk, cy, nets, emb, chunk=4096)
        if it >= args.warmup:
            sec_total += time.perf_counter() - t0
            rays_total += rbt.shape[0]
            frame_rays.append(jobs[it % N_POSES][1].shape[0])
    value = rays_total / sec_total
    rays_per_frame = float(np.mean(frame_rays)) if frame_rays else float(jobs[0][1].shape[0])
    line = {"impl": "reference", "metric": "rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_total / max(args.steps, 1) * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, float(np.mean([rb.shape[0] for _, rb in jobs]))),
            "cpu_baseline": {"value": value, "unit": "rays/s", "cores": threads, "kind": "port",
                             "sample": f"per step {args.cpu_rays} rays evenly strided from the bbox of one {args.res}x{args.res} frame of the "
                                       f"{N_POSES}-pose pool, chunk 4096"},
            "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "frames_per_sec_512": value / rays_per_frame}
    print(json.dumps(line), flush=True)


def workload_config(args, rays_per_frame):
    return {"workload": f"A-NeRF surreal.txt render, 1 synthetic SMPL pose per step at {args.res}x{args.res}, coarse+fine "
                        "(64+16 samples, 2x 8x256 MLP), cylinder-bbox rays, white_bkgd, random-init weights (x400 alpha head)",
            "rays_per_frame": int(round(rays_per_frame)), "res": args.res,
            "weights": "x400 alpha head: bf16 parity on these weights is held to the PSNR bound only (SURVEY.md §8d; tests/test_gpu_render.py)",
            "poses_cycled": N_POSES, "l2": "256 MiB buffer written between timed steps (L2 flush)",
            "parallelism": f"dp{args.gpus} (images sharded by rank, no data-path collective; one final all_gather of the frames)"}


# ------------------------------------------------------------------------------ our arm
def run_ours(args, rank, world, local):
    from posegen_b200 import dist as pdist, synthetic as syn
    from posegen_b200.raycaster import raycaster_from_checkpoint
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ckpt = syn.synthetic_raycaster_state(0, alpha_gain=400.)
    rc = raycaster_from_checkpoint(ckpt, device=dev, precision=args.precision)
    rc.return_alpha = False                    # alpha/alpha0 are unused by render_path and the trainer
    eng = rc.engine(dev)
    jobs = make_jobs(args.res, rank, world)
    # device-resident inputs for the kernel-throughput leg, pinned host copies for the end-to-end leg
    dev_in = [(torch.as_tensor(rb, device=dev), torch.as_tensor(f.pose.skts, device=dev), torch.as_tensor(f.pose.cyl, device=dev))
              for f, rb in jobs]
    host_in = [(torch.from_numpy(rb).pin_memory(), torch.from_numpy(f.pose.skts).pin_memory(),
                torch.from_numpy(f.pose.cyl).pin_memory()) for f, rb in jobs]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n_max = max(rb.shape[0] for _, rb in jobs)
    stash = torch.zeros((args.steps, n_max, 3), device=dev) if world > 1 else None            # this rank's finished frames
    gather_buf = torch.empty((world, args.steps, n_max, 3), device=dev) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step(i):
        rb, sk, cy = dev_in[i % N_POSES]
        ret = eng.render(rb, sk, cy, nanfill_chunk=4096, precision=args.precision, return_alpha=False)
        if world > 1 and i >= args.warmup:
            stash[i - args.warmup, :rb.shape[0]] = ret["rgb_map"]       # rows beyond this pose's ray count stay zero padding
        return rb.shape[0]

    host_out = torch.empty((max(h[0].shape[0] for h in host_in), 5), dtype=torch.float32).pin_memory()

    def e2e_step(i):
        rb_h, sk_h, cy_h = host_in[i % N_POSES]
        n = rb_h.shape[0]
        rb = rb_h.to(dev, non_blocking=True)
        sk = sk_h.to(dev, non_blocking=True)
        cy = cy_h.to(dev, non_blocking=True)
        # the call a PoseGen user makes: RayCaster.forward with the reference's kwargs (core/trainer.py:74)
        ret = rc(rb, N_samples=64, N_importance=16, kp_batch=None, skts=sk, cyls=cy, bones=None, cams=None,
                 subject_idxs=None, perturb=False, raw_noise_std=0., nanfill_chunk=4096)
        out = torch.cat([ret["rgb_map"], ret["acc_map"][:, None], ret["disp_map"][:, None]], 1)
        host = host_out[:n]
        host.copy_(out, non_blocking=True)                # device->host read of the step's result into pinned memory (on the
        return n, rb_h.numel() * 4 + sk_h.numel() * 4 + cy_h.numel() * 4, host.numel() * 4      # timed stream: inside e0..e1)

    step_ms = []

    def timed(step_fn, steps, warmup):
        for i in range(warmup):
            step_fn(i)
            flush.fill_(1)
        barrier()
        total_ms, rays = 0.0, 0
        extras = None
        step_ms.clear()
        for i in range(steps):
            flush.fill_(i & 0xFF)                      # L2 flush between timed iterations (not timed)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            r = step_fn(warmup + i)
            e1.record()
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
            step_ms.append(e0.elapsed_time(e1))
            if isinstance(r, tuple):
                rays += r[0]
                extras = r[1:]
            else:
                rays += r
        barrier()
        return total_ms, rays, extras

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count
    ms, rays, _ = timed(resident_step, args.steps, args.warmup)
    resident_step_ms = list(step_ms)
    launches = eng.launch_count - launches0 - 2 * args.warmup          # 2 kernels per render call
    if world > 1:                                                      # the final gather of every rank's frames
        dist.all_gather_into_tensor(gather_buf.view(-1, 3), stash.view(-1, 3))      # warm-up (NCCL channel setup)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather_into_tensor(gather_buf.view(-1, 3), stash.view(-1, 3))
        g1.record()
        torch.cuda.synchronize()
        ms += g0.elapsed_time(g1)
    eng.check_status()
    ms_e2e, rays_e2e, io = timed(e2e_step, args.steps, min(args.warmup, 3))
    if rank == 0:
        sampler.stop_flag.set()
        sampler.join(timeout=2)
    eng.check_status()

    ms_max = pdist.max_over_ranks(ms, dev)
    ms_e2e_max = pdist.max_over_ranks(ms_e2e, dev)
    tot = torch.tensor([rays, rays_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    total_rays, total_rays_e2e = float(tot[0]), float(tot[1])
    if rank != 0:
        return None
    value = total_rays / (ms_max * 1e-3)
    e2e_value = total_rays_e2e / (ms_e2e_max * 1e-3)
    peaks = measured_peaks()
    # dominant kernel = the fused render kernel; the near/far pre-pass is <0.1 % of the step
    per_gpu_tflops = (total_rays / world) * FLOP_PER_RAY / (ms_max * 1e-3) / 1e12
    rays_per_frame = total_rays / (args.steps * world)
    line = {
        "metric": "rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": workload_config(args, float(np.mean([rb.shape[0] for _, rb in jobs]))),      # mean of the pose pool, both arms
        "frames_per_sec_512": value / rays_per_frame,
        "step_ms_min_max": [round(min(resident_step_ms), 3), round(max(resident_step_ms), 3)],
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": int(io[0]), "d2h_bytes_per_step": int(io[1]),
                "frames_per_sec_512": e2e_value / rays_per_frame,
                "api": "posegen_b200.RayCaster.forward (pinned host ray_batch/skts/cyls -> device, result rows -> host)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": per_gpu_tflops, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": per_gpu_tflops / peaks["bf16_sustained"],
                     "traffic": ncu_traffic() if args.precision == "bf16" else None,
                     "traffic_note": f"DRAM bytes read+written by one launch (239,148-ray frame), ncu --set full, profiles/{NCU_CAPTURE}; "
                                     "algorithmic bytes per launch = 84 B/ray + 2 x 1.83 MB weights = 23.7 MB (HBM is not the bound)",
                     "peak_source": f"{peaks['src']} bf16 sustained (40-55 ms launches back to back inside a seconds-long step loop at the power cap); burst {peaks['bf16_burst']}",
                     "flop_per_ray": FLOP_PER_RAY, "kernel": "pgn_render_bf16_kernel" if args.precision == "bf16" else "pgn_render_fp32_kernel",
                     "note": "algorithmic FLOPs (reference nn.Linear MACs x2); whole-step device time (near/far pre-pass included)"},
        "clocks": sampler.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, jobs[0])
    return line


def run_aux(args, rank, world, local):
    """The other BASELINE.json configurations (configs[2..4]) as auxiliary legs; every rank runs them (they shard /
    all-reduce over the same process group), the dicts are identical on all ranks."""
    legs = [x for x in args.aux.split(",") if x and x != "none"]
    aux = {}
    peaks = measured_peaks()
    for leg in legs:
        try:
            if leg == "train":
                from tools import train_step_bench
                r = train_step_bench.run(rank, world, local, steps=20, warmup=3, deterministic=False, graph=True)
                r["roofline"] = {"bound": "tensor", "achieved": r["tflops_per_gpu"], "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                                 "frac": r["tflops_per_gpu"] / peaks["bf16_sustained"],
                                 "note": "algorithmic 3 x 248.2 MFLOP per ray (forward + dX + dW), whole step incl. all-reduce and Adam"}
                aux["train_step"] = r
            elif leg == "gan":
                from tools import gan_step_bench
                aux["gan_step"] = gan_step_bench.run(rank, world, local, steps=3, warmup=1, poses_per_gpu=args.gan_poses_per_gpu)
            elif leg == "genloop":
                from tools import generation_loop_bench
                aux[f"genloop_{args.genloop_poses}"] = generation_loop_bench.run(rank, world, local, poses=args.genloop_poses, res=args.res)
            else:
                aux[leg] = {"error": "unknown leg"}
        except Exception as e:  # noqa: BLE001  (a failing leg must not take the headline line down; it is reported as failed)
            import traceback
            aux[{"train": "train_step", "gan": "gan_step", "genloop": f"genloop_{args.genloop_poses}"}.get(leg, leg)] = \
                {"error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc().splitlines()[-3:]}
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
    return aux


def cpu_baseline(args, job):
    """Oracle port of the reference's CPU path on the host cores, bounded sample of the same frame."""
    from oracle import render_oracle as orc
    from posegen_b200 import synthetic as syn
    frame, rb = job
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    ckpt = syn.synthetic_raycaster_state(0, alpha_gain=400.)
    nets, emb = orc.nets_from_ckpt(ckpt), orc.embed_params_from_ckpt(ckpt)
    n_sample = min(3 * args.cpu_rays, rb.shape[0])                 # ~10 s of host work (one measurement, not K steps)
    sel = np.linspace(0, rb.shape[0] - 1, n_sample).astype(np.int64)
    rbt, sk, cy = torch.from_numpy(rb[sel]), torch.from_numpy(frame.pose.skts), torch.from_numpy(frame.pose.cyl)
    orc.render(rbt[:512], sk, cy, nets, emb, chunk=4096)          # warm-up
    t0 = time.perf_counter()
    orc.render(rbt, sk, cy, nets, emb, chunk=4096)
    sec = time.perf_counter() - t0
    return {"value": n_sample / sec, "unit": "rays/s", "cores": threads, "kind": "port",
            "sample": f"{n_sample} rays evenly strided from the {rb.shape[0]}-ray bbox of one {args.res}x{args.res} frame ({sec:.1f} s)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-rays", type=int, default=8192, help="rays in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--aux", default="genloop,train,gan", help="auxiliary legs (BASELINE.json configs[2..4]): comma list of "
                    "genloop,train,gan or 'none'")
    ap.add_argument("--genloop-poses", type=int, default=256)
    ap.add_argument("--gan-poses-per-gpu", type=int, default=2)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from posegen_b200 import dist as pdist
    rank, world, local = pdist.env_rank_world()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; posegen_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    pdist.init_process_group("nccl" if world > 1 else None)
    try:
        line = run_ours(args, rank, world, local)
        aux = run_aux(args, rank, world, local)
        if rank == 0:
            line["aux"] = aux
            print(json.dumps(line), flush=True)
    finally:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
