"""GPU: the drop-in driven by the reference's OWN callers, against the reference's OWN RayCaster on the same device.

`baseline/_ref/` holds a byte-for-byte staging of the reference's `core/`, `run_nerf.py` and configs (git-ignored,
written by `oracle/stage_reference.py` in the build container; it travels to the GPU box with the snapshot).  The tests
import `run_nerf.render_path` and `core.trainer.Trainer` from there under `oracle/ref_shim.py` (third-party stubs only:
with a GPU present the reference's hard-coded `.to('cuda')` calls run unmodified) and hand them the kwargs that
`posegen_b200.create_raycaster` returns - the three-line change of INTEGRATION.md.  The same calls are made with the
reference's own `core.raycasters.create_raycaster` (eager fp32 PyTorch on the B200) and the results compared.
"""
import contextlib
import io
import os
import threading

import numpy as np
import pytest
import torch

from posegen_b200 import raycaster as rcmod, synthetic as syn
from tests import parity_util as pu

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


@pytest.fixture(scope="module")
def ref(engine, tmp_path_factory):
    if not os.path.isdir(os.path.join(REF, "core")):
        pytest.skip("baseline/_ref not staged (python oracle/stage_reference.py in the build container)")
    from oracle import ref_shim
    ref_shim.REFERENCE_ROOT = REF
    ref_shim.install()
    with contextlib.redirect_stdout(io.StringIO()):
        import run_nerf
        import core.raycasters as ref_rc
        import core.trainer as ref_trainer
        from core.utils.skeleton_utils import SMPLSkeleton, smpl_rest_pose, get_per_joint_coords
    # torch >= 2.x compatibility of the reference itself, not of the path under test: core/trainer.py:8 divides a CUDA
    # tensor by the 1-element CPU tensor torch.Tensor([10.]), which current torch rejects for ANY ray caster (the
    # reference's own included).  Same formula, divisor on the operand's device.
    ref_trainer.mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.tensor(10., device=x.device))
    tmp = str(tmp_path_factory.mktemp("ref"))
    os.makedirs(os.path.join(tmp, "x"), exist_ok=True)
    args = run_nerf.config_parser().parse_args(["--config", os.path.join(REF, "configs/surreal/surreal.txt"),
                                                "--basedir", tmp, "--expname", "x", "--no_reload"])
    data_attrs = {"skel_type": SMPLSkeleton, "near": 60., "far": 100., "n_views": 1, "hwf": (64, 64, 125.),
                  "joint_coords": get_per_joint_coords(smpl_rest_pose.astype(np.float32))}
    return {"run_nerf": run_nerf, "rc": ref_rc, "trainer": ref_trainer, "args": args, "data_attrs": data_attrs}


def _both(ref, ckpt, precision, **arg_overrides):
    """(reference kwargs, drop-in kwargs) built by the two create_raycaster factories from the SAME args namespace."""
    import copy
    args = copy.deepcopy(ref["args"])
    for k, v in arg_overrides.items():
        setattr(args, k, v)
    tck = {k: ({kk: torch.as_tensor(vv) for kk, vv in v.items()} if isinstance(v, dict) else v) for k, v in ckpt.items()}
    with contextlib.redirect_stdout(io.StringIO()):
        theirs = ref["rc"].create_raycaster(args, ref["data_attrs"])
    theirs[1]["ray_caster"].load_state_dict(tck)
    theirs[1]["ray_caster"].to("cuda")
    ours = rcmod.create_raycaster(args, ref["data_attrs"], device="cuda", precision=precision)
    ours[1]["ray_caster"].load_state_dict(tck)
    return args, theirs, ours


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 2e-2)])
def test_reference_render_path_drives_the_dropin(ref, precision, tol):
    """run_nerf.render_path (run_nerf.py:27-147), unmodified, with the drop-in's render_kwargs_test."""
    g = pu.load_golden("d_64_calibrated")
    frame, ckpt, rb, cyl = pu.case_from_golden(g)
    args, theirs, ours = _both(ref, ckpt, precision)
    poses = torch.as_tensor(np.vstack([frame.c2w[:3, :4], [0, 0, 0, 1]]).astype(np.float32))[None]
    kw = dict(kp=torch.as_tensor(frame.pose.kps)[None], skts=torch.as_tensor(frame.pose.skts)[None],
              bones=torch.as_tensor(frame.pose.bones)[None], white_bkgd=True, ret_acc=True, ext_scale=args.ext_scale)
    out = {}
    for name, kwargs in (("ref", theirs[1]), ("ours", ours[1])):
        kwargs["ray_caster"].eval()
        with contextlib.redirect_stdout(io.StringIO()):
            out[name] = ref["run_nerf"].render_path(poses, (frame.H, frame.W, float(frame.focal)), args.chunk, kwargs, **kw)
    rgb_r, disp_r, acc_r, valid_r, _ = out["ref"]
    rgb_o, disp_o, acc_o, valid_o, _ = out["ours"]
    assert rgb_o.shape == rgb_r.shape == (1, 64, 64, 3)
    assert np.array_equal(np.asarray(valid_r[0]), np.asarray(valid_o[0]))
    assert pu.max_abs(rgb_o, rgb_r) <= tol and pu.max_abs(acc_o, acc_r) <= tol
    assert pu.psnr(rgb_o, rgb_r) >= 40.0
    # and the reference running on the B200 agrees with the committed CPU golden (the oracle's pin)
    flat = rgb_r[0].reshape(-1, 3)[np.asarray(valid_r[0])]
    assert pu.max_abs(flat, g["rgb_map"] + (1.0 - g["acc_map"][:, None])) <= 1e-4


def _train_batch(n=1024, seed=5):
    rng = np.random.RandomState(seed)
    rbs, sk, cy, kp, bn = [], [], [], [], []
    per = n // 8
    for p in range(8):
        frame = syn.synthetic_frame(300 + p, 64, 64)
        sel = rng.randint(0, frame.rays_o.shape[0], per)
        rbs.append(np.stack([frame.rays_o[sel], frame.rays_d[sel]]))
        sk.append(np.repeat(frame.pose.skts[None], per, 0)); cy.append(np.repeat(frame.pose.cyl[None], per, 0))
        kp.append(np.repeat(frame.pose.kps[None], per, 0)); bn.append(np.repeat(frame.pose.bones[None], per, 0))
    t = lambda a: torch.as_tensor(np.concatenate(a, 0).astype(np.float32))    # noqa: E731
    return {"rays": torch.as_tensor(np.concatenate(rbs, 1).astype(np.float32)), "target_s": torch.as_tensor(rng.rand(n, 3).astype(np.float32)),
            "kp3d": t(kp), "skts": t(sk), "bones": t(bn), "cyls": t(cy), "cam_idxs": torch.zeros(n, dtype=torch.long)}


def test_reference_trainer_train_batch_drives_the_dropin(ref):
    """core.trainer.Trainer.train_batch (core/trainer.py:232-275), unmodified: render -> loss -> backward -> Adam step ->
    lrate decay -> `ray_caster.module.update_embed_fns` / `.module.embed_fn.get_tau()` on the DataParallel wrapper."""
    ckpt = syn.synthetic_raycaster_state(4, alpha_gain=40.0)
    args, theirs, ours = _both(ref, ckpt, "bf16", perturb=0., raw_noise_std=0.)     # deterministic sampling: comparable losses
    batch = _train_batch()
    res = {}
    for name, (kw_train, kw_test, start, grad_vars, optimizer, _) in (("ref", theirs), ("ours", ours)):
        before = [p.detach().clone() for p in grad_vars]
        tr = ref["trainer"].Trainer(args, ref["data_attrs"], optimizer, None, kw_train, kw_test, device="cuda")
        kw_train["ray_caster"].train()
        with contextlib.redirect_stdout(io.StringIO()):
            loss_dict, stats = tr.train_batch(batch, i=0, global_step=1000)
            loss_dict2, stats2 = tr.train_batch(batch, i=1, global_step=1001)
        moved = float(torch.sqrt(sum(((a - b.to(a.device)) ** 2).sum() for a, b in zip(grad_vars, before))))
        res[name] = (float(loss_dict["total_loss"]), float(loss_dict2["total_loss"]), stats, moved)
    (l_r, l2_r, st_r, mv_r), (l_o, l2_o, st_o, mv_o) = res["ref"], res["ours"]
    assert np.isfinite([l_o, l2_o]).all() and mv_o > 0
    assert abs(l_o - l_r) <= 2e-2 * max(1.0, abs(l_r)), (l_o, l_r)
    assert abs(l2_o - l2_r) <= 3e-2 * max(1.0, abs(l2_r)), (l2_o, l2_r)      # after one Adam step each
    assert abs(mv_o - mv_r) <= 0.05 * mv_r                                      # Adam's first steps: |delta| = lr per element
    assert abs(st_o["cutoff"] - st_r["cutoff"]) < 1e-4 and st_o["cutoff"] > 20.  # tau schedule ran through .module
    assert abs(st_o["psnr"] - st_r["psnr"]) < 0.5 and st_o["total_norm"] > 0


def test_reference_checkpoint_tar_round_trip(ref, tmp_path):
    """A .tar written by the REFERENCE trainer's save_nerf (core/trainer.py:486-509) reloads into the drop-in through
    the reference's checkpoint discovery (newest *.tar in basedir/expname), optimizer state and global_step included."""
    ckpt = syn.synthetic_raycaster_state(2, alpha_gain=40.0)
    args, theirs, _ = _both(ref, ckpt, "bf16")
    kw_train, kw_test, _, grad_vars, optimizer, _ = theirs
    for p in grad_vars:
        p.grad = torch.full_like(p, 1e-3)
    optimizer.step()
    tr = ref["trainer"].Trainer(args, ref["data_attrs"], optimizer, None, kw_train, kw_test, device="cuda")
    exp = tmp_path / "exp"
    exp.mkdir()
    with contextlib.redirect_stdout(io.StringIO()):
        tr.save_nerf(str(exp / "012345.tar"), 12345)
    import copy
    a2 = copy.deepcopy(args)
    a2.basedir, a2.expname, a2.no_reload = str(tmp_path), "exp", False
    _, kw2, start, gv2, opt2, loaded = rcmod.create_raycaster(a2, ref["data_attrs"], device="cuda")
    assert start == 12345 and loaded["global_step"] == 12345
    for a, b in zip(kw_test["ray_caster"].parameters(), kw2["ray_caster"].parameters()):
        assert torch.equal(a.detach().cpu(), b.detach().cpu())
    assert int(opt2.state[gv2[0]]["step"]) == 1
    # the reloaded drop-in renders like the reference module the checkpoint came from
    g = pu.load_golden("a_32_boost_taps")
    frame, _, rb, cyl = pu.case_from_golden(g)
    n = rb.shape[0]
    args_in = dict(N_samples=64, N_importance=16, kp_batch=None, skts=torch.as_tensor(frame.pose.skts, device="cuda")[None].expand(n, 24, 4, 4),
                   cyls=torch.as_tensor(cyl, device="cuda")[None].expand(n, 5), bones=None, cams=None)
    kw2["ray_caster"].eval()
    ours = kw2["ray_caster"](torch.as_tensor(rb, device="cuda"), precision="fp32", **args_in)
    kw_test["ray_caster"].eval()
    with torch.no_grad():
        theirs_out = kw_test["ray_caster"](torch.as_tensor(rb, device="cuda"), kp_batch=torch.as_tensor(frame.pose.kps, device="cuda")[None].expand(n, 24, 3),
                                           **{k: v for k, v in args_in.items() if k != "kp_batch"}, **{k: v for k, v in kw_test.items() if k in ("preproc_kwargs",)})
    for k in ("rgb_map", "acc_map", "rgb0", "acc0"):
        assert pu.max_abs(ours[k].cpu().numpy(), theirs_out[k].cpu().numpy()) <= 1e-3, k


def test_two_threads_two_contexts_concurrently(engine):
    """SURVEY §8b threading contract: nn.DataParallel.parallel_apply calls replicas from one Python thread each; the C
    ABI is re-entrant per context and takes the stream explicitly.  Two contexts, two threads, two streams, interleaved
    launches: both results equal the single-threaded render bit for bit."""
    from posegen_b200.engine import Engine
    g = pu.load_golden("d_64_calibrated")
    frame, ckpt, rb, cyl = pu.case_from_golden(g)
    dev = torch.device("cuda", 0)
    ins = (torch.as_tensor(rb, device=dev), torch.as_tensor(frame.pose.skts, device=dev), torch.as_tensor(cyl, device=dev))
    engine.load_checkpoint(ckpt)
    want = {k: v.clone() for k, v in engine.render(*ins, nanfill_chunk=4096, precision="bf16").items()}
    torch.cuda.synchronize()
    results, errors = {}, []

    def worker(i):
        try:
            eng = Engine(dev)
            eng.load_checkpoint(ckpt)
            st = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(st):
                outs = [eng.render(*ins, nanfill_chunk=4096, precision="bf16") for _ in range(6)]
            st.synchronize()
            eng.check_status()
            results[i] = outs
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i in range(2):
        for out in results[i]:
            for k in ("rgb_map", "acc_map", "rgb0", "acc0", "alpha"):
                assert torch.equal(out[k], want[k]), (i, k)


def test_two_devices_one_process():
    """function attributes (opt-in shared memory) are per device and every entry point must leave the
    caller's current device alone.  Needs two GPUs (gpurun --gpus 2); skipped on a one-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 CUDA devices")
    from posegen_b200.engine import Engine
    g = pu.load_golden("d_64_calibrated")
    frame, ckpt, rb, cyl = pu.case_from_golden(g)
    outs = []
    torch.cuda.set_device(0)
    for d in (0, 1):
        dev = torch.device("cuda", d)
        eng = Engine(dev)
        eng.load_checkpoint(ckpt)
        assert torch.cuda.current_device() == 0
        ins = (torch.as_tensor(rb, device=dev), torch.as_tensor(frame.pose.skts, device=dev), torch.as_tensor(cyl, device=dev))
        outs.append(eng.render(*ins, nanfill_chunk=4096, precision="bf16"))
        nf = eng.near_far(*ins, nanfill_chunk=4096)
        assert nf.device == dev and torch.cuda.current_device() == 0
        torch.cuda.synchronize(dev)
        eng.check_status()
    for k in ("rgb_map", "acc_map"):
        assert torch.equal(outs[0][k].cpu(), outs[1][k].cpu())
    # nn.DataParallel over both devices, the reference's training wrapper (core/raycasters.py:157)
    kw_train, kw_test, _, grad_vars, _, _ = rcmod.create_raycaster(rcmod.surreal_args(perturb=0., raw_noise_std=0.), {"skel_type": None}, device="cuda:0")
    kw_test["ray_caster"].load_state_dict(syn.synthetic_raycaster_state(4, alpha_gain=40.0))      # semi-transparent volume: non-zero gradients
    dp = kw_train["ray_caster"]
    assert dp.device_ids == [0, 1]
    dp.train()
    n = 512
    rb0 = torch.as_tensor(rb[np.linspace(0, rb.shape[0] - 1, n).astype(np.int64)], device="cuda:0")
    sk = torch.as_tensor(frame.pose.skts, device="cuda:0")[None].expand(n, 24, 4, 4).contiguous()
    cy = torch.as_tensor(cyl, device="cuda:0")[None].expand(n, 5).contiguous()
    ret = dp(rb0, N_samples=64, N_importance=16, kp_batch=None, skts=sk, cyls=cy, bones=None, cams=None, perturb=0., raw_noise_std=0.)
    assert ret["rgb_map"].shape == (n, 3) and ret["rgb_map"].device.index == 0
    ((ret["rgb_map"] + (1 - ret["acc_map"][:, None]) - 0.25) ** 2).mean().backward()
    single = kw_test["ray_caster"]
    z = lambda p: torch.zeros_like(p) if p.grad is None else p.grad.clone()      # noqa: E731  (the loss reads the fine pass only)
    gdp = [z(p) for p in grad_vars]
    for p in grad_vars:
        p.grad = None
    r1 = single(rb0, N_samples=64, N_importance=16, kp_batch=None, skts=sk, cyls=cy, bones=None, cams=None, perturb=0., raw_noise_std=0.)
    ((r1["rgb_map"] + (1 - r1["acc_map"][:, None]) - 0.25) ** 2).mean().backward()
    g1 = [z(p) for p in grad_vars]
    num = sum(float((a * b).sum()) for a, b in zip(gdp, g1))
    den = np.sqrt(sum(float((a * a).sum()) for a in gdp) * sum(float((b ** 2).sum()) for b in g1))
    assert den > 0 and num / den >= 0.999
