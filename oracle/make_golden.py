"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

TEST INFRASTRUCTURE ONLY — run in the build container (needs /root/reference):

    python -m oracle.make_golden            # rewrites tests/golden/*.npz

For every case it (1) builds the reference ``RayCaster`` with ``create_raycaster``
(core/raycasters.py:17) for configs/surreal/surreal.txt, (2) loads the numpy-seeded
synthetic weights of ``posegen_b200.synthetic`` into it through the reference's own
``load_state_dict`` (core/raycasters.py:768), (3) renders the synthetic rays with the
reference's ``core.trainer.render`` (core/trainer.py:84) and records outputs and
stage taps, (4) checks that ``oracle/render_oracle.py`` reproduces the reference on
the same inputs and that the synthetic geometry helpers reproduce the reference's
host-side helpers (FK, cylinder, bbox rays), and (5) writes the fixture.
"""
from __future__ import annotations

import contextlib
import hashlib
import io
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim, render_oracle as orc          # noqa: E402
from posegen_b200 import synthetic as syn                  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def to_torch_ckpt(ckpt):
    out = {}
    for k, sd in ckpt.items():
        out[k] = {kk: torch.as_tensor(np.asarray(vv)) for kk, vv in sd.items()}
    return out


class Taps:
    """Monkey-patch wrappers that record what flows between the reference's stages."""

    def __init__(self):
        self.data = {}
        self._restore = []

    def _cat(self, key, t):
        self.data.setdefault(key, []).append(t.detach().cpu().clone())

    def install(self):
        import core.raycasters as rc
        import core.utils.ray_utils as ru
        taps = self

        orig_nf = rc.get_near_far_in_cylinder

        def nf(*a, **k):
            n, f = orig_nf(*a, **k)
            taps._cat("near", n)
            taps._cat("far", f)
            return n, f
        rc.get_near_far_in_cylinder = nf
        self._restore.append((rc, "get_near_far_in_cylinder", orig_nf))

        orig_ss = torch.searchsorted

        def ss(*a, **k):
            r = orig_ss(*a, **k)
            taps._cat("pdf_inds", r)
            taps._cat("cdf", a[0])
            return r
        torch.searchsorted = ss
        self._restore.append((torch, "searchsorted", orig_ss))

        orig_is = rc.isample_from_lineseg

        def isamp(*a, **k):
            z_all, z_s, idx = orig_is(*a, **k)
            taps._cat("z_fine", z_all)
            taps._cat("z_samples", z_s)
            taps._cat("sorted_idxs", idx)
            taps._cat("z_coarse", a[0])
            taps._cat("weights_coarse", a[1])
            return z_all, z_s, idx
        rc.isample_from_lineseg = isamp
        self._restore.append((rc, "isample_from_lineseg", orig_is))

        orig_run = rc.RayCaster.run_network

        def run(self_, encoded, network, *a, **k):
            out = orig_run(self_, encoded, network, *a, **k)
            key = "coarse" if out.shape[1] == 64 else "fine"
            taps._cat(f"raw_{key}", out)
            if key == "coarse":
                taps._cat("enc_coarse_head", torch.cat([encoded["v"], encoded["r"], encoded["d"]], -1)[:2])
            return out
        rc.RayCaster.run_network = run
        self._restore.append((rc.RayCaster, "run_network", orig_run))
        _ = ru
        return self

    def remove(self):
        for obj, name, fn in self._restore:
            setattr(obj, name, fn)

    def get(self, key):
        return torch.cat(self.data[key], 0).numpy()


def reference_render(render_kwargs, frame, ckpt, chunk=4096, cyl_override=None, cams=None, lindisp=None):
    from core.trainer import render
    rcast = render_kwargs["ray_caster"]
    rcast.load_state_dict(to_torch_ckpt(ckpt))
    rcast.eval()
    n = frame.rays_o.shape[0]
    rays = (torch.from_numpy(frame.rays_o), torch.from_numpy(frame.rays_d))
    cyl = frame.pose.cyl if cyl_override is None else cyl_override
    exp = lambda a: torch.from_numpy(np.ascontiguousarray(a))[None].expand(n, *a.shape).clone()  # noqa: E731
    if lindisp is not None:
        render_kwargs = dict(render_kwargs, lindisp=lindisp)
    taps = Taps().install()
    try:
        with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
            out = render(frame.H, frame.W, frame.focal, rays=rays, chunk=chunk,
                         kp_batch=exp(frame.pose.kps), skts=exp(frame.pose.skts), cyls=exp(cyl),
                         bones=exp(frame.pose.bones), cams=cams, subject_idxs=None, **render_kwargs)
    finally:
        taps.remove()
    return {k: v.cpu().numpy() for k, v in out.items()}, taps


def oracle_render(frame, ckpt, chunk=4096, cyl_override=None, dtype=torch.float32, cams=None, lindisp=False):
    rb = torch.from_numpy(syn.ray_batch(frame.rays_o, frame.rays_d)).to(dtype)
    cyl = frame.pose.cyl if cyl_override is None else cyl_override
    taps = {}
    nets = orc.nets_from_ckpt(ckpt, dtype)
    emb = orc.embed_params_from_ckpt(ckpt, dtype)
    if rb.shape[0] <= chunk:
        with torch.no_grad():
            out = orc.render_rays(rb, torch.from_numpy(frame.pose.skts).to(dtype)[None].expand(rb.shape[0], -1, -1, -1),
                                  torch.from_numpy(cyl).to(dtype)[None].expand(rb.shape[0], -1), nets, emb, taps=taps, cams=cams, lindisp=lindisp)
    else:
        out = orc.render(rb, torch.from_numpy(frame.pose.skts).to(dtype), torch.from_numpy(cyl).to(dtype),
                         nets, emb, chunk=chunk, cams=cams, lindisp=lindisp)
    return {k: v.numpy() for k, v in out.items()}, {k: v.numpy() for k, v in taps.items()}


def check_geometry(frame):
    """synthetic.py helpers vs the reference's host-side helpers."""
    from core.utils.skeleton_utils import get_smpl_l2ws, smpl_rest_pose, get_kp_bounding_cylinder
    from core.utils.ray_utils import kp_to_valid_rays
    assert np.array_equal(smpl_rest_pose, syn.SMPL_REST_POSE)
    l2ws = get_smpl_l2ws(frame.pose.bones.astype(np.float64), smpl_rest_pose * np.float32(syn.BODY_SCALE), 1.0)
    err_fk = np.abs(l2ws - syn.smpl_local_to_world(frame.pose.bones, smpl_rest_pose * np.float32(syn.BODY_SCALE))).max()
    skts_ref = np.linalg.inv(l2ws)
    err_skt = np.abs(skts_ref - frame.pose.skts).max()
    with contextlib.redirect_stdout(io.StringIO()):
        cyl_ref = get_kp_bounding_cylinder(frame.pose.kps[None], ext_scale=0.001, extend_mm=250,
                                           top_expand_ratio=1.6, bot_expand_ratio=1.1, head="-y")[0]
        rays, valid, cyl2, bboxes = kp_to_valid_rays(torch.from_numpy(frame.c2w)[None], frame.H, frame.W, frame.focal,
                                                     kps=torch.from_numpy(frame.pose.kps)[None], ext_scale=0.001)
    err_cyl = np.abs(cyl_ref - frame.pose.cyl).max()
    assert np.array_equal(valid[0].numpy(), frame.valid_idx), "bbox pixel set differs from kp_to_valid_rays"
    err_o = np.abs(rays[0][0].numpy() - frame.rays_o).max()
    err_d = np.abs(rays[0][1].numpy() - frame.rays_d).max()
    print(f"  geometry: FK {err_fk:.2e}  skts {err_skt:.2e}  cyl {err_cyl:.2e}  rays_o {err_o:.2e}  rays_d {err_d:.2e}")
    assert err_fk < 1e-6 and err_skt < 1e-5 and err_cyl < 1e-6 and err_o == 0 and err_d < 1e-6
    return {"geom_err_rays_d": err_d}


def compare(name, ref, got):
    worst = 0.0
    for k in ref:
        if k in got:
            d = float(np.abs(ref[k].astype(np.float64) - got[k].astype(np.float64)).max())
            worst = max(worst, d)
            if d != 0.0:
                print(f"    {name}: oracle vs reference {k}: max-abs {d:.3e}")
    return worst


def make_case(render_kwargs, name, pose_seed, res, weight_seed, alpha_gain, full_taps,
              calibrated=False, shrink_cyl=None, chunk=4096, lindisp=False):
    print(f"[{name}] pose_seed={pose_seed} res={res} weight_seed={weight_seed} gain={alpha_gain} "
          f"calibrated={calibrated} shrink_cyl={shrink_cyl}")
    frame = syn.synthetic_frame(pose_seed, res, res)
    check_geometry(frame)
    ckpt = syn.synthetic_raycaster_state(weight_seed, alpha_gain=None if calibrated else alpha_gain)
    cyl = None
    if shrink_cyl is not None:
        cyl = frame.pose.cyl.copy()
        cyl[2] *= np.float32(shrink_cyl)
    extra = {}
    if calibrated:
        # SURVEY.md §8d calibrated-head recipe: one zero-bias pass, theta per net
        for key, raw_key in (("network_fn_state_dict", "raw_coarse"), ("network_fine_state_dict", "raw_fine")):
            ckpt[key]["alpha_linear.bias"] = np.zeros_like(ckpt[key]["alpha_linear.bias"])
        _, taps0 = reference_render(render_kwargs, frame, ckpt, chunk, lindisp=lindisp)
        for key, raw_key in (("network_fn_state_dict", "raw_coarse"), ("network_fine_state_dict", "raw_fine")):
            sig_far = float(taps0.get(raw_key)[:, -1, 3].max())
            syn.calibrate_alpha_head(ckpt[key], sig_far)
            extra[f"sigma_far_max_{raw_key}"] = np.float32(sig_far)
    ref, taps = reference_render(render_kwargs, frame, ckpt, chunk, cyl, lindisp=lindisp)
    got, otaps = oracle_render(frame, ckpt, chunk, cyl, lindisp=lindisp)
    worst = compare(name, ref, got)
    for k in ("near", "far", "z_coarse", "z_samples", "z_fine", "sorted_idxs", "pdf_inds", "raw_coarse", "raw_fine"):
        if k in otaps:
            worst = max(worst, compare(name, {k: taps.get(k)}, otaps))
    print(f"  oracle vs reference worst max-abs: {worst:.3e}; acc mean {ref['acc_map'].mean():.4f} "
          f"max {ref['acc_map'].max():.4f}; rays {ref['acc_map'].shape[0]}")
    assert worst <= 2e-6, "oracle restatement deviates from the reference"

    rb = syn.ray_batch(frame.rays_o, frame.rays_d)
    fix = {
        "meta_pose_seed": np.int64(pose_seed), "meta_res": np.int64(res), "meta_weight_seed": np.int64(weight_seed),
        "meta_alpha_gain": np.float32(alpha_gain if alpha_gain else 0.0), "meta_calibrated": np.bool_(calibrated),
        "meta_chunk": np.int64(chunk), "meta_oracle_vs_ref_maxabs": np.float64(worst),
        "in_ray_batch_sha": np.array(sha(rb)), "in_skts_sha": np.array(sha(frame.pose.skts)),
        "in_cyl": (frame.pose.cyl if cyl is None else cyl).astype(np.float32),
        "in_ray_batch_head": rb[:16],
        "rgb_map": ref["rgb_map"], "disp_map": ref["disp_map"], "acc_map": ref["acc_map"],
        "rgb0": ref["rgb0"], "disp0": ref["disp0"], "acc0": ref["acc0"],
    }
    fix.update(extra)
    if lindisp:
        fix["meta_lindisp"] = np.bool_(True)
    if full_taps:
        fix.update({
            "alpha": ref["alpha"], "alpha0": ref["alpha0"],
            "near": taps.get("near"), "far": taps.get("far"),
            "z_samples": taps.get("z_samples"),
            "pdf_inds": taps.get("pdf_inds").astype(np.uint8),
            "sorted_idxs": taps.get("sorted_idxs").astype(np.uint8),
            "weights_coarse": taps.get("weights_coarse"),
            "cdf_last": taps.get("cdf")[:, -1].copy(),
            "enc_coarse_head": taps.get("enc_coarse_head")[:2],
            "raw_coarse_head": taps.get("raw_coarse")[:64], "raw_fine_head": taps.get("raw_fine")[:64],
        })
    elif shrink_cyl is not None:
        fix.update({"near": taps.get("near"), "far": taps.get("far")})
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(path, **fix)
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def build_framecode_raycaster(tmpdir, n_framecodes):
    """The reference's create_raycaster for configs/h36m/h36m_prot2.txt (opt_framecode = True: Optcodes frame codes,
    core/networks/embedding.py:4-46), which differs from surreal.txt in nothing else the render path reads."""
    import run_nerf
    from core.raycasters import create_raycaster
    from core.utils.skeleton_utils import SMPLSkeleton, smpl_rest_pose, get_per_joint_coords
    with contextlib.redirect_stdout(io.StringIO()):
        os.makedirs(os.path.join(tmpdir, "fc"), exist_ok=True)
        args = run_nerf.config_parser().parse_args(
            ["--config", os.path.join(ref_shim.REFERENCE_ROOT, "configs/h36m/h36m_prot2.txt"),
             "--basedir", tmpdir, "--expname", "fc", "--no_reload"])
        assert args.opt_framecode and args.framecode_size == 16
        data_attrs = {"skel_type": SMPLSkeleton, "near": 60., "far": 100., "n_views": n_framecodes,
                      "joint_coords": get_per_joint_coords(smpl_rest_pose.astype(np.float32))}
        _, render_kwargs_test, _, _, _, _ = create_raycaster(args, data_attrs)
    return render_kwargs_test


def make_framecode_case(render_kwargs, name, pose_seed, res, weight_seed, n_framecodes, chunk=4096):
    """Optcodes case (h36m_prot2-shaped model): calibrated head, one render with a real camera index per ray and one with
    cams = -1 (the mean code, embedding.py:23-24), plus the reference's density-only query on random points (pins
    `fwd_type='density'`, core/raycasters.py:597-648)."""
    print(f"[{name}] pose_seed={pose_seed} res={res} weight_seed={weight_seed} n_framecodes={n_framecodes}")
    frame = syn.synthetic_frame(pose_seed, res, res)
    check_geometry(frame)
    ckpt = syn.synthetic_raycaster_state(weight_seed, alpha_gain=None, n_framecodes=n_framecodes)
    n = frame.rays_o.shape[0]
    cams = torch.full((n,), 3, dtype=torch.long)
    cams[: n // 2] = 1                                   # two cameras in one batch (chunks see a mix of valid indices)
    extra = {}
    for key in ("network_fn_state_dict", "network_fine_state_dict"):
        ckpt[key]["alpha_linear.bias"] = np.zeros_like(ckpt[key]["alpha_linear.bias"])
    _, taps0 = reference_render(render_kwargs, frame, ckpt, chunk, cams=cams)
    for key, raw_key in (("network_fn_state_dict", "raw_coarse"), ("network_fine_state_dict", "raw_fine")):
        sig_far = float(taps0.get(raw_key)[:, -1, 3].max())
        syn.calibrate_alpha_head(ckpt[key], sig_far)
        extra[f"sigma_far_max_{raw_key}"] = np.float32(sig_far)
    worst = 0.0
    fix = {}
    for tag, c in (("", cams), ("_mean", torch.full((n,), -1, dtype=torch.long))):
        ref, taps = reference_render(render_kwargs, frame, ckpt, chunk, cams=c)
        got, otaps = oracle_render(frame, ckpt, chunk, cams=c)
        worst = max(worst, compare(name + tag, ref, got))
        for k in ("rgb_map", "disp_map", "acc_map", "rgb0", "disp0", "acc0"):
            fix[k + tag] = ref[k]
        print(f"  {tag or 'cams':>6}: acc mean {ref['acc_map'].mean():.4f} max {ref['acc_map'].max():.4f}")
    fix["rgb_delta_mean_vs_cam"] = np.float32(np.abs(fix["rgb_map"] - fix["rgb_map_mean"]).max())
    assert fix["rgb_delta_mean_vs_cam"] > 1e-4, "the frame code must be visible in the image"
    # density-only query of the reference on random points around the body
    rcast = render_kwargs["ray_caster"]
    rng = np.random.RandomState(5)
    pts = (frame.pose.kps[0] + (rng.rand(2000, 3).astype(np.float32) - 0.5) * 1.2).astype(np.float32)
    exp = lambda a: torch.from_numpy(np.ascontiguousarray(a))[None].expand(len(pts), *a.shape).clone()  # noqa: E731
    with contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
        dens = rcast(torch.from_numpy(pts).reshape(-1, 1, 3), exp(frame.pose.kps), exp(frame.pose.skts), exp(frame.pose.bones),
                     render_kwargs=render_kwargs["preproc_kwargs"], fwd_type="density")
    dens = dens.reshape(-1).cpu().numpy()
    from oracle import next_oracle as nxt
    nets, emb = orc.nets_from_ckpt(ckpt), orc.embed_params_from_ckpt(ckpt)
    dens_o = nxt.density_of_points(torch.from_numpy(pts), torch.from_numpy(frame.pose.skts), nets[1], emb).numpy()
    d_err = float(np.abs(dens - dens_o).max())
    print(f"  oracle vs reference worst max-abs: {worst:.3e}; density query {d_err:.3e} (|raw| max {np.abs(dens).max():.2f})")
    assert worst <= 2e-6 and d_err <= 2e-5 * max(1.0, float(np.abs(dens).max())), "oracle restatement deviates from the reference"
    rb = syn.ray_batch(frame.rays_o, frame.rays_d)
    fix.update({
        "meta_pose_seed": np.int64(pose_seed), "meta_res": np.int64(res), "meta_weight_seed": np.int64(weight_seed),
        "meta_alpha_gain": np.float32(0.0), "meta_calibrated": np.bool_(True), "meta_chunk": np.int64(chunk),
        "meta_n_framecodes": np.int64(n_framecodes), "meta_oracle_vs_ref_maxabs": np.float64(worst),
        "in_ray_batch_sha": np.array(sha(rb)), "in_skts_sha": np.array(sha(frame.pose.skts)),
        "in_cyl": frame.pose.cyl.astype(np.float32), "in_ray_batch_head": rb[:16], "in_cams": cams.numpy().astype(np.int32),
        "density_pts": pts, "density_raw": dens.astype(np.float32),
    })
    fix.update(extra)
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(path, **fix)
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def main():
    torch.manual_seed(0)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        render_kwargs, _ = ref_shim.build_reference_raycaster(tmp)
        # A: 32x32, boosted alpha head (x400, zero bias), full stage taps
        make_case(render_kwargs, "a_32_boost_taps", pose_seed=0, res=32, weight_seed=0, alpha_gain=400., full_taps=True)
        # B: BASELINE config 1 — 64x64 coarse+fine, boosted head (non-empty image)
        make_case(render_kwargs, "b_64_boost", pose_seed=1, res=64, weight_seed=0, alpha_gain=400., full_taps=False)
        # C: 64x64, plain nn.Linear-style init (near-empty volume)
        make_case(render_kwargs, "c_64_plain", pose_seed=2, res=64, weight_seed=1, alpha_gain=None, full_taps=False)
        # D: 64x64, calibrated head (the bf16-tier recipe)
        make_case(render_kwargs, "d_64_calibrated", pose_seed=1, res=64, weight_seed=0, alpha_gain=None,
                  full_taps=False, calibrated=True)
        # E: 32x32 with a shrunken cylinder so bbox-corner rays miss it: chunk-level NaN fill
        make_case(render_kwargs, "e_32_nanfill", pose_seed=3, res=32, weight_seed=0, alpha_gain=400., full_taps=False,
                  shrink_cyl=0.8)
        # G: 32x32, calibrated head, coarse samples linear in inverse depth (render_kwargs['lindisp'] = True)
        make_case(render_kwargs, "g_32_lindisp", pose_seed=5, res=32, weight_seed=3, alpha_gain=None, full_taps=False,
                  calibrated=True, lindisp=True)
        # F: 64x64, h36m_prot2-shaped model (opt_framecode = True, 5 frame codes), calibrated head, + density-only query
        make_framecode_case(build_framecode_raycaster(tmp, 5), "f_64_framecode", pose_seed=4, res=64, weight_seed=2, n_framecodes=5)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--lindisp-only":
        with tempfile.TemporaryDirectory() as tmp:
            rk, _ = ref_shim.build_reference_raycaster(tmp)
            make_case(rk, "g_32_lindisp", pose_seed=5, res=32, weight_seed=3, alpha_gain=None, full_taps=False, calibrated=True, lindisp=True)
    elif len(sys.argv) > 1 and sys.argv[1] == "--framecode-only":
        with tempfile.TemporaryDirectory() as tmp:
            ref_shim.install()
            make_framecode_case(build_framecode_raycaster(tmp, 5), "f_64_framecode", pose_seed=4, res=64, weight_seed=2, n_framecodes=5)
    else:
        main()
