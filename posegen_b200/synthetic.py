"""Synthetic SMPL poses, cameras, rays and random-init A-NeRF weights.

Everything the parity tests and ``bench.py`` feed to the renderer is produced here
with ``numpy.random.RandomState`` (bit-stable across numpy/torch versions), so the
GPU box can regenerate the exact inputs the golden fixtures were made from.

The geometry helpers restate, in vectorised numpy, what the reference does on the
host before it reaches the hot path:
  * forward kinematics        core/utils/skeleton_utils.py:334-375  (get_smpl_l2ws)
  * bounding cylinder         core/utils/skeleton_utils.py:635-685  (get_kp_bounding_cylinder, head='-y')
  * cylinder -> 2-D bbox      core/utils/skeleton_utils.py:700-787  (cylinder_to_box_2d)
  * pixel rays                core/utils/ray_utils.py:6-28          (get_rays)
  * bbox rays per image       core/utils/ray_utils.py:83-136        (kp_to_valid_rays)
  * camera                    run_gan.py:2023-2029 (fixed extrinsic), skeleton_utils.py:1412-1421
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

N_JOINTS = 24

# SMPL kinematic tree (parent of each joint), core/utils/skeleton_utils.py:98-104
SMPL_PARENTS = np.array([0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21],
                        dtype=np.int64)

# SMPL rest-pose joint locations (data table, core/utils/skeleton_utils.py:259-282)
SMPL_REST_POSE = np.array([
    [0.00000000e+00, 2.30003661e-09, -9.86228770e-08],
    [1.63832515e-01, -2.17391014e-01, -2.89178602e-02],
    [-1.57855421e-01, -2.14761734e-01, -2.09642015e-02],
    [-7.04505108e-03, 2.50450850e-01, -4.11837511e-02],
    [2.42021069e-01, -1.08830070e+00, -3.14962119e-02],
    [-2.47206554e-01, -1.10715497e+00, -3.06970738e-02],
    [3.95125849e-03, 5.94849110e-01, -4.03754264e-02],
    [2.12680623e-01, -1.99382353e+00, -1.29327580e-01],
    [-2.10857525e-01, -2.01218796e+00, -1.23002514e-01],
    [9.39484313e-03, 7.19204426e-01, 2.06931755e-02],
    [2.63385147e-01, -2.12222481e+00, 1.46775618e-01],
    [-2.51970559e-01, -2.12153077e+00, 1.60450473e-01],
    [3.83779174e-03, 1.22592449e+00, -9.78838727e-02],
    [1.91201791e-01, 1.00385976e+00, -6.21964522e-02],
    [-1.77145526e-01, 9.96228695e-01, -7.55542740e-02],
    [1.68482102e-02, 1.38698268e+00, 2.44048554e-02],
    [4.01985168e-01, 1.07928419e+00, -7.47655183e-02],
    [-3.98825467e-01, 1.07523870e+00, -9.96334553e-02],
    [1.00236952e+00, 1.05217218e+00, -1.35129794e-01],
    [-9.86728609e-01, 1.04515052e+00, -1.40235111e-01],
    [1.56646240e+00, 1.06961894e+00, -1.37338534e-01],
    [-1.56946480e+00, 1.05935931e+00, -1.53905824e-01],
    [1.75282109e+00, 1.04682994e+00, -1.68231070e-01],
    [-1.75758195e+00, 1.04255080e+00, -1.77773550e-01]], dtype=np.float32)

# The fixed extrinsic PoseGen renders every generated pose with (run_gan.py:2023-2028)
RUN_GAN_EXTRINSIC = np.array([
    [-5.29919172e-01, -5.56525674e-09, 8.48048140e-01, -1.34771157e-07],
    [1.47262004e-01, 9.84807813e-01, 9.20194958e-02, 1.26640154e-08],
    [-8.35164413e-01, 1.73648166e-01, -5.21868549e-01, 4.28571429e+00],
    [0.0, 0.0, 0.0, 1.0]], dtype=np.float32)

BODY_SCALE = 0.4          # run_gan.py:2032
EXT_SCALE = 0.001         # configs/surreal/surreal.txt:9


# --------------------------------------------------------------------------- FK
def rotvec_to_matrix(rotvec: np.ndarray) -> np.ndarray:
    """Rodrigues formula, [...,3] -> [...,3,3] (float64 internally)."""
    rv = np.asarray(rotvec, dtype=np.float64)
    theta = np.linalg.norm(rv, axis=-1, keepdims=True)
    safe = np.where(theta < 1e-12, 1.0, theta)
    k = rv / safe
    kx, ky, kz = k[..., 0], k[..., 1], k[..., 2]
    zero = np.zeros_like(kx)
    K = np.stack([zero, -kz, ky, kz, zero, -kx, -ky, kx, zero], axis=-1).reshape(rv.shape[:-1] + (3, 3))
    s = np.sin(theta)[..., None]
    c = np.cos(theta)[..., None]
    eye = np.broadcast_to(np.eye(3), K.shape)
    return eye + s * K + (1.0 - c) * (K @ K)


def smpl_local_to_world(bones: np.ndarray, rest_pose: np.ndarray | None = None,
                        parents: np.ndarray = SMPL_PARENTS) -> np.ndarray:
    """Kinematic chain: per-joint local->world 4x4 transforms, [24,4,4] float64.

    Root: [R_0 | rest_0].  Child i: l2w[parent] @ [R_i | rest_i - rest_parent].
    """
    rest = (SMPL_REST_POSE if rest_pose is None else rest_pose).astype(np.float64)
    rots = rotvec_to_matrix(bones)
    out = np.zeros((len(parents), 4, 4), dtype=np.float64)
    out[:, 3, 3] = 1.0
    out[0, :3, :3] = rots[0]
    out[0, :3, 3] = rest[0]
    for i in range(1, len(parents)):
        rel = np.eye(4)
        rel[:3, :3] = rots[i]
        rel[:3, 3] = rest[i] - rest[parents[i]]
        out[i] = out[parents[i]] @ rel
    return out


def rigid_inverse(mats: np.ndarray) -> np.ndarray:
    """Closed-form inverse of rigid 4x4 transforms ([R|t] -> [R^T | -R^T t])."""
    inv = np.zeros_like(mats)
    Rt = np.swapaxes(mats[..., :3, :3], -1, -2)
    inv[..., :3, :3] = Rt
    inv[..., :3, 3] = -(Rt @ mats[..., :3, 3:4])[..., 0]
    inv[..., 3, 3] = 1.0
    return inv


# --------------------------------------------------------------------- cylinder
def bounding_cylinder(kps: np.ndarray, ext_scale: float = EXT_SCALE, extend_mm: float = 250.,
                      top_expand_ratio: float = 1.6, bot_expand_ratio: float = 1.1) -> np.ndarray:
    """(cx, cz, R, top, bot) of the vertical cylinder around a pose, head='-y'.

    kps: [24,3] or [B,24,3].  Only (cx, cz, R) influence near/far; top/bot shape the bbox.
    """
    kps = np.asarray(kps)
    single = kps.ndim == 2
    k = kps[None] if single else kps
    root = k[:, 0]
    ext = extend_mm * ext_scale
    radial = np.linalg.norm(k[:, :, [0, 2]] - root[:, None, [0, 2]], axis=-1).max(-1)
    height = -k[:, :, 1]                       # flip = -1 for head '-y'
    top = -(height.max(-1) + ext * top_expand_ratio)
    bot = -(height.min(-1) - ext * bot_expand_ratio)
    cyl = np.stack([root[:, 0], root[:, 2], radial + ext, top, bot], axis=-1)
    return cyl[0] if single else cyl


def swap_yz_sign(mat: np.ndarray) -> np.ndarray:
    """Right-multiply by diag(1,-1,-1,1): OpenCV <-> NeRF camera axes."""
    out = np.array(mat, copy=True)
    out[..., 1] = -out[..., 1]
    out[..., 2] = -out[..., 2]
    return out


def run_gan_c2w() -> np.ndarray:
    """Camera-to-world of the PoseGen render camera: swap(inv(extrinsic)) (run_gan.py:2029)."""
    return swap_yz_sign(np.linalg.inv(RUN_GAN_EXTRINSIC.astype(np.float64))).astype(np.float32)


def cylinder_bbox_2d(cyl: np.ndarray, H: int, W: int, focal: float, c2w: np.ndarray):
    """Integer (tl, br) pixel box that encloses the projected cylinder caps
    (50 points per cap), clipped to the image."""
    w2c = np.linalg.inv(swap_yz_sign(np.vstack([c2w[:3, :4], [0, 0, 0, 1]]) if c2w.shape[0] == 3 else c2w))
    ang = np.linspace(0., 2 * np.pi, 50)
    x = cyl[0] + np.cos(ang) * cyl[2]
    z = cyl[1] + np.sin(ang) * cyl[2]
    one = np.ones_like(x)
    caps = np.concatenate([np.stack([x, cyl[3] * one, z, one], -1),
                           np.stack([x, cyl[4] * one, z, one], -1)], 0)
    cam = caps @ w2c.T
    K = np.array([[focal, 0, 0, 0], [0, focal, 0, 0], [0, 0, 1, 0]], dtype=np.float32)
    proj = cam @ K.T
    uv = proj[:, :2] / proj[:, 2:3]
    tl = np.floor(uv.min(0)).astype(np.int32) + np.array([int(W * .5), int(H * .5)], dtype=np.int32)
    br = np.ceil(uv.max(0)).astype(np.int32) + np.array([int(W * .5), int(H * .5)], dtype=np.int32)
    tl = np.clip(tl, 0, [W - 1, H - 1])
    br = np.clip(br, 0, [W - 1, H - 1])
    return tl, br


# ------------------------------------------------------------------------- rays
def pixel_rays(H: int, W: int, focal: float, c2w: np.ndarray):
    """rays_o, rays_d [H*W,3] (row-major pixels); d un-normalised, no half-pixel offset."""
    c2w = np.asarray(c2w, dtype=np.float32)
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    f = np.float32(focal)
    dirs = np.stack([(i - np.float32(W * .5)) / f, -(j - np.float32(H * .5)) / f, -np.ones_like(i)], -1)
    rays_d = (dirs[..., None, :] * c2w[:3, :3]).sum(-1, dtype=np.float32).reshape(-1, 3)
    rays_o = np.broadcast_to(c2w[:3, 3], rays_d.shape).copy()
    return rays_o.astype(np.float32), rays_d.astype(np.float32)


def bbox_pixel_indices(tl, br, W: int) -> np.ndarray:
    """Flat pixel indices of rows [tl.y, br.y) x cols [tl.x, br.x)."""
    hh = np.arange(tl[1], br[1], dtype=np.int64)
    ww = np.arange(tl[0], br[0], dtype=np.int64)
    return (hh[:, None] * W + ww[None, :]).reshape(-1)


# ------------------------------------------------------------------------ poses
@dataclass
class SyntheticPose:
    bones: np.ndarray      # [24,3] axis-angle
    kps: np.ndarray        # [24,3] world joint positions
    skts: np.ndarray       # [24,4,4] world->joint-local
    l2ws: np.ndarray       # [24,4,4]
    cyl: np.ndarray        # [5]


def synthetic_pose(seed: int, scale: float = BODY_SCALE) -> SyntheticPose:
    """bones ~ N(0,0.3^2), root orientation uniform in +-3.14 (run_gan.py:894)."""
    rng = np.random.RandomState(seed)
    bones = rng.randn(N_JOINTS, 3) * 0.3
    bones[0] = (rng.rand(3) - 0.5) * 6.28
    l2ws = smpl_local_to_world(bones, SMPL_REST_POSE * np.float32(scale))
    skts = rigid_inverse(l2ws)
    kps = l2ws[:, :3, 3]
    return SyntheticPose(bones.astype(np.float32), kps.astype(np.float32), skts.astype(np.float32),
                         l2ws.astype(np.float32), bounding_cylinder(kps.astype(np.float32)).astype(np.float32))


@dataclass
class SyntheticFrame:
    H: int
    W: int
    focal: float
    c2w: np.ndarray        # [4,4]
    pose: SyntheticPose
    tl: np.ndarray
    br: np.ndarray
    valid_idx: np.ndarray  # flat pixel indices inside the bbox
    rays_o: np.ndarray     # [n,3] bbox rays
    rays_d: np.ndarray


def synthetic_frame(seed: int, H: int = 512, W: int = 512, focal: float | None = None,
                    full_frame: bool = False) -> SyntheticFrame:
    """One PoseGen-style render job: pose(seed) seen by the run_gan camera.
    focal defaults to 1000 * H / 512 (run_gan.py:2034: 512x512 @ focal 1000)."""
    focal = float(1000.0 * H / 512.0) if focal is None else float(focal)
    pose = synthetic_pose(seed)
    c2w = run_gan_c2w()
    tl, br = cylinder_bbox_2d(pose.cyl, H, W, focal, c2w)
    if full_frame:
        tl, br = np.array([0, 0], np.int32), np.array([W, H], np.int32)
    idx = bbox_pixel_indices(tl, br, W)
    ro, rd = pixel_rays(H, W, focal, c2w)
    return SyntheticFrame(H, W, focal, c2w, pose, tl, br, idx, ro[idx], rd[idx])


def ray_batch(rays_o: np.ndarray, rays_d: np.ndarray) -> np.ndarray:
    """[n,11] = o(3) d(3) near=0 far=1 viewdirs(3) (core/trainer.py:118-137)."""
    vd = rays_d / np.linalg.norm(rays_d, axis=-1, keepdims=True)
    n = rays_o.shape[0]
    return np.concatenate([rays_o, rays_d, np.zeros((n, 1), np.float32), np.ones((n, 1), np.float32),
                           vd.astype(np.float32)], -1).astype(np.float32)


# ---------------------------------------------------------------------- weights
# (state_dict name, out_features, in_features) of one A-NeRF MLP for surreal.txt
# (core/networks/nerf.py:57-88)
NERF_LAYERS = (
    [("pts_linears.0", 256, 432)]
    + [(f"pts_linears.{i}", 256, 256) for i in range(1, 5)]
    + [("pts_linears.5", 256, 688), ("pts_linears.6", 256, 256), ("pts_linears.7", 256, 256),
       ("alpha_linear", 1, 256), ("views_linears.0", 128, 904), ("feature_linear", 256, 256),
       ("rgb_linear", 3, 128)]
)


def synthetic_nerf_state(seed: int) -> dict:
    """Random-init weights of one NeRF with nn.Linear's default distribution
    (weight, bias ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in))), drawn from a numpy
    RandomState so they are identical on every machine."""
    rng = np.random.RandomState(seed)
    state = {}
    for name, fan_out, fan_in in NERF_LAYERS:
        bound = 1.0 / math.sqrt(fan_in)
        state[f"{name}.weight"] = rng.uniform(-bound, bound, size=(fan_out, fan_in)).astype(np.float32)
        state[f"{name}.bias"] = rng.uniform(-bound, bound, size=(fan_out,)).astype(np.float32)
    return state


FRAMECODE_CH = 16


def add_framecodes(net_state: dict, n_framecodes: int, seed: int) -> dict:
    """Optcodes variant of a net (opt_framecode = True: h36m / mixamo / perfcap configs): views_linears.0 grows by 16
    input columns (core/networks/nerf.py:52-56) and the net owns `framecodes.codes.weight` [n,16] (xavier-normal,
    core/networks/embedding.py:38-41).  The first 904 columns keep the values of the frame-code-free state."""
    rng = np.random.RandomState(seed)
    w = net_state["views_linears.0.weight"]
    bound = 1.0 / math.sqrt(w.shape[1] + FRAMECODE_CH)
    extra = rng.uniform(-bound, bound, size=(w.shape[0], FRAMECODE_CH)).astype(np.float32)
    net_state["views_linears.0.weight"] = np.concatenate([w, extra], 1)
    std = math.sqrt(2.0 / (n_framecodes + FRAMECODE_CH))
    net_state["framecodes.codes.weight"] = (rng.randn(n_framecodes, FRAMECODE_CH) * std).astype(np.float32)
    return net_state


def synthetic_raycaster_state(seed: int, alpha_gain: float | None = None, n_framecodes: int = 0) -> dict:
    """Checkpoint-shaped dict (core/raycasters.py:752-766 key names) with two nets.

    alpha_gain: fp32-tier "boosted head" recipe of SURVEY.md §8d — multiply
    alpha_linear.weight by the gain and zero its bias so the volume is not empty.
    n_framecodes > 0: both nets get Optcodes frame codes (`add_framecodes`).
    """
    ckpt = {
        "network_fn_state_dict": synthetic_nerf_state(2 * seed + 1000),
        "network_fine_state_dict": synthetic_nerf_state(2 * seed + 1001),
        "embed_state_dict": {"cutoff_dist": np.full((N_JOINTS,), 0.5, np.float32),
                             "tau": np.float32(20.0)},
        "embedbones_state_dict": {},
        "embeddirs_state_dict": {"cutoff_dist": np.full((N_JOINTS,), 0.5, np.float32),
                                 "tau": np.float32(20.0)},
    }
    if alpha_gain is not None:
        for k in ("network_fn_state_dict", "network_fine_state_dict"):
            ckpt[k]["alpha_linear.weight"] = ckpt[k]["alpha_linear.weight"] * np.float32(alpha_gain)
            ckpt[k]["alpha_linear.bias"] = np.zeros_like(ckpt[k]["alpha_linear.bias"])
    if n_framecodes > 0:
        add_framecodes(ckpt["network_fn_state_dict"], n_framecodes, 7000 + 2 * seed)
        add_framecodes(ckpt["network_fine_state_dict"], n_framecodes, 7001 + 2 * seed)
    return ckpt


def calibrate_alpha_head(net_state: dict, sigma_far_max: float, gain: float = 400.0, margin: float = 0.005):
    """bf16-tier "calibrated head" recipe: given max over rays of the
    zero-bias sigma_raw at the last sample, scale the head by `gain` and bias it so the
    far sample is robustly empty.  Operates in place on a zero-bias state."""
    theta = np.float32(sigma_far_max + margin)
    net_state["alpha_linear.weight"] = net_state["alpha_linear.weight"] * np.float32(gain)
    net_state["alpha_linear.bias"] = np.full_like(net_state["alpha_linear.bias"], -np.float32(gain) * theta)
    return net_state
