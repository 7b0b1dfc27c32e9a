"""Drop-in for PoseGen's `core.raycasters` on the render path.

Same constructor, factory, call signature, return dict and checkpoint keys as the
reference (`core/raycasters.py:17-184,326-794`), but `forward` hands the whole
`render_rays` pipeline to the sm_100a library through the C ABI (posegen_b200.engine).

What is kept in Python (host-side only, mirrors the reference objects so that
`run_gan.py` / `run_nerf.py` callers keep working):
  * `NeRF`            parameter container with the reference's state_dict names
                      (core/networks/nerf.py:57-88); its math lives in the CUDA kernels;
  * `CutoffEmbedder`  tau buffer + cutoff_dist parameter and the tau schedule
                      (core/cutoff_embedder.py:83-95,176-183);
  * `Embedder`        the plain (identity for multires_bones=0) bone embedder;
  * `RayCaster`       forward / state_dict / load_state_dict / get_networks / get_embed_fns /
                      update_embed_fns (core/raycasters.py:326-794).
Unsupported reference options raise NotImplementedError (never a silent fallback).
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import Dict, Optional

import torch
import torch.nn as nn

from .engine import Engine, LINEAR_ORDER

N_JOINTS = 24


# ----------------------------------------------------------------------------- modules
class NeRF(nn.Module):
    """Parameter container of one A-NeRF MLP (surreal.txt shapes).  forward() is not
    implemented on purpose: the network only runs inside the fused CUDA kernels."""

    def __init__(self, D=8, W=256, input_ch=360, input_ch_bones=72, input_ch_views=648, output_ch=5,
                 skips=(4,), use_viewdirs=True, use_framecode=False, framecode_ch=16, n_framecodes=0,
                 skel_type=None, density_scale=1.0):
        super().__init__()
        if (D, W, input_ch, input_ch_bones, input_ch_views, tuple(skips), use_viewdirs) != (8, 256, 360, 72, 648, (4,), True):
            raise NotImplementedError("posegen_b200 implements the A-NeRF of the shipped configs only "
                                      "(8x256, skip 4, 360+72 | 648 inputs, viewdirs; optional 16-channel frame codes)")
        if use_framecode and (framecode_ch != 16 or n_framecodes <= 0):
            raise NotImplementedError("frame codes: framecode_size must be 16 (parser default, run_nerf.py:316) and n_framecodes > 0")
        self.D, self.W = D, W
        self.input_ch, self.input_ch_bones, self.input_ch_views = input_ch, input_ch_bones, input_ch_views
        self.skips, self.use_viewdirs, self.use_framecode = list(skips), use_viewdirs, use_framecode
        self.output_ch, self.skel_type, self.density_scale = output_ch, skel_type, density_scale
        self.framecode_ch, self.n_framecodes = framecode_ch, (n_framecodes if use_framecode else 0)
        self.cam_ch = 1 if use_framecode else 0
        dnet = input_ch + input_ch_bones
        layers = [nn.Linear(dnet, W)]
        for i in range(D - 1):
            layers.append(nn.Linear(W + dnet if i in self.skips else W, W))
        self.pts_linears = nn.ModuleList(layers)
        self.alpha_linear = nn.Linear(W, 1)
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + (framecode_ch if use_framecode else 0) + W, W // 2)])
        self.feature_linear = nn.Linear(W, W)
        self.rgb_linear = nn.Linear(W // 2, 3)
        if use_framecode:            # core/networks/nerf.py:87-88
            self.framecodes = Optcodes(n_framecodes, framecode_ch)

    def forward(self, *a, **k):
        raise NotImplementedError("NeRF.forward runs only inside posegen_b200's fused CUDA kernels "
                                  "(use RayCaster.forward or Engine.mlp)")


def net_tensors(net: nn.Module) -> Dict[str, torch.Tensor]:
    """'<layer>.weight' / '<layer>.bias' tensors of one NeRF by attribute access.  Unlike `named_parameters()` /
    `state_dict()` this also works on an `nn.DataParallel` replica, whose parameters are plain (non-leaf) tensor
    attributes produced by Broadcast (torch/nn/parallel/replicate.py) - gradients flow back through them."""
    out = {}
    for name in LINEAR_ORDER:
        mod = net
        for part in name.split("."):
            mod = mod[int(part)] if part.isdigit() else getattr(mod, part)
        out[f"{name}.weight"], out[f"{name}.bias"] = mod.weight, mod.bias
    if getattr(net, "use_framecode", False):
        out["framecodes.codes.weight"] = net.framecodes.codes.weight
    return out


class Optcodes(nn.Module):
    """Learned per-frame (per-camera) codes (core/networks/embedding.py:4-46): parameter container with the reference's
    state_dict name `codes.weight`; the gather / mean rule runs inside the kernels (pgn_render_inputs.cams)."""

    def __init__(self, n_codes, code_ch):
        super().__init__()
        self.n_codes, self.code_ch = n_codes, code_ch
        self.codes = nn.Embedding(n_codes, code_ch)
        nn.init.xavier_normal_(self.codes.weight)


class Embedder(nn.Module):
    """Plain positional embedder (core/cutoff_embedder.py:9-58); identity when multires=0."""

    def __init__(self, input_dims, multires):
        super().__init__()
        self.input_dims, self.multires = input_dims, multires
        self.out_dim = input_dims * (1 + 2 * multires)

    def update_threshold(self, *a, **k):
        pass

    def get_tau(self):
        return 0.0


class CutoffEmbedder(Embedder):
    """tau / cutoff_dist holder with the reference's schedule (core/cutoff_embedder.py:61-195)."""

    def __init__(self, input_dims, multires, cutoff_dist, cutoff_dim=N_JOINTS, dist_inputs=False):
        super().__init__(input_dims, multires)
        self.dist_inputs = dist_inputs
        self.cutoff_dist = nn.Parameter(torch.ones(cutoff_dim) * cutoff_dist, requires_grad=False)
        self.init_tau = 20.
        self.register_buffer("tau", torch.tensor(self.init_tau))

    def get_tau(self):
        return self.tau.item()

    def update_threshold(self, global_step, tau_step, tau_rate, alpha_step=None, alpha_target=None):
        self.update_tau(global_step, tau_step, tau_rate)

    def update_tau(self, global_step, step, rate):
        # tau = min(2000, 20 * rate^(step/(1000*cutoff_step)))  (cutoff_embedder.py:181-183)
        self.tau = (self.init_tau * torch.ones_like(self.tau) * rate ** (global_step / float(step * 1000))).clamp(max=2000.)


class _EncoderTag:
    """Stand-in for the reference encoder modules passed around in preproc_kwargs
    (core/encoders.py); only their names are observable by callers."""

    def __init__(self, name, dims):
        self.encoder_name, self.dims = name, dims


# ----------------------------------------------------------------------------- RayCaster
class RayCaster(nn.Module):
    """core.raycasters.RayCaster with the render path executed by the B200 library.

    precision: "bf16" (tcgen05 tensor-core path, default) or "fp32" (CUDA-core parity tier).
    """

    def __init__(self, network, embed_fn, embedbones_fn, embeddirs_fn, network_fine=None,
                 joint_coords=None, single_net=False, precision="bf16"):
        super().__init__()
        if single_net or network_fine is None or network_fine is network:
            raise NotImplementedError("single_net / no fine network is not implemented (surreal.txt uses two nets)")
        self.network, self.network_fine = network, network_fine
        self.embed_fn, self.embedbones_fn, self.embeddirs_fn = embed_fn, embedbones_fn, embeddirs_fn
        if joint_coords is not None:
            n_j = joint_coords.shape[-3]
            self.register_buffer("joint_coords", torch.as_tensor(joint_coords).reshape(-1, n_j, 3, 3))
        self.single_net = single_net
        self.precision = precision
        self.return_alpha = True
        self._engines = {}
        self._uploaded = {}
        self._dirty = {}

    # -- engine / weight sync ---------------------------------------------------
    @property
    def n_framecodes(self):
        return int(getattr(self.network, "n_framecodes", 0))

    def _param_signature(self):
        ps = list(net_tensors(self.network).values()) + list(net_tensors(self.network_fine).values())
        return tuple(p._version for p in ps) + tuple(p.data_ptr() for p in ps)

    def _scalar_signature(self):
        ts = (self.embed_fn.tau, self.embeddirs_fn.tau, self.embed_fn.cutoff_dist, self.embeddirs_fn.cutoff_dist)
        return tuple(t._version for t in ts) + tuple(t.data_ptr() for t in ts) + (float(self.network.density_scale),)

    def mark_weights_dirty(self):
        """Force a re-pack of the kernel-side weight copies on the next call (after in-place parameter edits that
        bypass autograd's version counters, e.g. `p.data.copy_()`)."""
        for key in self._engines:
            self._dirty[key] = True

    def engine(self, device) -> Engine:
        """The per-device C-ABI context with this module's current weights.  Weights are re-packed (device to device, on
        the current stream, no host sync) when a parameter's version counter moved (optimizer.step); the embedder
        scalars are read back to the host only when their buffers changed (update_embed_fns), so a training loop
        never blocks on the device here.  Version counters are not bumped by every optimizer (torch's fused Adam
        updates parameters without touching them), so a differentiable train-mode forward also marks the packed copy
        stale (`mark_weights_dirty`): the call after it always re-packs."""
        device = torch.device(device)
        key = device.index if device.index is not None else torch.cuda.current_device()
        if key not in self._engines:
            self._engines[key] = Engine(torch.device("cuda", key), n_framecodes=self.n_framecodes)
        eng = self._engines[key]
        sig_w, sig_s = self._param_signature(), self._scalar_signature()
        up = self._uploaded.get(key, (None, None))
        # a DataParallel replica receives freshly broadcast parameter copies every forward (possibly at recycled
        # addresses): its packed copy is always refreshed
        if up[0] != sig_w or self._dirty.pop(key, False) or getattr(self, "_is_replica", False):
            eng.upload_net(0, net_tensors(self.network))
            eng.upload_net(1, net_tensors(self.network_fine))
        if up[1] != sig_s:
            eng.set_scalars(float(self.embed_fn.tau), float(self.embeddirs_fn.tau),
                            self.embed_fn.cutoff_dist.detach().flatten().tolist(),
                            self.embeddirs_fn.cutoff_dist.detach().flatten().tolist(),
                            float(self.network.density_scale))
        self._uploaded[key] = (sig_w, sig_s)
        return eng

    # -- forward -------------------------------------------------------------------
    def forward(self, *args, fwd_type="", **kwargs):
        # routing of core/raycasters.py:349-355
        if fwd_type == "density":
            return self.render_pts_density(*args, **kwargs)
        if fwd_type == "mesh":
            return self.render_mesh_density(*args, **kwargs)
        if fwd_type == "density_color":
            raise NotImplementedError("fwd_type='density_color' needs texture_linears, which no shipped config has "
                                      "(core/raycasters.py:624-625)")
        return self.render_rays(*args, **kwargs)

    # -- density-only queries ---------------------------------------
    @torch.no_grad()
    def render_pts_density(self, pts, kps=None, skts=None, bones=None, render_kwargs=None, subject_idxs=None,
                           netchunk=1024 * 64, network=None, color=False, v=None, precision=None):
        """core/raycasters.py:597-648: raw density (alpha_linear output, before the ReLU) of arbitrary points.

        pts [N,S,3] (CUDA), skts [N,24,4,4] or [1,24,4,4] / [24,4,4] (one pose for all points).  Uses
        network_fine like the reference (`network` may be 'coarse'/'fine' or one of the two modules).
        The points run through the stage entry points pgn_encode (z = 0 along a zero direction, i.e. the
        point itself) and pgn_mlp; only the sigma channel is returned, shape [N,S,1]."""
        if color or v is not None:
            raise NotImplementedError("color / precomputed v are not part of the surreal.txt density query")
        if skts is None:
            raise ValueError("skts is required")
        if not pts.is_cuda:
            raise RuntimeError("posegen_b200.RayCaster needs CUDA tensors; there is no CPU fallback")
        dev = pts.device
        eng = self.engine(dev)
        net_id = 1
        if network is not None:
            net_id = 0 if (network is self.network or network == "coarse") else 1
        n, s = pts.shape[0], pts.shape[1]
        flat = pts.reshape(-1, 3).float().contiguous()
        sk = torch.as_tensor(skts, dtype=torch.float32, device=dev)
        shared = sk.dim() == 3 or sk.shape[0] == 1
        if shared:
            sk = sk.reshape(24, 4, 4)
        else:
            sk = sk[:, None].expand(n, s, 24, 4, 4).reshape(-1, 24, 4, 4)
        out = torch.empty((flat.shape[0], 1), dtype=torch.float32, device=dev)
        cyl = torch.zeros(5, dtype=torch.float32, device=dev)
        for i in range(0, flat.shape[0], netchunk):
            p = flat[i:i + netchunk]
            m = p.shape[0]
            rb = torch.zeros((m, 11), dtype=torch.float32, device=dev)
            rb[:, :3] = p
            rb[:, 7] = 1.0
            z = torch.zeros((m, 1), dtype=torch.float32, device=dev)
            enc = eng.encode(rb, sk if shared else sk[i:i + netchunk].contiguous(), cyl, z)
            out[i:i + netchunk] = eng.mlp(net_id, enc, precision=precision or self.precision).reshape(m, 4)[:, 3:4]
        return out.reshape(n, s, 1)

    @torch.no_grad()
    def render_mesh_density(self, kps, skts, bones=None, subject_idxs=None, radius=1.0, res=64,
                            render_kwargs=None, netchunk=1024 * 64, v=None, precision=None):
        """core/raycasters.py:579-595: raw density on a (res+1)^3 grid of half-width `radius` around the root joint."""
        import numpy as np
        kps = torch.as_tensor(kps)
        dev = next(self.parameters()).device
        t = np.linspace(-radius, radius, res + 1)
        grid = np.stack(np.meshgrid(t, t, t), axis=-1).astype(np.float32)
        sh = grid.shape
        pts = torch.tensor(grid.reshape(-1, 3)) + kps[0, 0].detach().cpu().float()
        raw = self.render_pts_density(pts.reshape(-1, 1, 3).to(dev), kps, skts, bones, render_kwargs, subject_idxs,
                                      netchunk, v=v, precision=precision)[..., :1]
        return raw.reshape(*sh[:-1]).transpose(1, 0)

    def render_rays(self, ray_batch, N_samples, kp_batch=None, skts=None, cyls=None, bones=None, cams=None,
                    subject_idxs=None, retraw=False, lindisp=False, perturb=0., N_importance=0,
                    network_fine=None, raw_noise_std=0., ray_noise_std=0., verbose=False, ext_scale=0.001,
                    pytest=False, preproc_kwargs=None, nerf_type="nerf", use_viewdirs=True,
                    precision=None, nanfill_chunk=None, **_ignored):
        train_step = self.training and torch.is_grad_enabled()
        if ray_noise_std:
            raise NotImplementedError("ray_noise_std (off in every shipped config) is not implemented: it makes a sample's "
                                      "position more than a function of (ray, z)")
        if (perturb or raw_noise_std) and not train_step:
            raise NotImplementedError("perturb / raw_noise_std (training-time sampling noise) need .train() and grad mode; "
                                      "the render path is deterministic (render_kwargs_test, core/raycasters.py:176-178)")
        if N_samples != 64 or N_importance != 16:
            raise NotImplementedError("only N_samples=64, N_importance=16 (surreal.txt) is implemented")
        if cams is not None and not self.n_framecodes:
            cams = None                # the reference appends the index column and a net without frame codes never reads it
        if skts is None or cyls is None:
            raise ValueError("skts and cyls are required (skeleton-relative encoding / cylinder near-far)")
        if not ray_batch.is_cuda:
            raise RuntimeError("posegen_b200.RayCaster needs CUDA tensors (the reference moves each chunk with "
                               ".to('cuda') in batchify_rays, core/trainer.py:70-74); there is no CPU fallback")
        if train_step:
            # training step (core/trainer.py:232-275): differentiable w.r.t. the two MLPs (posegen_b200/train.py)
            from .train import render_train
            if (precision or self.precision) != "bf16":
                raise NotImplementedError("the training step runs on the bf16 tensor-core path only")
            return render_train(self, ray_batch, skts.to(ray_batch.device), cyls.to(ray_batch.device), nanfill_chunk,
                                perturb=float(perturb), raw_noise_std=float(raw_noise_std), rand=_ignored.get("train_random"), cams=cams,
                                lindisp=bool(lindisp))
        eng = self.engine(ray_batch.device)
        n = ray_batch.shape[0]
        ret = eng.render(ray_batch.float(), skts.to(ray_batch.device).float(), cyls.to(ray_batch.device).float(),
                         nanfill_chunk=n if nanfill_chunk is None else nanfill_chunk,
                         precision=precision or self.precision, return_alpha=self.return_alpha, cams=cams, lindisp=bool(lindisp))
        # alpha / alpha0 are simply absent when not requested: the reference's batchify_rays concatenates every key
        # of the returned dict (core/trainer.py:75-80), so a None entry would break it
        eng.poll_status()
        return ret

    # -- reference API surface -------------------------------------------------------
    def get_networks(self):
        return self.network, self.network_fine

    def get_embed_fns(self):
        return self.embed_fn, self.embedbones_fn, self.embeddirs_fn

    def update_embed_fns(self, global_step, args):
        for fn in (self.embed_fn, self.embeddirs_fn, self.embedbones_fn):
            if fn is not None:
                fn.update_threshold(global_step, args.cutoff_step, args.cutoff_rate,
                                    getattr(args, "freq_schedule_step", None), args.multires - 1)

    def state_dict(self, *a, **k):
        # key naming rules of core/raycasters.py:752-766
        return {"network_fn_state_dict": self.network.state_dict(),
                "network_fine_state_dict": self.network_fine.state_dict(),
                "embed_state_dict": self.embed_fn.state_dict(),
                "embedbones_state_dict": self.embedbones_fn.state_dict() if self.embedbones_fn is not None else {},
                "embeddirs_state_dict": self.embeddirs_fn.state_dict()}

    def load_state_dict(self, ckpt, strict=True):
        """core/raycasters.py:768-788: a checkpoint dict with the five reference keys (extra keys such as global_step /
        optimizer_state_dict of a `Trainer.save_nerf` .tar are ignored here; `load_ckpt_from_path` reads them)."""
        def conv(sd):
            return {k: torch.as_tensor(v) for k, v in sd.items()}
        self.network.load_state_dict(conv(ckpt["network_fn_state_dict"]), strict=strict)
        self.network_fine.load_state_dict(conv(ckpt["network_fine_state_dict"]), strict=strict)
        for fn, key in ((self.embed_fn, "embed_state_dict"), (self.embeddirs_fn, "embeddirs_state_dict")):
            if key in ckpt:
                fn.load_state_dict(conv(ckpt[key]), strict=strict)
        self._uploaded.clear()


# ----------------------------------------------------------------------------- factory
SURREAL_ARGS = dict(
    netdepth=8, netwidth=256, multires=7, multires_views=4, multires_bones=0, i_embed=0,
    use_viewdirs=True, use_cutoff=True, cutoff_viewdir=True, cutoff_inputs=True, cutoff_bones=False,
    normalize_cutoff=False, opt_cutoff=False, cut_to_dist=False, cutoff_shift=False, freq_schedule=False,
    init_freq=0., cutoff_mm=500., ext_scale=0.001, kp_dist_type="reldist", bone_type="reldir", view_type="relray",
    pts_tr_type="local", density_type="relu", density_scale=1.0, opt_framecode=False, framecode_size=16,
    n_framecodes=None, single_net=False, N_samples=64, N_importance=16, perturb=1.0, raw_noise_std=1.0,
    ray_noise_std=0., lindisp=False, nerf_type="nerf", lrate=5e-4, chunk=4096, cutoff_step=250, cutoff_rate=10.,
    no_reload=True, ft_path=None, finetune=False, debug=False, weight_decay=None, white_bkgd=False,
)


def surreal_args(**overrides) -> SimpleNamespace:
    """The frozen configs/surreal/surreal.txt + run_nerf.py parser defaults."""
    d = dict(SURREAL_ARGS)
    d.update(overrides)
    return SimpleNamespace(**d)


def _require(args, **expected):
    for k, v in expected.items():
        got = getattr(args, k, v)
        if got != v:
            raise NotImplementedError(f"create_raycaster: {k}={got!r} is not implemented (surreal.txt uses {v!r})")


def create_raycaster(args, data_attrs, device=None, precision="bf16"):
    """core/raycasters.py:17-184: returns (render_kwargs_train, render_kwargs_test, start, grad_vars,
    optimizer, loaded_ckpt) with the B200 RayCaster inside."""
    _require(args, netdepth=8, netwidth=256, multires=7, multires_views=4, multires_bones=0, use_viewdirs=True,
             use_cutoff=True, cutoff_viewdir=True, cutoff_inputs=True, kp_dist_type="reldist", bone_type="reldir",
             view_type="relray", pts_tr_type="local", density_type="relu", single_net=False,
             N_importance=16, N_samples=64)
    use_fc = bool(getattr(args, "opt_framecode", False))
    n_fc = 0
    if use_fc:          # core/raycasters.py:22: one code per training view unless --n_framecodes overrides it
        n_fc = data_attrs["n_views"] if getattr(args, "n_framecodes", None) is None else args.n_framecodes
        _require(args, framecode_size=16)
    cutoff = args.cutoff_mm * args.ext_scale
    embed_fn = CutoffEmbedder(N_JOINTS, args.multires, cutoff, dist_inputs=False)
    embedbones_fn = Embedder(N_JOINTS * 3, args.multires_bones)
    embeddirs_fn = CutoffEmbedder(N_JOINTS * 3, args.multires_views, cutoff, dist_inputs=True)
    kw = dict(D=args.netdepth, W=args.netwidth, input_ch=embed_fn.out_dim, input_ch_bones=embedbones_fn.out_dim,
              input_ch_views=embeddirs_fn.out_dim, output_ch=5, skips=(4,), use_viewdirs=True,
              use_framecode=use_fc, framecode_ch=getattr(args, "framecode_size", 16), n_framecodes=n_fc,
              skel_type=data_attrs.get("skel_type"), density_scale=args.density_scale)
    model, model_fine = NeRF(**kw), NeRF(**kw)
    ray_caster = RayCaster(model, embed_fn, embedbones_fn, embeddirs_fn, network_fine=model_fine,
                           joint_coords=torch.as_tensor(data_attrs["joint_coords"]) if "joint_coords" in data_attrs else None,
                           single_net=False, precision=precision)
    if device is not None:
        ray_caster = ray_caster.to(device)
    grad_vars = [p for p in ray_caster.parameters() if p.requires_grad]
    optimizer = torch.optim.Adam(params=grad_vars, lr=args.lrate, betas=(0.9, 0.999))
    start, loaded_ckpt = 0, None
    # checkpoint discovery of core/raycasters.py:125-142: an explicit ft_path, else the newest *.tar of basedir/expname
    ft_path = getattr(args, "ft_path", None)
    if ft_path not in (None, "None"):
        ckpts = [ft_path]
    else:
        ckpts = []
        basedir, expname = getattr(args, "basedir", None), getattr(args, "expname", None)
        if basedir is not None and expname is not None and os.path.isdir(os.path.join(basedir, expname)):
            ckpts = [os.path.join(basedir, expname, f) for f in sorted(os.listdir(os.path.join(basedir, expname)))
                     if "tar" in f and "pose" not in f]
    if len(ckpts) > 0 and not getattr(args, "no_reload", False):
        start, ray_caster, optimizer, loaded_ckpt = load_ckpt_from_path(ray_caster, optimizer, ckpts[-1],
                                                                        getattr(args, "finetune", False))
        if getattr(args, "finetune", False):
            start = 0
    preproc_kwargs = {
        "pts_tr_fn": _EncoderTag("W2LEncoder", N_JOINTS), "kp_input_fn": _EncoderTag("RelDist", N_JOINTS),
        "view_input_fn": _EncoderTag("VecNorm", N_JOINTS * 3), "bone_input_fn": _EncoderTag("VecNorm", N_JOINTS * 3),
        "density_scale": args.density_scale, "density_fn": torch.nn.functional.relu,
    }
    # the reference hands the TRAINING kwargs an nn.DataParallel wrapper (core/raycasters.py:157) and its trainer
    # dereferences `.module` (core/trainer.py:267,272,506); the test kwargs keep the bare module (:172)
    # (without a CUDA device nn.DataParallel is a pass-through that still has `.module`)
    debug_one = getattr(args, "debug", False) and torch.cuda.is_available()
    wrapped = nn.DataParallel(ray_caster, device_ids=[0]) if debug_one else nn.DataParallel(ray_caster)
    render_kwargs_train = {
        "ray_caster": wrapped, "perturb": args.perturb, "N_importance": args.N_importance,
        "N_samples": args.N_samples, "use_viewdirs": args.use_viewdirs, "raw_noise_std": args.raw_noise_std,
        "ray_noise_std": args.ray_noise_std, "ext_scale": args.ext_scale, "preproc_kwargs": preproc_kwargs,
        "lindisp": args.lindisp, "nerf_type": args.nerf_type,
    }
    render_kwargs_test = dict(render_kwargs_train)
    render_kwargs_test.update(ray_caster=ray_caster, perturb=False, raw_noise_std=0., ray_noise_std=0.)
    optimizer.zero_grad()
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer, loaded_ckpt


def load_ckpt_from_path(ray_caster, optimizer, ckpt_path, finetune=False):
    """core/cutoff_embedder.py:227-238: weights + embedder buffers, and (unless fine-tuning) the optimizer state;
    returns (global_step, ray_caster, optimizer, ckpt)."""
    ckpt = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    global_step = ckpt["global_step"]
    ray_caster.load_state_dict(ckpt)
    if optimizer is not None and not finetune and "optimizer_state_dict" in ckpt:
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
    return global_step, ray_caster, optimizer, ckpt


def raycaster_from_checkpoint(ckpt: dict, device="cuda", precision="bf16") -> RayCaster:
    """Build a RayCaster and load a reference-format checkpoint dict (key names of
    core/raycasters.py:752-766), e.g. posegen_b200.synthetic.synthetic_raycaster_state()."""
    fc = ckpt["network_fn_state_dict"].get("framecodes.codes.weight")          # Optcodes checkpoint (h36m / mixamo / perfcap)
    args = surreal_args() if fc is None else surreal_args(opt_framecode=True, n_framecodes=int(fc.shape[0]))
    _, kw_test, _, _, _, _ = create_raycaster(args, {"skel_type": None}, device=device, precision=precision)
    rc = kw_test["ray_caster"]
    rc.load_state_dict(ckpt)
    rc.eval()
    return rc
