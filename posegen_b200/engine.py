"""Thin Python handle on a `pgn_context` (include/posegen_b200.h).

PyTorch is used only for device memory and streams: every tensor handed to the C ABI is a
contiguous fp32 CUDA tensor, passed as a raw pointer together with the current stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from .synthetic import NERF_LAYERS

T, S, I, J = 80, 64, 16, 24
# include/posegen_b200.h order of the 12 linear layers
LINEAR_ORDER = [f"pts_linears.{i}" for i in range(8)] + ["alpha_linear", "feature_linear", "views_linears.0", "rgb_linear"]
_SHAPES = {name: (o, i) for name, o, i in NERF_LAYERS}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _check_f32_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (posegen_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise ValueError(f"{name} must be float32, got {t.dtype}")


class Engine:
    """One renderer context on one CUDA device."""

    def __init__(self, device: Optional[torch.device] = None, n_framecodes: int = 0):
        """n_framecodes > 0: the nets carry Optcodes frame codes (views_linears.0 reads 920 inputs, `cams` selects a code)."""
        self.lib = _lib.load()
        self.n_framecodes = int(n_framecodes)
        if not torch.cuda.is_available():
            raise RuntimeError("posegen_b200 requires a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        cfg = _lib.Config(J, S, I, 7, 4, 8, 256, 4, self.device.index, self.n_framecodes, 16 if self.n_framecodes else 0)
        handle = C.c_void_p()
        _lib.check(self.lib.pgn_create(C.byref(cfg), C.byref(handle)))
        self.handle = handle
        self._status_poll = None       # (pinned int32 tensor, event) of the last asynchronous watchdog read-back

    def close(self):
        if getattr(self, "handle", None):
            self.lib.pgn_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ state
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def upload_net(self, net_id: int, state: Dict[str, torch.Tensor]):
        """state: '<layer>.weight' / '<layer>.bias' tensors (host or device, any float dtype)."""
        w = _lib.NetWeights()
        keep = []
        on_device = None
        for k, name in enumerate(LINEAR_ORDER):
            for kind, arr in (("weight", w.weight), ("bias", w.bias)):
                t = torch.as_tensor(state[f"{name}.{kind}"]).detach()
                exp = self._shape(name) if kind == "weight" else (_SHAPES[name][0],)
                if tuple(t.shape) != exp:
                    raise ValueError(f"{name}.{kind}: expected shape {exp}, got {tuple(t.shape)}")
                t = t.to(dtype=torch.float32).contiguous()
                if on_device is None:
                    on_device = t.is_cuda
                if t.is_cuda != on_device:
                    t = t.to(self.device) if on_device else t.cpu()
                keep.append(t)
                arr[k] = t.data_ptr()
        if self.n_framecodes:
            t = torch.as_tensor(state["framecodes.codes.weight"]).detach()
            if tuple(t.shape) != (self.n_framecodes, 16):
                raise ValueError(f"framecodes.codes.weight: expected {(self.n_framecodes, 16)}, got {tuple(t.shape)}")
            t = t.to(dtype=torch.float32).contiguous()
            t = (t.to(self.device) if on_device else t.cpu()) if t.is_cuda != on_device else t
            keep.append(t)
            w.framecodes = t.data_ptr()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pgn_upload_weights(self.handle, net_id, C.byref(w), 1 if on_device else 0, self._stream()))
            if not on_device:
                torch.cuda.current_stream(self.device).synchronize()   # host staging must outlive the copies

    def _shape(self, name):
        """nn.Linear weight shape of layer `name` (views_linears.0 has 16 more inputs with frame codes)."""
        sh = _SHAPES[name]
        return (sh[0], sh[1] + 16) if (name == "views_linears.0" and self.n_framecodes) else sh

    def set_scalars(self, tau_v: float, tau_d: float, cutoff_v, cutoff_d, density_scale: float = 1.0, rgb_eps: float = 1e-3):
        cv = (C.c_float * J)(*[float(x) for x in cutoff_v])
        cd = (C.c_float * J)(*[float(x) for x in cutoff_d])
        _lib.check(self.lib.pgn_set_embed_scalars(self.handle, float(tau_v), float(tau_d), cv, cd,
                                                  float(density_scale), float(rgb_eps)))

    def load_checkpoint(self, ckpt: dict, density_scale: float = 1.0):
        """ckpt with the reference key names (core/raycasters.py:752-766)."""
        self.upload_net(0, ckpt["network_fn_state_dict"])
        self.upload_net(1, ckpt["network_fine_state_dict"])
        e, d = ckpt["embed_state_dict"], ckpt["embeddirs_state_dict"]
        self.set_scalars(float(torch.as_tensor(e["tau"])), float(torch.as_tensor(d["tau"])),
                         torch.as_tensor(e["cutoff_dist"]).flatten().tolist(),
                         torch.as_tensor(d["cutoff_dist"]).flatten().tolist(), density_scale)

    @property
    def launch_count(self) -> int:
        return int(self.lib.pgn_launch_count(self.handle))

    def check_status(self):
        """Synchronous watchdog check (PGN_E_KERNEL if a bounded device-side wait gave up since the last check)."""
        _lib.check(self.lib.pgn_check_device_status(self.handle))

    def poll_status(self):
        """Asynchronous watchdog check for the hot paths: queues a 4-byte read-back of the device status word behind the
        work issued so far and raises for the PREVIOUS read-back if it has completed and latched a code.  Never blocks."""
        prev = self._status_poll
        if prev is not None and prev[1].query():
            code = int(prev[0][0])
            if code != 0:
                self._status_poll = None
                self.check_status()                       # clears the latch and raises with the library's message
                raise _lib.PosegenError(_lib.PGN_E_KERNEL, f"device watchdog tripped: pipeline wait code {code}")
            prev = None
        if prev is None:
            if torch.cuda.is_current_stream_capturing():
                return
            host = torch.empty(1, dtype=torch.int32, pin_memory=True)
            with torch.cuda.device(self.device):
                host.copy_(self._status_tensor(), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))
            self._status_poll = (host, ev)

    def _status_tensor(self):
        if getattr(self, "_status_view", None) is None:
            ptr = int(self.lib.pgn_device_status_ptr(self.handle))

            class _Raw:
                __cuda_array_interface__ = {"shape": (1,), "typestr": "<i4", "data": (ptr, False), "version": 2}
            self._status_view = torch.as_tensor(_Raw(), device=self.device)
        return self._status_view

    def _scratch(self, n_rays: int, dev) -> torch.Tensor:
        """The library's per-call workspace (8 B per ray: near/far).  Allocated per call from torch's caching allocator
        instead of a cached, growable buffer: a CUDA-graph capture bakes the pointer into the graph and gets it from
        the graph's private pool, so a later, larger eager call can never free memory a captured step still uses."""
        return torch.empty(int(self.lib.pgn_workspace_bytes(self.handle, n_rays)), dtype=torch.uint8, device=dev)

    # ------------------------------------------------------------------ inputs
    def _inputs(self, ray_batch, skts, cyls, pose_idx=None, nanfill_chunk=0, precision="bf16", chunk_starts=None, cams=None,
                lindisp=False):
        _check_f32_cuda(ray_batch, "ray_batch")
        n = ray_batch.shape[0]
        if ray_batch.dim() != 2 or ray_batch.shape[1] != 11:
            raise ValueError(f"ray_batch must be [N,11], got {tuple(ray_batch.shape)}")
        ray_batch = ray_batch.contiguous()
        keep = [ray_batch]
        inp = _lib.RenderInputs()
        inp.ray_batch = ray_batch.data_ptr()
        inp.n_rays = n

        def per_ray(t, tail, name):
            """-> (tensor, stride) accepting [*tail], [1,*tail] / stride-0 expands, or [N,*tail]."""
            _check_f32_cuda(t, name)
            numel = 1
            for x in tail:
                numel *= x
            if tuple(t.shape) == tuple(tail):
                return t.contiguous(), 0
            if tuple(t.shape[1:]) != tuple(tail):
                raise ValueError(f"{name}: expected [...,{tail}], got {tuple(t.shape)}")
            if t.shape[0] == 1 or (t.shape[0] == n and t.stride(0) == 0):
                return t[0].contiguous(), 0
            if t.shape[0] != n:
                raise ValueError(f"{name}: leading dim {t.shape[0]} != n_rays {n}")
            return t.contiguous(), numel

        if pose_idx is not None:
            _check_f32_cuda(skts, "skts")
            skts_t, cyls_t = skts.contiguous(), cyls.contiguous()
            pose_idx = pose_idx.to(device=ray_batch.device, dtype=torch.int32).contiguous()
            if pose_idx.numel() != n:
                raise ValueError("pose_idx must have one entry per ray")
            inp.pose_idx = pose_idx.data_ptr()
            keep.append(pose_idx)
            s_stride = c_stride = 0
        else:
            skts_t, s_stride = per_ray(skts, (J, 4, 4), "skts")
            cyls_t, c_stride = per_ray(cyls, (5,), "cyls")
        keep += [skts_t, cyls_t]
        inp.skts, inp.skts_stride = skts_t.data_ptr(), s_stride
        inp.cyls, inp.cyls_stride = cyls_t.data_ptr(), c_stride
        inp.nanfill_chunk = int(nanfill_chunk)
        if chunk_starts is not None:         # explicit chunk table (multi-image batches): int64 [n_chunks + 1] on the device
            if chunk_starts.dtype != torch.int64 or not chunk_starts.is_cuda or not chunk_starts.is_contiguous() or chunk_starts.numel() < 2:
                raise ValueError("chunk_starts must be a contiguous CUDA int64 [n_chunks + 1] tensor")
            inp.chunk_starts, inp.n_chunks = chunk_starts.data_ptr(), chunk_starts.numel() - 1
            keep.append(chunk_starts)
        if cams is not None and self.n_framecodes:      # frame-code index per ray (Optcodes); None = the mean code
            cams = torch.as_tensor(cams).reshape(-1).to(device=ray_batch.device, dtype=torch.int32).contiguous()
            if cams.numel() != n:
                raise ValueError("cams must have one entry per ray")
            inp.cams = cams.data_ptr()
            keep.append(cams)
        inp.lindisp = 1 if lindisp else 0
        inp.precision = {"fp32": _lib.PRECISION_FP32, "bf16": _lib.PRECISION_BF16}[precision]
        return inp, keep

    # ---------------------------------------------------------------- hot path
    def render(self, ray_batch, skts, cyls, pose_idx=None, nanfill_chunk=0, precision="bf16",
               return_alpha=True, taps=False, chunk_starts=None, cams=None, lindisp=False) -> Dict[str, torch.Tensor]:
        """pgn_render_forward: the reference's RayCaster.render_rays (eval path)."""
        inp, keep = self._inputs(ray_batch, skts, cyls, pose_idx, nanfill_chunk, precision, chunk_starts, cams, lindisp)
        n, dev = inp.n_rays, ray_batch.device
        f = lambda *sh: torch.empty(sh, dtype=torch.float32, device=dev)  # noqa: E731
        ret = {"rgb_map": f(n, 3), "disp_map": f(n), "acc_map": f(n), "rgb0": f(n, 3), "disp0": f(n), "acc0": f(n)}
        if return_alpha:
            ret["alpha"], ret["alpha0"] = f(n, T), f(n, S)
        if taps:
            ret.update(z_samples=f(n, I), z_fine=f(n, T), weights0=f(n, S), raw0=f(n, S, 4), raw=f(n, T, 4),
                       near_far=f(n, 2), pdf_inds=torch.empty((n, I), dtype=torch.int32, device=dev))
        out = _lib.RenderOutputs()
        for k, v in ret.items():
            setattr(out, k, v.data_ptr())
        ws = self._scratch(n, dev)
        with torch.cuda.device(dev):
            _lib.check(self.lib.pgn_render_forward(self.handle, C.byref(inp), C.byref(out),
                                                   C.c_void_p(ws.data_ptr()), ws.numel(),
                                                   self._stream()))
        return ret

    def render_train(self, ray_batch, skts, cyls, pose_idx=None, nanfill_chunk=0, rand=None, dump_coarse=True, cams=None,
                     lindisp=False):
        """pgn_render_forward_train: the fused bf16 forward + per-layer activation dump for the weight gradients.
        Returns (outputs incl. the taps the backward needs, {"c": dump, "f": dump}); dump[p] is a flat bf16 buffer of
        rows * 2304 elements: layers 0-7 tile-blocked [rows/128][32][128][8] each (the kernel's coalesced operand-image
        stores), then the view layer row-major [rows,128], then the ReLU mask bits of layers 0-7 (`train.act_layer`
        returns a row-major copy / view, `train.act_masks` the mask area), rows in (ray, sample) order.
        rand: optional dict of CUDA fp32 tensors t_rand [n,64], u_is [n,16], noise0 [n,64], noise [n,80] (training-time
        randomness drawn by the caller; missing keys = deterministic)."""
        inp, keep = self._inputs(ray_batch, skts, cyls, pose_idx, nanfill_chunk, "bf16", cams=cams, lindisp=lindisp)
        n, dev = inp.n_rays, ray_batch.device
        f = lambda *sh: torch.empty(sh, dtype=torch.float32, device=dev)  # noqa: E731
        ret = {"rgb_map": f(n, 3), "disp_map": f(n), "acc_map": f(n), "rgb0": f(n, 3), "disp0": f(n), "acc0": f(n),
               "z_fine": f(n, T), "raw0": f(n, S, 4), "raw": f(n, T, 4), "near_far": f(n, 2)}
        out = _lib.RenderOutputs()
        for k, v in ret.items():
            setattr(out, k, v.data_ptr())
        acts = {}
        for key, p in (("c", 0), ("f", 1)):
            if key == "c" and not dump_coarse:
                acts[key] = None
                continue
            nbytes = self.lib.pgn_activation_dump_bytes(n, p)
            acts[key] = torch.empty((nbytes // 2,), dtype=torch.bfloat16, device=dev)
        ws = self._scratch(n, dev)
        with torch.cuda.device(dev):
            rnd = None
            if rand:
                rnd = _lib.TrainRandom()
                for k, cols in (("t_rand", S), ("u_is", I), ("noise0", S), ("noise", T)):
                    v = rand.get(k)
                    if v is not None:
                        _check_f32_cuda(v, k)
                        if tuple(v.shape) != (n, cols) or not v.is_contiguous():
                            raise ValueError(f"rand[{k!r}] must be a contiguous [{n},{cols}] tensor")
                        setattr(rnd, k, v.data_ptr())
            _lib.check(self.lib.pgn_render_forward_train(self.handle, C.byref(inp), C.byref(out), _ptr(acts["c"]), _ptr(acts["f"]),
                                                         C.byref(rnd) if rnd is not None else None,
                                                         C.c_void_p(ws.data_ptr()), ws.numel(),
                                                         self._stream()))
        return ret, acts

    def render_masks(self, ray_batch, skts, cyls, pose_idx=None, nanfill_chunk=0):
        """pgn_render_forward_masks: the fused bf16 forward that keeps only the fine pass's ReLU masks (272 B per sample),
        for the pose gradient through a frozen network.  Returns (outputs incl. raw / z_fine, (trunk_mask, view_mask)):
        trunk_mask int32 [8, 8, rows] (word planes: word w of a layer = the bits [column 32 w + b > 0] of every row),
        view_mask int32 [4, rows]; rows >= n * 80, samples in (ray, sample) order."""
        inp, keep = self._inputs(ray_batch, skts, cyls, pose_idx, nanfill_chunk, "bf16")
        n, dev = inp.n_rays, ray_batch.device
        f = lambda *sh: torch.empty(sh, dtype=torch.float32, device=dev)  # noqa: E731
        ret = {"rgb_map": f(n, 3), "disp_map": f(n), "acc_map": f(n), "rgb0": f(n, 3), "disp0": f(n), "acc0": f(n),
               "z_fine": f(n, T), "raw": f(n, T, 4)}
        out = _lib.RenderOutputs()
        for k, v in ret.items():
            setattr(out, k, v.data_ptr())
        nbytes = self.lib.pgn_mask_dump_bytes(n)
        rows = nbytes // 272
        buf = torch.empty((nbytes // 4,), dtype=torch.int32, device=dev)
        ws = self._scratch(n, dev)
        with torch.cuda.device(dev):
            _lib.check(self.lib.pgn_render_forward_masks(self.handle, C.byref(inp), C.byref(out), _ptr(buf),
                                                         C.c_void_p(ws.data_ptr()), ws.numel(),
                                                         self._stream()))
        return ret, (buf[:rows * 64].view(8, 8, rows), buf[rows * 64:].view(4, rows))

    def view_delta_from_mask(self, d_raw, w_rgb, view_mask):
        """pgn_view_delta_from_mask: dG bf16 [m,128] = [g > 0] * (d_rgb W_rgb) from the view layer's mask bits."""
        m = d_raw.shape[0]
        if d_raw.dtype != torch.float32 or not d_raw.is_contiguous() or tuple(d_raw.shape) != (m, 4) or not d_raw.is_cuda:
            raise ValueError("d_raw must be contiguous CUDA fp32 [m,4]")
        if view_mask.dtype != torch.int32 or not view_mask.is_contiguous() or tuple(view_mask.shape) != (4, m):
            raise ValueError("view_mask must be contiguous int32 [4,m] (word planes)")
        w = w_rgb.detach().float().contiguous()
        dG = torch.empty((m, 128), dtype=torch.bfloat16, device=d_raw.device)
        with torch.cuda.device(d_raw.device):
            _lib.check(self.lib.pgn_view_delta_from_mask(self.handle, _ptr(dG), _ptr(d_raw), _ptr(w), _ptr(view_mask), m, self._stream()))
        return dG

    # ------------------------------------------------------ stage entry points
    def near_far(self, ray_batch, skts, cyls, nanfill_chunk=0):
        inp, keep = self._inputs(ray_batch, skts, cyls, None, nanfill_chunk)
        out = torch.empty((inp.n_rays, 2), dtype=torch.float32, device=ray_batch.device)
        _lib.check(self.lib.pgn_near_far(self.handle, C.byref(inp), _ptr(out), self._stream()))
        return out

    def encode(self, ray_batch, skts, cyls, z):
        inp, keep = self._inputs(ray_batch, skts, cyls)
        z = z.contiguous()
        enc = torch.empty((inp.n_rays, z.shape[1], 1080), dtype=torch.float32, device=z.device)
        _lib.check(self.lib.pgn_encode(self.handle, C.byref(inp), _ptr(z), z.shape[1], _ptr(enc), self._stream()))
        return enc

    def encode_bf16(self, ray_batch, skts, cyls, z):
        """The same encoding rounded to bf16 [n, n_z, 1080] (operand of the weight-gradient GEMMs)."""
        inp, keep = self._inputs(ray_batch, skts, cyls)
        z = z.contiguous()
        enc = torch.empty((inp.n_rays, z.shape[1], 1080), dtype=torch.bfloat16, device=z.device)
        _lib.check(self.lib.pgn_encode_bf16(self.handle, C.byref(inp), _ptr(z), z.shape[1], _ptr(enc), self._stream()))
        return enc

    def mlp_delta(self, dh, act, rs=None, wr=None, has_input=True, want_wsum=False):
        """pgn_mlp_delta: in place dh <- [act > 0] * ((has_input ? dh : 0) + rs @ wr); returns (column sums of the new dh,
        rs^T @ act or None).  dh / act: bf16 [m, 256 | 128] contiguous; rs: fp32 [m, nrs] view with unit inner stride
        (row stride arbitrary); wr: fp32 [nrs, cols] contiguous."""
        m, cols = dh.shape
        if dh.dtype != torch.bfloat16 or not dh.is_contiguous() or not dh.is_cuda:
            raise ValueError("dh must be a contiguous CUDA bf16 matrix")
        if act is not None and (act.dtype != torch.bfloat16 or not act.is_contiguous() or act.shape != dh.shape):
            raise ValueError("act must be a contiguous bf16 matrix of dh's shape")
        nrs = 0 if rs is None else rs.shape[1]
        if nrs:
            if rs.dtype != torch.float32 or rs.stride(1) != 1 or rs.shape[0] != m:
                raise ValueError("rs must be fp32 [m, nrs] with unit inner stride")
            wr = wr.detach().float().contiguous()
            if wr.shape != (nrs, cols):
                raise ValueError("wr must be [nrs, cols]")
        colsum = torch.empty((cols,), dtype=torch.float32, device=dh.device)
        wsum = torch.empty((nrs, cols), dtype=torch.float32, device=dh.device) if (want_wsum and nrs) else None
        with torch.cuda.device(dh.device):
            _lib.check(self.lib.pgn_mlp_delta(self.handle, _ptr(dh), 1 if has_input else 0, _ptr(act), m, cols,
                                              _ptr(rs) if nrs else None, rs.stride(0) if nrs else 0, nrs, _ptr(wr) if nrs else None,
                                              _ptr(colsum), _ptr(wsum), self._stream()))
        return colsum, wsum

    def mlp_delta_chain(self, dG, d_raw, mask, mask_rows, wstream, w_alpha, layer_mask=0xFF):
        """pgn_mlp_delta_chain: dG bf16 [m,128], d_raw fp32 [m,4], mask = the mask area of the activation dump ->
        (dz bf16 [8,m,256] with dz[l] = dZ_l, colsum fp32 [8,256] = the trunk's bias gradients); only the layers in
        layer_mask (bit l) are written / summed."""
        m = dG.shape[0]
        if dG.dtype != torch.bfloat16 or not dG.is_contiguous() or dG.shape[1] != 128 or not dG.is_cuda:
            raise ValueError("dG must be a contiguous CUDA bf16 [m,128] matrix")
        if d_raw.dtype != torch.float32 or not d_raw.is_contiguous() or tuple(d_raw.shape) != (m, 4):
            raise ValueError("d_raw must be contiguous fp32 [m,4]")
        if wstream.dtype != torch.bfloat16 or wstream.numel() != 120 * 4096 or not wstream.is_contiguous():
            raise ValueError("wstream must be the 120-slab bf16 weight stream (train.chain_wstream)")
        if mask.numel() * mask.element_size() < mask_rows * 256 or mask_rows < m:
            raise ValueError("mask area too small")
        dz = torch.empty((8, m, 256), dtype=torch.bfloat16, device=dG.device)
        colsum = torch.empty((8, 256), dtype=torch.float32, device=dG.device)
        with torch.cuda.device(dG.device):
            _lib.check(self.lib.pgn_mlp_delta_chain(self.handle, _ptr(dG), _ptr(d_raw), _ptr(mask), mask_rows, m, _ptr(wstream),
                                                    _ptr(w_alpha), _ptr(dz), _ptr(colsum), int(layer_mask), self._stream()))
        return dz, colsum

    def mlp_delta_chain_net(self, net_id, dG, d_raw, mask, mask_rows, layer_mask=0xFF):
        """pgn_mlp_delta_chain_net: the delta chain with the weights of uploaded net `net_id` (0 coarse, 1 fine)."""
        m = dG.shape[0]
        if dG.dtype != torch.bfloat16 or not dG.is_contiguous() or dG.shape[1] != 128 or not dG.is_cuda:
            raise ValueError("dG must be a contiguous CUDA bf16 [m,128] matrix")
        if d_raw.dtype != torch.float32 or not d_raw.is_contiguous() or tuple(d_raw.shape) != (m, 4):
            raise ValueError("d_raw must be contiguous fp32 [m,4]")
        if mask.numel() * mask.element_size() < mask_rows * 256 or mask_rows < m or not mask.is_contiguous():
            raise ValueError("mask area too small")
        dz = torch.empty((8, m, 256), dtype=torch.bfloat16, device=dG.device)
        colsum = torch.empty((8, 256), dtype=torch.float32, device=dG.device)
        with torch.cuda.device(dG.device):
            _lib.check(self.lib.pgn_mlp_delta_chain_net(self.handle, int(net_id), _ptr(dG), _ptr(d_raw), _ptr(mask), mask_rows, m,
                                                        _ptr(dz), _ptr(colsum), int(layer_mask), self._stream()))
        return dz, colsum

    def weight_grad_floats(self) -> int:
        """Floats of the flat weight-gradient buffer of one net (pgn_weight_grad_floats)."""
        return int(self.lib.pgn_weight_grad_floats(self.handle))

    def mlp_weight_grads(self, net_id, dz, dG, acts, enc, d_raw, bias_v, out=None):
        """pgn_mlp_weight_grads: every weight gradient of one NeRF MLP as one split-K tcgen05 kernel (+ two small fold
        kernels).  dz bf16 [8,m,256], dG bf16 [m,128], acts = the pass's activation dump (flat bf16), enc bf16 [m,1080],
        d_raw fp32 [m,4], bias_v fp32 [128].  Returns ({'<layer>.weight': fp32 view}, feature_linear.bias grad [256]); the
        views share one flat buffer in the C ABI's layer order (rgb_linear.weight is left zero: `mlp_delta` produces it);
        `out`: an fp32 CUDA buffer of `weight_grad_floats()` elements to use as that flat buffer (the training step's
        gradient arena, so that one all-reduce covers every gradient)."""
        m = dG.shape[0]
        for t, shape, dt in ((dz, (8, m, 256), torch.bfloat16), (dG, (m, 128), torch.bfloat16), (enc, (m, 1080), torch.bfloat16),
                             (d_raw, (m, 4), torch.float32), (bias_v, (128,), torch.float32)):
            if tuple(t.shape) != shape or t.dtype != dt or not t.is_contiguous() or not t.is_cuda:
                raise ValueError(f"mlp_weight_grads: expected a contiguous CUDA {dt} tensor of shape {shape}, got {tuple(t.shape)} {t.dtype}")
        if acts.dtype != torch.bfloat16 or not acts.is_contiguous():
            raise ValueError("acts must be the flat bf16 activation dump")
        rows = acts.numel() // 2304
        n = int(self.lib.pgn_weight_grad_floats(self.handle))
        if out is not None and (out.numel() != n or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dG.device):
            raise ValueError("out must be a contiguous fp32 buffer of weight_grad_floats() elements on the operands' device")
        flat = torch.empty((n,), dtype=torch.float32, device=dG.device) if out is None else out
        fb = torch.empty((256,), dtype=torch.float32, device=dG.device)
        _lib.check(self.lib.pgn_mlp_weight_grads(self.handle, int(net_id), _ptr(dz), _ptr(dG), _ptr(acts), rows, _ptr(enc), m,
                                                 _ptr(d_raw), _ptr(bias_v), _ptr(flat), _ptr(fb), self._stream()))
        out, o = {}, 0
        for name in LINEAR_ORDER:
            sh = self._shape(name)
            out[f"{name}.weight"] = flat[o:o + sh[0] * sh[1]].view(sh)
            o += sh[0] * sh[1]
        return out, fb

    def mlp_input_grads(self, net_id, dz, dG, tile_blocked=False):
        """pgn_mlp_input_grads: (g_xp bf16 [m,432], g_d bf16 [m,648]) = dL/d(network input) from dz bf16 [8,m,256] (layers 0
        and 5) and dG bf16 [m,128], with the uploaded weights of net `net_id` - tcgen05, no library GEMM.
        tile_blocked: the outputs are flat buffers of whole 128-row tiles, [ceil(m/128)][cols/8][128][8] (the operand form
        of `encode_backward_bf16(..., tile_blocked=True)`; `from_tile_blocked` turns one back into [m, cols])."""
        m = dG.shape[0]
        for t, shape in ((dz, (8, m, 256)), (dG, (m, 128))):
            if tuple(t.shape) != shape or t.dtype != torch.bfloat16 or not t.is_contiguous() or not t.is_cuda:
                raise ValueError(f"mlp_input_grads: expected a contiguous CUDA bf16 tensor of shape {shape}, got {tuple(t.shape)} {t.dtype}")
        rows = (m + 127) // 128 * 128 if tile_blocked else m
        g_xp = torch.empty((rows * 432,) if tile_blocked else (m, 432), dtype=torch.bfloat16, device=dG.device)
        g_d = torch.empty((rows * 648,) if tile_blocked else (m, 648), dtype=torch.bfloat16, device=dG.device)
        _lib.check(self.lib.pgn_mlp_input_grads(self.handle, int(net_id), _ptr(dz), _ptr(dG), m, _ptr(g_xp), _ptr(g_d),
                                                1 if tile_blocked else 0, self._stream()))
        return g_xp, g_d

    @staticmethod
    def from_tile_blocked(buf, m, cols):
        """[m, cols] view-copy of a tile-blocked buffer [tiles][cols/8][128][8]."""
        return buf.view(-1, cols // 8, 128, 8).permute(0, 2, 1, 3).reshape(-1, cols)[:m]

    def framecode_backward(self, net_id, dG, n_rays, n_z, cams, g_view_weight):
        """pgn_framecode_backward: adds dGr^T code[cam] into columns 904..919 of g_view_weight [128,920] (in place) and
        returns the gradient of framecodes.codes.weight [n_framecodes,16]."""
        if tuple(g_view_weight.shape) != (128, 920) or not g_view_weight.is_contiguous() or g_view_weight.dtype != torch.float32:
            raise ValueError("g_view_weight must be a contiguous fp32 [128,920] tensor")
        g_codes = torch.zeros((self.n_framecodes, 16), dtype=torch.float32, device=dG.device)
        c = None if cams is None else torch.as_tensor(cams).reshape(-1).to(device=dG.device, dtype=torch.int32).contiguous()
        _lib.check(self.lib.pgn_framecode_backward(self.handle, int(net_id), _ptr(dG), int(n_rays), int(n_z), _ptr(c), _ptr(g_view_weight),
                                                   _ptr(g_codes), self._stream()))
        return g_codes

    def debug_wgrad(self, A, B, Ma, Nb, n_ctas=8, out=None, b_tile_blocked=False):
        """pgn_debug_wgrad: out[Ma, Nb] += A[:, :Ma]^T B[:, :Nb] (bf16 operands, fp32 result) through the split-K kernel.
        A row-major [m, >=Ma]; B row-major [m, >=Nb], or with b_tile_blocked a flat 256-column buffer in the activation
        dump's tile-blocked layout (`train.to_tile_blocked`)."""
        m = A.shape[0]
        if out is None:
            out = torch.zeros((Ma, Nb), dtype=torch.float32, device=A.device)
        _lib.check(self.lib.pgn_debug_wgrad(self.handle, _ptr(A), A.stride(0), Ma, _ptr(B), 256 if b_tile_blocked else B.stride(0), Nb, m,
                                            _ptr(out), out.stride(0), int(n_ctas), 1 if b_tile_blocked else 0, self._stream()))
        return out

    def mlp(self, net_id, enc, precision="bf16"):
        _check_f32_cuda(enc, "enc")
        enc2 = enc.reshape(-1, 1080).contiguous()
        raw = torch.empty((enc2.shape[0], 4), dtype=torch.float32, device=enc.device)
        prec = {"fp32": _lib.PRECISION_FP32, "bf16": _lib.PRECISION_BF16}[precision]
        _lib.check(self.lib.pgn_mlp(self.handle, net_id, _ptr(enc2), enc2.shape[0], _ptr(raw), prec, self._stream()))
        return raw.reshape(*enc.shape[:-1], 4)

    def composite(self, ray_batch, skts, cyls, raw, z):
        inp, keep = self._inputs(ray_batch, skts, cyls)
        n, s = z.shape
        dev = z.device
        f = lambda *sh: torch.empty(sh, dtype=torch.float32, device=dev)  # noqa: E731
        ret = {"rgb_map": f(n, 3), "disp_map": f(n), "acc_map": f(n), "weights": f(n, s), "alpha": f(n, s)}
        _lib.check(self.lib.pgn_composite(self.handle, C.byref(inp), _ptr(raw.contiguous()), _ptr(z.contiguous()), s,
                                          _ptr(ret["rgb_map"]), _ptr(ret["disp_map"]), _ptr(ret["acc_map"]),
                                          _ptr(ret["weights"]), _ptr(ret["alpha"]), self._stream()))
        return ret

    def composite_backward(self, ray_batch, skts, cyls, raw, z, g_rgb, g_acc=None, noise=None):
        """dL/d raw [n,s,4] of raw2outputs given dL/d rgb_map [n,3] and dL/d acc_map [n]."""
        inp, keep = self._inputs(ray_batch, skts, cyls)
        d_raw = torch.empty_like(raw, memory_format=torch.contiguous_format)
        _lib.check(self.lib.pgn_composite_backward(self.handle, C.byref(inp), _ptr(raw.contiguous()), _ptr(z.contiguous()), z.shape[1],
                                                   _ptr(g_rgb.contiguous()), _ptr(g_acc.contiguous()) if g_acc is not None else None,
                                                   _ptr(noise.contiguous()) if noise is not None else None,
                                                   _ptr(d_raw), self._stream()))
        return d_raw

    def encode_backward(self, ray_batch, skts, cyls, z, g_enc):
        """dL/d skts [n,24,4,4] (per ray) given dL/d(network input) [n, n_z, 1080]."""
        inp, keep = self._inputs(ray_batch, skts, cyls)
        z = z.contiguous()
        d = torch.empty((inp.n_rays, 24, 4, 4), dtype=torch.float32, device=z.device)
        _lib.check(self.lib.pgn_encode_backward(self.handle, C.byref(inp), _ptr(z), z.shape[1], _ptr(g_enc.contiguous()), _ptr(d),
                                                self._stream()))
        return d

    def encode_backward_bf16(self, ray_batch, skts, cyls, z, g_xp, g_d, tile_blocked=False):
        """dL/d skts [n,24,4,4] (per ray) given dL/d(network input) as two bf16 matrices [n * n_z, 432] and [n * n_z, 648]
        (tile_blocked: the flat whole-tile buffers `mlp_input_grads(..., tile_blocked=True)` returns)."""
        inp, keep = self._inputs(ray_batch, skts, cyls)
        z = z.contiguous()
        rows = inp.n_rays * z.shape[1]
        for t, cols in ((g_xp, 432), (g_d, 648)):
            want = ((rows + 127) // 128 * 128 * cols,) if tile_blocked else (rows, cols)
            if t.dtype != torch.bfloat16 or not t.is_contiguous() or tuple(t.shape) != want:
                raise ValueError(f"expected a contiguous bf16 {want} tensor, got {tuple(t.shape)} {t.dtype}")
        d = torch.empty((inp.n_rays, 24, 4, 4), dtype=torch.float32, device=z.device)
        _lib.check(self.lib.pgn_encode_backward_bf16(self.handle, C.byref(inp), _ptr(z), z.shape[1], _ptr(g_xp), _ptr(g_d),
                                                     1 if tile_blocked else 0, _ptr(d), self._stream()))
        return d

    def sample_pdf(self, z, weights):
        _check_f32_cuda(z, "z")
        n, dev = z.shape[0], z.device
        ret = {"z_samples": torch.empty((n, I), dtype=torch.float32, device=dev),
               "z_sorted": torch.empty((n, T), dtype=torch.float32, device=dev),
               "pdf_inds": torch.empty((n, I), dtype=torch.int32, device=dev),
               "sorted_idxs": torch.empty((n, T), dtype=torch.int32, device=dev)}
        _lib.check(self.lib.pgn_sample_pdf(self.handle, _ptr(z.contiguous()), _ptr(weights.contiguous()), n,
                                           _ptr(ret["z_samples"]), _ptr(ret["z_sorted"]), _ptr(ret["pdf_inds"]),
                                           _ptr(ret["sorted_idxs"]), self._stream()))
        return ret

    def generate_rays(self, H, W, focal, c2w, x0, y0, x1, y1):
        c = (C.c_float * 12)(*[float(v) for v in torch.as_tensor(c2w)[:3, :4].reshape(-1).tolist()])
        n = max(x1 - x0, 0) * max(y1 - y0, 0)
        rb = torch.empty((n, 11), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.pgn_generate_rays(self.handle, H, W, float(focal), c, x0, y0, x1, y1, _ptr(rb), self._stream()))
        return rb

    def compose_frame(self, H, W, x0, y0, x1, y1, rgb_map, acc_map, bg=1.0):
        img = torch.empty((H * W, 3), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.pgn_compose_frame(self.handle, H, W, x0, y0, x1, y1, _ptr(rgb_map), _ptr(acc_map), float(bg),
                                              _ptr(img), self._stream()))
        return img.view(H, W, 3)

    PHASES = ["issuer_wait_weights", "issuer_wait_staging", "issuer_wait_act", "issuer_total", "producer_wait_slot",
              "producer_total", "encode_x", "encode_d", "epilogue", "wait_acc", "wait_stg_free", "composite", "compute_total",
              "wait_act_free", "tables", "store_arrive",
              "wacc_L0", "wacc_L1", "wacc_L2", "wacc_L3", "wacc_L4", "wacc_L5", "wacc_L6", "wacc_L7", "wacc_V",
              "pre_L0", "pre_L5", "pre_V", "hidden_issue", "hidden_complete", "-", "-"]

    def pose_to_skts(self, bones, rest_pose, ext_scale=0.001, extend_mm=250., top_expand_ratio=1.6, bot_expand_ratio=1.1,
                     return_l2ws=False):
        """Device FK: bones [B,24,3] axis-angle (CUDA) -> skts [B,24,4,4], kps [B,24,3], cyls [B,5]."""
        _check_f32_cuda(bones, "bones")
        b = bones.contiguous()
        n = b.shape[0]
        rest = (C.c_float * 72)(*[float(v) for v in torch.as_tensor(rest_pose, dtype=torch.float32).reshape(-1).tolist()])
        skts = torch.empty((n, 24, 4, 4), dtype=torch.float32, device=b.device)
        kps = torch.empty((n, 24, 3), dtype=torch.float32, device=b.device)
        cyls = torch.empty((n, 5), dtype=torch.float32, device=b.device)
        l2ws = torch.empty((n, 24, 4, 4), dtype=torch.float32, device=b.device) if return_l2ws else None
        _lib.check(self.lib.pgn_pose_to_skts(self.handle, _ptr(b), rest, n, float(extend_mm * ext_scale), float(top_expand_ratio),
                                             float(bot_expand_ratio), _ptr(skts), _ptr(kps), _ptr(cyls), _ptr(l2ws), self._stream()))
        return (skts, kps, cyls, l2ws) if return_l2ws else (skts, kps, cyls)

    def pose_fk_backward(self, bones, rest_pose, g_skts, g_kps=None):
        """pgn_pose_fk_backward: dL/d bones [B,24,3] from dL/d skts [B,24,4,4] (+ optional dL/d kps [B,24,3])."""
        _check_f32_cuda(bones, "bones")
        b = bones.contiguous()
        n = b.shape[0]
        gs = g_skts.float().contiguous()
        gk = g_kps.float().contiguous() if g_kps is not None else None
        if tuple(gs.shape) != (n, 24, 4, 4) or (gk is not None and tuple(gk.shape) != (n, 24, 3)):
            raise ValueError("g_skts must be [B,24,4,4] and g_kps [B,24,3]")
        rest = (C.c_float * 72)(*[float(v) for v in torch.as_tensor(rest_pose, dtype=torch.float32).reshape(-1).tolist()])
        out = torch.empty((n, 24, 3), dtype=torch.float32, device=b.device)
        _lib.check(self.lib.pgn_pose_fk_backward(self.handle, _ptr(b), rest, n, _ptr(gs), _ptr(gk), _ptr(out), self._stream()))
        return out

    def cylinder_bboxes(self, cyls, c2w, H, W, focal):
        """pgn_cylinder_bboxes: cyls [B,5] (CUDA) -> int32 [B,4] (x0, y0, x1, y1) on the device (cylinder_to_box_2d)."""
        import numpy as np
        _check_f32_cuda(cyls, "cyls")
        cy = cyls.contiguous()
        c = np.asarray(c2w, dtype=np.float32)
        full = np.vstack([c[:3, :4], [0, 0, 0, 1]]).astype(np.float32) if c.shape[0] == 3 else c
        sw = np.array(full, copy=True)
        sw[..., 1] = -sw[..., 1]
        sw[..., 2] = -sw[..., 2]
        w2c = np.linalg.inv(sw).astype(np.float64)          # nerf_c2w_to_extrinsic (skeleton_utils.py:1412-1421), float32 inverse
        m = (C.c_double * 16)(*w2c.reshape(-1).tolist())
        out = torch.empty((cy.shape[0], 4), dtype=torch.int32, device=cy.device)
        _lib.check(self.lib.pgn_cylinder_bboxes(self.handle, _ptr(cy), cy.shape[0], m, int(H), int(W), float(focal), _ptr(out), self._stream()))
        return out

    def generate_rays_batch(self, H, W, focal, c2w, bboxes, offsets, n_total, max_rays):
        """pgn_generate_rays_batch: (ray_batch [n_total,11], pose_idx int32 [n_total]) for B bboxes (device int32 [B,4]),
        offsets device int64 [B+1]."""
        c = (C.c_float * 12)(*[float(v) for v in torch.as_tensor(c2w)[:3, :4].reshape(-1).tolist()])
        rb = torch.empty((n_total, 11), dtype=torch.float32, device=self.device)
        pidx = torch.empty((n_total,), dtype=torch.int32, device=self.device)
        _lib.check(self.lib.pgn_generate_rays_batch(self.handle, int(H), int(W), float(focal), c, _ptr(bboxes), _ptr(offsets),
                                                    bboxes.shape[0], int(max_rays), _ptr(rb), _ptr(pidx), self._stream()))
        return rb, pidx

    def compose_frames_batch(self, H, W, bboxes, offsets, rgb_map, acc_map, bg=1.0):
        """pgn_compose_frames_batch: [B,H,W,3] white-background frames from the batch's rgb / acc rows."""
        n = bboxes.shape[0]
        img = torch.empty((n, H, W, 3), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.pgn_compose_frames_batch(self.handle, int(H), int(W), _ptr(bboxes), _ptr(offsets), n, _ptr(rgb_map.contiguous()),
                                                     _ptr(acc_map.contiguous()), float(bg), _ptr(img), self._stream()))
        return img

    def gather_ray_rows(self, src, idx, n_planes=1):
        """pgn_gather_ray_rows: rows `idx` (int64, CUDA) of `src` viewed as [n_planes, n_rows, ...]; `src` [n_rows, ...] when
        n_planes == 1.  A row (everything behind the row dimension) must be a multiple of 16 bytes."""
        src = src.contiguous()
        shape = tuple(src.shape)
        rows = shape[1] if n_planes > 1 else shape[0]
        tail = shape[2:] if n_planes > 1 else shape[1:]
        row_bytes = src.element_size()
        for d in tail:
            row_bytes *= int(d)
        n = int(idx.numel())
        out = torch.empty(((n_planes, n) if n_planes > 1 else (n,)) + tuple(tail), dtype=src.dtype, device=src.device)
        if n:
            _lib.check(self.lib.pgn_gather_ray_rows(self.handle, _ptr(src), _ptr(out), _ptr(idx.contiguous()), n, row_bytes, int(n_planes),
                                                    rows * row_bytes, n * row_bytes, self._stream()))
        return out

    def frame_to_hmr_input(self, image, crop=(100, 100, 412, 412), out_res=224, mean=(0.485, 0.456, 0.406),
                           std=(0.485, 0.456, 0.406), quantize_u8=True):
        """Rendered frame [H,W,3] in [0,1] (CUDA) -> HMR input [3,out_res,out_res].
        crop = (x0, y0, x1, y1); the reference normalises with std = mean (run_gan.py:2343)."""
        _check_f32_cuda(image, "image")
        img = image.contiguous()
        H, W = img.shape[0], img.shape[1]
        out = torch.empty((3, out_res, out_res), dtype=torch.float32, device=img.device)
        m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
        _lib.check(self.lib.pgn_frame_to_hmr_input(self.handle, _ptr(img), H, W, crop[0], crop[1], crop[2], crop[3], out_res,
                                                   m3, s3, 1 if quantize_u8 else 0, _ptr(out), self._stream()))
        return out

    def phase_timers(self, enable=True, read=False):
        """Enable/disable the bf16 kernel's phase timers; read=True returns the last launch's averages (cycles)."""
        buf = (C.c_uint64 * 32)()
        _lib.check(self.lib.pgn_debug_phase_timers(self.handle, 1 if enable else 0, buf if read else None))
        return {n: int(buf[i]) for i, n in enumerate(self.PHASES) if n != "-"} if read else None

    def debug_umma_gemm(self, A, B, variant=0):
        K, N = A.shape[1], B.shape[0]
        D = torch.empty(((256 if variant & 2 else 128) + (1 if variant & 4 else 0), N), dtype=torch.float32, device=A.device)
        _lib.check(self.lib.pgn_debug_umma_gemm(self.handle, _ptr(A.contiguous()), _ptr(B.contiguous()), _ptr(D), K, N,
                                                variant, self._stream()))
        return D
