"""Training step of the A-NeRF renderer (BASELINE.json configs[3]; reference core/trainer.py:232-275).

Forward: the fused bf16 tensor-core kernel (`pgn_render_forward_train`) with the post-ReLU activations of every MLP
layer dumped row-major in bf16 plus their 1-bit ReLU masks.  Backward, per pass: `pgn_composite_backward` turns
dL/d(rgb_map, acc_map) into dL/d raw; `pgn_mlp_delta` forms the view layer's delta (+ bias and rgb-head gradients);
`pgn_mlp_delta_chain` runs the trunk's delta chain dG -> dZ_7 .. dZ_0 as one tcgen05 kernel (bias gradients included);
the weight gradients are plain dense GEMMs over all samples of the batch (deltas^T x activations; `pgn_encode_bf16`
regenerates x_p / d_emb) and go through cuBLAS (`torch.mm`, bf16 inputs, fp32 accumulation).  When `skts` requires grad
(the pose generator / pose optimisation, configs[4]) dL/d(network input) is formed by three more GEMMs and
`pgn_encode_backward_bf16` turns it into dL/d skts.  No gradient flows through the sample positions (the importance
samples are detached in the reference, core/utils/ray_utils.py:286).

Sampling: deterministic (perturb = 0, raw_noise_std = 0: the reference's parity setting)
or the reference's training-time randomness (perturb > 0: stratified jitter of the coarse samples and random
importance quantiles, ray_utils.py:169-170,236-246; raw_noise_std > 0: density noise, nerf.py:176-186).  The
random numbers are drawn here with torch's CUDA generator and handed to the kernel as explicit arrays.

Multi-GPU: data-parallel over rays; `allreduce_gradients` is one NCCL all-reduce of the flattened 1.73 M-element
gradient bucket.
"""
from __future__ import annotations

import functools
from typing import Callable, Dict, List

import torch
import torch.distributed as dist

from .raycaster import net_tensors

S, T = 64, 80
USE_DELTA_CHAIN = True        # trunk backward through pgn_mlp_delta_chain (False: layer by layer, for A/B runs)
USE_WGRAD_KERNEL = True       # weight gradients through pgn_mlp_weight_grads (False: torch.mm, for A/B runs)
USE_INPUT_GRAD_KERNEL = True  # dL/d(network input) through pgn_mlp_input_grads (False: torch.mm, for A/B runs)
INPUT_GRADS_TILE_BLOCKED = True   # ... handed to pgn_encode_backward_bf16 in 128-row tiles (coalesced epilogue stores)
PARAM_ORDER = [f"pts_linears.{i}.{k}" for i in range(8) for k in ("weight", "bias")] + \
    [f"{m}.{k}" for m in ("alpha_linear", "feature_linear", "views_linears.0", "rgb_linear") for k in ("weight", "bias")]


def _mm32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a [m,k] @ b [k,n] with bf16 inputs and an fp32 result (cuBLAS accumulates in fp32)."""
    if a.is_cuda:
        try:
            return torch.mm(a, b, out_dtype=torch.float32)
        except (TypeError, RuntimeError):                  # older torch: no out_dtype
            pass
    return torch.mm(a.float(), b.float())


ACT_ROW_ELEMS = 2304          # bf16 elements per dump row: 8 x 256 + 128 activations + 8 x 16 (the ReLU mask bits)


def act_layer(acts: torch.Tensor, l: int, m: int) -> torch.Tensor:
    """Activation matrix [m, cols] bf16 of layer l (0..7: 256 columns, 8: view layer, 128) for the first m rows of the
    kernel's dump.  The trunk layers are stored tile-blocked ([rows/128][32][128][8]: the UMMA operand image, coalesced
    stores in the forward, one bulk copy per tile in the weight-gradient kernel) and come back as a row-major COPY
    (tests, the matrix-product formulation); the view layer is a row-major view."""
    rows = acts.numel() // ACT_ROW_ELEMS
    if l < 8:
        tb = acts[l * rows * 256:(l + 1) * rows * 256].view(rows // 128, 32, 128, 8)
        return tb.permute(0, 2, 1, 3).reshape(rows, 256)[:m]
    return acts[8 * rows * 256:rows * 2176].view(rows, 128)[:m]


def to_tile_blocked(x: torch.Tensor) -> torch.Tensor:
    """Row-major [rows, 256] -> the dump's tile-blocked layout (rows padded to a multiple of 128 with zeros), flat."""
    rows = x.shape[0]
    pad = (-rows) % 128
    if pad:
        x = torch.cat([x, x.new_zeros((pad, x.shape[1]))], 0)
    return x.reshape(-1, 128, 32, 8).permute(0, 2, 1, 3).contiguous().reshape(-1)


def act_masks(acts: torch.Tensor):
    """(mask area of the dump, rows): word planes [8 layers][8 words][rows] (word w = the bits [column 32 w + b > 0]),
    the operand of `pgn_mlp_delta_chain`."""
    rows = acts.numel() // ACT_ROW_ELEMS
    return acts[rows * 2176:], rows


def chain_wstream(P: Dict[str, torch.Tensor]) -> torch.Tensor:
    """The eight weights of the delta chain in the fill layout `pgn_mlp_delta_chain` streams (include/posegen_b200.h):
    W'_0 = (W_v[:, :256] W_f)^T [256,128], W'_j = W_l^T [256,256] for l = 7..1 (layer 5 without its 432 skip columns),
    each as [K/32 fills][2 N halves][2 K-steps][2][128][8] bf16 (one 8 KB bulk copy per fill and CTA of the pair)."""
    fold = P["views_linears.0.weight"][:, :256] @ P["feature_linear.weight"]
    mats = [fold.t()] + [(P[f"pts_linears.{l}.weight"][:, 432:] if l == 5 else P[f"pts_linears.{l}.weight"]).t()
                         for l in range(7, 0, -1)]
    return torch.cat([w.to(torch.bfloat16).reshape(2, 128, -1, 2, 2, 8).permute(2, 0, 3, 4, 1, 5).reshape(-1) for w in mats])


def mlp_backward(params: Dict[str, torch.Tensor], enc: torch.Tensor, acts: torch.Tensor, d_raw: torch.Tensor,
                 fuse: Callable, want_input_grad: bool = False, want_weight_grad: bool = True,
                 chain: Callable | None = None, mask_dump=None, view_delta: Callable | None = None,
                 wgrad: Callable | None = None, input_grads: Callable | None = None) -> Dict[str, torch.Tensor]:
    """Weight gradients of one NeRF MLP (core/networks/nerf.py:94-148).

    params: fp32 nn.Linear tensors; enc [m,1080] bf16 network input (`pgn_encode_bf16`; may be None when
    want_weight_grad is False: a frozen network, the GAN step, only passes dL/d(network input) on); acts: the kernel's activation
    dump of this pass; d_raw [m,4] fp32 = dL/d(rgb_raw, sigma_raw).  `fuse(dh, act, rs, wr, has_input, want_wsum)` is
    `Engine.mlp_delta` (`pgn_mlp_delta`): the ReLU backward, the bias gradient and the two skinny heads in one pass
    over each delta matrix.  Deltas and activations are bf16, every GEMM accumulates in fp32 and the weight
    gradients are produced in fp32.  `feature_linear` has no activation (nerf.py:125-128), so it never shows up at
    batch size: with T = dG^T h7 its gradients and those of the feature block of `views_linears.0` are
    [128,256]-sized products, and dL/d h7 reads the folded weight W_v[:, :256] @ W_f.
    `chain(dG, d_raw, mask, mask_rows, layer_mask)` is `Engine.mlp_delta_chain_net` bound to this net
    (`pgn_mlp_delta_chain_net`; the library packs the chain's weight stream when the weights are uploaded): the
    whole trunk chain dG -> dZ_7 .. dZ_0 (+ bias gradients) as one tcgen05 kernel; without it the chain runs layer by
    layer (a cuBLAS GEMM and a `fuse` pass per layer), which is also what the host-logic test exercises.
    `wgrad(dz, dG, acts, enc, d_raw, bias_v)` is `Engine.mlp_weight_grads` bound to this net (`pgn_mlp_weight_grads`):
    with it (and `chain`) every weight gradient is formed by the library's split-K tcgen05 kernel and no library GEMM
    runs in the step; without it the products below go through `torch.mm` (the CPU host-logic test, A/B runs).
    `input_grads(dz, dG)` is `Engine.mlp_input_grads` bound to this net (`pgn_mlp_input_grads`): dL/d(network input) on
    tcgen05 instead of three library GEMMs (needs `chain`).
    mask_dump = (trunk_mask int32 [8,8,m], view_mask int32 [4,m]; word planes) with `view_delta` = `Engine.view_delta_from_mask`
    replaces `acts` for a frozen network (want_weight_grad False, `chain` required): the masks-only dump of
    `pgn_render_forward_masks` is all the input-gradient chain reads.
    Returns {name: fp32 gradient}; with want_input_grad also dL/d(network input) as "_g_xp" [m,432] (v-embed | r
    channels) and "_g_d" [m,648] (view embed), bf16 (the operands of `pgn_encode_backward_bf16`; with "_g_tb" set they are
    the flat 128-row-tile buffers of `Engine.mlp_input_grads(..., tile_blocked=True)`)."""
    m = d_raw.shape[0]
    bf = torch.bfloat16
    if want_weight_grad:
        x_p, d_emb = enc[:, :432], enc[:, 432:]
    if mask_dump is not None and (want_weight_grad or chain is None or view_delta is None):
        raise ValueError("a masks-only dump serves the input-gradient chain only (want_weight_grad=False, chain, view_delta)")
    fused_w = want_weight_grad and wgrad is not None and chain is not None
    if mask_dump is None:
        # (row-major copies of the tile-blocked trunk activations: only the matrix-product formulation reads them)
        H = [act_layer(acts, l, m) for l in range(8)] if (want_weight_grad and not fused_w) or chain is None else None
        G = act_layer(acts, 8, m)
    P = {k: v.detach() for k, v in params.items()}
    # bf16 trunk weights: only the input-gradient GEMMs and the layer-by-layer fallback read them
    need_w = (0, 5) if (chain is not None and want_input_grad and input_grads is None) else (range(8) if chain is None else ())
    W = {f"pts_linears.{l}.weight": P[f"pts_linears.{l}.weight"].to(bf) for l in need_w}
    g: Dict[str, torch.Tensor] = {}
    # rgb head + view layer:  g = relu(W_v [f | d_emb] + b_v),  f = W_f h7 + b_f,  rgb_raw = W_rgb g + b_rgb
    wg = want_weight_grad
    if mask_dump is not None:
        dG = view_delta(d_raw, P["rgb_linear.weight"], mask_dump[1])
    else:
        dG = torch.empty((m, 128), dtype=bf, device=d_raw.device)
        bias_v, g_rgb = fuse(dG, G, d_raw[:, :3], P["rgb_linear.weight"], False, wg)
    W_v, W_f, b_f = P["views_linears.0.weight"], P["feature_linear.weight"], P["feature_linear.bias"]
    W_vf = W_v[:, :256]
    if wg:
        g["rgb_linear.weight"] = g_rgb
        g["rgb_linear.bias"] = d_raw[:, :3].sum(0)
    if wg and not fused_w:
        dGt = dG.t()
        Tm = _mm32(dGt, H[7])                                                # [128,256] = dG^T h7
        g["views_linears.0.weight"] = torch.cat([Tm @ W_f.t() + bias_v[:, None] * b_f[None, :], _mm32(dGt, d_emb)], 1)
        g["views_linears.0.bias"] = bias_v
        g["feature_linear.weight"] = W_vf.t() @ Tm
        g["feature_linear.bias"] = W_vf.t() @ bias_v
    fused_in = want_input_grad and input_grads is not None and chain is not None
    if want_input_grad and not fused_in:        # dL/d(network input) stays two bf16 GEMM outputs: x_p part and view part
        g["_g_d"] = torch.mm(dG, W_v[:, 256:904].to(bf))
    if chain is not None:
        # trunk: one fused kernel for the eight deltas; the weight gradients are GEMMs over (dZ_l, h_{l-1})
        mask, mask_rows = (mask_dump[0], m) if mask_dump is not None else act_masks(acts)
        dz, colsum = chain(dG, d_raw, mask, mask_rows, 0xFF if wg else 0x21)   # a frozen network's pose gradient reads only dZ_0, dZ_5
        if fused_w:
            gw, feat_b = wgrad(dz, dG, acts, enc, d_raw, bias_v)
            for k, v in gw.items():
                if k != "rgb_linear.weight":
                    g[k] = v
            g["views_linears.0.bias"] = bias_v
            g["feature_linear.bias"] = feat_b
            g["_dG"] = dG
            g["alpha_linear.bias"] = d_raw[:, 3:4].sum(0)
            for l in range(8):
                g[f"pts_linears.{l}.bias"] = colsum[l]
        elif wg:
            g["alpha_linear.weight"] = _mm32(d_raw[:, 3:4].to(bf).t(), H[7])
            g["alpha_linear.bias"] = d_raw[:, 3:4].sum(0)
            for l in range(8):
                dZt = dz[l].t()
                if l == 0:
                    g["pts_linears.0.weight"] = _mm32(dZt, x_p)
                elif l == 5:
                    g["pts_linears.5.weight"] = torch.cat([_mm32(dZt, x_p), _mm32(dZt, H[4])], 1)
                else:
                    g[f"pts_linears.{l}.weight"] = _mm32(dZt, H[l - 1])
                g[f"pts_linears.{l}.bias"] = colsum[l]
        if fused_in:
            g["_g_xp"], g["_g_d"] = input_grads(dz, dG, tile_blocked=INPUT_GRADS_TILE_BLOCKED)
            g["_g_tb"] = INPUT_GRADS_TILE_BLOCKED
        elif want_input_grad:
            g["_g_xp"] = torch.mm(dz[5], W["pts_linears.5.weight"][:, :432]).addmm_(dz[0], W["pts_linears.0.weight"])
        return g
    # sigma head + last trunk layer: dL/d h7 = dG (W_vf W_f) + d_sigma w_alpha
    dH = torch.mm(dG, (W_vf @ W_f).to(bf))
    bias, g_alpha = fuse(dH, H[7], d_raw[:, 3:4], P["alpha_linear.weight"], True, wg)
    if wg:
        g["alpha_linear.weight"] = g_alpha
        g["alpha_linear.bias"] = d_raw[:, 3:4].sum(0)
    # trunk, last layer first; layer 5 reads [x_p | h4] (skip after layer index 4, nerf.py:100-101)
    for l in range(7, -1, -1):
        dZ = dH                                                              # masked in place by `fuse`
        if wg:
            dZt = dZ.t()
            if l == 0:
                g["pts_linears.0.weight"] = _mm32(dZt, x_p)
            elif l == 5:
                g["pts_linears.5.weight"] = torch.cat([_mm32(dZt, x_p), _mm32(dZt, H[4])], 1)
            else:
                g[f"pts_linears.{l}.weight"] = _mm32(dZt, H[l - 1])
            g[f"pts_linears.{l}.bias"] = bias
        Wl = W[f"pts_linears.{l}.weight"]
        if want_input_grad and l in (0, 5):
            g["_g_xp"] = torch.mm(dZ, Wl[:, :432]) if l == 5 else g["_g_xp"].addmm_(dZ, Wl)
        if l > 0:
            dH = torch.mm(dZ, Wl[:, 432:] if l == 5 else Wl)
            bias, _ = fuse(dH, H[l - 1], None, None, True, False)
    return g


class _RenderTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rc, ray_batch, skts, cyls, nanfill_chunk, rand, cams, lindisp, *params):
        eng = rc.engine(ray_batch.device)
        ret, acts = eng.render_train(ray_batch, skts, cyls, nanfill_chunk=nanfill_chunk, rand=rand, cams=cams, lindisp=lindisp)
        ctx.lindisp = lindisp
        rc.mark_weights_dirty()            # an optimizer step follows; not every optimizer bumps the version counters
        ctx.rc, ctx.eng, ctx.acts, ctx.rand, ctx.cams = rc, eng, acts, (rand or {}), cams
        ctx.set_materialize_grads(False)   # outputs the loss does not read arrive as None: their pass is skipped
        ctx.save_for_backward(ray_batch, skts, cyls, ret["raw0"], ret["raw"], ret["z_fine"], ret["near_far"])
        ctx.mark_non_differentiable(ret["disp_map"], ret["disp0"])
        return ret["rgb_map"], ret["acc_map"], ret["rgb0"], ret["acc0"], ret["disp_map"], ret["disp0"]

    @staticmethod
    def backward(ctx, g_rgb, g_acc, g_rgb0, g_acc0, _gd, _gd0):
        rb, sk, cy, raw0, raw, z_fine, near_far = ctx.saved_tensors
        rc, eng = ctx.rc, ctx.eng      # the context as the forward left it (rc.engine() would re-pack the weights here)
        n = rb.shape[0]
        zero3, zero1 = torch.zeros((n, 3), device=rb.device), torch.zeros(n, device=rb.device)
        t = torch.linspace(0., 1., S, device=rb.device)                       # sample_from_lineseg, ray_utils.py:204-251
        if ctx.lindisp:
            z_c = 1. / (1. / near_far[:, :1] * (1. - t) + 1. / near_far[:, 1:2] * t)
        else:
            z_c = near_far[:, :1] * (1. - t) + near_far[:, 1:2] * t
        rand = ctx.rand
        if rand.get("t_rand") is not None:                                    # stratified jitter, ray_utils.py:236-246
            mids = .5 * (z_c[:, 1:] + z_c[:, :-1])
            upper, lower = torch.cat([mids, z_c[:, -1:]], -1), torch.cat([z_c[:, :1], mids], -1)
            z_c = lower + (upper - lower) * rand["t_rand"]
        grads: List[torch.Tensor] = []
        want_sk = ctx.needs_input_grad[2]
        want_w = any(ctx.needs_input_grad[8:])               # False for a frozen NeRF (the GAN step)
        order = param_order(rc.network)
        d_skts = None
        # Gradient arena: every weight / bias gradient of both nets lives in ONE fp32 buffer (the weight-gradient kernel
        # writes its flat result straight into it, the small tensors are copied behind), so that the data-parallel step
        # needs one in-place all-reduce and no flatten / scatter copies (`allreduce_gradients`).
        fused = want_w and USE_DELTA_CHAIN and USE_WGRAD_KERNEL
        n_w = eng.weight_grad_floats() if fused else 0
        n_small = 8 * 256 + 1 + 256 + 128 + 3 + 3 * 128 + rc.n_framecodes * 16
        arena = torch.empty((2 * (n_w + n_small),), dtype=torch.float32, device=rb.device) if fused else None
        net_no = -1
        for net, acts, z, raw_p, gr, ga, nz in ((rc.network, ctx.acts["c"], z_c, raw0, g_rgb0, g_acc0, rand.get("noise0")),
                                                (rc.network_fine, ctx.acts["f"], z_fine, raw, g_rgb, g_acc, rand.get("noise"))):
            net_no += 1
            if (gr is None and ga is None) or not (want_w or want_sk):      # the loss does not read this pass
                grads += [None] * len(order)
                continue
            gr = zero3 if gr is None else gr.contiguous().float()
            ga = zero1 if ga is None else ga.contiguous().float()
            z = z.contiguous()
            d_raw = eng.composite_backward(rb, sk, cy, raw_p, z, gr, ga, noise=nz)
            enc = eng.encode_bf16(rb, sk, cy, z).reshape(-1, 1080) if want_w else None
            pd = net_tensors(net)
            net_id = 0 if net is rc.network else 1
            gd = mlp_backward(pd, enc, acts, d_raw.reshape(-1, 4), eng.mlp_delta, want_input_grad=want_sk, want_weight_grad=want_w,
                              chain=functools.partial(eng.mlp_delta_chain_net, net_id) if USE_DELTA_CHAIN else None,
                              wgrad=functools.partial(eng.mlp_weight_grads, net_id,
                                                      out=arena[net_no * (n_w + n_small):net_no * (n_w + n_small) + n_w] if fused else None)
                              if (USE_DELTA_CHAIN and USE_WGRAD_KERNEL) else None,
                              input_grads=functools.partial(eng.mlp_input_grads, net_id) if (USE_DELTA_CHAIN and USE_INPUT_GRAD_KERNEL) else None)
            if want_w and rc.n_framecodes:      # Optcodes: the 16 frame-code columns of views_linears.0 and the codes themselves
                gv = gd["views_linears.0.weight"]
                if tuple(gv.shape) != (128, 920) or not gv.is_contiguous():
                    raise RuntimeError("frame-code gradients need the fused weight-gradient path (USE_WGRAD_KERNEL)")
                gd["framecodes.codes.weight"] = eng.framecode_backward(net_id, gd["_dG"], n, z.shape[1], ctx.cams, gv)
            if fused:            # small tensors (biases, rgb head, frame codes) -> the arena's tail of this net
                o = net_no * (n_w + n_small) + n_w
                base = arena.untyped_storage().data_ptr()
                for k in order:
                    t = gd[k]
                    if t.untyped_storage().data_ptr() != base:
                        slot = arena[o:o + t.numel()].view(t.shape)
                        slot.copy_(t)
                        gd[k] = slot
                        o += t.numel()
            grads += [gd[k].reshape(pd[k].shape).to(pd[k].dtype) for k in order] if want_w else [None] * len(order)
            if want_sk:          # pose gradient: dL/d(network input) -> dL/d skts (per ray), both passes add up
                d = eng.encode_backward_bf16(rb, sk, cy, z, gd["_g_xp"], gd["_g_d"], tile_blocked=gd.get("_g_tb", False))
                d_skts = d if d_skts is None else d_skts + d
        ctx.acts = None
        if want_sk and d_skts is not None and sk.dim() == 3:
            d_skts = d_skts.sum(0)                     # one pose shared by every ray of the batch
        return (None, None, d_skts, None, None, None, None, None) + tuple(grads)


def draw_train_random(n: int, device, perturb: float = 0., raw_noise_std: float = 0., density_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """The random numbers one training forward consumes (what the reference draws with torch.rand / torch.randn)."""
    rand: Dict[str, torch.Tensor] = {}
    if perturb > 0.:
        rand["t_rand"] = torch.rand((n, S), device=device)
        rand["u_is"] = torch.rand((n, 16), device=device)
    if raw_noise_std > 0.:
        rand["noise0"] = torch.randn((n, S), device=device) * (raw_noise_std * density_scale)
        rand["noise"] = torch.randn((n, T), device=device) * (raw_noise_std * density_scale)
    return rand


def param_order(net) -> List[str]:
    """PARAM_ORDER, plus the frame codes of an Optcodes net."""
    return PARAM_ORDER + (["framecodes.codes.weight"] if getattr(net, "use_framecode", False) else [])


def render_train(rc, ray_batch, skts, cyls, nanfill_chunk=None, perturb: float = 0., raw_noise_std: float = 0.,
                 rand: Dict[str, torch.Tensor] | None = None, cams=None, lindisp: bool = False) -> Dict[str, torch.Tensor]:
    """Differentiable (w.r.t. the two MLPs' parameters and `skts`) render of a ray batch: the train-mode body of
    RayCaster.forward.  Returns the reference's dict (core/raycasters.py:711-724) without alpha/alpha0."""
    params = [net_tensors(net)[k] for net in (rc.network, rc.network_fine) for k in param_order(net)]
    n = ray_batch.shape[0]
    if rand is None:
        rand = draw_train_random(n, ray_batch.device, perturb, raw_noise_std, float(rc.network.density_scale))
    out = _RenderTrainFn.apply(rc, ray_batch.float().contiguous(), skts if skts.dtype == torch.float32 else skts.float(), cyls.float(),
                               n if nanfill_chunk is None else nanfill_chunk, rand or None, cams, bool(lindisp), *params)
    return {"rgb_map": out[0], "acc_map": out[1], "rgb0": out[2], "acc0": out[3], "disp_map": out[4], "disp0": out[5]}


def allreduce_gradients(parameters, average: bool = True):
    """One all-reduce of the gradient bucket (NCCL on GPUs, gloo in the CPU tests).  The training backward hands out
    gradients that are views of one arena buffer: then the all-reduce runs in place on that buffer - one collective, no
    flatten / scatter copies.  Gradients from anywhere else are flattened, reduced and copied back."""
    ps = [p for p in parameters if p.grad is not None]
    if not ps or not dist.is_initialized() or dist.get_world_size() == 1:
        return
    g0 = ps[0].grad
    if all(p.grad.is_contiguous() and p.grad.dtype == g0.dtype and p.grad.device == g0.device and
           p.grad.untyped_storage().data_ptr() == g0.untyped_storage().data_ptr() for p in ps):
        lo = min(p.grad.storage_offset() for p in ps)
        hi = max(p.grad.storage_offset() + p.grad.numel() for p in ps)
        flat = torch.empty(0, dtype=g0.dtype, device=g0.device).set_(g0.untyped_storage(), lo, (hi - lo,))
        dist.all_reduce(flat)
        if average:
            flat /= dist.get_world_size()
        return
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat)
    if average:
        flat /= dist.get_world_size()
    o = 0
    for p in ps:
        k = p.grad.numel()
        p.grad.copy_(flat[o:o + k].view_as(p.grad))
        o += k


class GraphedTrainStep:
    """One training step (forward, loss, backward, optimizer step) captured into a CUDA graph.

    The step is ~150 kernel launches of 10-400 us; issued from Python the launch queue runs dry between them (about
    1 ms of a 5.5 ms step).  `GraphedTrainStep(rc, optimizer, loss_fn, inputs)` warms the step up on a side stream,
    captures it once and `__call__(**new_inputs)` replays it after copying the new batch into the static input
    tensors.  The optimizer must be capturable (`torch.optim.Adam(..., capturable=True)`); the random numbers of the
    reference's training-time sampling are drawn inside the graph by torch's graph-safe generator.

    inputs: dict with `ray_batch [n,11]`, `skts`, `cyls`, `target` (CUDA tensors; shapes are frozen by the capture);
    loss_fn(ret, target) -> scalar; render_kwargs are passed to `rc(...)` (perturb, raw_noise_std, ...)."""

    def __init__(self, rc, optimizer, loss_fn, inputs: Dict[str, torch.Tensor], warmup: int = 3, **render_kwargs):
        self.rc, self.opt, self.loss_fn, self.kw = rc, optimizer, loss_fn, render_kwargs
        self.static = {k: v.clone() for k, v in inputs.items()}
        # data-parallel runs keep the gradient all-reduce out of the capture (NCCL's watchdog and a capturing stream do
        # not mix): graph A = forward + backward, eager all-reduce, graph B = optimizer step
        self.split = dist.is_initialized() and dist.get_world_size() > 1
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._fwd_bwd()
                allreduce_gradients(self.rc.parameters())
                self.opt.step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        self.graph_opt = None
        if not self.split:
            with torch.cuda.graph(self.graph):
                self.loss = self._fwd_bwd()
                self.opt.step()
        else:
            with torch.cuda.graph(self.graph):
                self.loss = self._fwd_bwd()
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt, pool=self.graph.pool()):
                self.opt.step()

    def _fwd_bwd(self):
        # gradients are re-created by every backward of the captured graph at the same addresses
        self.opt.zero_grad(set_to_none=True)
        s = self.static
        ret = self.rc(s["ray_batch"], N_samples=S, N_importance=T - S, kp_batch=None, skts=s["skts"], cyls=s["cyls"], bones=None,
                      cams=None, **self.kw)
        loss = self.loss_fn(ret, s["target"])
        loss.backward()
        return loss.detach()

    def __call__(self, **inputs):
        for k, v in inputs.items():
            self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        if self.split:
            allreduce_gradients(self.rc.parameters())
            self.graph_opt.replay()
        return self.loss
