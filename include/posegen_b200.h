/*
 * posegen_b200 — C ABI of the B200-native A-NeRF volumetric renderer.
 *
 * This is the drop-in boundary for ONE hot path of mgholamikn/PoseGen:
 * `RayCaster.render_rays` (reference core/raycasters.py:361-474) and the stages it
 * calls.  The reference is pure Python/PyTorch and has no FFI of its own; the entry
 * points below are what a binding for that path would bind (see INTEGRATION.md for
 * the ctypes stub that replaces `core.raycasters.RayCaster.forward`).
 *
 * Conventions
 *   - plain C, no torch types: device pointers + sizes + a CUDA stream handle
 *     (`void*`, i.e. a `cudaStream_t`; NULL = legacy default stream);
 *   - every function returns 0 on success, a negative PGN_E_* code on failure;
 *     `pgn_last_error()` returns a thread-local message for the last failure;
 *   - no allocation on the hot path: the caller passes a workspace of
 *     `pgn_workspace_bytes()` bytes; outputs are caller-allocated;
 *   - all tensors are fp32, row-major, contiguous unless a stride is given;
 *   - re-entrant per context; one context per device per thread
 *     (nn.DataParallel calls replicas from one Python thread per GPU,
 *     reference core/raycasters.py:157).
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     PGN_E_CUDA.
 */
#ifndef POSEGEN_B200_H
#define POSEGEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGN_ABI_VERSION 2

#if defined(__GNUC__)
#define PGN_API __attribute__((visibility("default")))
#else
#define PGN_API
#endif

#define PGN_OK          0
#define PGN_E_INVALID  -1   /* bad argument / unsupported configuration */
#define PGN_E_CUDA     -2   /* CUDA runtime error (message in pgn_last_error) */
#define PGN_E_STATE    -3   /* weights / scalars not uploaded yet */
#define PGN_E_KERNEL   -4   /* device-side watchdog tripped (pipeline stall) */

#define PGN_N_JOINTS     24
#define PGN_RAY_STRIDE   11  /* o(3) d(3) near far viewdirs(3): core/trainer.py:118-137 */
#define PGN_NET_COARSE    0
#define PGN_NET_FINE      1
#define PGN_N_LINEAR     12  /* linear layers per NeRF, order below */

/* precision of the MLP stage */
#define PGN_PRECISION_FP32  0   /* fp32 CUDA-core path: 1e-3 parity tier        */
#define PGN_PRECISION_BF16  1   /* bf16 tcgen05 tensor-core path: 2e-2 / 40 dB  */

typedef struct pgn_context pgn_context;

/* Model hyper-parameters (frozen table: SURVEY.md §8d; reference
 * configs/surreal/surreal.txt + run_nerf.py:186-490 defaults).  Only the surreal.txt
 * architecture is implemented; pgn_create rejects anything else. */
typedef struct pgn_config {
  int32_t n_joints;        /* 24 */
  int32_t n_samples;       /* 64  (N_samples)     */
  int32_t n_importance;    /* 16  (N_importance)  */
  int32_t multires;        /* 7   (distance PE frequencies) */
  int32_t multires_views;  /* 4   (view PE frequencies)     */
  int32_t net_depth;       /* 8 */
  int32_t net_width;       /* 256 */
  int32_t skip_layer;      /* 4 */
  int32_t device;          /* CUDA device ordinal */
  /* ABI v2: Optcodes frame codes (core/networks/embedding.py:4-46; `opt_framecode = True` in the h36m / mixamo /
   * perfcap configs): every NeRF owns n_framecodes learned codes of framecode_ch (16) channels and views_linears.0
   * reads [feature(256) | input_views(648) | code(16)] = 920 inputs (core/networks/nerf.py:52-56,104-131).  0 = off. */
  int32_t n_framecodes;
  int32_t framecode_ch;
} pgn_config;

/* One NeRF MLP's parameters in nn.Linear layout ([out,in] row-major weight, [out] bias),
 * in this order (reference core/networks/nerf.py:57-88):
 *   0..7  pts_linears.0-7   (432->256, 4x 256->256, 688->256, 2x 256->256)
 *   8     alpha_linear      (256->1)
 *   9     feature_linear    (256->256)
 *   10    views_linears.0   (904->128; 920->128 with frame codes)
 *   11    rgb_linear        (128->3)
 * Pointers may be host or device memory (flag in pgn_upload_weights). */
typedef struct pgn_net_weights {
  const float* weight[PGN_N_LINEAR];
  const float* bias[PGN_N_LINEAR];
  const float* framecodes;   /* ABI v2: framecodes.codes.weight [n_framecodes,16] (NULL when the config has none) */
} pgn_net_weights;

/* Inputs of one render call.  Replaces the arguments of RayCaster.render_rays
 * (core/raycasters.py:361-381).  kp_batch / bones / subject_idxs / cams are not
 * part of the ABI: the surreal.txt encoders ignore them (core/encoders.py:110-122,
 * 181-193) and opt_framecode is off. */
typedef struct pgn_render_inputs {
  const float*   ray_batch;    /* device [n_rays, 11] */
  int64_t        n_rays;
  const float*   skts;         /* device; world->joint-local 4x4 per joint        */
  int64_t        skts_stride;  /* floats between consecutive rays' skts: 384 for the
                                  reference's per-ray layout [N,24,4,4], 0 when one
                                  pose is shared (the expand() view of run_nerf.py:63-90) */
  const float*   cyls;         /* device; (cx, cz, R, top, bot) per ray            */
  int64_t        cyls_stride;  /* 5 or 0 */
  const int32_t* pose_idx;     /* optional device [n_rays]: when non-NULL the pose of
                                  ray i is pose_idx[i] and skts/cyls are indexed
                                  [pose,24,4,4] / [pose,5] (strides ignored)          */
  int64_t        nanfill_chunk;/* rays per chunk for the missed-cylinder near/far fill
                                  (core/utils/ray_utils.py:328-342 takes the mean over
                                  the batchify chunk); <=0 => whole call               */
  int32_t        precision;    /* PGN_PRECISION_* */
  /* ABI v2: optional explicit chunk table for the near/far fill, device int64 [n_chunks + 1] ascending ray indices with
   * chunk c = rays [chunk_starts[c], chunk_starts[c+1]).  When non-NULL it overrides nanfill_chunk: a batch that holds
   * the rays of several images keeps the reference's per-image batchify chunks (run_nerf.py:77-95). */
  const int64_t* chunk_starts;
  int64_t        n_chunks;
  /* ABI v2: the `cams` kwarg of RayCaster.render_rays (core/raycasters.py:368,429,453): optional device int32 [n_rays]
   * frame-code index per ray.  NULL or an index outside [0, n_framecodes) selects the mean code, the reference's
   * evaluation rule for idx < 0 (core/networks/embedding.py:23-24).  Ignored by a context without frame codes. */
  const int32_t* cams;
  /* ABI v2: the `lindisp` kwarg (core/raycasters.py:650-663 -> sample_from_lineseg, core/utils/ray_utils.py:224-227):
   * non-zero = the 64 coarse samples are linear in inverse depth, z = 1 / (1/near (1 - t) + 1/far t). */
  int32_t        lindisp;
} pgn_render_inputs;

/* Outputs of one render call (core/raycasters.py:711-724).  Any pointer may be NULL
 * to skip that output.  Optional taps are for stage-level parity tests. */
typedef struct pgn_render_outputs {
  float* rgb_map;    /* [n,3]  */
  float* disp_map;   /* [n]    */
  float* acc_map;    /* [n]    */
  float* alpha;      /* [n,80] per-sample fine alpha   */
  float* rgb0;       /* [n,3]  */
  float* disp0;      /* [n]    */
  float* acc0;       /* [n]    */
  float* alpha0;     /* [n,64] per-sample coarse alpha */
  /* taps */
  float*   z_samples;  /* [n,16] importance samples (before the merge sort) */
  float*   z_fine;     /* [n,80] merged, sorted z                           */
  int32_t* pdf_inds;   /* [n,16] searchsorted(cdf,u,right) bin indices      */
  float*   weights0;   /* [n,64] coarse compositing weights                 */
  float*   raw0;       /* [n,64,4] coarse network output (rgb_raw, sigma_raw) */
  float*   raw;        /* [n,80,4] fine network output                      */
  float*   near_far;   /* [n,2] near/far after the cylinder intersection    */
} pgn_render_outputs;

/* ------------------------------------------------------------------ lifecycle */
PGN_API int  pgn_abi_version(void);
PGN_API const char* pgn_last_error(void);

/* replaces create_raycaster's module construction (core/raycasters.py:17-109) */
PGN_API int  pgn_create(const pgn_config* cfg, pgn_context** out);
PGN_API void pgn_destroy(pgn_context* ctx);

/* replaces RayCaster.load_state_dict / parameter refresh after optimizer.step
 * (core/raycasters.py:768-788): repacks one net into the kernel layouts
 * (fp32 transposed for the CUDA-core path; bf16 UMMA K-major slabs for tcgen05). */
PGN_API int  pgn_upload_weights(pgn_context* ctx, int net_id, const pgn_net_weights* w,
                        int pointers_are_device, void* stream);

/* embedder scalars: tau buffers and cutoff distances of embed_fn / embeddirs_fn
 * (core/cutoff_embedder.py:83-95,181-183), density_scale and rgb_eps
 * (core/networks/nerf.py:150-151).  cutoff arrays are HOST pointers [24]. */
PGN_API int  pgn_set_embed_scalars(pgn_context* ctx, float tau_v, float tau_d,
                           const float* cutoff_v, const float* cutoff_d,
                           float density_scale, float rgb_eps);

/* ------------------------------------------------------------------- hot path */
PGN_API size_t pgn_workspace_bytes(const pgn_context* ctx, int64_t n_rays);

/* replaces RayCaster.render_rays (eval path: perturb=0, raw_noise_std=0,
 * ray_noise_std=0; core/raycasters.py:361-474).  Asynchronous on `stream`. */
PGN_API int  pgn_render_forward(pgn_context* ctx, const pgn_render_inputs* in,
                        const pgn_render_outputs* out,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Training-step forward (reference core/trainer.py:232-275, eval-style sampling: perturb = 0, no noise): the same
 * fused bf16 pipeline as pgn_render_forward that additionally stores the post-ReLU activations of every MLP layer
 * (bf16) for the weight-gradient kernel (pgn_mlp_weight_grads).  Per pass the dump holds: layers 0-7 (pts_linears),
 * each TILE-BLOCKED as [rows / 128][256 / 8][128][8] (element (r, c) of a layer at bf16 index
 * ((r / 128) * 32 + c / 8) * 1024 + (r % 128) * 8 + c % 8: the UMMA operand image of a 128-row tile, so that the
 * forward's stores are fully coalesced and the weight-gradient kernel loads a tile with one 64 KB bulk copy); then
 * layer 8 (views_linears.0) row-major [rows,128]; then the ReLU masks of layers 0-7 as bits in WORD PLANES,
 * uint32 [layer][word 0..7][rows]: bit b of word w of row r = [activation 32 w + b of row r > 0] (a warp of the forward
 * writes one 128-byte line per word; the delta chain reads four words per thread); rows are samples in (ray, sample) order, padded to
 * rows = pgn_activation_dump_bytes(n, pass) / 4608 (a multiple of 128).  act_coarse / act_fine: device buffers of
 * pgn_activation_dump_bytes(n_rays, 0 / 1) bytes; one of them may be NULL (that pass is not dumped: a loss that reads
 * only the fine outputs sends no gradient into the coarse network).  Request out->raw0 / raw / z_fine / near_far for the backward.
 * rnd (may be NULL = deterministic eval sampling) carries the training-time randomness as explicit device arrays so
 * that a run is reproducible and checkable: the caller draws them with its own generator. */
typedef struct pgn_train_random {
  const float* t_rand;   /* [n,64] U(0,1): stratified jitter of the coarse samples, perturb > 0 (ray_utils.py:236-246) */
  const float* u_is;     /* [n,16] U(0,1): importance-sampling quantiles, det = False (ray_utils.py:169-170)          */
  const float* noise0;   /* [n,64] N(0,1) * raw_noise_std * B: density noise of the coarse pass (nerf.py:176-186)     */
  const float* noise;    /* [n,80] same for the fine pass                                                              */
} pgn_train_random;
PGN_API size_t pgn_activation_dump_bytes(int64_t n_rays, int32_t pass);
PGN_API int  pgn_render_forward_train(pgn_context* ctx, const pgn_render_inputs* in, const pgn_render_outputs* out,
                                      void* act_coarse, void* act_fine, const pgn_train_random* rnd,
                                      void* workspace, size_t workspace_bytes, void* stream);

/* The same forward for a FROZEN network whose pose gradient is wanted (GAN step, run_gan.py:159-160): only the ReLU
 * masks of the fine pass are kept - 272 B per sample instead of 4,608 B - which is all the input-gradient backward
 * needs (pgn_view_delta_from_mask, pgn_mlp_delta_chain).  masks_fine: device buffer of pgn_mask_dump_bytes(n_rays)
 * bytes laid out as word planes: uint32 [layer 0..7][word 0..7][rows] (bit b of word w = [activation 32 w + b > 0]), then
 * uint32 [word 0..3][rows] for views_linears.0; rows = bytes / 272, samples in (ray, sample) order.  Sampling is the deterministic eval sampling; request out->raw / z_fine for the backward. */
PGN_API size_t pgn_mask_dump_bytes(int64_t n_rays);
PGN_API int  pgn_render_forward_masks(pgn_context* ctx, const pgn_render_inputs* in, const pgn_render_outputs* out,
                                      void* masks_fine, void* workspace, size_t workspace_bytes, void* stream);

/* dG [m,128] (bf16) = [g > 0] * (d_rgb W_rgb): the view layer's delta from its mask bits (vmask: word planes uint32 [4][m]),
 * d_raw fp32 [m,4] (columns 0-2 = d_rgb), w_rgb fp32 [3,128] (rgb_linear.weight; core/networks/nerf.py:129-131). */
PGN_API int  pgn_view_delta_from_mask(pgn_context* ctx, void* dG, const float* d_raw, const float* w_rgb, const void* vmask,
                                      int64_t m, void* stream);

/* number of kernel launches issued by this context since creation
 * (bench.py reports it as gpu_launches) */
PGN_API int64_t pgn_launch_count(const pgn_context* ctx);

/* after a stream sync: 0 if no device-side watchdog tripped since the last call */
PGN_API int  pgn_check_device_status(pgn_context* ctx);
/* device address of the int32 status word pgn_check_device_status reads (0 = healthy): lets a host binding queue an
 * asynchronous 4-byte read-back behind its launches instead of synchronising (Engine.poll_status). */
PGN_API const int32_t* pgn_device_status_ptr(pgn_context* ctx);

/* ------------------------------------------------- stage-level entry points
 * (unit parity against the oracle; each replaces one reference function) */

/* get_near_far_in_cylinder (core/utils/ray_utils.py:292-344) incl. NaN fill.
 * near_far: device [n,2]. */
PGN_API int  pgn_near_far(pgn_context* ctx, const pgn_render_inputs* in, float* near_far, void* stream);

/* encode_inputs (core/raycasters.py:476-555) for explicit z values:
 * z [n, n_z] -> enc [n, n_z, 1080] in the reference channel order
 * ([0,360) v_emb k*24+j | [360,432) r j*3+c | [432,1080) d_emb k*72+j*3+c). */
PGN_API int  pgn_encode(pgn_context* ctx, const pgn_render_inputs* in, const float* z, int32_t n_z,
                float* enc, void* stream);

/* the same encoding rounded to bf16, [n, n_z, 1080] row-major: the operand the weight-gradient GEMMs of the training
 * step read (enc: device bf16 buffer of n * n_z * 1080 * 2 bytes). */
PGN_API int  pgn_encode_bf16(pgn_context* ctx, const pgn_render_inputs* in, const float* z, int32_t n_z,
                             void* enc, void* stream);

/* One fused pass over a layer's delta matrix in the MLP backward of the training step (what autograd does with
 * threshold_backward + sum + an outer product + a skinny GEMM; core/networks/nerf.py:94-148):
 *   dh[m, n_cols] (bf16, in/out)  <-  [act > 0] * ((has_input ? dh : 0) + rs @ wr)
 *   colsum[n_cols]                <-  column sums of the new dh            (bias gradient of the layer)
 *   wsum[nrs, n_cols]             <-  rs^T @ act                           (weight gradient of the head reading act)
 * act: bf16 [m, n_cols] post-ReLU activations of the layer (NULL = no mask); rs: fp32 per-row head deltas, row r at
 * rs + r * rs_stride, nrs values (0: no head; 1: alpha_linear on h7; 3: rgb_linear on the view layer);
 * wr: fp32 [nrs, n_cols] head weights; wsum may be NULL.  n_cols is 256 (nrs 0|1) or 128 (nrs 0|3). */
PGN_API int  pgn_mlp_delta(pgn_context* ctx, void* dh, int32_t has_input, const void* act, int64_t m, int32_t n_cols,
                           const float* rs, int32_t rs_stride, int32_t nrs, const float* wr,
                           float* colsum, float* wsum, void* stream);

/* The delta chain of the trunk backward as ONE tcgen05 kernel (what autograd runs as 8 x (GEMM + threshold_backward +
 * sum) through HBM; core/networks/nerf.py:94-102 backwards):
 *   dL/dh7 = dG W_fold + d_sigma w_alpha;  dZ_l = [h_l > 0] * dL/dh_l;  dL/dh_{l-1} = dZ_l W_l   (l = 7..1)
 * dG: bf16 [m,128] = dL/d(pre-activation of views_linears.0) (from pgn_mlp_delta); d_raw: fp32 [m,4] (column 3 =
 * d_sigma); mask: the ReLU-mask area of the activation dump of pgn_render_forward_train (starts
 * rows * 4352 bytes into the pass's buffer, rows = mask_rows = pgn_activation_dump_bytes / 4608), i.e. word planes
 * uint32 [8 layers][8 words][mask_rows]; w_alpha fp32 [256].
 * wstream: the eight weights W'_j [256, K_j] bf16 (W'_0 = (W_v[:, :256] W_f)^T with K = 128; W'_j = W_l^T for
 * l = 8 - j, K = 256, the skip layer l = 5 without its 432 input columns), each cut into fills of two K = 16 steps laid
 * out [K/32][2 N halves][2 K-steps][2][128][8] (UMMA K-major core matrices; one N half per CTA of the pair that works
 * on a 512-row block, 8 KB per fill and CTA), concatenated: 60 fills of 16 KB.
 * Outputs: dz bf16 [8][m][256] (dz[l] = dZ_l, row-major) and colsum fp32 [8][256] (bias gradients), for the layers
 * whose bit is set in layer_mask (0xFF: all; a frozen network's pose gradient reads only dZ_0 and dZ_5: 0x21). */
PGN_API int  pgn_mlp_delta_chain(pgn_context* ctx, const void* dG, const float* d_raw, const void* mask, int64_t mask_rows,
                                 int64_t m, const void* wstream, const float* w_alpha, void* dz, float* colsum,
                                 uint32_t layer_mask, void* stream);

/* the same with the weights of an uploaded net (pgn_upload_weights also packs the chain's weight stream from its fp32
 * copies): what the training / GAN step call, no per-step packing on the host side */
PGN_API int  pgn_mlp_delta_chain_net(pgn_context* ctx, int32_t net_id, const void* dG, const float* d_raw, const void* mask,
                                     int64_t mask_rows, int64_t m, void* dz, float* colsum, uint32_t layer_mask, void* stream);

/* The weight gradients of one NeRF MLP (what autograd computes as dW_l = dZ_l^T h_{l-1} for every nn.Linear of
 * core/networks/nerf.py:94-148) as ONE split-K tcgen05 kernel over all samples of the pass - no library GEMM:
 *   dz   bf16 [8][m][256]   the trunk deltas written by pgn_mlp_delta_chain (all eight layers)
 *   dG   bf16 [m][128]      the view layer's delta (pgn_mlp_delta)
 *   act  the pass's activation dump of pgn_render_forward_train (dump_rows = pgn_activation_dump_bytes / 4608)
 *   enc  bf16 [m][1080]     the network input (pgn_encode_bf16)
 *   d_raw fp32 [m][4] (column 3 = d_sigma), bias_v fp32 [128] = column sums of dG (pgn_mlp_delta's colsum)
 * Outputs: flat fp32 [pgn_weight_grad_floats(ctx)] = the 12 weight gradients back to back in the pgn_net_weights order and
 * nn.Linear layouts ([out,in] row-major; slot 11, rgb_linear.weight, is left zero: pgn_mlp_delta produces it), and
 * feat_bias fp32 [256] = feature_linear.bias' gradient.  feature_linear has no activation, so its gradients and the
 * feature block of views_linears.0 are formed from T = dG^T h7 [128,256] with the uploaded weights of `net_id`. */
PGN_API size_t pgn_weight_grad_floats(const pgn_context* ctx);
PGN_API int  pgn_mlp_weight_grads(pgn_context* ctx, int32_t net_id, const void* dz, const void* dG, const void* act, int64_t dump_rows,
                                  const void* enc, int64_t m, const float* d_raw, const float* bias_v, float* flat,
                                  float* feat_bias, void* stream);

/* Backward of the frame-code term (Optcodes): dG bf16 [n_rays * n_z, 128] (the view layer's delta), cams int32 [n_rays] or
 * NULL -> g_view_weight fp32 [128, 920]: columns 904..919 += dGr^T code[cam]; g_codes fp32 [n_framecodes,16] +=
 * scatter over cam of dGr W_v[:, 904:920], with dGr the per-ray sum of dG over the ray's samples.  Both outputs are
 * accumulated into (the caller zeroes g_codes; g_view_weight is the views_linears.0 slot of pgn_mlp_weight_grads). */
PGN_API int  pgn_framecode_backward(pgn_context* ctx, int32_t net_id, const void* dG, int64_t n_rays, int32_t n_z,
                                    const int32_t* cams, float* g_view_weight, float* g_codes, void* stream);

/* dL/d(network input) of one NeRF MLP on tcgen05 (the pose gradient, BASELINE.json configs[4]; what autograd computes as
 * grad_input of pts_linears.0 / .5 and views_linears.0, core/networks/nerf.py:94-131 backwards) with the uploaded
 * weights of `net_id`:  g_xp bf16 [m,432] = dZ_5 W_5[:, :432] + dZ_0 W_0,  g_d bf16 [m,648] = dG W_v[:, 256:904];
 * dz bf16 [8][m][256] as written by pgn_mlp_delta_chain (layers 0 and 5 are read), dG bf16 [m,128].  The outputs are
 * the operands of pgn_encode_backward_bf16.  tile_blocked != 0: both outputs in 128-row tiles,
 * [ceil(m / 128)][columns / 8][128][8] (element (r, c) at (r / 128) * 128 * columns + (c / 8) * 1024 + (r % 128) * 8 + c % 8;
 * the buffers hold whole tiles, rows beyond m are written as zeros) - the form the epilogue stores with full lines
 * and pgn_encode_backward_bf16 reads back; 0: row-major. */
PGN_API int  pgn_mlp_input_grads(pgn_context* ctx, int32_t net_id, const void* dz, const void* dG, int64_t m, void* g_xp, void* g_d,
                                 int32_t tile_blocked, void* stream);

/* the split-K kernel on one explicit product, for unit tests: out[Ma, Nb] (fp32, row stride ld_out) += A[m, :Ma]^T B[m, :Nb],
 * A / B bf16 row-major with row strides lda / ldb (elements), Ma in {128, 256}, Nb a multiple of 8 <= 256, n_ctas CTAs
 * share the rows (split-K; out must be zero-initialised by the caller).  b_tile_blocked != 0: B is a 256-column matrix in
 * the training forward's tile-blocked dump layout [row / 128][column / 8][128][8] (ldb ignored; rows padded to 128). */
PGN_API int  pgn_debug_wgrad(pgn_context* ctx, const void* A, int32_t lda, int32_t Ma, const void* B, int32_t ldb, int32_t Nb,
                             int64_t m, float* out, int32_t ld_out, int32_t n_ctas, int32_t b_tile_blocked, void* stream);

/* NeRF.forward (core/networks/nerf.py:133-148) on explicit encodings:
 * enc [m,1080] -> raw [m,4].  precision selects the MLP engine. */
PGN_API int  pgn_mlp(pgn_context* ctx, int net_id, const float* enc, int64_t m, float* raw,
             int32_t precision, void* stream);

/* NeRF.raw2outputs (core/networks/nerf.py:150-205): raw [n,s,4], z [n,s], rays_d taken
 * from in->ray_batch.  Any output may be NULL. */
PGN_API int  pgn_composite(pgn_context* ctx, const pgn_render_inputs* in, const float* raw, const float* z,
                   int32_t s, float* rgb_map, float* disp_map, float* acc_map,
                   float* weights, float* alpha, void* stream);

/* backward of encode_inputs (core/raycasters.py:476-555) w.r.t. the world->joint transforms, the gradient the pose
 * generator / pose optimisation receives (run_gan.py GAN step; core/pose_opt.py): g_enc [n, n_z, 1080] = dL/d(network
 * input) in the reference channel order, z [n, n_z] -> d_skts [n,24,4,4] (per-ray gradient, bottom rows zero). */
PGN_API int  pgn_encode_backward(pgn_context* ctx, const pgn_render_inputs* in, const float* z, int32_t n_z,
                                 const float* g_enc, float* d_skts, void* stream);

/* the same with dL/d(network input) as the training backward produces it: two bf16 GEMM outputs, g_xp [n * n_z, 432]
 * (channels [0,432): v-embed | r) and g_d [n * n_z, 648] (view embed), no fp32 [.,1080] matrix in between;
 * tile_blocked != 0: in the 128-row tile layout of pgn_mlp_input_grads. */
PGN_API int  pgn_encode_backward_bf16(pgn_context* ctx, const pgn_render_inputs* in, const float* z, int32_t n_z,
                                      const void* g_xp, const void* g_d, int32_t tile_blocked, float* d_skts, void* stream);

/* backward of NeRF.raw2outputs for the training step (core/trainer.py:321-370 reads rgb_map and acc_map):
 * g_rgb [n,3] = dL/d rgb_map, g_acc [n] = dL/d acc_map (may be NULL) -> d_raw [n,s,4] = dL/d raw.
 * No gradient flows through z (the importance samples are detached, core/utils/ray_utils.py:286). */
PGN_API int  pgn_composite_backward(pgn_context* ctx, const pgn_render_inputs* in, const float* raw, const float* z,
                                    int32_t s, const float* g_rgb, const float* g_acc, const float* noise /* [n,s] or NULL */,
                                    float* d_raw, void* stream);

/* isample_from_lineseg + sample_pdf, det=True (core/utils/ray_utils.py:157-201,255-289):
 * z [n,64], weights [n,64] -> z_samples [n,16], z_sorted [n,80], pdf_inds [n,16],
 * sorted_idxs [n,80] (position in cat[z, z_samples] of each sorted element). */
PGN_API int  pgn_sample_pdf(pgn_context* ctx, const float* z, const float* weights, int64_t n,
                    float* z_samples, float* z_sorted, int32_t* pdf_inds, int32_t* sorted_idxs,
                    void* stream);

/* "next" row 1: on-device pixel rays for a bbox
 * (get_rays + kp_to_valid_rays, core/utils/ray_utils.py:6-28,83-136).
 * c2w: HOST [3,4] row-major.  Writes ray_batch [ (y1-y0)*(x1-x0), 11 ]. */
PGN_API int  pgn_generate_rays(pgn_context* ctx, int32_t H, int32_t W, float focal, const float* c2w,
                       int32_t x0, int32_t y0, int32_t x1, int32_t y1, float* ray_batch, void* stream);

/* white/constant background composite + scatter into the full frame
 * (run_nerf.py:100-133): image[valid] = rgb + (1-acc)*bg.  image: device [H*W,3]
 * pre-filled by this call with `bg`. */
PGN_API int  pgn_compose_frame(pgn_context* ctx, int32_t H, int32_t W, int32_t x0, int32_t y0, int32_t x1, int32_t y1,
                       const float* rgb_map, const float* acc_map, float bg, float* image, void* stream);

/* "next" row 2: forward kinematics on the device.  bones: device [n_poses,24,3] axis-angle,
 * rest_pose: HOST [24,3] (already scaled, e.g. smpl_rest_pose * 0.4, run_gan.py:2032).  Replaces
 * get_smpl_l2ws (core/utils/skeleton_utils.py:334-377), skts = inv(l2ws) / kps = l2ws[..., :3, 3]
 * (run_gan.py:447-449) and get_kp_bounding_cylinder (skeleton_utils.py:635-685; cyl_extend = extend_mm * ext_scale,
 * ratios 1.6 / 1.1 in the reference).  Outputs (device): skts [n,24,4,4], kps [n,24,3] (may be NULL),
 * cyls [n,5] (may be NULL), l2ws [n,24,4,4] (may be NULL). */
PGN_API int  pgn_pose_to_skts(pgn_context* ctx, const float* bones, const float* rest_pose, int32_t n_poses,
                              float cyl_extend, float top_expand_ratio, float bot_expand_ratio,
                              float* skts, float* kps, float* cyls, float* l2ws, void* stream);

/* Backward of pgn_pose_to_skts (core/utils/skeleton_utils.py:379-463 get_smpl_l2ws_torch + the rigid inverse, as torch
 * autograd differentiates them for the pose generator / core/pose_opt.py:372-445): g_skts device [n,24,4,4] = dL/d skts
 * (bottom rows ignored), g_kps device [n,24,3] = dL/d kps or NULL, rest_pose HOST [24,3] -> g_bones device [n,24,3]. */
PGN_API int  pgn_pose_fk_backward(pgn_context* ctx, const float* bones, const float* rest_pose, int32_t n_poses,
                                  const float* g_skts, const float* g_kps, float* g_bones, void* stream);

/* "next" row 1, batched (run_nerf.py:27-147 renders image by image; here B images share launches):
 * pgn_cylinder_bboxes: cylinder_to_box_2d (core/utils/skeleton_utils.py:700-787) for n cylinders, one camera:
 *   cyls device [n,5], w2c HOST double[16] (row-major inverse of the OpenCV-convention c2w), -> bbox device int32 [n,4]
 *   = (x0, y0, x1, y1), rows [y0,y1) x cols [x0,x1) are rendered (kp_to_valid_rays, core/utils/ray_utils.py:124-130).
 * pgn_generate_rays_batch: get_rays of every bbox into ONE ray batch; offsets device int64 [n_poses + 1] (prefix sums of
 *   the bbox areas, computed by the caller from the bboxes), c2w HOST [3,4]; writes ray_batch [offsets[n],11] and
 *   pose_idx int32 [offsets[n]] (the pose_idx form of pgn_render_inputs).
 * pgn_compose_frames_batch: images device [n_poses,H,W,3] = bg, bbox pixels = rgb + (1 - acc) * bg (run_nerf.py:100-133). */
PGN_API int  pgn_cylinder_bboxes(pgn_context* ctx, const float* cyls, int32_t n, const double* w2c, int32_t H, int32_t W,
                                 float focal, int32_t* bbox, void* stream);
PGN_API int  pgn_generate_rays_batch(pgn_context* ctx, int32_t H, int32_t W, float focal, const float* c2w, const int32_t* bbox,
                                     const int64_t* offsets, int32_t n_poses, int64_t max_rays_per_pose, float* ray_batch,
                                     int32_t* pose_idx, void* stream);
PGN_API int  pgn_compose_frames_batch(pgn_context* ctx, int32_t H, int32_t W, const int32_t* bbox, const int64_t* offsets,
                                      int32_t n_poses, const float* rgb_map, const float* acc_map, float bg, float* images,
                                      void* stream);

/* Rows of selected rays out of per-sample planes: dst[p][i] = src[p][idx[i]], n_planes planes of rows of row_bytes bytes
 * (a multiple of 16; planes 16-byte aligned), idx device int64 [n_idx].  The pose gradient of a frame (run_gan.py:2040-2091
 * made differentiable, BASELINE.json configs[4]) walks the rays the HMR crop reads in chunks and selects their samples out
 * of the frame's mask dump / raw / z_fine with it (torch.index_select took 6 ms per image on these shapes).  An index
 * outside [0, src_plane_bytes / row_bytes) yields a zero row and latches the device status (pgn_check_device_status). */
PGN_API int  pgn_gather_ray_rows(pgn_context* ctx, const void* src, void* dst, const int64_t* idx, int64_t n_idx, int64_t row_bytes,
                                 int32_t n_planes, int64_t src_plane_bytes, int64_t dst_plane_bytes, void* stream);

/* "next" row 4: rendered frame -> HMR input without the PNG round trip
 * (run_gan.py:2057-2071, 2326, 2433-2445): optional uint8 quantisation, crop [y0:y1, x0:x1], /255,
 * Normalize(mean, std), skimage.transform.resize(..., (3,R,R), anti_aliasing=True).
 * image: device [H,W,3] in [0,1]; mean3/std3: HOST [3]; out: device [3,R,R]. */
PGN_API int  pgn_frame_to_hmr_input(pgn_context* ctx, const float* image, int32_t H, int32_t W, int32_t x0, int32_t y0,
                                    int32_t x1, int32_t y1, int32_t out_res, const float* mean3, const float* std3,
                                    int32_t quantize_u8, float* out, void* stream);

/* per-role phase timers of the bf16 render kernel (cycles of pipeline slot 0, averaged over CTAs, of
 * the LAST launch made while enabled): out32 (32 values) may be NULL.  Slots: 3 issuer total, 4 producer-wait-slot,
 * 5 producer total, 6 encode_x, 7 encode_d, 8 epilogue, 9 compute-wait-accumulator,
 * 10 compute-wait-staging-free, 11 compositing, 12 compute total, 13 compute-wait-act-free,
 * 15 chunk store + arrive, 16..24 compute-wait-accumulator per layer L0..L7,V, 25..27 PRE elapsed of L0, L5, V. */
PGN_API int  pgn_debug_phase_timers(pgn_context* ctx, int32_t enable, uint64_t* out32);

/* bring-up probe of the tcgen05 plumbing: D[128,N] = A[128,K] * B[N,K]^T with bf16 inputs
 * and fp32 accumulation, one CTA.  variant bit 1 selects the CTA-pair form (cta_group::2, UMMA M=256):
 * A is [256,K], D is [256,N], two CTAs of one cluster each stage half of A and half of B.
 * (variant bit 0 is rejected with PGN_E_INVALID: it was a bring-up experiment that is not reachable any more.) */
PGN_API int  pgn_debug_umma_gemm(pgn_context* ctx, const float* A, const float* B, float* D, int32_t K, int32_t N,
                                 int32_t variant, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* POSEGEN_B200_H */
