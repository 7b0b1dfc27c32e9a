// Internal (C++) interface between the C-ABI layer (pgn_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "pgn_common.cuh"

// One-time launch configuration (cudaFuncSetAttribute) is PER DEVICE: a process may hold contexts on several GPUs
// (nn.DataParallel, core/raycasters.py:157), so the "configured" latch is indexed by the current device ordinal.
struct PgnPerDeviceOnce {
  bool done[64] = {};
  int dev() const { int d = 0; cudaGetDevice(&d); return d & 63; }
  bool need() const { return !done[dev()]; }
  void set() { done[dev()] = true; }
};

// ---- fp32 CUDA-core engine -------------------------------------------------
// wt[l]: transposed weights [K][N] of linear l (PGN linear order, include/posegen_b200.h);
// small heads keep nn.Linear layout.
struct PgnFp32Net {
  const float* wt[12];
  const float* b[12];
  const float* w_alpha;  // [1][256]
  const float* b_alpha;
  const float* w_rgb;    // [3][128]
  const float* b_rgb;
  const float* codes_ext;   // frame codes [n_codes + 1][16] (last row = mean code) or NULL: the 16 extra input columns of views_linears.0
};

size_t pgn_fp32_smem_bytes();
cudaError_t pgn_launch_render_fp32(const PgnRayRefs& rays, const PgnOutputs& out, const PgnFp32Net& nc,
                                   const PgnFp32Net& nf, const PgnScalars* sc_dev, const float* near_far,
                                   int num_sms, cudaStream_t stream);
cudaError_t pgn_launch_mlp_fp32(const PgnFp32Net& net, const float* enc, long long m, float* raw,
                                int num_sms, int n_codes, cudaStream_t stream);

// ---- bf16 tcgen05 engine ---------------------------------------------------
// The packed weight stream of one net: UMMA K-major slabs in consumption order
// (layout documented in pgn_render_bf16.cu / DESIGN.md) + fp32 epilogue vectors.
struct PgnBf16Net {
  const __nv_bfloat16* wstream;   // packed slabs
  const float* bias;              // [9][256] biases of the 9 MMA layers (view layer uses 128)
  const float* w_alpha;           // [256]
  const float* w_rgb;             // [3][128]
  const float* b_alpha;           // [1]
  const float* b_rgb;             // [3]
  const float* fc_table;          // [n_codes + 1][128] = W_v[:, 904:920] codes^T (last row: mean code) or NULL (no frame codes):
                                  // a per-ray additive term of the view layer's pre-activation
};

// Training forward: post-ReLU activations of the 8 trunk layers (256 columns) and of the view layer (128), bf16,
// per pass laid out row-major per layer: layers 0-7 [rows,256] each, then layer 8 [rows,128]; rows in (ray, sample)
// order, pgn_bf16_dump_rows() rows per pass.
struct PgnActDump {
  __nv_bfloat16* c;     // coarse pass
  __nv_bfloat16* f;     // fine pass
  long long rows_c, rows_f;
  int masks_only;       // 1: c / f receive only ReLU masks: [8][rows][256 bits] (trunk) then [rows][128 bits] (view layer)
  // training-time randomness, all optional (device pointers, NULL = the deterministic eval sampling):
  const float* t_rand;   // [n,64] U(0,1): stratified jitter of the coarse samples (perturb > 0)
  const float* u_is;     // [n,16] U(0,1): importance-sampling quantiles (det = False)
  const float* noise0;   // [n,64] raw-density noise of the coarse pass, already scaled by raw_noise_std * B
  const float* noise;    // [n,80] same for the fine pass
};
long long pgn_bf16_dump_rows(long long n_rays, int samples_per_ray, int masks_only);
int pgn_bf16_group_rays(long long n_rays, int masks_only);

size_t pgn_bf16_wstream_elems();
// pack one net (device fp32 nn.Linear tensors) into wstream/bias; runs on `stream`
cudaError_t pgn_pack_bf16_net(const float* const* w_dev, const float* const* b_dev, __nv_bfloat16* wstream,
                              float* bias, float* w_alpha, float* w_rgb, float* fold_tmp, int view_ld, cudaStream_t stream);
// frame codes (Optcodes): codes_ext [n+1][16] (rows 0..n-1 given, row n <- mean) and fc_table [n+1][128] = codes_ext W_v[:, 904:920]^T
cudaError_t pgn_launch_framecode_tables(float* codes_ext, int n_codes, const float* w_view, int view_ld, float* fc_table, cudaStream_t stream);
// backward of the frame-code term: dG bf16 [n_rays * nz][128] -> g_wvc [128][16] (+=), g_codes [n_codes][16] (+=)
cudaError_t pgn_launch_framecode_backward(const void* dG, long long n_rays, int nz, const int* cams, int n_codes, const float* codes_ext,
                                          const float* w_view, int view_ld, float* g_wvc, int g_wvc_ld, float* g_codes, cudaStream_t stream);
cudaError_t pgn_launch_render_bf16(const PgnRayRefs& rays, const PgnOutputs& out, const PgnBf16Net& nc,
                                   const PgnBf16Net& nf, const PgnScalars* sc_dev, const float* near_far,
                                   int* status, unsigned long long* prof, const PgnActDump* dump, int num_sms, cudaStream_t stream);
cudaError_t pgn_launch_mlp_bf16(const PgnBf16Net& net, const float* enc, long long m, float* raw,
                                const PgnScalars* sc_dev, int* status, int num_sms, int n_codes, cudaStream_t stream);

// ---- stage kernels ---------------------------------------------------------
cudaError_t pgn_launch_near_far(const PgnRayRefs& rays, long long chunk, float* near_far, cudaStream_t stream);
cudaError_t pgn_launch_encode(const PgnRayRefs& rays, const PgnScalars* sc_dev, const float* z, int n_z,
                              float* enc, cudaStream_t stream);
cudaError_t pgn_launch_composite(const PgnRayRefs& rays, const PgnScalars* sc_dev, const float* raw, const float* z,
                                 int s, float* rgb, float* disp, float* acc, float* weights, float* alpha,
                                 cudaStream_t stream);
cudaError_t pgn_launch_sample_pdf(const PgnScalars* sc_dev, const float* z, const float* weights, long long n,
                                  float* z_samples, float* z_sorted, int* pdf_inds, int* sorted_idxs,
                                  cudaStream_t stream);
cudaError_t pgn_launch_transpose(const float* w, int out_f, int in_f, float* wt, cudaStream_t stream);
cudaError_t pgn_launch_generate_rays(int H, int W, float focal, const float* c2w12_dev, int x0, int y0, int x1, int y1,
                                     float* ray_batch, cudaStream_t stream);
cudaError_t pgn_launch_compose_frame(int H, int W, int x0, int y0, int x1, int y1, const float* rgb, const float* acc,
                                     float bg, float* image, cudaStream_t stream);

cudaError_t pgn_launch_composite_backward(const PgnRayRefs& rays, const PgnScalars* sc_dev, const float* raw, const float* z,
                                          int s, const float* g_rgb, const float* g_acc, const float* noise, float* d_raw,
                                          cudaStream_t stream);
cudaError_t pgn_launch_encode_backward(const PgnRayRefs& rays, const PgnScalars* sc_dev, const float* z, int n_z,
                                       const float* g_enc, float* d_skts, cudaStream_t stream);
cudaError_t pgn_launch_encode_backward_bf16(const PgnRayRefs& rays, const PgnScalars* sc_dev, const float* z, int n_z,
                                            const __nv_bfloat16* g_xp, const __nv_bfloat16* g_d, int tile_blocked, float* d_skts, cudaStream_t stream);
cudaError_t pgn_launch_encode_bf16(const PgnRayRefs& rays, const PgnScalars* sc_dev, const float* z, int n_z,
                                   __nv_bfloat16* enc, cudaStream_t stream);
cudaError_t pgn_launch_mlp_delta(void* dh, int has_in, const void* act, long long m, int C, const float* rs, int rs_stride,
                                 int nrs, const float* wr, float* colsum, float* wsum, int num_sms, cudaStream_t stream);
cudaError_t pgn_launch_view_delta_bits(void* dG, const float* d_raw, const float* w_rgb, const void* vmask, long long m,
                                       int num_sms, cudaStream_t stream);
cudaError_t pgn_launch_pack_chain_weights(const float* const* w_dev, const float* fold, __nv_bfloat16* out, cudaStream_t stream);
cudaError_t pgn_launch_delta_chain(const void* dG, const float* d_raw, const void* mask, long long mask_rows, long long m,
                                   const void* wstream, const float* w_alpha, void* dz, float* colsum, unsigned layer_mask,
                                   int* status, int num_sms, cudaStream_t stream);
cudaError_t pgn_launch_pose_fk(const float* bones, const float* rest, int n_poses, float ext, float top_ratio, float bot_ratio,
                               float* skts, float* kps, float* cyls, float* l2ws, cudaStream_t stream);
cudaError_t pgn_launch_hmr_input(const float* image, int H, int W, int x0, int y0, int x1, int y1, int R,
                                 const float* mean3, const float* std3, int quantize, float* out, cudaStream_t stream);

// ---- batched generation-loop ends + FK backward (pgn_batch_kernels.cu) ----
cudaError_t pgn_launch_cyl_bboxes(const float* cyls, int n, const double* w2c16, int H, int W, double focal, int* bbox, cudaStream_t stream);
cudaError_t pgn_launch_generate_rays_batch(int H, int W, float focal, const float* c2w12_dev, const int* bbox, const long long* offsets,
                                           int n_poses, long long max_rays_per_pose, float* ray_batch, int* pose_idx, cudaStream_t stream);
cudaError_t pgn_launch_gather_ray_rows(const void* src, void* dst, const long long* idx, long long n_idx, long long row_bytes,
                                       int n_planes, long long src_plane_bytes, long long dst_plane_bytes, int* status, int num_sms,
                                       cudaStream_t stream);
cudaError_t pgn_launch_compose_frames_batch(int H, int W, const int* bbox, const long long* offsets, int n_poses, const float* rgb,
                                            const float* acc, float bg, float* images, cudaStream_t stream);
cudaError_t pgn_launch_pose_fk_backward(const float* bones, const float* rest, int n_poses, const float* g_skts, const float* g_kps,
                                        float* g_bones, cudaStream_t stream);
cudaError_t pgn_launch_near_far_chunks(const PgnRayRefs& rays, const long long* chunk_starts, long long n_chunks, float* near_far,
                                       cudaStream_t stream);

// ---- split-K tcgen05 weight gradients (pgn_wgrad.cu) ----
size_t pgn_wgrad_flat_floats();
cudaError_t pgn_launch_weight_grads(const void* dz, const void* dG, const void* act, long long dump_rows, const void* enc,
                                    long long m, const float* d_raw, const float* bias_v, const float* w_f, const float* b_f,
                                    const float* w_v, int view_ld, float* flat, float* feat_bias, float* tm_scratch, int* epoch_ctr,
                                    int* status, int num_sms, cudaStream_t stream);
#define PGN_WGRAD_MAX_EPOCHS 65536        // epochs of 2,048 rows: 134 M rows per launch
size_t pgn_wgrad_flat_floats_ld(int view_ld);
cudaError_t pgn_launch_wgrad_single(const void* A, int lda, int Ma, const void* B, int ldb, int Nb, long long m, float* out, int ld_out,
                                    int n_ctas, int b_tile_blocked, int* epoch_ctr, int* status, cudaStream_t stream);

// ---- input gradients of the MLP on tcgen05 (pgn_input_grads.cu) ----
size_t pgn_input_grad_weight_elems();
cudaError_t pgn_launch_pack_input_grad_weights(const float* w5, const float* w0, const float* wv, int view_ld, __nv_bfloat16* out,
                                               cudaStream_t stream);
cudaError_t pgn_launch_input_grads(const void* dz, const void* dG, long long m, const __nv_bfloat16* wpack, void* g_xp, void* g_d, int tile_blocked,
                                   int* status, int num_sms, cudaStream_t stream);

cudaError_t pgn_launch_gather_params(const float* const* src_w, const float* const* src_b, float* const* dst_w, float* const* dst_b,
                                     const int* n_w, const int* n_b, cudaStream_t stream);

// bring-up probe (pgn_probe.cu)
cudaError_t pgn_launch_probe_umma(const float* A, const float* B, float* D, int K, int N, int variant, int* status,
                                  cudaStream_t stream);
