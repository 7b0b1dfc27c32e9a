// C ABI of posegen_b200 (include/posegen_b200.h): context, weight repacking, launches.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <new>
#include "../../include/posegen_b200.h"
#include "pgn_kernels.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define PGN_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) return fail(PGN_E_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                       __FILE__, __LINE__);                                         \
  } while (0)

// Every entry point runs on its context's device and leaves the caller's current device as it found it (torch keeps
// its own notion of the current device; a DataParallel replica thread must not move it for the others).
struct DeviceGuard {
  int prev = -1;
  cudaError_t err;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev); else if (err == cudaSuccess) prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define PGN_ON_DEVICE(c)                                                                              \
  DeviceGuard _guard((c)->cfg.device);                                                               \
  if (_guard.err != cudaSuccess) return fail(PGN_E_CUDA, "cudaSetDevice(%d): %s", (c)->cfg.device, cudaGetErrorString(_guard.err))

// (out, in) of the 12 linear layers, include/posegen_b200.h order
const int kOut[PGN_N_LINEAR] = {256, 256, 256, 256, 256, 256, 256, 256, 1, 256, 128, 3};
const int kIn[PGN_N_LINEAR]  = {432, 256, 256, 256, 256, 688, 256, 256, 256, 256, 904, 128};

// views_linears.0 reads 16 more columns (the frame code) when the model has Optcodes (core/networks/nerf.py:52-56)
int in_of(int l, int view_in) { return l == 10 ? view_in : kIn[l]; }
size_t weight_floats(int view_in) { size_t t = 0; for (int i = 0; i < PGN_N_LINEAR; ++i) t += (size_t)kOut[i] * in_of(i, view_in); return t; }
size_t bias_floats() { size_t t = 0; for (int i = 0; i < PGN_N_LINEAR; ++i) t += kOut[i]; return t; }

}  // namespace

struct pgn_context {
  pgn_config cfg;
  int num_sms;
  PgnScalars h_sc;
  PgnScalars* d_sc;
  bool have_sc;
  bool have_w[2];
  // raw copies of the parameters (nn.Linear layout) on the device
  float* d_w[2];
  float* d_b[2];
  const float* w_ptr[2][PGN_N_LINEAR];
  const float* b_ptr[2][PGN_N_LINEAR];
  // fp32 engine: transposed weights
  float* d_wt[2];
  PgnFp32Net fp32[2];
  // bf16 engine
  __nv_bfloat16* d_wstream[2];
  float* d_bf16_aux[2];     // bias[9*256] | w_alpha[256] | w_rgb[384]
  float* d_fold;
  __nv_bfloat16* d_chain_w[2] = {nullptr, nullptr};   // per net: the delta chain's weight stream (60 fills of 16 KB)
  PgnBf16Net bf16[2];
  int* d_status;
  bool fp32_stale[2] = {false, false};   // fp32-tier transposes pending since the last pgn_upload_weights
  unsigned long long* d_prof;   // optional phase timers [num_sms][32]
  bool prof_on;
  float* d_tm = nullptr;    // T = dG^T h7 [128,256] scratch of pgn_mlp_weight_grads
  int* d_epoch = nullptr;   // epoch counters of the weight-gradient kernel's soft lock-step
  __nv_bfloat16* d_igw[2] = {nullptr, nullptr};   // per net: bf16 [W_5[:, :432] | W_0 | W_v[:, 256:904]] for pgn_mlp_input_grads
  int view_in = 904;        // input width of views_linears.0: 904, + framecode_ch with Optcodes
  int n_codes = 0;          // frame codes per net (0: none)
  float* d_codes_ext[2] = {nullptr, nullptr};   // [n_codes + 1][16]: the codes + their mean
  float* d_fc_table[2] = {nullptr, nullptr};    // [n_codes + 1][128]: W_v[:, 904:920] codes^T
  float* d_c2w;
  float* d_rest;            // rest pose [24,3] of the last pgn_pose_to_skts call
  int64_t launches;
};

extern "C" {

int pgn_abi_version(void) { return PGN_ABI_VERSION; }
const char* pgn_last_error(void) { return g_err; }

int pgn_create(const pgn_config* cfg, pgn_context** out) {
  if (!cfg || !out) return fail(PGN_E_INVALID, "pgn_create: null argument");
  if (cfg->n_joints != PGN_J || cfg->n_samples != PGN_S || cfg->n_importance != PGN_I || cfg->multires != PGN_LV ||
      cfg->multires_views != PGN_LD || cfg->net_depth != 8 || cfg->net_width != PGN_W || cfg->skip_layer != 4)
    return fail(PGN_E_INVALID, "pgn_create: only the surreal.txt architecture is supported "
                "(24 joints, 64+16 samples, multires 7/4, 8x256 MLP, skip 4)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(PGN_E_CUDA, "pgn_create: no CUDA device (%s); posegen_b200 has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (cfg->n_framecodes < 0 || (cfg->n_framecodes > 0 && cfg->framecode_ch != 16))
    return fail(PGN_E_INVALID, "pgn_create: frame codes (Optcodes) need n_framecodes >= 0 and framecode_ch == 16");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(PGN_E_INVALID, "pgn_create: bad device ordinal %d", cfg->device);
  DeviceGuard _guard(cfg->device);
  if (_guard.err != cudaSuccess) return fail(PGN_E_CUDA, "cudaSetDevice(%d): %s", cfg->device, cudaGetErrorString(_guard.err));
  cudaDeviceProp prop;
  PGN_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return fail(PGN_E_INVALID, "pgn_create: device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
  pgn_context* c = new (std::nothrow) pgn_context();
  if (!c) return fail(PGN_E_INVALID, "pgn_create: out of host memory");
  memset(c, 0, sizeof(*c));
  c->cfg = *cfg;
  c->num_sms = prop.multiProcessorCount;
  c->n_codes = cfg->n_framecodes;
  c->view_in = 904 + (cfg->n_framecodes > 0 ? cfg->framecode_ch : 0);
  for (int i = 0; i < PGN_S; ++i) c->h_sc.t_coarse[i] = pgn_linspace01(i, PGN_S);
  for (int i = 0; i < PGN_I; ++i) c->h_sc.u_det[i] = pgn_linspace01(i, PGN_I);
  PGN_CUDA(cudaMalloc(&c->d_sc, sizeof(PgnScalars)));
  PGN_CUDA(cudaMalloc(&c->d_status, sizeof(int)));
  PGN_CUDA(cudaMemset(c->d_status, 0, sizeof(int)));
  PGN_CUDA(cudaMalloc(&c->d_prof, (size_t)c->num_sms * 32 * sizeof(unsigned long long)));
  PGN_CUDA(cudaMemset(c->d_prof, 0, (size_t)c->num_sms * 32 * sizeof(unsigned long long)));
  PGN_CUDA(cudaMalloc(&c->d_c2w, 12 * sizeof(float)));
  PGN_CUDA(cudaMalloc(&c->d_rest, PGN_J * 3 * sizeof(float)));
  PGN_CUDA(cudaMalloc(&c->d_fold, (128 * 256 + 128) * sizeof(float)));
  PGN_CUDA(cudaMalloc(&c->d_tm, 128 * 256 * sizeof(float)));
  PGN_CUDA(cudaMalloc(&c->d_epoch, PGN_WGRAD_MAX_EPOCHS * sizeof(int)));
  for (int n = 0; n < 2; ++n) PGN_CUDA(cudaMalloc(&c->d_chain_w[n], (size_t)120 * 4096 * sizeof(__nv_bfloat16)));
  for (int n = 0; n < 2; ++n) PGN_CUDA(cudaMalloc(&c->d_igw[n], pgn_input_grad_weight_elems() * sizeof(__nv_bfloat16)));
  for (int n = 0; n < 2; ++n) {
    PGN_CUDA(cudaMalloc(&c->d_w[n], weight_floats(c->view_in) * sizeof(float)));
    PGN_CUDA(cudaMalloc(&c->d_b[n], bias_floats() * sizeof(float)));
    PGN_CUDA(cudaMalloc(&c->d_wt[n], weight_floats(c->view_in) * sizeof(float)));
    if (c->n_codes > 0) {
      PGN_CUDA(cudaMalloc(&c->d_codes_ext[n], (size_t)(c->n_codes + 1) * 16 * sizeof(float)));
      PGN_CUDA(cudaMalloc(&c->d_fc_table[n], (size_t)(c->n_codes + 1) * 128 * sizeof(float)));
    }
    PGN_CUDA(cudaMalloc(&c->d_wstream[n], pgn_bf16_wstream_elems() * sizeof(__nv_bfloat16)));
    PGN_CUDA(cudaMalloc(&c->d_bf16_aux[n], (9 * 256 + 256 + 384) * sizeof(float)));
    size_t wo = 0, bo = 0;
    for (int l = 0; l < PGN_N_LINEAR; ++l) {
      c->w_ptr[n][l] = c->d_w[n] + wo;
      c->b_ptr[n][l] = c->d_b[n] + bo;
      c->fp32[n].wt[l] = c->d_wt[n] + wo;
      c->fp32[n].b[l] = c->d_b[n] + bo;
      wo += (size_t)kOut[l] * in_of(l, c->view_in);
      bo += kOut[l];
    }
    c->fp32[n].w_alpha = c->w_ptr[n][8];
    c->fp32[n].b_alpha = c->b_ptr[n][8];
    c->fp32[n].w_rgb = c->w_ptr[n][11];
    c->fp32[n].b_rgb = c->b_ptr[n][11];
    c->fp32[n].codes_ext = c->d_codes_ext[n];
    c->bf16[n].fc_table = c->d_fc_table[n];
    c->bf16[n].wstream = c->d_wstream[n];
    c->bf16[n].bias = c->d_bf16_aux[n];
    c->bf16[n].w_alpha = c->d_bf16_aux[n] + 9 * 256;
    c->bf16[n].w_rgb = c->d_bf16_aux[n] + 9 * 256 + 256;
    c->bf16[n].b_alpha = c->b_ptr[n][8];
    c->bf16[n].b_rgb = c->b_ptr[n][11];
  }
  *out = c;
  return PGN_OK;
}

void pgn_destroy(pgn_context* c) {
  if (!c) return;
  DeviceGuard _guard(c->cfg.device);
  cudaFree(c->d_sc); cudaFree(c->d_status); cudaFree(c->d_c2w); cudaFree(c->d_rest); cudaFree(c->d_fold); cudaFree(c->d_tm); cudaFree(c->d_epoch); cudaFree(c->d_chain_w[0]); cudaFree(c->d_chain_w[1]); cudaFree(c->d_igw[0]); cudaFree(c->d_igw[1]);
  for (int n = 0; n < 2; ++n) {
    cudaFree(c->d_codes_ext[n]); cudaFree(c->d_fc_table[n]);
    cudaFree(c->d_w[n]); cudaFree(c->d_b[n]); cudaFree(c->d_wt[n]); cudaFree(c->d_wstream[n]); cudaFree(c->d_bf16_aux[n]);
  }
  delete c;
}

int pgn_upload_weights(pgn_context* c, int net_id, const pgn_net_weights* w, int pointers_are_device, void* stream_) {
  if (!c || !w || net_id < 0 || net_id > 1) return fail(PGN_E_INVALID, "pgn_upload_weights: bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  PGN_ON_DEVICE(c);
  const cudaMemcpyKind kind = pointers_are_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  for (int l = 0; l < PGN_N_LINEAR; ++l)
    if (!w->weight[l] || !w->bias[l]) return fail(PGN_E_INVALID, "pgn_upload_weights: null tensor %d", l);
  if (pointers_are_device) {          // the per-step refresh: one gather launch for the 24 tensors
    float* dw[PGN_N_LINEAR]; float* db[PGN_N_LINEAR];
    int nw[PGN_N_LINEAR], nb[PGN_N_LINEAR];
    for (int l = 0; l < PGN_N_LINEAR; ++l) {
      dw[l] = (float*)c->w_ptr[net_id][l]; db[l] = (float*)c->b_ptr[net_id][l];
      nw[l] = kOut[l] * in_of(l, c->view_in); nb[l] = kOut[l];
    }
    PGN_CUDA(pgn_launch_gather_params(w->weight, w->bias, dw, db, nw, nb, stream));
    c->launches++;
  } else {
    for (int l = 0; l < PGN_N_LINEAR; ++l) {
      PGN_CUDA(cudaMemcpyAsync((void*)c->w_ptr[net_id][l], w->weight[l], (size_t)kOut[l] * in_of(l, c->view_in) * sizeof(float), kind, stream));
      PGN_CUDA(cudaMemcpyAsync((void*)c->b_ptr[net_id][l], w->bias[l], (size_t)kOut[l] * sizeof(float), kind, stream));
    }
  }
  // the fp32 CUDA-core tier's transposed copies are rebuilt lazily, by the first fp32 call after an upload
  // (ensure_fp32_tier): a bf16 training loop re-packs every step and never reads them
  c->fp32_stale[net_id] = true;
  PGN_CUDA(pgn_pack_bf16_net(c->w_ptr[net_id], c->b_ptr[net_id], c->d_wstream[net_id], c->d_bf16_aux[net_id],
                             c->d_bf16_aux[net_id] + 9 * 256, c->d_bf16_aux[net_id] + 9 * 256 + 256, c->d_fold, c->view_in, stream));
  if (c->n_codes > 0) {         // Optcodes: the codes, their mean and the per-code view-layer term
    if (!w->framecodes) return fail(PGN_E_INVALID, "pgn_upload_weights: the context was created with frame codes but framecodes is NULL");
    PGN_CUDA(cudaMemcpyAsync(c->d_codes_ext[net_id], w->framecodes, (size_t)c->n_codes * 16 * sizeof(float), kind, stream));
    PGN_CUDA(pgn_launch_framecode_tables(c->d_codes_ext[net_id], c->n_codes, c->w_ptr[net_id][10], c->view_in, c->d_fc_table[net_id], stream));
    c->launches++;
  }
  // the backward's delta-chain weight stream (reads the fold the pack above left in d_fold)
  PGN_CUDA(pgn_launch_pack_chain_weights(c->w_ptr[net_id], c->d_fold, c->d_chain_w[net_id], stream));
  // bf16 weight blocks of the input-gradient kernel (pose gradient)
  PGN_CUDA(pgn_launch_pack_input_grad_weights(c->w_ptr[net_id][5], c->w_ptr[net_id][0], c->w_ptr[net_id][10], c->view_in, c->d_igw[net_id], stream));
  c->launches += 4;
  c->have_w[net_id] = true;
  return PGN_OK;
}

// fp32 tier: transposed weights from the context's fp32 copies, on the calling stream (which must be ordered after the
// upload's stream; both are the caller's current stream in the Python binding)
static int ensure_fp32_tier(pgn_context* c, cudaStream_t stream) {
  for (int n = 0; n < 2; ++n) {
    if (!c->fp32_stale[n] || !c->have_w[n]) continue;
    for (int l = 0; l < PGN_N_LINEAR; ++l) {
      if (l == 8 || l == 11) continue;   // small heads keep nn.Linear layout
      PGN_CUDA(pgn_launch_transpose(c->w_ptr[n][l], kOut[l], in_of(l, c->view_in), (float*)c->fp32[n].wt[l], stream));
      c->launches++;
    }
    c->fp32_stale[n] = false;
  }
  return PGN_OK;
}

int pgn_set_embed_scalars(pgn_context* c, float tau_v, float tau_d, const float* cutoff_v, const float* cutoff_d,
                          float density_scale, float rgb_eps) {
  if (!c || !cutoff_v || !cutoff_d) return fail(PGN_E_INVALID, "pgn_set_embed_scalars: null argument");
  if (!(density_scale > 0.f)) return fail(PGN_E_INVALID, "pgn_set_embed_scalars: density_scale must be > 0");
  PGN_ON_DEVICE(c);
  c->h_sc.tau_v = tau_v; c->h_sc.tau_d = tau_d;
  memcpy(c->h_sc.cutoff_v, cutoff_v, PGN_J * sizeof(float));
  memcpy(c->h_sc.cutoff_d, cutoff_d, PGN_J * sizeof(float));
  c->h_sc.density_scale = density_scale; c->h_sc.rgb_eps = rgb_eps;
  PGN_CUDA(cudaMemcpy(c->d_sc, &c->h_sc, sizeof(PgnScalars), cudaMemcpyHostToDevice));
  c->have_sc = true;
  return PGN_OK;
}

size_t pgn_workspace_bytes(const pgn_context* c, int64_t n_rays) {
  (void)c;
  if (n_rays < 0) n_rays = 0;
  return ((size_t)n_rays * 2 * sizeof(float) + 255) / 256 * 256 + 256;
}

static int check_inputs(const pgn_context* c, const pgn_render_inputs* in, const char* who) {
  if (!c || !in) return fail(PGN_E_INVALID, "%s: null argument", who);
  if (in->n_rays < 0) return fail(PGN_E_INVALID, "%s: negative n_rays", who);
  if (in->n_rays > 0 && (!in->ray_batch || !in->skts || !in->cyls)) return fail(PGN_E_INVALID, "%s: null input tensor", who);
  if (!in->pose_idx) {
    if (in->skts_stride != 0 && in->skts_stride != PGN_J * 16) return fail(PGN_E_INVALID, "%s: skts_stride must be 0 or 384", who);
    if (in->cyls_stride != 0 && in->cyls_stride != 5) return fail(PGN_E_INVALID, "%s: cyls_stride must be 0 or 5", who);
  }
  return PGN_OK;
}

static PgnRayRefs make_refs(const pgn_context* c, const pgn_render_inputs* in) {
  PgnRayRefs r;
  r.ray_batch = in->ray_batch; r.n_rays = in->n_rays; r.skts = in->skts; r.skts_stride = in->skts_stride;
  r.cyls = in->cyls; r.cyls_stride = in->cyls_stride; r.pose_idx = in->pose_idx;
  r.cams = c->n_codes > 0 ? in->cams : nullptr; r.n_codes = c->n_codes;
  r.lindisp = in->lindisp != 0;
  return r;
}

static int render_forward_impl(pgn_context* c, const pgn_render_inputs* in, const pgn_render_outputs* out,
                               void* workspace, size_t workspace_bytes, void* stream_, const PgnActDump* dump);

int pgn_render_forward(pgn_context* c, const pgn_render_inputs* in, const pgn_render_outputs* out,
                       void* workspace, size_t workspace_bytes, void* stream_) {
  return render_forward_impl(c, in, out, workspace, workspace_bytes, stream_, nullptr);
}

size_t pgn_activation_dump_bytes(int64_t n_rays, int32_t pass) {
  if (n_rays < 0 || (pass != 0 && pass != 1)) return 0;
  // activations (8 x 256 + 128 bf16 per row) followed by the ReLU masks of the 8 trunk layers (256 bits per row each)
  return (size_t)pgn_bf16_dump_rows(n_rays, pass == 0 ? PGN_S : PGN_T, 0) * ((8 * 256 + 128) * sizeof(__nv_bfloat16) + 8 * 32);
}

int pgn_render_forward_train(pgn_context* c, const pgn_render_inputs* in, const pgn_render_outputs* out,
                             void* act_coarse, void* act_fine, const pgn_train_random* rnd,
                             void* workspace, size_t workspace_bytes, void* stream_) {
  if (!in || in->precision != PGN_PRECISION_BF16) return fail(PGN_E_INVALID, "pgn_render_forward_train: the bf16 tensor-core path only");
  if (!act_coarse && !act_fine) return fail(PGN_E_INVALID, "pgn_render_forward_train: null activation dump");
  PgnActDump d;
  d.c = (__nv_bfloat16*)act_coarse; d.f = (__nv_bfloat16*)act_fine;
  d.masks_only = 0;
  d.rows_c = pgn_bf16_dump_rows(in->n_rays, PGN_S, 0); d.rows_f = pgn_bf16_dump_rows(in->n_rays, PGN_T, 0);
  d.t_rand = rnd ? rnd->t_rand : nullptr; d.u_is = rnd ? rnd->u_is : nullptr;
  d.noise0 = rnd ? rnd->noise0 : nullptr; d.noise = rnd ? rnd->noise : nullptr;
  return render_forward_impl(c, in, out, workspace, workspace_bytes, stream_, &d);
}

size_t pgn_mask_dump_bytes(int64_t n_rays) {
  if (n_rays < 0) return 0;
  return (size_t)pgn_bf16_dump_rows(n_rays, PGN_T, 1) * (8 * 32 + 16);
}

int pgn_render_forward_masks(pgn_context* c, const pgn_render_inputs* in, const pgn_render_outputs* out, void* masks_fine,
                             void* workspace, size_t workspace_bytes, void* stream_) {
  if (!in || in->precision != PGN_PRECISION_BF16) return fail(PGN_E_INVALID, "pgn_render_forward_masks: the bf16 tensor-core path only");
  if (!masks_fine) return fail(PGN_E_INVALID, "pgn_render_forward_masks: null mask buffer");
  PgnActDump d;
  d.c = nullptr; d.f = (__nv_bfloat16*)masks_fine;
  d.rows_c = pgn_bf16_dump_rows(in->n_rays, PGN_S, 1); d.rows_f = pgn_bf16_dump_rows(in->n_rays, PGN_T, 1);
  d.masks_only = 1;
  d.t_rand = nullptr; d.u_is = nullptr; d.noise0 = nullptr; d.noise = nullptr;
  return render_forward_impl(c, in, out, workspace, workspace_bytes, stream_, &d);
}

static int render_forward_impl(pgn_context* c, const pgn_render_inputs* in, const pgn_render_outputs* out,
                               void* workspace, size_t workspace_bytes, void* stream_, const PgnActDump* dump) {
  int rc = check_inputs(c, in, "pgn_render_forward");
  if (rc) return rc;
  if (!out) return fail(PGN_E_INVALID, "pgn_render_forward: null outputs");
  if (!c->have_sc || !c->have_w[0] || !c->have_w[1]) return fail(PGN_E_STATE, "pgn_render_forward: weights/scalars not uploaded");
  if (in->precision != PGN_PRECISION_FP32 && in->precision != PGN_PRECISION_BF16) return fail(PGN_E_INVALID, "pgn_render_forward: bad precision");
  if (in->n_rays == 0) return PGN_OK;
  if (!workspace || workspace_bytes < pgn_workspace_bytes(c, in->n_rays)) return fail(PGN_E_INVALID, "pgn_render_forward: workspace too small");
  cudaStream_t stream = (cudaStream_t)stream_;
  PGN_ON_DEVICE(c);
  const PgnRayRefs refs = make_refs(c, in);
  float* near_far = (float*)workspace;
  if (in->chunk_starts) {
    if (in->n_chunks <= 0) return fail(PGN_E_INVALID, "pgn_render_forward: chunk_starts without n_chunks");
    PGN_CUDA(pgn_launch_near_far_chunks(refs, (const long long*)in->chunk_starts, in->n_chunks, near_far, stream));
  } else
  PGN_CUDA(pgn_launch_near_far(refs, in->nanfill_chunk, near_far, stream));
  c->launches++;
  PgnOutputs o;
  o.rgb_map = out->rgb_map; o.disp_map = out->disp_map; o.acc_map = out->acc_map; o.alpha = out->alpha;
  o.rgb0 = out->rgb0; o.disp0 = out->disp0; o.acc0 = out->acc0; o.alpha0 = out->alpha0;
  o.z_samples = out->z_samples; o.z_fine = out->z_fine; o.pdf_inds = out->pdf_inds; o.weights0 = out->weights0;
  o.raw0 = out->raw0; o.raw = out->raw; o.near_far = out->near_far;
  if (out->near_far)
    PGN_CUDA(cudaMemcpyAsync(out->near_far, near_far, (size_t)in->n_rays * 2 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  if (in->precision == PGN_PRECISION_FP32) {
    int rc2 = ensure_fp32_tier(c, stream);
    if (rc2) return rc2;
  }
  if (in->precision == PGN_PRECISION_FP32)
    PGN_CUDA(pgn_launch_render_fp32(refs, o, c->fp32[0], c->fp32[1], c->d_sc, near_far, c->num_sms, stream));
  else
    PGN_CUDA(pgn_launch_render_bf16(refs, o, c->bf16[0], c->bf16[1], c->d_sc, near_far, c->d_status, c->prof_on ? c->d_prof : nullptr, dump, c->num_sms, stream));
  c->launches++;
  return PGN_OK;
}

int64_t pgn_launch_count(const pgn_context* c) { return c ? c->launches : 0; }

int pgn_check_device_status(pgn_context* c) {
  if (!c) return fail(PGN_E_INVALID, "pgn_check_device_status: null context");
  PGN_ON_DEVICE(c);
  int st = 0;
  PGN_CUDA(cudaMemcpy(&st, c->d_status, sizeof(int), cudaMemcpyDeviceToHost));
  if (st != 0) {
    cudaMemset(c->d_status, 0, sizeof(int));
    return fail(PGN_E_KERNEL, st == 950 ? "device status %d: pgn_gather_ray_rows index out of range"
                                         : "device watchdog tripped: pipeline wait code %d", st);
  }
  return PGN_OK;
}

const int32_t* pgn_device_status_ptr(pgn_context* c) { return c ? c->d_status : nullptr; }

int pgn_near_far(pgn_context* c, const pgn_render_inputs* in, float* near_far, void* stream) {
  int rc = check_inputs(c, in, "pgn_near_far");
  if (rc) return rc;
  if (!near_far) return fail(PGN_E_INVALID, "pgn_near_far: null output");
  PGN_ON_DEVICE(c);
  if (in->chunk_starts)
    PGN_CUDA(pgn_launch_near_far_chunks(make_refs(c, in), (const long long*)in->chunk_starts, in->n_chunks, near_far, (cudaStream_t)stream));
  else
  PGN_CUDA(pgn_launch_near_far(make_refs(c, in), in->nanfill_chunk, near_far, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_encode(pgn_context* c, const pgn_render_inputs* in, const float* z, int32_t n_z, float* enc, void* stream) {
  int rc = check_inputs(c, in, "pgn_encode");
  if (rc) return rc;
  if (!z || !enc || n_z <= 0) return fail(PGN_E_INVALID, "pgn_encode: bad argument");
  if (!c->have_sc) return fail(PGN_E_STATE, "pgn_encode: scalars not set");
  PGN_ON_DEVICE(c);
  PGN_CUDA(pgn_launch_encode(make_refs(c, in), c->d_sc, z, n_z, enc, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_encode_bf16(pgn_context* c, const pgn_render_inputs* in, const float* z, int32_t n_z, void* enc, void* stream) {
  int rc = check_inputs(c, in, "pgn_encode_bf16");
  if (rc) return rc;
  if (!z || !enc || n_z <= 0) return fail(PGN_E_INVALID, "pgn_encode_bf16: bad argument");
  if (!c->have_sc) return fail(PGN_E_STATE, "pgn_encode_bf16: scalars not set");
  PGN_ON_DEVICE(c);
  PGN_CUDA(pgn_launch_encode_bf16(make_refs(c, in), c->d_sc, z, n_z, reinterpret_cast<__nv_bfloat16*>(enc), (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_mlp_delta(pgn_context* c, void* dh, int32_t has_input, const void* act, int64_t m, int32_t n_cols, const float* rs,
                  int32_t rs_stride, int32_t nrs, const float* wr, float* colsum, float* wsum, void* stream) {
  if (!c || !dh || !colsum || m < 0) return fail(PGN_E_INVALID, "pgn_mlp_delta: bad argument");
  if (!((n_cols == 256 && (nrs == 0 || nrs == 1)) || (n_cols == 128 && (nrs == 0 || nrs == 3))))
    return fail(PGN_E_INVALID, "pgn_mlp_delta: supported shapes are 256 columns with 0/1 head rows, 128 columns with 0/3");
  if (nrs > 0 && (!rs || !wr || rs_stride < nrs)) return fail(PGN_E_INVALID, "pgn_mlp_delta: head deltas / weights missing");
  if (!has_input && nrs == 0) return fail(PGN_E_INVALID, "pgn_mlp_delta: nothing to do");
  PGN_ON_DEVICE(c);
  PGN_CUDA(pgn_launch_mlp_delta(dh, has_input, act, m, n_cols, rs, rs_stride, nrs, wr, colsum, wsum, c->num_sms, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_view_delta_from_mask(pgn_context* c, void* dG, const float* d_raw, const float* w_rgb, const void* vmask, int64_t m,
                             void* stream) {
  if (!c || !dG || !d_raw || !w_rgb || !vmask || m < 0) return fail(PGN_E_INVALID, "pgn_view_delta_from_mask: bad argument");
  PGN_ON_DEVICE(c);
  PGN_CUDA(pgn_launch_view_delta_bits(dG, d_raw, w_rgb, vmask, m, c->num_sms, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_mlp_delta_chain(pgn_context* c, const void* dG, const float* d_raw, const void* mask, int64_t mask_rows, int64_t m,
                        const void* wstream, const float* w_alpha, void* dz, float* colsum, uint32_t layer_mask, void* stream) {
  if (!c || !dG || !d_raw || !mask || !wstream || !w_alpha || !dz || !colsum || m < 0 || mask_rows < m)
    return fail(PGN_E_INVALID, "pgn_mlp_delta_chain: bad argument");
  PGN_ON_DEVICE(c);
  PGN_CUDA(pgn_launch_delta_chain(dG, d_raw, mask, mask_rows, m, wstream, w_alpha, dz, colsum, layer_mask, c->d_status, c->num_sms, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_mlp_delta_chain_net(pgn_context* c, int32_t net_id, const void* dG, const float* d_raw, const void* mask,
                            int64_t mask_rows, int64_t m, void* dz, float* colsum, uint32_t layer_mask, void* stream) {
  if (!c || net_id < 0 || net_id > 1) return fail(PGN_E_INVALID, "pgn_mlp_delta_chain_net: bad argument");
  if (!c->have_w[net_id]) return fail(PGN_E_STATE, "pgn_mlp_delta_chain_net: weights not uploaded");
  return pgn_mlp_delta_chain(c, dG, d_raw, mask, mask_rows, m, c->d_chain_w[net_id], c->w_ptr[net_id][8], dz, colsum, layer_mask, stream);
}

size_t pgn_weight_grad_floats(const pgn_context* c) { return pgn_wgrad_flat_floats_ld(c ? c->view_in : 904); }

int pgn_framecode_backward(pgn_context* c, int32_t net_id, const void* dG, int64_t n_rays, int32_t n_z, const int32_t* cams,
                           float* g_view_weight, float* g_codes, void* stream) {
  if (!c || net_id < 0 || net_id > 1 || !dG || !g_view_weight || !g_codes || n_rays < 0 || (n_z != PGN_S && n_z != PGN_T))
    return fail(PGN_E_INVALID, "pgn_framecode_backward: bad argument");
  if (c->n_codes <= 0) return fail(PGN_E_STATE, "pgn_framecode_backward: the context has no frame codes");
  if (!c->have_w[net_id]) return fail(PGN_E_STATE, "pgn_framecode_backward: weights not uploaded");
  PGN_ON_DEVICE(c);
  PGN_CUDA(pgn_launch_framecode_backward(dG, n_rays, n_z, cams, c->n_codes, c->d_codes_ext[net_id], c->w_ptr[net_id][10], c->view_in,
                                         g_view_weight + 904, c->view_in, g_codes, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_mlp_weight_grads(pgn_context* c, int32_t net_id, const void* dz, const void* dG, const void* act, int64_t dump_rows,
                         const void* enc, int64_t m, const float* d_raw, const float* bias_v, float* flat, float* feat_bias,
                         void* stream) {
  if (!c || net_id < 0 || net_id > 1 || !dz || !dG || !act || !enc || !d_raw || !bias_v || !flat || !feat_bias || m < 0 || dump_rows < m)
    return fail(PGN_E_INVALID, "pgn_mlp_weight_grads: bad argument");
  if (!c->have_w[net_id]) return fail(PGN_E_STATE, "pgn_mlp_weight_grads: weights not uploaded");
  PGN_ON_DEVICE(c);
  PGN_CUDA(pgn_launch_weight_grads(dz, dG, act, dump_rows, enc, m, d_raw, bias_v, c->w_ptr[net_id][9], c->b_ptr[net_id][9],
                                   c->w_ptr[net_id][10], c->view_in, flat, feat_bias, c->d_tm, c->d_epoch, c->d_status, c->num_sms, (cudaStream_t)stream));
  c->launches += 3;
  return PGN_OK;
}

int pgn_mlp_input_grads(pgn_context* c, int32_t net_id, const void* dz, const void* dG, int64_t m, void* g_xp, void* g_d,
                        int32_t tile_blocked, void* stream) {
  if (!c || net_id < 0 || net_id > 1 || !dz || !dG || !g_xp || !g_d || m < 0) return fail(PGN_E_INVALID, "pgn_mlp_input_grads: bad argument");
  if (!c->have_w[net_id]) return fail(PGN_E_STATE, "pgn_mlp_input_grads: weights not uploaded");
  PGN_ON_DEVICE(c);
  PGN_CUDA(pgn_launch_input_grads(dz, dG, m, c->d_igw[net_id], g_xp, g_d, tile_blocked ? 1 : 0, c->d_status, c->num_sms, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_debug_wgrad(pgn_context* c, const void* A, int32_t lda, int32_t Ma, const void* B, int32_t ldb, int32_t Nb, int64_t m,
                    float* out, int32_t ld_out, int32_t n_ctas, int32_t b_tile_blocked, void* stream) {
  if (!c || !A || !B || !out || (Ma != 128 && Ma != 256) || Nb <= 0 || Nb > 256 || Nb % 8 || lda % 8 || ldb % 8 || ld_out % 4 || m < 0 || n_ctas <= 0)
    return fail(PGN_E_INVALID, "pgn_debug_wgrad: bad argument");
  PGN_ON_DEVICE(c);
  if (b_tile_blocked && Nb != 256) return fail(PGN_E_INVALID, "pgn_debug_wgrad: a tile-blocked B is a 256-column matrix");
  PGN_CUDA(pgn_launch_wgrad_single(A, lda, Ma, B, ldb, Nb, m, out, ld_out, n_ctas, b_tile_blocked, c->d_epoch, c->d_status, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_mlp(pgn_context* c, int net_id, const float* enc, int64_t m, float* raw, int32_t precision, void* stream) {
  if (!c || !enc || !raw || net_id < 0 || net_id > 1 || m < 0) return fail(PGN_E_INVALID, "pgn_mlp: bad argument");
  if (!c->have_w[net_id]) return fail(PGN_E_STATE, "pgn_mlp: weights not uploaded");
  PGN_ON_DEVICE(c);
  if (precision == PGN_PRECISION_FP32) {
    int rc2 = ensure_fp32_tier(c, (cudaStream_t)stream);
    if (rc2) return rc2;
  }
  if (precision == PGN_PRECISION_FP32)
    PGN_CUDA(pgn_launch_mlp_fp32(c->fp32[net_id], enc, m, raw, c->num_sms, c->n_codes, (cudaStream_t)stream));
  else if (precision == PGN_PRECISION_BF16)
    PGN_CUDA(pgn_launch_mlp_bf16(c->bf16[net_id], enc, m, raw, c->d_sc, c->d_status, c->num_sms, c->n_codes, (cudaStream_t)stream));
  else return fail(PGN_E_INVALID, "pgn_mlp: bad precision");
  c->launches++;
  return PGN_OK;
}

int pgn_composite(pgn_context* c, const pgn_render_inputs* in, const float* raw, const float* z, int32_t s,
                  float* rgb_map, float* disp_map, float* acc_map, float* weights, float* alpha, void* stream) {
  int rc = check_inputs(c, in, "pgn_composite");
  if (rc) return rc;
  if (!raw || !z || (s != PGN_S && s != PGN_T)) return fail(PGN_E_INVALID, "pgn_composite: s must be 64 or 80");
  if (!c->have_sc) return fail(PGN_E_STATE, "pgn_composite: scalars not set");
  PGN_ON_DEVICE(c);
  PGN_CUDA(pgn_launch_composite(make_refs(c, in), c->d_sc, raw, z, s, rgb_map, disp_map, acc_map, weights, alpha, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_encode_backward(pgn_context* c, const pgn_render_inputs* in, const float* z, int32_t n_z, const float* g_enc,
                        float* d_skts, void* stream) {
  int rc = check_inputs(c, in, "pgn_encode_backward");
  if (rc) return rc;
  if (!z || !g_enc || !d_skts || n_z <= 0) return fail(PGN_E_INVALID, "pgn_encode_backward: bad argument");
  if (!c->have_sc) return fail(PGN_E_STATE, "pgn_encode_backward: scalars not set");
  PGN_ON_DEVICE(c);
  PGN_CUDA(pgn_launch_encode_backward(make_refs(c, in), c->d_sc, z, n_z, g_enc, d_skts, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_encode_backward_bf16(pgn_context* c, const pgn_render_inputs* in, const float* z, int32_t n_z, const void* g_xp,
                             const void* g_d, int32_t tile_blocked, float* d_skts, void* stream) {
  int rc = check_inputs(c, in, "pgn_encode_backward_bf16");
  if (rc) return rc;
  if (!z || !g_xp || !g_d || !d_skts || n_z <= 0) return fail(PGN_E_INVALID, "pgn_encode_backward_bf16: bad argument");
  if (!c->have_sc) return fail(PGN_E_STATE, "pgn_encode_backward_bf16: scalars not set");
  PGN_ON_DEVICE(c);
  PGN_CUDA(pgn_launch_encode_backward_bf16(make_refs(c, in), c->d_sc, z, n_z, reinterpret_cast<const __nv_bfloat16*>(g_xp),
                                           reinterpret_cast<const __nv_bfloat16*>(g_d), tile_blocked ? 1 : 0, d_skts, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_composite_backward(pgn_context* c, const pgn_render_inputs* in, const float* raw, const float* z, int32_t s,
                           const float* g_rgb, const float* g_acc, const float* noise, float* d_raw, void* stream) {
  int rc = check_inputs(c, in, "pgn_composite_backward");
  if (rc) return rc;
  if (!raw || !z || !g_rgb || !d_raw || (s != PGN_S && s != PGN_T)) return fail(PGN_E_INVALID, "pgn_composite_backward: bad argument");
  if (!c->have_sc) return fail(PGN_E_STATE, "pgn_composite_backward: scalars not set");
  PGN_ON_DEVICE(c);
  PGN_CUDA(pgn_launch_composite_backward(make_refs(c, in), c->d_sc, raw, z, s, g_rgb, g_acc, noise, d_raw, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_sample_pdf(pgn_context* c, const float* z, const float* weights, int64_t n, float* z_samples, float* z_sorted,
                   int32_t* pdf_inds, int32_t* sorted_idxs, void* stream) {
  if (!c || !z || !weights || n < 0) return fail(PGN_E_INVALID, "pgn_sample_pdf: bad argument");
  PGN_ON_DEVICE(c);
  if (!c->have_sc) {  // only the linspace tables are needed
    PGN_CUDA(cudaMemcpy(c->d_sc, &c->h_sc, sizeof(PgnScalars), cudaMemcpyHostToDevice));
  }
  PGN_CUDA(pgn_launch_sample_pdf(c->d_sc, z, weights, n, z_samples, z_sorted, pdf_inds, sorted_idxs, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_generate_rays(pgn_context* c, int32_t H, int32_t W, float focal, const float* c2w, int32_t x0, int32_t y0,
                      int32_t x1, int32_t y1, float* ray_batch, void* stream_) {
  if (!c || !c2w || !ray_batch || x1 < x0 || y1 < y0) return fail(PGN_E_INVALID, "pgn_generate_rays: bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  PGN_ON_DEVICE(c);
  PGN_CUDA(cudaMemcpyAsync(c->d_c2w, c2w, 12 * sizeof(float), cudaMemcpyHostToDevice, stream));
  PGN_CUDA(pgn_launch_generate_rays(H, W, focal, c->d_c2w, x0, y0, x1, y1, ray_batch, stream));
  c->launches++;
  return PGN_OK;
}

int pgn_compose_frame(pgn_context* c, int32_t H, int32_t W, int32_t x0, int32_t y0, int32_t x1, int32_t y1,
                      const float* rgb_map, const float* acc_map, float bg, float* image, void* stream) {
  if (!c || !image || ((x1 > x0 && y1 > y0) && (!rgb_map || !acc_map))) return fail(PGN_E_INVALID, "pgn_compose_frame: bad argument");
  PGN_ON_DEVICE(c);
  PGN_CUDA(pgn_launch_compose_frame(H, W, x0, y0, x1, y1, rgb_map, acc_map, bg, image, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_pose_to_skts(pgn_context* c, const float* bones, const float* rest_pose, int32_t n_poses, float cyl_extend,
                     float top_expand_ratio, float bot_expand_ratio, float* skts, float* kps, float* cyls, float* l2ws,
                     void* stream_) {
  if (!c || !bones || !rest_pose || !skts || n_poses < 0) return fail(PGN_E_INVALID, "pgn_pose_to_skts: bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  PGN_ON_DEVICE(c);
  PGN_CUDA(cudaMemcpyAsync(c->d_rest, rest_pose, PGN_J * 3 * sizeof(float), cudaMemcpyHostToDevice, stream));
  PGN_CUDA(pgn_launch_pose_fk(bones, c->d_rest, n_poses, cyl_extend, top_expand_ratio, bot_expand_ratio, skts, kps, cyls, l2ws, stream));
  c->launches++;
  return PGN_OK;
}

int pgn_pose_fk_backward(pgn_context* c, const float* bones, const float* rest_pose, int32_t n_poses, const float* g_skts,
                         const float* g_kps, float* g_bones, void* stream_) {
  if (!c || !bones || !rest_pose || !g_skts || !g_bones || n_poses < 0) return fail(PGN_E_INVALID, "pgn_pose_fk_backward: bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  PGN_ON_DEVICE(c);
  PGN_CUDA(cudaMemcpyAsync(c->d_rest, rest_pose, PGN_J * 3 * sizeof(float), cudaMemcpyHostToDevice, stream));
  PGN_CUDA(pgn_launch_pose_fk_backward(bones, c->d_rest, n_poses, g_skts, g_kps, g_bones, stream));
  c->launches++;
  return PGN_OK;
}

int pgn_cylinder_bboxes(pgn_context* c, const float* cyls, int32_t n, const double* w2c, int32_t H, int32_t W, float focal,
                        int32_t* bbox, void* stream) {
  if (!c || !cyls || !w2c || !bbox || n < 0 || H <= 0 || W <= 0) return fail(PGN_E_INVALID, "pgn_cylinder_bboxes: bad argument");
  PGN_ON_DEVICE(c);
  PGN_CUDA(pgn_launch_cyl_bboxes(cyls, n, w2c, H, W, (double)focal, bbox, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_generate_rays_batch(pgn_context* c, int32_t H, int32_t W, float focal, const float* c2w, const int32_t* bbox,
                            const int64_t* offsets, int32_t n_poses, int64_t max_rays_per_pose, float* ray_batch,
                            int32_t* pose_idx, void* stream_) {
  if (!c || !c2w || !bbox || !offsets || !ray_batch || !pose_idx || n_poses < 0) return fail(PGN_E_INVALID, "pgn_generate_rays_batch: bad argument");
  cudaStream_t stream = (cudaStream_t)stream_;
  PGN_ON_DEVICE(c);
  PGN_CUDA(cudaMemcpyAsync(c->d_c2w, c2w, 12 * sizeof(float), cudaMemcpyHostToDevice, stream));
  PGN_CUDA(pgn_launch_generate_rays_batch(H, W, focal, c->d_c2w, bbox, (const long long*)offsets, n_poses, max_rays_per_pose,
                                          ray_batch, pose_idx, stream));
  c->launches++;
  return PGN_OK;
}

int pgn_gather_ray_rows(pgn_context* c, const void* src, void* dst, const int64_t* idx, int64_t n_idx, int64_t row_bytes,
                        int32_t n_planes, int64_t src_plane_bytes, int64_t dst_plane_bytes, void* stream) {
  if (!c || !src || !dst || !idx || n_idx < 0 || n_planes < 0 || row_bytes <= 0 || (row_bytes & 15) || (src_plane_bytes & 15) ||
      (dst_plane_bytes & 15) || ((uintptr_t)src & 15) || ((uintptr_t)dst & 15))
    return fail(PGN_E_INVALID, "pgn_gather_ray_rows: bad argument (rows and planes are multiples of 16 bytes, 16-byte aligned)");
  PGN_ON_DEVICE(c);
  if (src_plane_bytes % row_bytes) return fail(PGN_E_INVALID, "pgn_gather_ray_rows: a source plane is a whole number of rows");
  PGN_CUDA(pgn_launch_gather_ray_rows(src, dst, (const long long*)idx, n_idx, row_bytes, n_planes, src_plane_bytes, dst_plane_bytes,
                                      c->d_status, c->num_sms, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_compose_frames_batch(pgn_context* c, int32_t H, int32_t W, const int32_t* bbox, const int64_t* offsets, int32_t n_poses,
                             const float* rgb_map, const float* acc_map, float bg, float* images, void* stream) {
  if (!c || !bbox || !offsets || !images || !rgb_map || !acc_map || n_poses < 0) return fail(PGN_E_INVALID, "pgn_compose_frames_batch: bad argument");
  PGN_ON_DEVICE(c);
  PGN_CUDA(pgn_launch_compose_frames_batch(H, W, bbox, (const long long*)offsets, n_poses, rgb_map, acc_map, bg, images, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_frame_to_hmr_input(pgn_context* c, const float* image, int32_t H, int32_t W, int32_t x0, int32_t y0, int32_t x1,
                           int32_t y1, int32_t out_res, const float* mean3, const float* std3, int32_t quantize_u8,
                           float* out, void* stream) {
  if (!c || !image || !out || !mean3 || !std3 || x0 < 0 || y0 < 0 || x1 > W || y1 > H || x1 <= x0 || y1 <= y0 || out_res <= 0)
    return fail(PGN_E_INVALID, "pgn_frame_to_hmr_input: bad argument");
  PGN_ON_DEVICE(c);
  PGN_CUDA(pgn_launch_hmr_input(image, H, W, x0, y0, x1, y1, out_res, mean3, std3, quantize_u8, out, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

int pgn_debug_phase_timers(pgn_context* c, int32_t enable, uint64_t* out32) {
  if (!c) return fail(PGN_E_INVALID, "pgn_debug_phase_timers: null context");
  PGN_ON_DEVICE(c);
  if (out32) {
    const size_t n = (size_t)c->num_sms * 32;
    unsigned long long* h = new unsigned long long[n];
    PGN_CUDA(cudaMemcpy(h, c->d_prof, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 32; ++k) { unsigned long long s = 0; for (int b = 0; b < c->num_sms; ++b) s += h[(size_t)b * 32 + k]; out32[k] = s / (unsigned long long)c->num_sms; }
    delete[] h;
  }
  c->prof_on = enable != 0;
  return PGN_OK;
}

int pgn_debug_umma_gemm(pgn_context* c, const float* A, const float* B, float* D, int32_t K, int32_t N,
                        int32_t variant, void* stream) {
  if (!c || !A || !B || !D || K <= 0 || K % 16 || N < 16 || N > 256 || N % 16) return fail(PGN_E_INVALID, "pgn_debug_umma_gemm: bad argument");
  if ((size_t)(K / 8) * 2048 + (size_t)(K / 8) * N * 16 > 200 * 1024) return fail(PGN_E_INVALID, "pgn_debug_umma_gemm: K too large");
  if (variant & 1) return fail(PGN_E_INVALID, "pgn_debug_umma_gemm: variant bit 0 (swapped LBO/SBO bring-up experiment) is not part of the shipped ABI");
  PGN_ON_DEVICE(c);
  PGN_CUDA(pgn_launch_probe_umma(A, B, D, K, N, variant, c->d_status, (cudaStream_t)stream));
  c->launches++;
  return PGN_OK;
}

}  // extern "C"
