// Shared device code of the posegen_b200 renderer: parameter blocks and the
// per-ray / per-sample math every kernel uses (near/far, z tables, joint-frame
// geometry, cutoff window, compositing, inverse-CDF resampling).
//
// Each function cites the reference lines it re-implements (paths relative to the
// reference repo).  fp32 throughout; the places where the reference's unfused
// multiply/add order is observable (near/far, z, pts) use explicit _rn intrinsics so
// nvcc cannot contract them into FMAs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define PGN_J 24          // joints
#define PGN_S 64          // coarse samples
#define PGN_I 16          // importance samples
#define PGN_T 80          // S + I
#define PGN_LV 7          // multires (distance PE)
#define PGN_LD 4          // multires_views
#define PGN_ENC 1080      // 360 + 72 + 648
#define PGN_ENC_P 432     // density-net input
#define PGN_ENC_D 648     // view input
#define PGN_W 256

struct PgnScalars {
  float tau_v, tau_d;
  float cutoff_v[PGN_J];
  float cutoff_d[PGN_J];
  float density_scale, rgb_eps;
  float t_coarse[PGN_S];   // torch.linspace(0,1,64)
  float u_det[PGN_I];      // torch.linspace(0,1,16)
};

struct PgnRayRefs {
  const float*   ray_batch;     // [n,11]
  long long      n_rays;
  const float*   skts;
  long long      skts_stride;
  const float*   cyls;
  long long      cyls_stride;
  const int*     pose_idx;
  const int*     cams;          // optional [n] frame-code index per ray (Optcodes, core/networks/embedding.py:4-46); NULL or
                                // an index outside [0, n_codes) selects the mean code (the reference's eval rule for idx < 0)
  int            n_codes;       // rows of the per-net frame-code tables (0: the model has no frame codes)
  int            lindisp;       // coarse samples linear in inverse depth (sample_from_lineseg(lindisp=True), ray_utils.py:224-227)
};

// row of the frame-code tables for ray i: the code of its camera, or row n_codes = the mean code
__device__ __forceinline__ int pgn_ray_code_row(const PgnRayRefs& r, long long i) {
  if (!r.cams) return r.n_codes;
  const int c = r.cams[i];
  return (c < 0 || c >= r.n_codes) ? r.n_codes : c;
}

struct PgnOutputs {
  float *rgb_map, *disp_map, *acc_map, *alpha, *rgb0, *disp0, *acc0, *alpha0;
  float *z_samples, *z_fine; int *pdf_inds; float *weights0, *raw0, *raw, *near_far;
};

__device__ __forceinline__ const float* pgn_ray_skts(const PgnRayRefs& r, long long i) {
  if (r.pose_idx) return r.skts + (long long)r.pose_idx[i] * (PGN_J * 16);
  return r.skts + i * r.skts_stride;
}
__device__ __forceinline__ const float* pgn_ray_cyl(const PgnRayRefs& r, long long i) {
  if (r.pose_idx) return r.cyls + (long long)r.pose_idx[i] * 5;
  return r.cyls + i * r.cyls_stride;
}

// torch.linspace(start=0,end=1,steps=n) on CPU: symmetric evaluation around the midpoint; the
// upper half is end - step*(n-1-i) evaluated with ONE rounding (the vectorised kernel fuses it).
__host__ __device__ inline float pgn_linspace01(int i, int n) {
  float step = 1.0f / (float)(n - 1);
  return (i < n / 2) ? step * (float)i : fmaf(-step, (float)(n - 1 - i), 1.0f);
}

// ---------------------------------------------------------------------------
// get_near_far_in_cylinder, core/utils/ray_utils.py:292-344 (without the NaN fill)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pgn_near_far_ray(const float* __restrict__ rb, const float* __restrict__ cyl,
                                                 float& new_near, float& new_far) {
  const float ox = rb[0], oz = rb[2], dx = rb[3], dz = rb[5], near = rb[6], far = rb[7];
  const float rnx = __fadd_rn(ox, __fmul_rn(dx, near)), rnz = __fadd_rn(oz, __fmul_rn(dz, near));
  const float rfx = __fadd_rn(ox, __fmul_rn(dx, far)),  rfz = __fadd_rn(oz, __fmul_rn(dz, far));
  const float ncx = __fsub_rn(cyl[0], rnx), ncz = __fsub_rn(cyl[1], rnz);
  const float nfx = __fsub_rn(rfx, rnx),    nfz = __fsub_rn(rfz, rnz);
  const float nf_norm = __fsqrt_rn(__fadd_rn(__fmul_rn(nfx, nfx), __fmul_rn(nfz, nfz)));
  const float scale   = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dz, dz)));
  const float cross = __fsub_rn(__fmul_rn(ncx, nfz), __fmul_rn(ncz, nfx));
  const float dist = __fdiv_rn(fabsf(cross), nf_norm);
  const float R = cyl[2];
  const float Q = __fsqrt_rn(__fsub_rn(__fmul_rn(R, R), __fmul_rn(dist, dist)));   // NaN if the ray misses
  const float K = __fdiv_rn(__fadd_rn(__fmul_rn(ncx, nfx), __fmul_rn(ncz, nfz)), nf_norm);
  const float mask = (Q < K) ? 1.0f : 0.0f;
  new_near = __fadd_rn(near, __fdiv_rn(__fmul_rn(mask, __fsub_rn(K, Q)), scale));
  new_far  = __fadd_rn(near, __fdiv_rn(__fadd_rn(K, Q), scale));
}

// sample_from_lineseg, core/utils/ray_utils.py:204-251 (perturb=0):
//   z = near (1 - t) + far t,   or with lindisp   z = 1 / (1/near (1 - t) + 1/far t)
__device__ __forceinline__ float pgn_coarse_z(float near, float far, float t, int lindisp = 0) {
  if (lindisp) return __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(__fdiv_rn(1.0f, near), __fsub_rn(1.0f, t)), __fmul_rn(__fdiv_rn(1.0f, far), t)));
  return __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, t)), __fmul_rn(far, t));
}

// stratified jitter of the coarse samples (training, perturb > 0; ray_utils.py:236-246):
//   mids = .5 (z[1:] + z[:-1]); upper = [mids, z[-1]]; lower = [z[0], mids]; z = lower + (upper - lower) t_rand
__device__ __forceinline__ float pgn_coarse_z_jitter(float near, float far, const float* __restrict__ t, int i, float t_rand, int lindisp = 0) {
  const float zi = pgn_coarse_z(near, far, t[i], lindisp);
  const float lower = i > 0 ? __fmul_rn(0.5f, __fadd_rn(zi, pgn_coarse_z(near, far, t[i - 1], lindisp))) : zi;
  const float upper = i + 1 < PGN_S ? __fmul_rn(0.5f, __fadd_rn(pgn_coarse_z(near, far, t[i + 1], lindisp), zi)) : zi;
  return __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t_rand));
}

// ---------------------------------------------------------------------------
// Joint-frame geometry of one (sample, joint):
//   pts = o + d*z                                   core/raycasters.py:658
//   pts_t = R_j pts + t_j                           core/encoders.py:8-23
//   v = ||pts_t||, r = pts_t / max(v, 1e-12)        core/encoders.py:110-122,181-193
//   w = 1 - sigmoid(tau * (v - cutoff_j))           core/cutoff_embedder.py:148-156
// skt points at the joint's 4x4 (row-major).
// ---------------------------------------------------------------------------
struct PgnJointGeom { float v, w, rx, ry, rz; };

template <bool kFast>
__device__ __forceinline__ float pgn_window(float v, float tau, float cutoff) {
  const float x = tau * (v - cutoff);
  if (kFast) return 1.0f - __fdividef(1.0f, 1.0f + __expf(-x));
  return 1.0f - 1.0f / (1.0f + expf(-x));
}

template <bool kFast>
__device__ __forceinline__ PgnJointGeom pgn_joint_geom(const float4 m0, const float4 m1, const float4 m2,
                                                       float px, float py, float pz, float tau, float cutoff) {
  PgnJointGeom g;
  const float x = fmaf(m0.z, pz, fmaf(m0.y, py, m0.x * px)) + m0.w;
  const float y = fmaf(m1.z, pz, fmaf(m1.y, py, m1.x * px)) + m1.w;
  const float z = fmaf(m2.z, pz, fmaf(m2.y, py, m2.x * px)) + m2.w;
  const float n2 = fmaf(z, z, fmaf(y, y, x * x));
  g.v = sqrtf(n2);
  if (kFast) {
    const float inv = __fdividef(1.0f, fmaxf(g.v, 1e-12f));
    g.rx = x * inv; g.ry = y * inv; g.rz = z * inv;
  } else {
    const float den = fmaxf(g.v, 1e-12f);
    g.rx = x / den; g.ry = y / den; g.rz = z / den;
  }
  g.w = pgn_window<kFast>(g.v, tau, cutoff);
  return g;
}

__device__ __forceinline__ void pgn_sample_point(const float* o, const float* d, float z, float& px, float& py, float& pz) {
  px = __fadd_rn(o[0], __fmul_rn(d[0], z));
  py = __fadd_rn(o[1], __fmul_rn(d[1], z));
  pz = __fadd_rn(o[2], __fmul_rn(d[2], z));
}

// normalised ray direction in a joint frame: F.normalize(R_j d), core/encoders.py:25-37,181-193
__device__ __forceinline__ void pgn_joint_dir(const float4 m0, const float4 m1, const float4 m2,
                                              const float* d, float& x, float& y, float& z) {
  x = fmaf(m0.z, d[2], fmaf(m0.y, d[1], m0.x * d[0]));
  y = fmaf(m1.z, d[2], fmaf(m1.y, d[1], m1.x * d[0]));
  z = fmaf(m2.z, d[2], fmaf(m2.y, d[1], m2.x * d[0]));
  const float den = fmaxf(sqrtf(fmaf(z, z, fmaf(y, y, x * x))), 1e-12f);
  x = x / den; y = y / den; z = z / den;
}

// ---------------------------------------------------------------------------
// Reference channel order of the 1080-vector:
//   [0,360)    v-embed   c = k*24 + j,        k in {x, sin(2^0 x), cos(2^0 x), ... sin(2^6 x), cos(2^6 x)}, all * w_j
//   [360,432)  r         c = 360 + j*3 + a
//   [432,1080) d-embed   c = 432 + k*72 + j*3 + a, k in {x, sin(2^0 x), cos(2^0 x), ... cos(2^3 x)}, all * w_j
// pgn_pe_term(x, k): k-th row of the positional encoding of scalar x (accurate sinf/cosf;
// x * 2^f is exact in fp32, matching cutoff_embedder.py:139 inputs_freq = freq_bands * inputs).
// ---------------------------------------------------------------------------
__device__ __forceinline__ float pgn_pe_term(float x, int k) {
  if (k == 0) return x;
  const int f = (k - 1) >> 1;
  const float a = x * (float)(1 << f);
  return ((k - 1) & 1) ? cosf(a) : sinf(a);
}

// ---------------------------------------------------------------------------
// Warp-level compositing of one ray (raw2outputs, core/networks/nerf.py:150-205).
// The calling warp owns the ray; sample i is handled by lane i / CH (CH consecutive
// samples per lane) so the transmittance is a per-lane sequential product followed by
// a warp-level exclusive prefix product over the lane totals.
//   raw: smem/global [S][4] (rgb_raw, sigma_raw);  z: [S]
// Outputs through pointers (any may be null); weights_out/alpha_out indexed [S].
// ---------------------------------------------------------------------------
template <int S>
__device__ __forceinline__ void pgn_composite_warp(const float* __restrict__ raw, const float* __restrict__ z,
                                                   float dnorm, float density_scale, float rgb_eps, int lane,
                                                   float* rgb3, float* disp, float* acc,
                                                   float* weights_out, float* alpha_out) {
  constexpr int CH = (S + 31) / 32;
  float a[CH], p[CH];
  float lane_prod = 1.0f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int i = lane * CH + c;
    float al = 0.0f;
    if (i < S) {
      float dist = (i + 1 < S) ? __fsub_rn(z[i + 1], z[i]) : 1e10f;
      dist = __fmul_rn(dist, dnorm);
      const float sig = fmaxf(raw[i * 4 + 3] / density_scale, 0.0f);
      al = 1.0f - expf(-__fmul_rn(sig, dist));
    }
    a[c] = al;
    p[c] = lane_prod;                                   // product of earlier samples in this lane
    if (i < S) lane_prod = lane_prod * (__fadd_rn(__fsub_rn(1.0f, al), 1e-10f));
  }
  // exclusive prefix product of lane totals
  float incl = lane_prod;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float up = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl *= up;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 1.0f;
  float sr = 0.f, sg = 0.f, sb = 0.f, sw = 0.f, sd = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int i = lane * CH + c;
    if (i < S) {
      const float w = a[c] * (excl * p[c]);
      if (weights_out) weights_out[i] = w;
      if (alpha_out) alpha_out[i] = a[c];
      const float k = 1.0f + 2.0f * rgb_eps;
      const float r = (1.0f / (1.0f + expf(-raw[i * 4 + 0]))) * k - rgb_eps;
      const float g = (1.0f / (1.0f + expf(-raw[i * 4 + 1]))) * k - rgb_eps;
      const float b = (1.0f / (1.0f + expf(-raw[i * 4 + 2]))) * k - rgb_eps;
      sr = fmaf(w, r, sr); sg = fmaf(w, g, sg); sb = fmaf(w, b, sb);
      sw += w; sd = fmaf(w, z[i], sd);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    sr += __shfl_xor_sync(0xffffffffu, sr, off);
    sg += __shfl_xor_sync(0xffffffffu, sg, off);
    sb += __shfl_xor_sync(0xffffffffu, sb, off);
    sw += __shfl_xor_sync(0xffffffffu, sw, off);
    sd += __shfl_xor_sync(0xffffffffu, sd, off);
  }
  if (lane == 0) {
    if (rgb3) { rgb3[0] = sr; rgb3[1] = sg; rgb3[2] = sb; }
    if (disp) {
      // 1/max(1e-10, depth/(acc+1e-10)), zeroed where isclose(acc, 0) (atol 1e-8)
      float dv = 1.0f / fmaxf(1e-10f, sd / (sw + 1e-10f));
      if (fabsf(sw) <= 1e-8f) dv = 0.0f;
      *disp = dv;
    }
    if (acc) *acc = fminf(sw, 1.0f);
  }
}

// ---------------------------------------------------------------------------
// Incremental form of the same compositing (raw2outputs) for samples [s0, s1) of one ray whose
// earlier samples were composited before: `carry` = {T before s0, sum w*r, sum w*g, sum w*b,
// sum w, sum w*z}.  Used by the tensor-core kernel, whose 128-row tiles cut fine rays (80
// samples) at arbitrary positions, so no per-ray raw buffer has to be kept.
//   raw_seg: rows of samples s0..s1-1 ([i - s0][4]);  z: the ray's full z array [S].
// All lanes of the warp must call it; carry is updated by lane 0 (then __syncwarp()).
// ---------------------------------------------------------------------------
// kFast (bf16 tensor-core tier only): ex2/rcp approximations instead of expf and IEEE division.
template <int S, bool kFast = false, int CH = 3>       // CH samples per lane: up to 32*CH samples per call
__device__ __forceinline__ void pgn_composite_segment_warp(const float* __restrict__ raw_seg, const float* __restrict__ z,
                                                           int s0, int s1, float dnorm, float density_scale, float rgb_eps,
                                                           int lane, float* carry, float* weights_out, float* alpha_out,
                                                           const float* __restrict__ noise = nullptr) {
  // noise (training, nerf.py:176-186): per-sample additive term on raw_sigma / B before the ReLU, indexed by sample
  float a[CH], p[CH];
  float lane_prod = 1.0f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int i = s0 + lane * CH + c;
    float al = 0.0f;
    if (i < s1) {
      float dist = (i + 1 < S) ? __fsub_rn(z[i + 1], z[i]) : 1e10f;
      dist = __fmul_rn(dist, dnorm);
      const float sig = fmaxf((kFast ? __fdividef(raw_seg[(i - s0) * 4 + 3], density_scale) : raw_seg[(i - s0) * 4 + 3] / density_scale) +
                              (noise ? noise[i] : 0.0f), 0.0f);
      al = 1.0f - (kFast ? __expf(-__fmul_rn(sig, dist)) : expf(-__fmul_rn(sig, dist)));
    }
    a[c] = al;
    p[c] = lane_prod;
    if (i < s1) lane_prod = lane_prod * (__fadd_rn(__fsub_rn(1.0f, al), 1e-10f));
  }
  float incl = lane_prod;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float up = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl *= up;
  }
  float excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 1.0f;
  const float t_in = carry[0];
  const float seg_prod = __shfl_sync(0xffffffffu, incl, 31);
  float sr = 0.f, sg = 0.f, sb = 0.f, sw = 0.f, sd = 0.f;
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const int i = s0 + lane * CH + c;
    if (i < s1) {
      const float w = a[c] * (t_in * (excl * p[c]));
      if (weights_out) weights_out[i] = w;
      if (alpha_out) alpha_out[i] = a[c];
      const float k = 1.0f + 2.0f * rgb_eps;
      const float* rw = raw_seg + (i - s0) * 4;
      const float r = (kFast ? __fdividef(1.0f, 1.0f + __expf(-rw[0])) : 1.0f / (1.0f + expf(-rw[0]))) * k - rgb_eps;
      const float g = (kFast ? __fdividef(1.0f, 1.0f + __expf(-rw[1])) : 1.0f / (1.0f + expf(-rw[1]))) * k - rgb_eps;
      const float b = (kFast ? __fdividef(1.0f, 1.0f + __expf(-rw[2])) : 1.0f / (1.0f + expf(-rw[2]))) * k - rgb_eps;
      sr = fmaf(w, r, sr); sg = fmaf(w, g, sg); sb = fmaf(w, b, sb);
      sw += w; sd = fmaf(w, z[i], sd);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    sr += __shfl_xor_sync(0xffffffffu, sr, off);
    sg += __shfl_xor_sync(0xffffffffu, sg, off);
    sb += __shfl_xor_sync(0xffffffffu, sb, off);
    sw += __shfl_xor_sync(0xffffffffu, sw, off);
    sd += __shfl_xor_sync(0xffffffffu, sd, off);
  }
  __syncwarp();
  if (lane == 0) {
    carry[0] = t_in * seg_prod;
    carry[1] += sr; carry[2] += sg; carry[3] += sb; carry[4] += sw; carry[5] += sd;
  }
  __syncwarp();
}

// outputs of a finished ray from its carry (nerf.py:188-203)
__device__ __forceinline__ void pgn_composite_finalize(const float* carry, float* rgb3, float* disp, float* acc) {
  const float sw = carry[4], sd = carry[5];
  rgb3[0] = carry[1]; rgb3[1] = carry[2]; rgb3[2] = carry[3];
  float dv = 1.0f / fmaxf(1e-10f, sd / (sw + 1e-10f));
  if (fabsf(sw) <= 1e-8f) dv = 0.0f;
  *disp = dv;
  *acc = fminf(sw, 1.0f);
}

// ---------------------------------------------------------------------------
// Warp-level inverse-CDF resampling of one ray (det=True):
//   sample_pdf            core/utils/ray_utils.py:157-201
//   isample_from_lineseg  core/utils/ray_utils.py:255-289
// z[64], weights[64] -> z_samples[16], z_sorted[80] (+ optional pdf_inds[16], sorted_idxs[80]).
// scratch: 2*64 floats of shared memory private to the warp (cdf | bins).
// Summation order (documented in DESIGN.md): sum and cumsum are accumulated in fp64 and
// rounded once to fp32 (torch CPU cumsum accumulates float in double).
// The search is a warp ballot: ind = #{k : cdf[k] <= u}.
// ---------------------------------------------------------------------------
// part 1: pdf -> cdf[63] | bins[63] in scratch (2 x 64 floats)
__device__ __forceinline__ void pgn_sample_pdf_cdf_warp(const float* __restrict__ z, const float* __restrict__ weights,
                                                        int lane, float* scratch) {
  float* cdf = scratch;        // [63] (+1 pad)
  float* bins = scratch + 64;  // [63]
  // pdf over the 62 interior weights; lane handles k = lane and lane+32
  float w0 = (lane < 62) ? __fadd_rn(weights[1 + lane], 1e-5f) : 0.0f;
  float w1 = (lane + 32 < 62) ? __fadd_rn(weights[1 + lane + 32], 1e-5f) : 0.0f;
  double s = (double)w0 + (double)w1;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float total = (float)s;
  const float p0 = (lane < 62) ? __fdiv_rn(w0, total) : 0.0f;
  const float p1 = (lane + 32 < 62) ? __fdiv_rn(w1, total) : 0.0f;
  // inclusive scan in double over index order: first all p0 (k=0..31), then p1 (k=32..61)
  double c0 = (double)p0;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double up = __shfl_up_sync(0xffffffffu, c0, off);
    if (lane >= off) c0 += up;
  }
  const double first_half = __shfl_sync(0xffffffffu, c0, 31);
  double c1 = (double)p1;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double up = __shfl_up_sync(0xffffffffu, c1, off);
    if (lane >= off) c1 += up;
  }
  c1 += first_half;
  if (lane == 0) cdf[0] = 0.0f;
  cdf[1 + lane] = (float)c0;                         // k = lane      -> cdf index lane+1 (1..32)
  if (lane + 32 < 62) cdf[33 + lane] = (float)c1;    // k = lane+32   -> cdf index 33..62
  // bins = midpoints of z (63)
  bins[lane] = __fmul_rn(0.5f, __fadd_rn(z[lane + 1], z[lane]));
  if (lane + 32 < 63) bins[lane + 32] = __fmul_rn(0.5f, __fadd_rn(z[lane + 33], z[lane + 32]));
  __syncwarp();
}

// part 2: inverse-CDF draw at u_det, merge with the coarse z (scratch as left by part 1)
__device__ __forceinline__ void pgn_sample_pdf_draw_warp(const float* __restrict__ z, const float* __restrict__ u_det, int lane,
                                                         float* scratch, float* z_samples, float* z_sorted,
                                                         int* pdf_inds, int* sorted_idxs) {
  float* cdf = scratch;
  float* bins = scratch + 64;
  const float ca = cdf[lane];
  const float cb = (lane + 32 < 63) ? cdf[lane + 32] : INFINITY;
  float my_sample = 0.0f;
#pragma unroll
  for (int k = 0; k < PGN_I; ++k) {
    const float u = u_det[k];
    const int ind = __popc(__ballot_sync(0xffffffffu, ca <= u)) + __popc(__ballot_sync(0xffffffffu, cb <= u));
    if (lane == k) {
      const int below = max(ind - 1, 0), above = min(ind, 62);
      const float cdb = cdf[below], cda = cdf[above];
      float denom = __fsub_rn(cda, cdb);
      if (denom < 1e-5f) denom = 1.0f;
      const float t = __fdiv_rn(__fsub_rn(u, cdb), denom);
      my_sample = __fadd_rn(bins[below], __fmul_rn(t, __fsub_rn(bins[above], bins[below])));
      if (pdf_inds) pdf_inds[k] = ind;
    }
  }
  if (lane < PGN_I && z_samples) z_samples[lane] = my_sample;
  // merge: rank of coarse z_i = i + #{samples < z_i}; rank of sample k = k + #{z_i <= sample_k}
  __syncwarp();
  float* samp = scratch;        // reuse (cdf no longer needed): samples in scratch[0..15]
  __syncwarp();
  if (lane < PGN_I) samp[lane] = my_sample;
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = lane + 32 * h;
    const float zi = z[i];
    int r = i;
#pragma unroll
    for (int k = 0; k < PGN_I; ++k) r += (samp[k] < zi) ? 1 : 0;
    z_sorted[r] = zi;
    if (sorted_idxs) sorted_idxs[r] = i;
  }
  if (lane < PGN_I) {
    int r = lane;
    for (int i = 0; i < PGN_S; ++i) r += (z[i] <= my_sample) ? 1 : 0;
    z_sorted[r] = my_sample;
    if (sorted_idxs) sorted_idxs[r] = PGN_S + lane;
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------
// bf16-tier variants of the two parts (same algorithm, cheaper arithmetic): the pdf sum / cumsum run in
// fp32, and the 16 draws are independent per-lane binary searches instead of 16 warp-wide ballots; the
// merge ranks use binary searches over the (sorted) coarse z and the (non-decreasing) samples, which give
// the same counts as the exhaustive comparisons above.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pgn_sample_pdf_cdf_warp_fast(const float* __restrict__ z, const float* __restrict__ weights,
                                                             int lane, float* scratch) {
  float* cdf = scratch;        // [63] (+1 pad)
  float* bins = scratch + 64;  // [63]
  const float w0 = (lane < 62) ? weights[1 + lane] + 1e-5f : 0.0f;
  const float w1 = (lane + 32 < 62) ? weights[1 + lane + 32] + 1e-5f : 0.0f;
  float s = w0 + w1;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float inv = __fdividef(1.0f, s);
  float c0 = w0 * inv, c1 = w1 * inv;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const float u0 = __shfl_up_sync(0xffffffffu, c0, off), u1 = __shfl_up_sync(0xffffffffu, c1, off);
    if (lane >= off) { c0 += u0; c1 += u1; }
  }
  c1 += __shfl_sync(0xffffffffu, c0, 31);
  if (lane == 0) cdf[0] = 0.0f;
  cdf[1 + lane] = c0;
  if (lane + 32 < 62) cdf[33 + lane] = c1;
  bins[lane] = 0.5f * (z[lane + 1] + z[lane]);
  if (lane + 32 < 63) bins[lane + 32] = 0.5f * (z[lane + 33] + z[lane + 32]);
  __syncwarp();
}

__device__ __forceinline__ void pgn_sample_pdf_draw_warp_fast(const float* __restrict__ z, const float* __restrict__ u_det, int lane,
                                                              float* scratch, float* z_samples, float* z_sorted, int* pdf_inds) {
  float* cdf = scratch;
  float* bins = scratch + 64;
  float my_sample = 0.0f;
  if (lane < PGN_I) {
    const float u = u_det[lane];
    int lo = 0, hi = 63;                                // ind = #{k in [0,63) : cdf[k] <= u}  (cdf is non-decreasing)
#pragma unroll
    for (int it = 0; it < 6; ++it) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    const int ind = lo;
    const int below = max(ind - 1, 0), above = min(ind, 62);
    const float cdb = cdf[below], cda = cdf[above];
    float denom = cda - cdb;
    if (denom < 1e-5f) denom = 1.0f;
    const float t = __fdividef(u - cdb, denom);
    my_sample = fmaf(t, bins[above] - bins[below], bins[below]);
    if (pdf_inds) pdf_inds[lane] = ind;
    if (z_samples) z_samples[lane] = my_sample;
  }
  __syncwarp();
  float* samp = scratch;        // reuse (cdf no longer needed): samples in scratch[0..15]
  if (lane < PGN_I) samp[lane] = my_sample;
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = lane + 32 * h;
    const float zi = z[i];
    int lo = 0, hi = PGN_I;                             // #{k : samp[k] < zi}
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const int mid = (lo + hi) >> 1;
      if (lo < hi) { if (samp[mid] < zi) lo = mid + 1; else hi = mid; }
    }
    z_sorted[i + lo] = zi;
  }
  if (lane < PGN_I) {
    int lo = 0, hi = PGN_S;                             // #{i : z[i] <= my_sample}
#pragma unroll
    for (int it = 0; it < 7; ++it) {
      const int mid = (lo + hi) >> 1;
      if (lo < hi) { if (z[mid] <= my_sample) lo = mid + 1; else hi = mid; }
    }
    z_sorted[lane + lo] = my_sample;
  }
  __syncwarp();
}

// training variant of the draw (det = False, ray_utils.py:169-170): 16 arbitrary, unsorted u per ray; the merge is
// a full rank computation over the 80 values (ties: coarse samples first, then importance samples in index order)
__device__ __forceinline__ void pgn_sample_pdf_draw_warp_rand(const float* __restrict__ z, const float* __restrict__ u_ray, int lane,
                                                              float* scratch, float* z_samples, float* z_sorted, int* pdf_inds) {
  float* cdf = scratch;
  float* bins = scratch + 64;
  float my_sample = 0.0f;
  if (lane < PGN_I) {
    const float u = u_ray[lane];
    int lo = 0, hi = 63;                                // ind = #{k in [0,63) : cdf[k] <= u}
#pragma unroll
    for (int it = 0; it < 6; ++it) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    const int ind = lo;
    const int below = max(ind - 1, 0), above = min(ind, 62);
    const float cdb = cdf[below], cda = cdf[above];
    float denom = cda - cdb;
    if (denom < 1e-5f) denom = 1.0f;
    const float t = __fdividef(u - cdb, denom);
    my_sample = fmaf(t, bins[above] - bins[below], bins[below]);
    if (pdf_inds) pdf_inds[lane] = ind;
    if (z_samples) z_samples[lane] = my_sample;
  }
  __syncwarp();
  float* samp = scratch;
  if (lane < PGN_I) samp[lane] = my_sample;
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = lane + 32 * h;
    const float zi = z[i];
    int r = i;
#pragma unroll
    for (int k = 0; k < PGN_I; ++k) r += (samp[k] < zi) ? 1 : 0;
    z_sorted[r] = zi;
  }
  if (lane < PGN_I) {
    int r = 0;
    for (int i = 0; i < PGN_S; ++i) r += (z[i] <= my_sample) ? 1 : 0;
#pragma unroll
    for (int k = 0; k < PGN_I; ++k) r += (samp[k] < my_sample || (samp[k] == my_sample && k < lane)) ? 1 : 0;
    z_sorted[r] = my_sample;
  }
  __syncwarp();
}

__device__ __forceinline__ void pgn_sample_pdf_warp(const float* __restrict__ z, const float* __restrict__ weights,
                                                    const float* __restrict__ u_det, int lane, float* scratch,
                                                    float* z_samples, float* z_sorted,
                                                    int* pdf_inds, int* sorted_idxs) {
  pgn_sample_pdf_cdf_warp(z, weights, lane, scratch);
  pgn_sample_pdf_draw_warp(z, u_det, lane, scratch, z_samples, z_sorted, pdf_inds, sorted_idxs);
}
