// Stage-level kernels: the cheap HBM-bound stages of the render path that run
// outside the fused MLP kernel (near/far pre-pass with the chunk-level NaN fill,
// ray generation, frame composition) and the debug/parity entry points that expose
// one reference function each (encode dump, compositing, inverse-CDF resampling).
#include "pgn_common.cuh"
#include "pgn_kernels.h"

// ---------------------------------------------------------------------------
// near/far pre-pass: get_near_far_in_cylinder (core/utils/ray_utils.py:292-344).
// One CTA per batchify chunk (core/trainer.py:64-81) because rays whose Q is NaN take
// the nan-mean near/far of *their chunk* (ray_utils.py:328-342).  The mean is
// accumulated in fp64 (numpy.nanmean is a float32 pairwise sum; difference <= 1 ulp).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
pgn_near_far_kernel(PgnRayRefs rays, long long chunk, float* __restrict__ near_far) {
  __shared__ double s_sum[2][32];
  __shared__ int s_cnt[2][32];
  __shared__ float s_mean[2];
  __shared__ int s_any_nan;
  const long long r0 = (long long)blockIdx.x * chunk;
  const long long r1 = min(r0 + chunk, rays.n_rays);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_any_nan = 0;
  __syncthreads();
  double sum_n = 0.0, sum_f = 0.0;
  int cnt_n = 0, cnt_f = 0;
  bool any_nan = false;
  for (long long i = r0 + threadIdx.x; i < r1; i += blockDim.x) {
    float nn, ff;
    pgn_near_far_ray(rays.ray_batch + i * 11, pgn_ray_cyl(rays, i), nn, ff);
    near_far[i * 2] = nn;
    near_far[i * 2 + 1] = ff;
    if (nn == nn) { sum_n += nn; ++cnt_n; } else any_nan = true;
    if (ff == ff) { sum_f += ff; ++cnt_f; }
  }
  if (any_nan) s_any_nan = 1;     // reference trigger: torch.isnan(new_near).any()
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    sum_n += __shfl_xor_sync(0xffffffffu, sum_n, off);
    sum_f += __shfl_xor_sync(0xffffffffu, sum_f, off);
    cnt_n += __shfl_xor_sync(0xffffffffu, cnt_n, off);
    cnt_f += __shfl_xor_sync(0xffffffffu, cnt_f, off);
  }
  if (lane == 0) { s_sum[0][warp] = sum_n; s_sum[1][warp] = sum_f; s_cnt[0][warp] = cnt_n; s_cnt[1][warp] = cnt_f; }
  __syncthreads();
  if (!s_any_nan) return;
  if (threadIdx.x < 2) {
    double s = 0.0; int c = 0;
    const int nw = (blockDim.x + 31) >> 5;
    for (int w = 0; w < nw; ++w) { s += s_sum[threadIdx.x][w]; c += s_cnt[threadIdx.x][w]; }
    s_mean[threadIdx.x] = c > 0 ? (float)(s / (double)c) : NAN;
  }
  __syncthreads();
  const float mean_n = s_mean[0], mean_f = s_mean[1];
  for (long long i = r0 + threadIdx.x; i < r1; i += blockDim.x) {
    // rows where Q is NaN <=> new_far is NaN (new_far = near + (K+Q)/scale)
    const float ff = near_far[i * 2 + 1];
    const float nn = near_far[i * 2];
    const bool q_nan = !(ff == ff) || !(nn == nn);
    if (q_nan) {
      near_far[i * 2]     = (mean_n == mean_n) ? mean_n : rays.ray_batch[i * 11 + 6];
      near_far[i * 2 + 1] = (mean_f == mean_f) ? mean_f : rays.ray_batch[i * 11 + 7];
    }
  }
}

cudaError_t pgn_launch_near_far(const PgnRayRefs& rays, long long chunk, float* near_far, cudaStream_t stream) {
  if (rays.n_rays == 0) return cudaSuccess;
  if (chunk <= 0 || chunk > rays.n_rays) chunk = rays.n_rays;
  const long long n_chunks = (rays.n_rays + chunk - 1) / chunk;
  pgn_near_far_kernel<<<(unsigned)n_chunks, 1024, 0, stream>>>(rays, chunk, near_far);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// encode dump: encode_inputs (core/raycasters.py:476-555) for explicit z.
// One thread per (ray, sample, joint); writes the joint's 45 channels.
// ---------------------------------------------------------------------------
__global__ void pgn_encode_kernel(PgnRayRefs rays, const PgnScalars* __restrict__ scp, const float* __restrict__ z,
                                  int n_z, float* __restrict__ enc) {
  const PgnScalars& sc = *scp;
  const long long total = rays.n_rays * n_z * PGN_J;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % PGN_J);
    const long long rs = idx / PGN_J;
    const long long ray = rs / n_z;
    const float* rb = rays.ray_batch + ray * 11;
    const float4* m = reinterpret_cast<const float4*>(pgn_ray_skts(rays, ray) + j * 16);
    const float4 m0 = __ldg(m), m1 = __ldg(m + 1), m2 = __ldg(m + 2);
    float px, py, pz;
    pgn_sample_point(rb, rb + 3, z[rs], px, py, pz);
    const PgnJointGeom g = pgn_joint_geom<false>(m0, m1, m2, px, py, pz, sc.tau_v, sc.cutoff_v[j]);
    const float wd = pgn_window<false>(g.v, sc.tau_d, sc.cutoff_d[j]);
    float* e = enc + rs * PGN_ENC;
    for (int k = 0; k < 1 + 2 * PGN_LV; ++k) e[k * PGN_J + j] = g.w * pgn_pe_term(g.v, k);
    e[360 + j * 3 + 0] = g.rx; e[360 + j * 3 + 1] = g.ry; e[360 + j * 3 + 2] = g.rz;
    float dj[3];
    pgn_joint_dir(m0, m1, m2, rb + 3, dj[0], dj[1], dj[2]);
    for (int k = 0; k < 1 + 2 * PGN_LD; ++k)
      for (int a = 0; a < 3; ++a) e[PGN_ENC_P + k * 72 + j * 3 + a] = wd * pgn_pe_term(dj[a], k);
  }
}

cudaError_t pgn_launch_encode(const PgnRayRefs& rays, const PgnScalars* sc_dev, const float* z, int n_z,
                              float* enc, cudaStream_t stream) {
  const long long total = rays.n_rays * n_z * PGN_J;
  if (total == 0) return cudaSuccess;
  const int block = 128;
  const long long grid = min((total + block - 1) / block, (long long)148 * 16);
  pgn_encode_kernel<<<(unsigned)grid, block, 0, stream>>>(rays, sc_dev, z, n_z, enc);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// compositing: NeRF.raw2outputs (core/networks/nerf.py:150-205), one warp per ray
// ---------------------------------------------------------------------------
template <int S>
__global__ void pgn_composite_kernel(PgnRayRefs rays, const PgnScalars* __restrict__ scp, const float* __restrict__ raw,
                                     const float* __restrict__ z, float* rgb, float* disp, float* acc,
                                     float* weights, float* alpha) {
  const PgnScalars& sc = *scp;
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rays.n_rays; r += nwarps) {
    const float* d = rays.ray_batch + r * 11 + 3;
    const float dn = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    float rgb3[3], dv, av;
    pgn_composite_warp<S>(raw + r * S * 4, z + r * S, dn, sc.density_scale, sc.rgb_eps, lane, rgb3, &dv, &av,
                          weights ? weights + r * S : nullptr, alpha ? alpha + r * S : nullptr);
    if (lane == 0) {
      if (rgb) { rgb[r * 3] = rgb3[0]; rgb[r * 3 + 1] = rgb3[1]; rgb[r * 3 + 2] = rgb3[2]; }
      if (disp) disp[r] = dv;
      if (acc) acc[r] = av;
    }
  }
}

cudaError_t pgn_launch_composite(const PgnRayRefs& rays, const PgnScalars* sc_dev, const float* raw, const float* z,
                                 int s, float* rgb, float* disp, float* acc, float* weights, float* alpha,
                                 cudaStream_t stream) {
  if (rays.n_rays == 0) return cudaSuccess;
  const int block = 256;
  const long long grid = min((rays.n_rays * 32 + block - 1) / block, (long long)148 * 8);
  if (s == PGN_S) pgn_composite_kernel<PGN_S><<<(unsigned)grid, block, 0, stream>>>(rays, sc_dev, raw, z, rgb, disp, acc, weights, alpha);
  else if (s == PGN_T) pgn_composite_kernel<PGN_T><<<(unsigned)grid, block, 0, stream>>>(rays, sc_dev, raw, z, rgb, disp, acc, weights, alpha);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// inverse-CDF resampling + merge, one warp per ray
// ---------------------------------------------------------------------------
__global__ void pgn_sample_pdf_kernel(const PgnScalars* __restrict__ scp, const float* __restrict__ z,
                                      const float* __restrict__ weights, long long n, float* z_samples,
                                      float* z_sorted, int* pdf_inds, int* sorted_idxs) {
  __shared__ float scratch[8][128];
  __shared__ float zs[8][PGN_S], ws[8][PGN_S], zo[8][PGN_T];
  const PgnScalars& sc = *scp;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < n; r += nwarps) {
    zs[wib][lane] = z[r * PGN_S + lane]; zs[wib][lane + 32] = z[r * PGN_S + lane + 32];
    ws[wib][lane] = weights[r * PGN_S + lane]; ws[wib][lane + 32] = weights[r * PGN_S + lane + 32];
    __syncwarp();
    pgn_sample_pdf_warp(zs[wib], ws[wib], sc.u_det, lane, scratch[wib],
                        z_samples ? z_samples + r * PGN_I : nullptr, zo[wib],
                        pdf_inds ? pdf_inds + r * PGN_I : nullptr,
                        sorted_idxs ? sorted_idxs + r * PGN_T : nullptr);
    if (z_sorted) for (int i = lane; i < PGN_T; i += 32) z_sorted[r * PGN_T + i] = zo[wib][i];
    __syncwarp();
  }
}

cudaError_t pgn_launch_sample_pdf(const PgnScalars* sc_dev, const float* z, const float* weights, long long n,
                                  float* z_samples, float* z_sorted, int* pdf_inds, int* sorted_idxs,
                                  cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  const int block = 256;
  const long long grid = min((n * 32 + block - 1) / block, (long long)148 * 8);
  pgn_sample_pdf_kernel<<<(unsigned)grid, block, 0, stream>>>(sc_dev, z, weights, n, z_samples, z_sorted, pdf_inds, sorted_idxs);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// weight transpose [out][in] -> [in][out] (fp32 engine layout)
// ---------------------------------------------------------------------------
__global__ void pgn_transpose_kernel(const float* __restrict__ w, int out_f, int in_f, float* __restrict__ wt) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;   // bx over in_f, by over out_f
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int o = by + i, k = bx + threadIdx.x;
    tile[i][threadIdx.x] = (o < out_f && k < in_f) ? w[(size_t)o * in_f + k] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = bx + i, o = by + threadIdx.x;
    if (k < in_f && o < out_f) wt[(size_t)k * out_f + o] = tile[threadIdx.x][i];
  }
}

cudaError_t pgn_launch_transpose(const float* w, int out_f, int in_f, float* wt, cudaStream_t stream) {
  dim3 grid((in_f + 31) / 32, (out_f + 31) / 32), block(32, 8);
  pgn_transpose_kernel<<<grid, block, 0, stream>>>(w, out_f, in_f, wt);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// on-device pixel rays of a bbox: get_rays (core/utils/ray_utils.py:6-28) restricted to
// rows [y0,y1) x cols [x0,x1) (kp_to_valid_rays :124-130) + the [n,11] batch layout of
// core/trainer.py:118-137.  c2w12: device [3][4].
// ---------------------------------------------------------------------------
__global__ void pgn_generate_rays_kernel(int H, int W, float focal, const float* __restrict__ c2w, int x0, int y0,
                                         int x1, int y1, float* __restrict__ rb) {
  const int bw = x1 - x0;
  const long long n = (long long)bw * (y1 - y0);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    const int px = x0 + (int)(idx % bw), py = y0 + (int)(idx / bw);
    const float dx = __fdiv_rn(__fsub_rn((float)px, W * 0.5f), focal);
    const float dy = -__fdiv_rn(__fsub_rn((float)py, H * 0.5f), focal);
    const float dz = -1.0f;
    float d[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)   // torch.sum(dirs[..., None, :] * c2w[:3,:3], -1): ((dx*c0 + dy*c1) + dz*c2)
      d[r] = __fadd_rn(__fadd_rn(__fmul_rn(dx, c2w[r * 4 + 0]), __fmul_rn(dy, c2w[r * 4 + 1])), __fmul_rn(dz, c2w[r * 4 + 2]));
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
    float* o = rb + idx * 11;
    o[0] = c2w[3]; o[1] = c2w[7]; o[2] = c2w[11];
    o[3] = d[0]; o[4] = d[1]; o[5] = d[2];
    o[6] = 0.f; o[7] = 1.f;
    o[8] = __fdiv_rn(d[0], nrm); o[9] = __fdiv_rn(d[1], nrm); o[10] = __fdiv_rn(d[2], nrm);
  }
}

cudaError_t pgn_launch_generate_rays(int H, int W, float focal, const float* c2w12_dev, int x0, int y0, int x1, int y1,
                                     float* ray_batch, cudaStream_t stream) {
  const long long n = (long long)(x1 - x0) * (y1 - y0);
  if (n <= 0) return cudaSuccess;
  const int block = 256;
  const long long grid = min((n + block - 1) / block, (long long)148 * 8);
  pgn_generate_rays_kernel<<<(unsigned)grid, block, 0, stream>>>(H, W, focal, c2w12_dev, x0, y0, x1, y1, ray_batch);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// frame composition (run_nerf.py:100-133): image = bg everywhere; inside the bbox
// image[valid] = rgb + (1 - acc) * bg
// ---------------------------------------------------------------------------
__global__ void pgn_compose_frame_kernel(int H, int W, int x0, int y0, int x1, int y1, const float* __restrict__ rgb,
                                         const float* __restrict__ acc, float bg, float* __restrict__ image) {
  const long long n = (long long)H * W;
  const int bw = x1 - x0;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
    const int px = (int)(p % W), py = (int)(p / W);
    float r = bg, g = bg, b = bg;
    if (px >= x0 && px < x1 && py >= y0 && py < y1) {
      const long long i = (long long)(py - y0) * bw + (px - x0);
      const float back = __fmul_rn(__fsub_rn(1.0f, acc[i]), bg);
      r = __fadd_rn(rgb[i * 3], back); g = __fadd_rn(rgb[i * 3 + 1], back); b = __fadd_rn(rgb[i * 3 + 2], back);
    }
    image[p * 3] = r; image[p * 3 + 1] = g; image[p * 3 + 2] = b;
  }
}

cudaError_t pgn_launch_compose_frame(int H, int W, int x0, int y0, int x1, int y1, const float* rgb, const float* acc,
                                     float bg, float* image, cudaStream_t stream) {
  const long long n = (long long)H * W;
  if (n <= 0) return cudaSuccess;
  const int block = 256;
  const long long grid = min((n + block - 1) / block, (long long)148 * 8);
  pgn_compose_frame_kernel<<<(unsigned)grid, block, 0, stream>>>(H, W, x0, y0, x1, y1, rgb, acc, bg, image);
  return cudaGetLastError();
}
