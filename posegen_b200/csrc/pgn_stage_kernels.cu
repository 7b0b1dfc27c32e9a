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
pgn_near_far_kernel(PgnRayRefs rays, long long chunk, const long long* __restrict__ chunk_starts, float* __restrict__ near_far) {
  __shared__ double s_sum[2][32];
  __shared__ int s_cnt[2][32];
  __shared__ float s_mean[2];
  __shared__ int s_any_nan;
  // chunk_starts (optional): explicit chunk table [n_chunks + 1] for multi-image batches, so that a chunk never
  // straddles two images (the reference chunks every image separately, run_nerf.py:77-95 -> core/trainer.py:64-81)
  const long long r0 = chunk_starts ? chunk_starts[blockIdx.x] : (long long)blockIdx.x * chunk;
  const long long r1 = chunk_starts ? min(chunk_starts[blockIdx.x + 1], rays.n_rays) : min(r0 + chunk, rays.n_rays);
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
  pgn_near_far_kernel<<<(unsigned)n_chunks, 1024, 0, stream>>>(rays, chunk, nullptr, near_far);
  return cudaGetLastError();
}

cudaError_t pgn_launch_near_far_chunks(const PgnRayRefs& rays, const long long* chunk_starts, long long n_chunks, float* near_far,
                                       cudaStream_t stream) {
  if (rays.n_rays == 0 || n_chunks <= 0) return cudaSuccess;
  pgn_near_far_kernel<<<(unsigned)n_chunks, 1024, 0, stream>>>(rays, 0, chunk_starts, near_far);
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

// ---------------------------------------------------------------------------
// forward kinematics on the device.
//   axis-angle -> rotation (scipy Rotation.from_rotvec = Rodrigues)      core/utils/skeleton_utils.py:352
//   kinematic chain l2w[i] = l2w[parent] @ [R_i | rest_i - rest_parent]  core/utils/skeleton_utils.py:334-377
//   skts = inverse(l2ws) (closed-form rigid inverse), kps = l2ws[:, :3, 3]      run_gan.py:447-449
//   cylinder (cx, cz, R, top, bot), head '-y'                             core/utils/skeleton_utils.py:635-685
// One thread per pose; the chain is evaluated in fp64 like the numpy reference and rounded once.
// ---------------------------------------------------------------------------
__constant__ int c_smpl_parents[PGN_J] = {0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21};

__global__ void pgn_pose_fk_kernel(const float* __restrict__ bones, const float* __restrict__ rest, int n_poses,
                                   float ext, float top_ratio, float bot_ratio,
                                   float* __restrict__ skts, float* __restrict__ kps, float* __restrict__ cyls,
                                   float* __restrict__ l2ws_out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_poses) return;
  double R[PGN_J][9], T[PGN_J][3];
  for (int i = 0; i < PGN_J; ++i) {
    const double rx = bones[(p * PGN_J + i) * 3], ry = bones[(p * PGN_J + i) * 3 + 1], rz = bones[(p * PGN_J + i) * 3 + 2];
    const double th = sqrt(rx * rx + ry * ry + rz * rz);
    double r[9];
    if (th < 1e-12) {
      r[0] = 1; r[1] = 0; r[2] = 0; r[3] = 0; r[4] = 1; r[5] = 0; r[6] = 0; r[7] = 0; r[8] = 1;
    } else {
      const double kx = rx / th, ky = ry / th, kz = rz / th, s = sin(th), c = cos(th), v = 1.0 - c;
      r[0] = c + kx * kx * v;      r[1] = kx * ky * v - kz * s; r[2] = kx * kz * v + ky * s;
      r[3] = ky * kx * v + kz * s; r[4] = c + ky * ky * v;      r[5] = ky * kz * v - kx * s;
      r[6] = kz * kx * v - ky * s; r[7] = kz * ky * v + kx * s; r[8] = c + kz * kz * v;
    }
    if (i == 0) {
      for (int k = 0; k < 9; ++k) R[0][k] = r[k];
      for (int k = 0; k < 3; ++k) T[0][k] = rest[k];
    } else {
      const int pa = c_smpl_parents[i];
      const double d[3] = {(double)rest[i * 3] - (double)rest[pa * 3], (double)rest[i * 3 + 1] - (double)rest[pa * 3 + 1],
                           (double)rest[i * 3 + 2] - (double)rest[pa * 3 + 2]};
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b)
          R[i][a * 3 + b] = R[pa][a * 3] * r[b] + R[pa][a * 3 + 1] * r[3 + b] + R[pa][a * 3 + 2] * r[6 + b];
        T[i][a] = R[pa][a * 3] * d[0] + R[pa][a * 3 + 1] * d[1] + R[pa][a * 3 + 2] * d[2] + T[pa][a];
      }
    }
  }
  double rad = 0.0, hmax = -1e300, hmin = 1e300;
  for (int i = 0; i < PGN_J; ++i) {
    float* s = skts + ((size_t)p * PGN_J + i) * 16;
    for (int a = 0; a < 3; ++a) {           // [R^T | -R^T t]
      for (int b = 0; b < 3; ++b) s[a * 4 + b] = (float)R[i][b * 3 + a];
      s[a * 4 + 3] = (float)(-(R[i][a] * T[i][0] + R[i][3 + a] * T[i][1] + R[i][6 + a] * T[i][2]));
    }
    s[12] = 0.f; s[13] = 0.f; s[14] = 0.f; s[15] = 1.f;
    if (l2ws_out) {
      float* l = l2ws_out + ((size_t)p * PGN_J + i) * 16;
      for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) l[a * 4 + b] = (float)R[i][a * 3 + b]; l[a * 4 + 3] = (float)T[i][a]; }
      l[12] = 0.f; l[13] = 0.f; l[14] = 0.f; l[15] = 1.f;
    }
    if (kps) { kps[(p * PGN_J + i) * 3] = (float)T[i][0]; kps[(p * PGN_J + i) * 3 + 1] = (float)T[i][1]; kps[(p * PGN_J + i) * 3 + 2] = (float)T[i][2]; }
    // the cylinder is computed from the fp32 key points like the reference (skeleton_utils.py:635-685 on kps arrays)
    const float kx = (float)T[i][0] - (float)T[0][0], kz = (float)T[i][2] - (float)T[0][2];
    rad = fmax(rad, (double)sqrtf(kx * kx + kz * kz));
    const double h = -(double)(float)T[i][1];
    hmax = fmax(hmax, h); hmin = fmin(hmin, h);
  }
  if (cyls) {
    float* c = cyls + (size_t)p * 5;
    c[0] = (float)T[0][0]; c[1] = (float)T[0][2]; c[2] = (float)rad + ext;
    c[3] = -((float)hmax + ext * top_ratio); c[4] = -((float)hmin - ext * bot_ratio);
  }
}

cudaError_t pgn_launch_pose_fk(const float* bones, const float* rest, int n_poses, float ext, float top_ratio, float bot_ratio,
                               float* skts, float* kps, float* cyls, float* l2ws, cudaStream_t stream) {
  if (n_poses <= 0) return cudaSuccess;
  pgn_pose_fk_kernel<<<(n_poses + 63) / 64, 64, 0, stream>>>(bones, rest, n_poses, ext, top_ratio, bot_ratio, skts, kps, cyls, l2ws);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// rendered frame -> HMR input on the device (run_gan.py:2057-2071,2433-2445):
//   uint8 quantisation of the saved PNG ((rgb*255).astype(uint8), run_gan.py:2326) -> crop [y0:y1, x0:x1] -> /255 ->
//   Normalize(mean, std) -> skimage.transform.resize(img, (3, R, R), anti_aliasing=True) -> [3,R,R] fp32.
// skimage's resize (>= 0.19) = scipy.ndimage.gaussian_filter(sigma = (scale-1)/2 per axis, truncate 4, mode 'mirror')
// followed by scipy.ndimage.zoom(order=1, mode='mirror', grid_mode=True): x_in = (x_out + 0.5) * scale - 0.5.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int pgn_mirror(int i, int n) {        // scipy 'mirror': d c b | a b c d | c b a
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  i = i % period;
  if (i < 0) i += period;
  return i < n ? i : period - i;
}

__global__ void pgn_hmr_input_kernel(const float* __restrict__ image, int H, int W, int x0, int y0, int cw, int ch, int R,
                                     float m0, float m1, float m2, float s0, float s1, float s2, int quantize,
                                     float* __restrict__ out) {
  const int n = 3 * R * R;
  const double sy = (double)ch / R, sx = (double)cw / R;
  const double sig_y = fmax(0.0, (sy - 1.0) * 0.5), sig_x = fmax(0.0, (sx - 1.0) * 0.5);
  const int ry = (int)(4.0 * sig_y + 0.5), rx = (int)(4.0 * sig_x + 0.5);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += gridDim.x * blockDim.x) {
    const int c = idx / (R * R), oy = (idx / R) % R, ox = idx % R;
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), stdv = c == 0 ? s0 : (c == 1 ? s1 : s2);
    auto pix = [&](int yy, int xx) -> double {            // normalised crop pixel with mirror extension
      const int y = y0 + pgn_mirror(yy, ch), x = x0 + pgn_mirror(xx, cw);
      float v = image[((size_t)y * W + x) * 3 + c];
      if (quantize) v = floorf(fminf(fmaxf(v * 255.0f, 0.0f), 255.0f));      // astype(uint8) truncates
      else v = v * 255.0f;
      return (double)((v / 255.0f - mean) / stdv);
    };
    auto blurred = [&](int yy, int xx) -> double {        // separable Gaussian, evaluated at one (mirrored) location
      yy = pgn_mirror(yy, ch); xx = pgn_mirror(xx, cw);
      double acc = 0.0, wsum_y = 0.0;
      for (int dy = -ry; dy <= ry; ++dy) {
        const double wy = sig_y > 0.0 ? exp(-0.5 * dy * dy / (sig_y * sig_y)) : 1.0;
        double row = 0.0, wsum_x = 0.0;
        for (int dx = -rx; dx <= rx; ++dx) {
          const double wx = sig_x > 0.0 ? exp(-0.5 * dx * dx / (sig_x * sig_x)) : 1.0;
          row += wx * pix(yy + dy, xx + dx);
          wsum_x += wx;
        }
        acc += wy * row / wsum_x;
        wsum_y += wy;
      }
      return acc / wsum_y;
    };
    const double fy = (oy + 0.5) * sy - 0.5, fx = (ox + 0.5) * sx - 0.5;
    const int iy = (int)floor(fy), ix = (int)floor(fx);
    const double ty = fy - iy, tx = fx - ix;
    const double v = (1.0 - ty) * ((1.0 - tx) * blurred(iy, ix) + tx * blurred(iy, ix + 1)) +
                     ty * ((1.0 - tx) * blurred(iy + 1, ix) + tx * blurred(iy + 1, ix + 1));
    out[idx] = (float)v;
  }
}

cudaError_t pgn_launch_hmr_input(const float* image, int H, int W, int x0, int y0, int x1, int y1, int R,
                                 const float* mean3, const float* std3, int quantize, float* out, cudaStream_t stream) {
  const int n = 3 * R * R;
  if (n <= 0) return cudaSuccess;
  pgn_hmr_input_kernel<<<(n + 255) / 256, 256, 0, stream>>>(image, H, W, x0, y0, x1 - x0, y1 - y0, R, mean3[0], mean3[1], mean3[2],
                                                            std3[0], std3[1], std3[2], quantize, out);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Backward of NeRF.raw2outputs (core/networks/nerf.py:150-205) for the training step
// (core/trainer.py:321-370: the loss reads rgb_map and acc_map of both passes): one warp per ray.
//   w_i = a_i T_i,  T_i = prod_{k<i} (1 - a_k + 1e-10),  a_i = 1 - exp(-relu(raw_s / B) d_i)
//   dL/da_i = T_i G_i - (sum_{j>i} w_j G_j) / (1 - a_i + 1e-10),   G_i = g_rgb . c_i + g_acc [acc < 1]
//   dL/draw_s = dL/da_i d_i exp(-s_i d_i) [raw_s > 0] / B ;  dL/draw_c = w_i g_rgb (1 + 2 eps) sig (1 - sig)
// (no gradient flows through z: the importance samples are detached, ray_utils.py:286.)
// ---------------------------------------------------------------------------
template <int S>
__global__ void pgn_composite_backward_kernel(PgnRayRefs rays, const PgnScalars* __restrict__ scp, const float* __restrict__ raw,
                                              const float* __restrict__ z, const float* __restrict__ g_rgb,
                                              const float* __restrict__ g_acc, const float* __restrict__ noise,
                                              float* __restrict__ d_raw) {
  constexpr int CH = (S + 31) / 32;
  const PgnScalars& sc = *scp;
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const float keps = 1.0f + 2.0f * sc.rgb_eps;
  for (long long r = warp; r < rays.n_rays; r += nwarps) {
    const float* d = rays.ray_batch + r * 11 + 3;
    const float dn = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const float* rw = raw + r * S * 4;
    const float* zz = z + r * S;
    const float gr = g_rgb[r * 3], gg = g_rgb[r * 3 + 1], gb = g_rgb[r * 3 + 2];
    float al[CH], om[CH], dist[CH], p[CH], c0[CH], c1[CH], c2[CH], s0[CH], s1[CH], s2[CH];
    float lane_prod = 1.0f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int i = lane * CH + c;
      al[c] = 0.f; om[c] = 1.f; dist[c] = 0.f; c0[c] = c1[c] = c2[c] = 0.f; s0[c] = s1[c] = s2[c] = 0.f;
      p[c] = lane_prod;
      if (i < S) {
        float di = (i + 1 < S) ? __fsub_rn(zz[i + 1], zz[i]) : 1e10f;
        di = __fmul_rn(di, dn);
        const float pre = rw[i * 4 + 3] / sc.density_scale + (noise ? noise[r * S + i] : 0.0f);     // nerf.py:165: act(raw / B + noise)
        const float sig = fmaxf(pre, 0.0f);
        al[c] = 1.0f - expf(-__fmul_rn(sig, di));
        om[c] = __fadd_rn(__fsub_rn(1.0f, al[c]), 1e-10f);
        dist[c] = pre > 0.0f ? di : 0.0f;                  // ReLU mask folded into the distance factor
        s0[c] = 1.0f / (1.0f + expf(-rw[i * 4 + 0])); s1[c] = 1.0f / (1.0f + expf(-rw[i * 4 + 1])); s2[c] = 1.0f / (1.0f + expf(-rw[i * 4 + 2]));
        c0[c] = s0[c] * keps - sc.rgb_eps; c1[c] = s1[c] * keps - sc.rgb_eps; c2[c] = s2[c] * keps - sc.rgb_eps;
        lane_prod *= om[c];
      }
    }
    float incl = lane_prod;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl *= up;
    }
    float excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 1.0f;
    float w[CH], G[CH], T[CH];
    float sw = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) { T[c] = excl * p[c]; w[c] = al[c] * T[c]; sw += w[c]; }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sw += __shfl_xor_sync(0xffffffffu, sw, off);
    const float ga = (g_acc && sw < 1.0f) ? g_acc[r] : 0.0f;          // acc_map = min(sum w, 1)
    float lane_sum = 0.f, pre[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      G[c] = gr * c0[c] + gg * c1[c] + gb * c2[c] + ga;
      lane_sum += w[c] * G[c];
      pre[c] = lane_sum;                                               // inclusive within the lane
    }
    float incl_s = lane_sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, incl_s, off);
      if (lane >= off) incl_s += up;
    }
    const float total = __shfl_sync(0xffffffffu, incl_s, 31);
    const float before = incl_s - lane_sum;                            // sum over earlier lanes
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int i = lane * CH + c;
      if (i < S) {
        const float after = total - (before + pre[c]);                 // sum_{j>i} w_j G_j
        const float dalpha = T[c] * G[c] - after / om[c];
        const float dsig = dalpha * dist[c] * (1.0f - al[c]);
        float* o = d_raw + (r * S + i) * 4;
        o[0] = w[c] * gr * keps * s0[c] * (1.0f - s0[c]);
        o[1] = w[c] * gg * keps * s1[c] * (1.0f - s1[c]);
        o[2] = w[c] * gb * keps * s2[c] * (1.0f - s2[c]);
        o[3] = dsig / sc.density_scale;
      }
    }
  }
}

cudaError_t pgn_launch_composite_backward(const PgnRayRefs& rays, const PgnScalars* sc_dev, const float* raw, const float* z,
                                          int s, const float* g_rgb, const float* g_acc, const float* noise, float* d_raw,
                                          cudaStream_t stream) {
  if (rays.n_rays == 0) return cudaSuccess;
  const int block = 256;
  const long long grid = min((rays.n_rays * 32 + block - 1) / block, (long long)148 * 8);
  if (s == PGN_S) pgn_composite_backward_kernel<PGN_S><<<(unsigned)grid, block, 0, stream>>>(rays, sc_dev, raw, z, g_rgb, g_acc, noise, d_raw);
  else if (s == PGN_T) pgn_composite_backward_kernel<PGN_T><<<(unsigned)grid, block, 0, stream>>>(rays, sc_dev, raw, z, g_rgb, g_acc, noise, d_raw);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Backward of encode_inputs (core/raycasters.py:476-555; encoders.py:8-37,110-122,181-193;
// cutoff_embedder.py:111-174) w.r.t. the world->joint transforms: the gradient the pose generator / pose
// optimisation receives (BASELINE.json configs[4]: "differentiable render, grad w.r.t. pose/bone transforms").
// One thread per (ray, joint) loops over the ray's samples; g_enc [n, n_z, 1080] = dL/d(network input) in the
// reference channel order; d_skts [n, 24, 4, 4] receives the per-ray gradient (bottom row 0).  Sample positions
// carry no gradient (rays are inputs, importance samples are detached).
//   pts_t = R p + t, v = |pts_t|, r = pts_t / v, w = 1 - sigmoid(tau (v - c)),  w' = -tau w (1 - w)
//   e_k = phi_k(v) w (v-embed), q_k,a = psi_k(u_a) w_d(v) with u = R d / |R d| (view embed)
// ---------------------------------------------------------------------------
__device__ __forceinline__ float pgn_ldg_f(const float* p) { return __ldg(p); }
__device__ __forceinline__ float pgn_ldg_f(const __nv_bfloat16* p) {
  return __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(p)) << 16);
}

// G = float: one [rows,1080] matrix (g_xp = g_enc, g_d = g_enc + 432, both row strides 1080);
// G = __nv_bfloat16: the two GEMM outputs of the training backward as they are, [rows,432] and [rows,648].
// kTB (bf16 only): both matrices tile-blocked, [row / 128][column / 8][128][8] - the layout pgn_mlp_input_grads' epilogue
// writes with fully coalesced stores; element (row, col) sits at (row / 128) * 128 * stride + (col / 8) * 1024 +
// (row % 128) * 8 + col % 8.
template <typename G, bool kTB = false>
__global__ void pgn_encode_backward_kernel(PgnRayRefs rays, const PgnScalars* __restrict__ scp, const float* __restrict__ z,
                                           int n_z, const G* __restrict__ g_xp, int stride_xp, const G* __restrict__ g_d,
                                           int stride_d, float* __restrict__ d_skts) {
  const PgnScalars& sc = *scp;
  const long long total = rays.n_rays * PGN_J;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % PGN_J);
    const long long ray = idx / PGN_J;
    const float* rb = rays.ray_batch + ray * 11;
    const float4* m = reinterpret_cast<const float4*>(pgn_ray_skts(rays, ray) + j * 16);
    const float4 m0 = __ldg(m), m1 = __ldg(m + 1), m2 = __ldg(m + 2);
    const float dd[3] = {rb[3], rb[4], rb[5]};
    // joint-frame view direction (per ray)
    float dj[3] = {fmaf(m0.z, dd[2], fmaf(m0.y, dd[1], m0.x * dd[0])), fmaf(m1.z, dd[2], fmaf(m1.y, dd[1], m1.x * dd[0])),
                   fmaf(m2.z, dd[2], fmaf(m2.y, dd[1], m2.x * dd[0]))};
    const float nd = sqrtf(dj[0] * dj[0] + dj[1] * dj[1] + dj[2] * dj[2]);
    const float ndc = fmaxf(nd, 1e-12f);
    const float u[3] = {dj[0] / ndc, dj[1] / ndc, dj[2] / ndc};
    float gR[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, gt[3] = {0, 0, 0}, gu[3] = {0, 0, 0};
    // sin / cos of the view-direction frequencies (per ray): one sincosf + double-angle steps
    float usn[3][PGN_LD], ucs[3][PGN_LD];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      float sn, cs;
      sincosf(u[a], &sn, &cs);
#pragma unroll
      for (int f = 0; f < PGN_LD; ++f) {
        usn[a][f] = sn; ucs[a][f] = cs;
        const float t2 = cs + cs;
        sn = t2 * sn;
        cs = fmaf(t2, cs, -1.0f);
      }
    }
    for (int s = 0; s < n_z; ++s) {
      const long long rs = ray * n_z + s;
      float p[3];
      pgn_sample_point(rb, rb + 3, z[rs], p[0], p[1], p[2]);
      const float x = fmaf(m0.z, p[2], fmaf(m0.y, p[1], m0.x * p[0])) + m0.w;
      const float y = fmaf(m1.z, p[2], fmaf(m1.y, p[1], m1.x * p[0])) + m1.w;
      const float zz = fmaf(m2.z, p[2], fmaf(m2.y, p[1], m2.x * p[0])) + m2.w;
      const float v = sqrtf(x * x + y * y + zz * zz);
      const float vc = fmaxf(v, 1e-12f);
      const float r[3] = {x / vc, y / vc, zz / vc};
      const float w = pgn_window<false>(v, sc.tau_v, sc.cutoff_v[j]);
      const float wd = pgn_window<false>(v, sc.tau_d, sc.cutoff_d[j]);
      const float dw = -sc.tau_v * w * (1.0f - w), dwd = -sc.tau_d * wd * (1.0f - wd);
      // row bases and the element offset of column c inside a row
      const G* ge = kTB ? g_xp + (rs >> 7) * (128ll * stride_xp) + (rs & 127) * 8 : g_xp + rs * stride_xp;
      const G* gdr = kTB ? g_d + (rs >> 7) * (128ll * stride_d) + (rs & 127) * 8 : g_d + rs * stride_d;
      auto co = [](int c) -> int { return kTB ? ((c >> 3) << 10) + (c & 7) : c; };
      // v-embed: k = 0 -> v, k = 1 + 2f -> sin(2^f v), 2 + 2f -> cos(2^f v)
      float gv = pgn_ldg_f(ge + co(j)) * (w + v * dw);
      float sn, cs;
      sincosf(v, &sn, &cs);
#pragma unroll
      for (int f = 0; f < PGN_LV; ++f) {
        const float fr = (float)(1 << f);
        gv += pgn_ldg_f(ge + co((1 + 2 * f) * PGN_J + j)) * (fr * cs * w + sn * dw);
        gv += pgn_ldg_f(ge + co((2 + 2 * f) * PGN_J + j)) * (-fr * sn * w + cs * dw);
        const float t2 = cs + cs;                 // double angle
        sn = t2 * sn;
        cs = fmaf(t2, cs, -1.0f);
      }
      // view embed: every channel is psi(u_a) * wd(v)
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const int cq = j * 3 + a;
        const float g0 = pgn_ldg_f(gdr + co(cq));
        gv += g0 * u[a] * dwd;
        gu[a] += g0 * wd;
#pragma unroll
        for (int f = 0; f < PGN_LD; ++f) {
          const float fr = (float)(1 << f);
          const float g1 = pgn_ldg_f(gdr + co(cq + (1 + 2 * f) * 72)), g2 = pgn_ldg_f(gdr + co(cq + (2 + 2 * f) * 72));
          gv += (g1 * usn[a][f] + g2 * ucs[a][f]) * dwd;
          gu[a] += (g1 * ucs[a][f] - g2 * usn[a][f]) * fr * wd;
        }
      }
      // r = pts_t / v
      const float gr[3] = {pgn_ldg_f(ge + co(360 + j * 3)), pgn_ldg_f(ge + co(360 + j * 3 + 1)), pgn_ldg_f(ge + co(360 + j * 3 + 2))};
      float gp[3];
      if (v > 1e-12f) {
        const float rg = r[0] * gr[0] + r[1] * gr[1] + r[2] * gr[2];
#pragma unroll
        for (int a = 0; a < 3; ++a) gp[a] = gv * r[a] + (gr[a] - r[a] * rg) / v;
      } else {
#pragma unroll
        for (int a = 0; a < 3; ++a) gp[a] = gr[a] / 1e-12f;
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        gR[a * 3 + 0] += gp[a] * p[0]; gR[a * 3 + 1] += gp[a] * p[1]; gR[a * 3 + 2] += gp[a] * p[2];
        gt[a] += gp[a];
      }
    }
    // u = R d / |R d|
    float gdj[3];
    if (nd > 1e-12f) {
      const float ug = u[0] * gu[0] + u[1] * gu[1] + u[2] * gu[2];
#pragma unroll
      for (int a = 0; a < 3; ++a) gdj[a] = (gu[a] - u[a] * ug) / nd;
    } else {
#pragma unroll
      for (int a = 0; a < 3; ++a) gdj[a] = gu[a] / 1e-12f;
    }
    float* o = d_skts + (ray * PGN_J + j) * 16;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      o[a * 4 + 0] = gR[a * 3 + 0] + gdj[a] * dd[0];
      o[a * 4 + 1] = gR[a * 3 + 1] + gdj[a] * dd[1];
      o[a * 4 + 2] = gR[a * 3 + 2] + gdj[a] * dd[2];
      o[a * 4 + 3] = gt[a];
    }
    o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 0.f;
  }
}

cudaError_t pgn_launch_encode_backward(const PgnRayRefs& rays, const PgnScalars* sc_dev, const float* z, int n_z,
                                       const float* g_enc, float* d_skts, cudaStream_t stream) {
  const long long total = rays.n_rays * PGN_J;
  if (total == 0) return cudaSuccess;
  const int block = 96;                                   // 4 rays x 24 joints: the 24 threads of a ray read contiguous channels
  const long long grid = min((total + block - 1) / block, (long long)148 * 32);
  pgn_encode_backward_kernel<float><<<(unsigned)grid, block, 0, stream>>>(rays, sc_dev, z, n_z, g_enc, PGN_ENC, g_enc + PGN_ENC_P,
                                                                          PGN_ENC, d_skts);
  return cudaGetLastError();
}

cudaError_t pgn_launch_encode_backward_bf16(const PgnRayRefs& rays, const PgnScalars* sc_dev, const float* z, int n_z,
                                            const __nv_bfloat16* g_xp, const __nv_bfloat16* g_d, int tile_blocked, float* d_skts,
                                            cudaStream_t stream) {
  const long long total = rays.n_rays * PGN_J;
  if (total == 0) return cudaSuccess;
  const int block = 96;
  const long long grid = min((total + block - 1) / block, (long long)148 * 32);
  if (tile_blocked)
    pgn_encode_backward_kernel<__nv_bfloat16, true><<<(unsigned)grid, block, 0, stream>>>(rays, sc_dev, z, n_z, g_xp, PGN_ENC_P, g_d,
                                                                                          PGN_ENC - PGN_ENC_P, d_skts);
  else
    pgn_encode_backward_kernel<__nv_bfloat16><<<(unsigned)grid, block, 0, stream>>>(rays, sc_dev, z, n_z, g_xp, PGN_ENC_P, g_d,
                                                                                    PGN_ENC - PGN_ENC_P, d_skts);
  return cudaGetLastError();
}
