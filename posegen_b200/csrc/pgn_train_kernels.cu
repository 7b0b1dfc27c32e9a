// Training-step glue kernels of the A-NeRF MLP backward (HBM-bound, CUDA cores).
//
//   pgn_encode_bf16_kernel   encode_inputs (core/raycasters.py:476-555) straight to the bf16 operand the weight-gradient
//                            GEMMs read (x_p | d_emb, reference channel order), rows staged in shared memory and stored
//                            with 16-byte coalesced writes.
//   pgn_mlp_delta_kernel     one pass over a layer's delta matrix that fuses what autograd does in four:
//                            dZ = [act > 0] * (dH + rs @ wr)        ReLU backward (core/networks/nerf.py:96-99,129)
//                            colsum += sum_rows dZ                   bias gradient of the layer
//                            wsum   += rs^T @ act                    weight gradient of a 1- or 3-row head that reads `act`
//                            (alpha_linear on h7, rgb_linear on the view layer; nerf.py:102,131)
//                            with `rs @ wr` the head's contribution to dL/d act (outer product of the per-row head deltas
//                            and the head weights).
#include "pgn_common.cuh"
#include "pgn_kernels.h"

namespace {

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

constexpr int kEncRows = 8;                                  // rows (samples) per block iteration
constexpr int kEncThreads = kEncRows * PGN_J;                // one thread per (row, joint)

__global__ void __launch_bounds__(kEncThreads) pgn_encode_bf16_kernel(PgnRayRefs rays, const PgnScalars* __restrict__ scp,
                                                                      const float* __restrict__ z, int n_z,
                                                                      __nv_bfloat16* __restrict__ enc) {
  __shared__ __align__(16) __nv_bfloat16 tile[kEncRows][PGN_ENC];
  const PgnScalars& sc = *scp;
  const long long rows = rays.n_rays * n_z;
  const int j = threadIdx.x % PGN_J, rl = threadIdx.x / PGN_J;
  for (long long row0 = (long long)blockIdx.x * kEncRows; row0 < rows; row0 += (long long)gridDim.x * kEncRows) {
    const long long rs = row0 + rl;
    if (rs < rows) {
      const long long ray = rs / n_z;
      const float* rb = rays.ray_batch + ray * 11;
      const float4* m = reinterpret_cast<const float4*>(pgn_ray_skts(rays, ray) + j * 16);
      const float4 m0 = __ldg(m), m1 = __ldg(m + 1), m2 = __ldg(m + 2);
      float px, py, pz;
      pgn_sample_point(rb, rb + 3, z[rs], px, py, pz);
      // same arithmetic as the operand generator of the fused forward kernel (pgn_render_bf16.cu encode_x): fast
      // window, one __sincosf + double-angle steps; the values are rounded to bf16 anyway
      const PgnJointGeom g = pgn_joint_geom<true>(m0, m1, m2, px, py, pz, sc.tau_v, sc.cutoff_v[j]);
      const float wd = pgn_window<true>(g.v, sc.tau_d, sc.cutoff_d[j]);
      __nv_bfloat16* e = tile[rl];
      float sn, cs;
      __sincosf(g.v, &sn, &cs);
      e[j] = __float2bfloat16_rn(g.v * g.w);
#pragma unroll
      for (int f = 0; f < PGN_LV; ++f) {
        e[(1 + 2 * f) * PGN_J + j] = __float2bfloat16_rn(sn * g.w);
        e[(2 + 2 * f) * PGN_J + j] = __float2bfloat16_rn(cs * g.w);
        const float t2 = cs + cs;
        sn = t2 * sn;
        cs = fmaf(t2, cs, -1.0f);
      }
      e[360 + j * 3 + 0] = __float2bfloat16_rn(g.rx);
      e[360 + j * 3 + 1] = __float2bfloat16_rn(g.ry);
      e[360 + j * 3 + 2] = __float2bfloat16_rn(g.rz);
      float dj[3];
      pgn_joint_dir(m0, m1, m2, rb + 3, dj[0], dj[1], dj[2]);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        __sincosf(dj[a], &sn, &cs);                 // |dj| <= 1
        e[PGN_ENC_P + j * 3 + a] = __float2bfloat16_rn(wd * dj[a]);
#pragma unroll
        for (int f = 0; f < PGN_LD; ++f) {
          e[PGN_ENC_P + (1 + 2 * f) * 72 + j * 3 + a] = __float2bfloat16_rn(wd * sn);
          e[PGN_ENC_P + (2 + 2 * f) * 72 + j * 3 + a] = __float2bfloat16_rn(wd * cs);
          const float t2 = cs + cs;
          sn = t2 * sn;
          cs = fmaf(t2, cs, -1.0f);
        }
      }
    }
    __syncthreads();
    // 8 consecutive rows are one contiguous 17,280-byte block of the output
    const long long nrow = min((long long)kEncRows, rows - row0);
    const int n16 = (int)(nrow * (PGN_ENC * 2 / 16));
    uint4* dst = reinterpret_cast<uint4*>(enc + row0 * PGN_ENC);
    const uint4* src = reinterpret_cast<const uint4*>(&tile[0][0]);
    for (int i = threadIdx.x; i < n16; i += kEncThreads) dst[i] = src[i];
    __syncthreads();
  }
}

template <int C, int NRS>
__global__ void __launch_bounds__(256) pgn_mlp_delta_kernel(uint4* __restrict__ dh, int has_in, const uint4* __restrict__ act,
                                                            long long m, const float* __restrict__ rs, int rs_stride,
                                                            const float* __restrict__ wr, float* __restrict__ colsum,
                                                            float* __restrict__ wsum) {
  constexpr int TPR = C / 8;            // threads per row (8 columns = one 16-byte word each)
  constexpr int RPB = 256 / TPR;        // rows per block iteration
  constexpr int NR = NRS > 0 ? NRS : 1;
  __shared__ float red[RPB][C + 1];
  const int cg = threadIdx.x % TPR, rl = threadIdx.x / TPR;
  float cs[8], ws[NR][8], wv[NR][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    cs[i] = 0.f;
#pragma unroll
    for (int k = 0; k < NR; ++k) { ws[k][i] = 0.f; wv[k][i] = NRS > 0 ? __ldg(wr + k * C + cg * 8 + i) : 0.f; }
  }
  for (long long row = (long long)blockIdx.x * RPB + rl; row < m; row += (long long)gridDim.x * RPB) {
    const size_t off = (size_t)row * TPR + cg;
    uint4 a = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);      // no mask: every column "active"
    if (act) a = __ldg(act + off);
    uint4 d = make_uint4(0u, 0u, 0u, 0u);
    if (has_in) d = dh[off];
    float r[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) r[k] = NRS > 0 ? __ldg(rs + (size_t)row * rs_stride + k) : 0.f;
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
    uint32_t ow[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float a0 = bf16_lo(aw[w]), a1 = bf16_hi(aw[w]);
      float p0 = bf16_lo(dw[w]), p1 = bf16_hi(dw[w]);
#pragma unroll
      for (int k = 0; k < NRS; ++k) {
        p0 = fmaf(r[k], wv[k][2 * w], p0);
        p1 = fmaf(r[k], wv[k][2 * w + 1], p1);
        ws[k][2 * w] = fmaf(r[k], a0, ws[k][2 * w]);
        ws[k][2 * w + 1] = fmaf(r[k], a1, ws[k][2 * w + 1]);
      }
      p0 = a0 > 0.f ? p0 : 0.f;
      p1 = a1 > 0.f ? p1 : 0.f;
      cs[2 * w] += p0;
      cs[2 * w + 1] += p1;
      ow[w] = pack_bf16x2(p0, p1);
    }
    dh[off] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
  // block reduction over the row lanes, then one atomic per column per block
  for (int q = 0; q < 1 + NRS; ++q) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) red[rl][cg * 8 + i] = q == 0 ? cs[i] : ws[q > 0 ? q - 1 : 0][i];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
      float s = 0.f;
#pragma unroll 4
      for (int r2 = 0; r2 < RPB; ++r2) s += red[r2][c];
      if (q == 0) atomicAdd(colsum + c, s);
      else if (wsum) atomicAdd(wsum + (q - 1) * C + c, s);
    }
  }
}

// dG[m,128] = [g > 0] * (d_rgb W_rgb) with the view layer's ReLU mask given as bits (masks-only dump of the forward):
// one thread per (row, 8 columns)
__global__ void __launch_bounds__(256) pgn_view_delta_bits_kernel(uint4* __restrict__ dG, const float* __restrict__ d_raw,
                                                                  const float* __restrict__ w_rgb,
                                                                  const uint32_t* __restrict__ vmask, long long m) {
  const int cg = threadIdx.x & 15;
  float wv[3][8];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) wv[k][i] = __ldg(w_rgb + k * 128 + cg * 8 + i);
  for (long long row = (long long)blockIdx.x * 16 + (threadIdx.x >> 4); row < m; row += (long long)gridDim.x * 16) {
    const float r0 = __ldg(d_raw + row * 4), r1 = __ldg(d_raw + row * 4 + 1), r2 = __ldg(d_raw + row * 4 + 2);
    const uint32_t bits = __ldg(vmask + (size_t)(cg >> 2) * (size_t)m + row) >> ((cg & 3) * 8);      // word planes [4][m]
    uint32_t ow[4];
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      float p0 = fmaf(r2, wv[2][2 * w], fmaf(r1, wv[1][2 * w], r0 * wv[0][2 * w]));
      float p1 = fmaf(r2, wv[2][2 * w + 1], fmaf(r1, wv[1][2 * w + 1], r0 * wv[0][2 * w + 1]));
      if (!(bits & (1u << (2 * w)))) p0 = 0.f;
      if (!(bits & (1u << (2 * w + 1)))) p1 = 0.f;
      ow[w] = pack_bf16x2(p0, p1);
    }
    dG[(size_t)row * 16 + cg] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  }
}

template <int C, int NRS>
cudaError_t launch_delta(void* dh, int has_in, const void* act, long long m, const float* rs, int rs_stride, const float* wr,
                         float* colsum, float* wsum, int num_sms, cudaStream_t stream) {
  constexpr int RPB = 256 / (C / 8);
  const long long grid = min((m + RPB - 1) / RPB, (long long)num_sms * 8);
  pgn_mlp_delta_kernel<C, NRS><<<(unsigned)grid, 256, 0, stream>>>(reinterpret_cast<uint4*>(dh), has_in,
                                                                    reinterpret_cast<const uint4*>(act), m, rs, rs_stride, wr,
                                                                    colsum, wsum);
  return cudaGetLastError();
}

}  // namespace

cudaError_t pgn_launch_encode_bf16(const PgnRayRefs& rays, const PgnScalars* sc_dev, const float* z, int n_z,
                                   __nv_bfloat16* enc, cudaStream_t stream) {
  const long long rows = rays.n_rays * n_z;
  if (rows == 0) return cudaSuccess;
  const long long grid = min((rows + kEncRows - 1) / kEncRows, (long long)148 * 32);      // (x10: 134 us per pass, x32: 124 us)
  pgn_encode_bf16_kernel<<<(unsigned)grid, kEncThreads, 0, stream>>>(rays, sc_dev, z, n_z, enc);
  return cudaGetLastError();
}

cudaError_t pgn_launch_view_delta_bits(void* dG, const float* d_raw, const float* w_rgb, const void* vmask, long long m,
                                       int num_sms, cudaStream_t stream) {
  if (m == 0) return cudaSuccess;
  const long long grid = min((m + 15) / 16, (long long)num_sms * 8);
  pgn_view_delta_bits_kernel<<<(unsigned)grid, 256, 0, stream>>>(reinterpret_cast<uint4*>(dG), d_raw, w_rgb,
                                                                reinterpret_cast<const uint32_t*>(vmask), m);
  return cudaGetLastError();
}

cudaError_t pgn_launch_mlp_delta(void* dh, int has_in, const void* act, long long m, int C, const float* rs, int rs_stride,
                                 int nrs, const float* wr, float* colsum, float* wsum, int num_sms, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(colsum, 0, sizeof(float) * C, stream);
  if (e != cudaSuccess) return e;
  if (wsum && nrs > 0) {
    e = cudaMemsetAsync(wsum, 0, sizeof(float) * C * nrs, stream);
    if (e != cudaSuccess) return e;
  }
  if (m == 0) return cudaSuccess;
  if (C == 256 && nrs == 0) return launch_delta<256, 0>(dh, has_in, act, m, rs, rs_stride, wr, colsum, wsum, num_sms, stream);
  if (C == 256 && nrs == 1) return launch_delta<256, 1>(dh, has_in, act, m, rs, rs_stride, wr, colsum, wsum, num_sms, stream);
  if (C == 128 && nrs == 0) return launch_delta<128, 0>(dh, has_in, act, m, rs, rs_stride, wr, colsum, wsum, num_sms, stream);
  if (C == 128 && nrs == 3) return launch_delta<128, 3>(dh, has_in, act, m, rs, rs_stride, wr, colsum, wsum, num_sms, stream);
  return cudaErrorInvalidValue;
}


// ---------------------------------------------------------------------------------------------------------------
// Optcodes frame codes (core/networks/embedding.py:4-46; core/networks/nerf.py:104-131: views_linears.0 reads
// [feature | input_views | framecodes(frame_idx)], 16 extra input columns).  A ray's code is constant along the ray,
// so its contribution W_v[:, 904:920] code[cam] is a per-ray additive term of the view layer's pre-activation:
//   codes_ext[n]   <- mean over the n codes (the reference's eval rule for idx < 0, embedding.py:23-24)
//   fc_table[i][c] <- sum_k W_v[c][904 + k] codes_ext[i][k]            (fp32; the tensor-core tier adds it in its epilogue)
// ---------------------------------------------------------------------------------------------------------------
__global__ void pgn_framecode_tables_kernel(float* __restrict__ codes_ext, int n, const float* __restrict__ w_view, int view_ld,
                                            float* __restrict__ fc_table) {
  __shared__ float code[16];
  const int i = blockIdx.x;                 // 0 .. n (row n: the mean code)
  if (threadIdx.x < 16) {
    float v;
    if (i < n) v = codes_ext[i * 16 + threadIdx.x];
    else {
      float sacc = 0.f;
      for (int r = 0; r < n; ++r) sacc += codes_ext[r * 16 + threadIdx.x];
      v = sacc / (float)n;
      codes_ext[n * 16 + threadIdx.x] = v;
    }
    code[threadIdx.x] = v;
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    const float* w = w_view + (size_t)threadIdx.x * view_ld + (view_ld - 16);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(w[k], code[k], acc);
    fc_table[(size_t)i * 128 + threadIdx.x] = acc;
  }
}

cudaError_t pgn_launch_framecode_tables(float* codes_ext, int n_codes, const float* w_view, int view_ld, float* fc_table, cudaStream_t stream) {
  if (n_codes <= 0) return cudaSuccess;
  pgn_framecode_tables_kernel<<<n_codes + 1, 128, 0, stream>>>(codes_ext, n_codes, w_view, view_ld, fc_table);
  return cudaGetLastError();
}

// backward: with dGr[ray] = sum over the ray's samples of dG (the code is shared by them)
//   g_wvc[c][k]        += dGr[ray][c] code[row(ray)][k]        (the 16 extra columns of views_linears.0.weight)
//   g_codes[cam][k]    += sum_c dGr[ray][c] W_v[c][904 + k]    (rays with a real camera index; in eval the mean row gets none)
// one warp per ray: lane owns 4 of the 128 columns.
__global__ void __launch_bounds__(256) pgn_framecode_backward_kernel(const uint2* __restrict__ dG, long long n_rays, int nz,
                                                                     const int* __restrict__ cams, int n_codes,
                                                                     const float* __restrict__ codes_ext, const float* __restrict__ w_view,
                                                                     int view_ld, float* __restrict__ g_wvc, int g_ld, float* __restrict__ g_codes) {
  __shared__ float s_w[128 * 16];            // block-local partial of g_wvc
  for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) s_w[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float wv[4][16];                            // this lane's 4 rows of W_v[:, 904:920]
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < 16; ++k) wv[c][k] = __ldg(w_view + (size_t)(lane * 4 + c) * view_ld + (view_ld - 16) + k);
  for (long long ray = (long long)blockIdx.x * 8 + wib; ray < n_rays; ray += (long long)gridDim.x * 8) {
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    for (int sidx = 0; sidx < nz; ++sidx) {
      const uint2 v = __ldg(dG + (ray * nz + sidx) * 32 + lane);      // 4 bf16 = this lane's columns
      g[0] += __uint_as_float(v.x << 16); g[1] += __uint_as_float(v.x & 0xffff0000u);
      g[2] += __uint_as_float(v.y << 16); g[3] += __uint_as_float(v.y & 0xffff0000u);
    }
    int cam = cams ? cams[ray] : -1;
    const int row = (cam < 0 || cam >= n_codes) ? n_codes : cam;
    float gc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float ck = __ldg(codes_ext + row * 16 + k);
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        atomicAdd(&s_w[(lane * 4 + c) * 16 + k], g[c] * ck);
        acc = fmaf(g[c], wv[c][k], acc);
      }
      gc[k] = acc;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) gc[k] += __shfl_xor_sync(0xffffffffu, gc[k], off);
    }
    if (row < n_codes && lane < 16) {
      float mine = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) if (k == lane) mine = gc[k];
      atomicAdd(g_codes + row * 16 + lane, mine);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) atomicAdd(g_wvc + (size_t)(i >> 4) * g_ld + (i & 15), s_w[i]);
}

cudaError_t pgn_launch_framecode_backward(const void* dG, long long n_rays, int nz, const int* cams, int n_codes, const float* codes_ext,
                                          const float* w_view, int view_ld, float* g_wvc, int g_wvc_ld, float* g_codes, cudaStream_t stream) {
  if (n_rays <= 0 || n_codes <= 0) return cudaSuccess;
  const long long blocks = (n_rays + 7) / 8;
  pgn_framecode_backward_kernel<<<(unsigned)(blocks < 148 * 2 ? blocks : 148 * 2), 256, 0, stream>>>(
      reinterpret_cast<const uint2*>(dG), n_rays, nz, cams, n_codes, codes_ext, w_view, view_ld, g_wvc, g_wvc_ld, g_codes);
  return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------------------------------
// Parameter hand-off (pgn_upload_weights with device pointers: the per-step refresh after optimizer.step): the 24 tensors
// of a net are gathered from the nn.Parameters' own storages into the context's contiguous copies by ONE launch
// instead of 24 x 2 cudaMemcpyAsync (48 x 1.5 us per step, serialised, in the eager and the graphed step alike).
// ---------------------------------------------------------------------------------------------------------------
struct GatherSegs { const float* src[24]; float* dst[24]; int n[24]; };

__global__ void __launch_bounds__(256) pgn_gather_params_kernel(GatherSegs g) {
  const int seg = blockIdx.y;
  const float* __restrict__ s = g.src[seg];
  float* __restrict__ d = g.dst[seg];
  const int n = g.n[seg];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) d[i] = s[i];
}

cudaError_t pgn_launch_gather_params(const float* const* src_w, const float* const* src_b, float* const* dst_w, float* const* dst_b,
                                     const int* n_w, const int* n_b, cudaStream_t stream) {
  GatherSegs g;
  for (int l = 0; l < 12; ++l) {
    g.src[l] = src_w[l]; g.dst[l] = dst_w[l]; g.n[l] = n_w[l];
    g.src[12 + l] = src_b[l]; g.dst[12 + l] = dst_b[l]; g.n[12 + l] = n_b[l];
  }
  pgn_gather_params_kernel<<<dim3(24, 24), 256, 0, stream>>>(g);
  return cudaGetLastError();
}
