// Weight gradients of the A-NeRF MLP backward (training step, BASELINE.json configs[3]) as ONE split-K tcgen05 kernel.
//
// What autograd does for every nn.Linear of core/networks/nerf.py:94-148 is dW_l = dZ_l^T h_{l-1}: a [256 x K_in] GEMM
// whose contraction runs over ALL samples of the batch (2-2.5 x 10^5 rows per pass).  Both operands already sit in HBM
// as row-major bf16 matrices with the contraction dimension (rows) outermost: the deltas dZ_l [rows,256] written by
// pgn_delta_chain, the activations h_l [rows,256] dumped by the training forward, the regenerated network input
// [rows,1080].  In UMMA terms both are "MN-major" operands (8 consecutive M/N elements = 16 contiguous bytes, K strided),
// so no transpose is ever materialised:
//
//   * a stage is 32 rows of A (<= 256 columns) and of B (<= 256 columns), staged with 16-byte cp.async into the
//     SWIZZLE_NONE MN-major canonical image [column / 8][row][8] (core matrix = 8 rows x 16 B contiguous; LBO = 128 B to
//     the next 8 rows, SBO = 512 B to the next 8 columns).  A warp copy covers 8 rows x 4 column chunks: 64-byte
//     global segments (whole sectors) and four conflict-free 128-byte shared-memory wavefronts.
//   * one elected thread issues tcgen05.mma kind::f16 (M = 128, N <= 256, K = 16) with a_major = b_major = MN:
//     per stage 2 K-steps x (1 or 2) M-halves, fp32 accumulators in TMEM (2 x 256 columns = a 256 x 256 output tile).
//   * split-K: the units (one output tile each: an (A matrix, B column range) pair) get a number of CTAs proportional
//     to the bytes they stream; CTA k of a unit takes stages k, k + n, k + 2n, ... so that every unit sweeps the rows at
//     the same rate and operands shared by several units (dZ_5, dG, x_p are each read by 2-4 units) come from the 126 MB
//     L2 instead of HBM a second time.  At the end every CTA adds its partial tile to the fp32 gradient with
//     red.global.add.v4.f32 (the gradient buffer is zeroed by the launch wrapper).
//
// Roofline: HBM-bound.  Algorithmic bytes per row and pass = 8 x 512 (dZ) + 256 (dG) + 8 x 512 (h) + 2,160 (input)
// = 10.6 KB against 1.72 MFLOP (162 FLOP/B; the machine balance is ~210), i.e. ~0.37 ms per 245,760-row pass at the
// measured 6.5 TB/s; the tensor pipe is <= 45 % busy by construction.
//
// Roles (192 threads): warps 0-3 = cp.async loaders, then the TMEM -> red.add epilogue; warp 4 = MMA issuer (one lane);
// warp 5 owns the TMEM allocation.  6 stages x 32 KB in flight per SM.
#include <cuda_bf16.h>
#include "pgn_common.cuh"
#include "pgn_kernels.h"
#include "pgn_umma.cuh"

using namespace pgn;

namespace {

constexpr int kRows = 32;                      // rows (K of the GEMM) per stage
constexpr int kStages = 6;
constexpr int kOpBytes = 32 * kRows * 16;      // one operand of a stage: [32 column chunks][32 rows][16 B] = 16 KB
constexpr int kThreads = 192;
constexpr int kLoaders = 128;
constexpr int kMaxUnits = 16;

struct Unit {
  const __nv_bfloat16* A;    // [rows, lda] row-major, Ma columns used from column 0
  const __nv_bfloat16* B;    // [rows, ldb] row-major, Nb columns used from column 0 of this pointer
  float* out;                // fp32 [Ma, ld_out] tile origin (row = A column, column = B column)
  int lda, ldb, ld_out;
  int Ma;                    // 128 | 256
  int Nb;                    // valid B columns (multiple of 8)
  int Nmma;                  // UMMA N: Nb rounded up to 16 (the extra columns are zero-filled, never stored)
  int cta0, ncta;            // CTAs [cta0, cta0 + ncta) work on this unit
};
struct Params {
  Unit u[kMaxUnits];
  int n_units;
  long long m;               // rows
};

struct __align__(1024) Smem {
  uint8_t a[kStages][kOpBytes];
  uint8_t b[kStages][kOpBytes];
  uint64_t full[kStages], empty[kStages], acc_full;
  uint32_t tmem_slot;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// MN-major SWIZZLE_NONE descriptor of this kernel's stage image: LBO = 128 B (next 8 rows = K), SBO = 512 B (next 8 columns)
__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr) { return umma_smem_desc(saddr, 128, kRows * 16); }

__global__ void __launch_bounds__(kThreads, 1) pgn_wgrad_kernel(const __grid_constant__ Params p, int* __restrict__ status_g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  volatile int* status = status_g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int ui = 0;
  for (int i = 0; i < p.n_units; ++i)
    if ((int)blockIdx.x >= p.u[i].cta0 && (int)blockIdx.x < p.u[i].cta0 + p.u[i].ncta) ui = i;
  const Unit& u = p.u[ui];
  const bool active = (int)blockIdx.x >= u.cta0 && (int)blockIdx.x < u.cta0 + u.ncta;
  const int k = (int)blockIdx.x - u.cta0;
  const long long n_stages_total = (p.m + kRows - 1) / kRows;
  const long long n_mine = (active && n_stages_total > k) ? (n_stages_total - k + u.ncta - 1) / u.ncta : 0;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&sm.full[s], kLoaders); mbar_init(&sm.empty[s], 1); }
    mbar_init(&sm.acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc(&sm.tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_slot;
  const int halves = u.Ma / 128;

  if (warp < 4) {
    // ------------------------------------------------------------ loaders: warp w owns rows 8w .. 8w+7 of every stage
    const int r = warp * 8 + (lane & 7), c4 = lane >> 3;
    const int a_groups = u.Ma / 32, b_groups = (u.Nmma + 31) / 32;
    for (long long it = 0; it < n_mine; ++it) {
      const int s = (int)(it % kStages);
      const uint32_t use = (uint32_t)(it / kStages);
      if (use > 0 && !mbar_wait(&sm.empty[s], (use - 1) & 1, status, 801)) break;
      const long long row = ((long long)k + it * u.ncta) * kRows + r;
      const bool in = row < p.m;
      const long long rr = in ? row : 0;
      const uint8_t* ga = reinterpret_cast<const uint8_t*>(u.A + rr * u.lda);
      const uint8_t* gb = reinterpret_cast<const uint8_t*>(u.B + rr * u.ldb);
      const uint32_t da = smem_u32(sm.a[s]) + r * 16, db = smem_u32(sm.b[s]) + r * 16;
#pragma unroll 4
      for (int g = 0; g < a_groups; ++g) {
        const int ch = g * 4 + c4;
        cp_async16(da + ch * (kRows * 16), ga + ch * 16, in ? 16u : 0u);
      }
#pragma unroll 4
      for (int g = 0; g < b_groups; ++g) {
        const int ch = g * 4 + c4;
        if (ch * 8 < u.Nmma) {
          const bool ok = in && ch * 8 < u.Nb;
          cp_async16(db + ch * (kRows * 16), ok ? gb + ch * 16 : gb, ok ? 16u : 0u);     // zero-fill beyond Nb / beyond the last row
        }
      }
      cp_async_arrive_noinc(smem_u32(&sm.full[s]));
    }
    // ------------------------------------------------------------ epilogue: partial tile -> fp32 gradient (atomic adds)
    if (n_mine > 0 && mbar_wait(&sm.acc_full, 0, status, 803)) {
      tc_fence_after_sync();
      for (int h = 0; h < halves; ++h) {
        float* orow = u.out + (size_t)(h * 128 + warp * 32 + lane) * u.ld_out;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)h * 256u;
        for (int c0 = 0; c0 < u.Nmma; c0 += 16) {
          uint32_t v[16];
          tmem_ld_32x16(taddr + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (c0 + 4 * q < u.Nb)
              red_add_v4(orow + c0 + 4 * q, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                         __uint_as_float(v[4 * q + 3]));
        }
      }
    }
  } else if (warp == 4) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0 && n_mine > 0) {
      const uint32_t idesc = umma_idesc_bf16(128, u.Nmma) | (1u << 15) | (1u << 16);      // A and B MN-major
      bool ok = true;
      for (long long it = 0; it < n_mine && ok; ++it) {
        const int s = (int)(it % kStages);
        const uint32_t use = (uint32_t)(it / kStages);
        if (!mbar_wait(&sm.full[s], use & 1, status, 802)) { ok = false; break; }
        fence_proxy_async_smem();                 // cp.async wrote through the generic proxy; the tensor core reads through the async one
        tc_fence_after_sync();
        const uint32_t a0 = smem_u32(sm.a[s]), b0 = smem_u32(sm.b[s]);
#pragma unroll
        for (int kk = 0; kk < kRows / 16; ++kk) {
          const uint64_t bd = mn_desc(b0 + kk * 256);
          for (int h = 0; h < halves; ++h)
            umma_bf16(tmem + (uint32_t)h * 256u, mn_desc(a0 + h * (16 * kRows * 16) + kk * 256), bd, idesc, (it > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&sm.empty[s]);
      }
      if (ok) umma_commit(&sm.acc_full);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 5) { tc_fence_after_sync(); tmem_dealloc(tmem, 512); }
}

// alpha_linear.weight gradient: out[c] = sum_r w[r * w_stride] * X[r, c]  (d_sigma^T h7; a [1 x 256] "GEMM"), bf16 X
__global__ void __launch_bounds__(256) pgn_weighted_colsum_kernel(const __nv_bfloat16* __restrict__ X, long long m, const float* __restrict__ w,
                                                                  int w_stride, float* __restrict__ out) {
  // thread t owns columns 8 * (t & 31) .. +7 of rows (t >> 5) + 8 i: a warp reads one 512-byte row per instruction
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  float acc[8] = {};
  for (long long r = (long long)blockIdx.x * 8 + rl; r < m; r += (long long)gridDim.x * 8) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(X + r * 256) + cg);
    const float wr = __ldg(w + r * w_stride);
    const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      acc[2 * e] = fmaf(wr, __uint_as_float(q[e] << 16), acc[2 * e]);
      acc[2 * e + 1] = fmaf(wr, __uint_as_float(q[e] & 0xffff0000u), acc[2 * e + 1]);
    }
  }
  __shared__ float red[8][256];
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rl][cg * 8 + e] = acc[e];
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
  atomicAdd(out + threadIdx.x, s);
}

// The feature_linear / views_linears.0 block that never exists at batch size (feature_linear has no activation,
// core/networks/nerf.py:125-128): with T = dG^T h7 [128,256] and bv = column sums of dG [128]
//   g_views[:, :256]  = T W_f^T + bv b_f^T      g_feature.weight = W_v[:, :256]^T T      g_feature.bias = W_v[:, :256]^T bv
// fp32 CUDA cores, 33 MFLOP.  grid = 128 + 256 blocks of 256 threads.
__global__ void __launch_bounds__(256) pgn_view_fold_grads_kernel(const float* __restrict__ T, const float* __restrict__ bv,
                                                                  const float* __restrict__ w_f, const float* __restrict__ b_f,
                                                                  const float* __restrict__ w_v, float* __restrict__ g_views,
                                                                  float* __restrict__ g_feat_w, float* __restrict__ g_feat_b) {
  const int t = threadIdx.x;
  if (blockIdx.x < 128) {
    // g_views[n][c] = sum_k T[n][k] W_f[c][k] + bv[n] b_f[c]; this block: row n, thread: column c
    const int n = blockIdx.x;
    __shared__ float trow[256];
    trow[t] = T[n * 256 + t];
    __syncthreads();
    float acc = 0.f;
    const float4* wf = reinterpret_cast<const float4*>(w_f + (size_t)t * 256);
#pragma unroll 8
    for (int k4 = 0; k4 < 64; ++k4) {
      const float4 w = __ldg(wf + k4);
      acc = fmaf(trow[4 * k4], w.x, acc); acc = fmaf(trow[4 * k4 + 1], w.y, acc);
      acc = fmaf(trow[4 * k4 + 2], w.z, acc); acc = fmaf(trow[4 * k4 + 3], w.w, acc);
    }
    g_views[(size_t)n * 904 + t] = fmaf(bv[n], b_f[t], acc);
  } else {
    // g_feature.weight[j][c] = sum_n W_v[n][j] T[n][c]; this block: row j, thread: column c
    const int j = blockIdx.x - 128;
    __shared__ float wcol[128];
    if (t < 128) wcol[t] = w_v[(size_t)t * 904 + j];
    __syncthreads();
    float acc = 0.f;
#pragma unroll 8
    for (int n = 0; n < 128; ++n) acc = fmaf(wcol[n], __ldg(T + n * 256 + t), acc);
    g_feat_w[(size_t)j * 256 + t] = acc;
    if (t == 0) {
      float b = 0.f;
      for (int n = 0; n < 128; ++n) b = fmaf(wcol[n], bv[n], b);
      g_feat_b[j] = b;
    }
  }
}

}  // namespace

// Offsets (floats) of the 12 weight gradients inside the flat buffer: include/posegen_b200.h linear order, nn.Linear layouts
static const int kW_out[12] = {256, 256, 256, 256, 256, 256, 256, 256, 1, 256, 128, 3};
static const int kW_in[12] = {432, 256, 256, 256, 256, 688, 256, 256, 256, 256, 904, 128};

size_t pgn_wgrad_flat_floats() {
  size_t t = 0;
  for (int i = 0; i < 12; ++i) t += (size_t)kW_out[i] * kW_in[i];
  return t;
}

cudaError_t pgn_launch_weight_grads(const void* dz_, const void* dG_, const void* act_, long long dump_rows, const void* enc_,
                                    long long m, const float* d_raw, const float* bias_v, const float* w_f, const float* b_f,
                                    const float* w_v, float* flat, float* feat_bias, float* tm_scratch, int* status, int num_sms,
                                    cudaStream_t stream) {
  const __nv_bfloat16* dz = reinterpret_cast<const __nv_bfloat16*>(dz_);
  const __nv_bfloat16* dG = reinterpret_cast<const __nv_bfloat16*>(dG_);
  const __nv_bfloat16* act = reinterpret_cast<const __nv_bfloat16*>(act_);
  const __nv_bfloat16* enc = reinterpret_cast<const __nv_bfloat16*>(enc_);
  size_t off[12], o = 0;
  for (int i = 0; i < 12; ++i) { off[i] = o; o += (size_t)kW_out[i] * kW_in[i]; }
  cudaError_t e = cudaMemsetAsync(flat, 0, o * sizeof(float), stream);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(tm_scratch, 0, 128 * 256 * sizeof(float), stream);
  if (e != cudaSuccess || m == 0) return e;
  auto H = [&](int l) { return act + (size_t)l * dump_rows * 256; };        // activation dump: layers 0-7 [dump_rows,256] each
  auto DZ = [&](int l) { return dz + (size_t)l * m * 256; };
  Params p;
  int n = 0;
  auto add = [&](const __nv_bfloat16* A, int lda, int Ma, const __nv_bfloat16* B, int ldb, int Nb, float* out, int ld_out) {
    Unit& u = p.u[n++];
    u.A = A; u.B = B; u.out = out; u.lda = lda; u.ldb = ldb; u.ld_out = ld_out; u.Ma = Ma; u.Nb = Nb; u.Nmma = (Nb + 15) / 16 * 16;
    u.cta0 = 0; u.ncta = 0;
  };
  // pts_linears.0: dZ_0^T x_p (432 = 256 + 176 columns)
  add(DZ(0), 256, 256, enc, 1080, 256, flat + off[0], 432);
  add(DZ(0), 256, 256, enc + 256, 1080, 176, flat + off[0] + 256, 432);
  for (int l = 1; l < 8; ++l) {
    if (l == 5) {       // pts_linears.5 reads [x_p | h4] (skip connection, nerf.py:100-101)
      add(DZ(5), 256, 256, enc, 1080, 256, flat + off[5], 688);
      add(DZ(5), 256, 256, enc + 256, 1080, 176, flat + off[5] + 256, 688);
      add(DZ(5), 256, 256, H(4), 256, 256, flat + off[5] + 432, 688);
    } else {
      add(DZ(l), 256, 256, H(l - 1), 256, 256, flat + off[l], 256);
    }
  }
  // views_linears.0: dG^T [h7 -> T scratch | d_emb (648 = 256 + 256 + 136 columns)]
  add(dG, 128, 128, H(7), 256, 256, tm_scratch, 256);
  add(dG, 128, 128, enc + 432, 1080, 256, flat + off[10] + 256, 904);
  add(dG, 128, 128, enc + 688, 1080, 256, flat + off[10] + 512, 904);
  add(dG, 128, 128, enc + 944, 1080, 136, flat + off[10] + 768, 904);
  p.n_units = n;
  p.m = m;
  // CTAs per unit proportional to the bytes a stage streams (Ma + Nmma columns): every unit then sweeps the rows at the
  // same rate and shared operands are served by L2.  Largest-remainder rounding, at least one CTA per unit.
  const long long n_stages = (m + kRows - 1) / kRows;
  int grid = num_sms;
  if (grid < n) grid = n;
  double tot = 0;
  for (int i = 0; i < n; ++i) tot += p.u[i].Ma + p.u[i].Nmma;
  int used = 0;
  double frac[kMaxUnits];
  for (int i = 0; i < n; ++i) {
    const double share = (p.u[i].Ma + p.u[i].Nmma) / tot * grid;
    int c = (int)share;
    if (c < 1) c = 1;
    frac[i] = share - c;
    p.u[i].ncta = c;
    used += c;
  }
  while (used < grid) {
    int best = 0;
    for (int i = 1; i < n; ++i) if (frac[i] > frac[best]) best = i;
    p.u[best].ncta++; frac[best] -= 1.0; ++used;
  }
  int c0 = 0;
  for (int i = 0; i < n; ++i) {
    if ((long long)p.u[i].ncta > n_stages) p.u[i].ncta = (int)n_stages;
    p.u[i].cta0 = c0; c0 += p.u[i].ncta;
  }
  const size_t smem = sizeof(Smem) + 1024;
  static PgnPerDeviceOnce configured;
  if (configured.need()) {
    e = cudaFuncSetAttribute(pgn_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.set();
  }
  pgn_wgrad_kernel<<<c0, kThreads, smem, stream>>>(p, status);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // alpha_linear.weight = d_sigma^T h7
  pgn_weighted_colsum_kernel<<<num_sms * 2, 256, 0, stream>>>(H(7), m, d_raw + 3, 4, flat + off[8]);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  pgn_view_fold_grads_kernel<<<128 + 256, 256, 0, stream>>>(tm_scratch, bias_v, w_f, b_f, w_v, flat + off[10], flat + off[9], feat_bias);
  return cudaGetLastError();
}

// Generic entry for tests: out[Ma, Nb] (fp32, ld_out) += A[m, :Ma]^T B[m, :Nb]
cudaError_t pgn_launch_wgrad_single(const void* A, int lda, int Ma, const void* B, int ldb, int Nb, long long m, float* out, int ld_out,
                                    int n_ctas, int* status, cudaStream_t stream) {
  if (m == 0) return cudaSuccess;
  Params p;
  Unit& u = p.u[0];
  u.A = reinterpret_cast<const __nv_bfloat16*>(A); u.B = reinterpret_cast<const __nv_bfloat16*>(B); u.out = out;
  u.lda = lda; u.ldb = ldb; u.ld_out = ld_out; u.Ma = Ma; u.Nb = Nb; u.Nmma = (Nb + 15) / 16 * 16;
  const long long n_stages = (m + kRows - 1) / kRows;
  u.cta0 = 0; u.ncta = (int)(n_ctas < n_stages ? n_ctas : n_stages);
  p.n_units = 1; p.m = m;
  const size_t smem = sizeof(Smem) + 1024;
  static PgnPerDeviceOnce configured;
  if (configured.need()) {
    cudaError_t e = cudaFuncSetAttribute(pgn_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.set();
  }
  pgn_wgrad_kernel<<<u.ncta, kThreads, smem, stream>>>(p, status);
  return cudaGetLastError();
}
