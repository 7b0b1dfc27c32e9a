// dL/d(network input) of one A-NeRF MLP on tcgen05 (the pose gradient of BASELINE.json configs[4]: what autograd computes as
// grad_input of pts_linears.0 / pts_linears.5 / views_linears.0, core/networks/nerf.py:94-131 backwards):
//
//     g_xp [m,432] = dZ_5 [m,256] W_5[:, :432] + dZ_0 [m,256] W_0          (v-embed | r channels of the network input)
//     g_d  [m,648] = dG   [m,128] W_v[:, 256:904]                          (view-embed channels)
//
// Both are plain [rows x K] x [K x N] products with K <= 512 and the rows outermost, so a CTA keeps a 128-row tile of the
// delta matrices in shared memory (K-major SWIZZLE_128B boxes straight from the row-major deltas) and streams the bf16
// weights - nn.Linear layout [out = K][in = N], i.e. MN-major B operands, no transpose anywhere - through a ring of 64-row
// K stages of 32 KB.  The 1,080 output columns of a tile are produced as five jobs of <= 256 columns that ping-pong
// between two TMEM accumulators, so the epilogue of one job (TMEM -> bf16 -> HBM) runs under the MMAs of the next.  The
// jobs run as two passes over the tiles (two launches: g_xp, then g_d; PassCfg below) so that each pass has room for a
// weight ring that covers the L2 latency.  All operand movement is TMA (cp.async.bulk.tensor, SASS UTMALDG).
//
// Algorithmic bytes per row: 1,024 + 256 read, 2,160 written = 3.4 KB against 0.61 MFLOP (tensor-bound: ~180 FLOP/B).
// Roles (192 threads, one persistent CTA per SM): warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue.
#include <cuda.h>
#include <cuda_bf16.h>
#include "pgn_common.cuh"
#include "pgn_kernels.h"
#include "pgn_umma.cuh"
#include "pgn_tma.h"

using namespace pgn;

namespace {

constexpr int kTile = 128;                      // rows per tile (UMMA M)
constexpr int kABox = kTile * 128;              // A box: 64 K-columns (128 B) x 128 rows = 16 KB
constexpr int kKStage = 64;                     // K rows of B per ring stage
constexpr int kBBox = kKStage * 128;            // B box: 64 N-columns x 64 K-rows = 8 KB
constexpr int kBStageBytes = 4 * kBBox;         // up to 256 N-columns
constexpr int kThreads = 192;
constexpr int kMaxBStages = 5;
// Two passes over the tiles (two launches), so that each has the shared memory for a weight ring that covers the L2
// latency: with all five jobs in one kernel the three delta matrices (160 KB) left room for 2 x 32 KB of weights in flight -
// 43 B/clk against the 62 B/clk the MMAs consume - and the issuer waited ~0.8 k cycles per 64-row stage
// (40 k cycles per tile, tensor pipe 25 %).
//   pass 0: g_xp (jobs 0, 1): A = dZ_5 | dZ_0 (128 KB, one buffer), 3 weight stages
//   pass 1: g_d  (jobs 2-4):  A = dG (32 KB, two buffers: the next tile loads under this tile's MMAs), 5 weight stages
template <int PASS> struct PassCfg;
template <> struct PassCfg<0> { static constexpr int kABoxes = 8, kABufs = 1, kBStages = 3, kJob0 = 0, kNJobs = 2; };
template <> struct PassCfg<1> { static constexpr int kABoxes = 2, kABufs = 2, kBStages = 5, kJob0 = 2, kNJobs = 3; };

struct Params {
  CUtensorMap a_dz;      // [8][m][256] bf16 deltas of the trunk (layers 0 and 5 are read)
  CUtensorMap a_dg;      // [1][m][128] bf16 view-layer delta
  CUtensorMap b_w5, b_w0, b_wv;   // bf16 weights [K][N]: W_5[:, :432] (ld 432), W_0 (ld 432), W_v[:, 256:904] (ld 648)
  __nv_bfloat16* g_xp;   // [m,432]  (tile_blocked: [ceil(m / 128)][54][128][8])
  __nv_bfloat16* g_d;    // [m,648]  (tile_blocked: [ceil(m / 128)][81][128][8])
  long long m;
  int tile_blocked;
};

template <int PASS>
struct __align__(1024) Smem {
  uint8_t a[PassCfg<PASS>::kABufs][PassCfg<PASS>::kABoxes * kABox];
  uint8_t b[PassCfg<PASS>::kBStages][kBStageBytes];
  uint64_t a_full[2], a_empty[2], b_full[kMaxBStages], b_empty[kMaxBStages], acc_full[2], acc_empty[2];
  uint32_t tmem_slot;
};
static_assert(sizeof(Smem<0>) + 1024 <= 232448 && sizeof(Smem<1>) + 1024 <= 232448, "shared memory budget exceeded");

// job j: output columns [col0, col0 + nb) of g_xp (j < 2) or g_d; UMMA N = nmma
__device__ __forceinline__ void job_of(int j, int& col0, int& nb, int& nmma) {
  col0 = j == 0 ? 0 : (j == 1 ? 256 : (j - 2) * 256);
  nb = j == 0 ? 256 : (j == 1 ? 176 : (j == 4 ? 136 : 256));
  nmma = j == 4 ? 144 : nb;
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
// A: K-major SWIZZLE_128B (rows of 128 B = 64 K elements, 8-row groups 1 KB apart); B: MN-major SWIZZLE_128B (64 N-columns
// per box, boxes 8 KB apart, 8 K-rows = 1 KB)
__device__ __forceinline__ uint64_t a_desc(uint32_t saddr) { return umma_smem_desc(saddr, 16, 1024) | (2ull << 61); }
__device__ __forceinline__ uint64_t b_desc(uint32_t saddr) { return umma_smem_desc(saddr, kBBox, 1024) | (2ull << 61); }

template <int PASS>
__global__ void __launch_bounds__(kThreads, 1) pgn_input_grads_kernel(const __grid_constant__ Params p, int* __restrict__ status_g) {
  using Cfg = PassCfg<PASS>;
  constexpr int kBStages = Cfg::kBStages, kABufs = Cfg::kABufs;
  constexpr uint32_t kABytes = (uint32_t)Cfg::kABoxes * kABox;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem<PASS>& sm = *reinterpret_cast<Smem<PASS>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  volatile int* status = status_g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long n_tiles = (p.m + kTile - 1) / kTile;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.a_full[i], 1); mbar_init(&sm.a_empty[i], 1); }
    for (int s = 0; s < kBStages; ++s) { mbar_init(&sm.b_full[s], 1); mbar_init(&sm.b_empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&sm.acc_full[i], 1); mbar_init(&sm.acc_empty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&sm.tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_slot;
  const uint32_t a_s0 = smem_u32(sm.a[0]), b0 = smem_u32(sm.b[0]);
  const uint32_t a_full0 = smem_u32(&sm.a_full[0]), a_empty0 = smem_u32(&sm.a_empty[0]);
  const uint32_t b_full0 = smem_u32(&sm.b_full[0]), b_empty0 = smem_u32(&sm.b_empty[0]);
  const uint32_t acc_full0 = smem_u32(&sm.acc_full[0]), acc_empty0 = smem_u32(&sm.acc_empty[0]);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t ab = 0, a_ph = 1, bs = 0, b_ph = 1;      // "empty" barriers start released
      for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int row = (int)(t * kTile);
        if (!mbar_wait_s(a_empty0 + ab * 8, a_ph, status, 901)) return;
        const uint32_t abar = a_full0 + ab * 8, adst = a_s0 + ab * kABytes;
        mbar_arrive_expect_tx_s(abar, kABytes);
        if (PASS == 0) {
          for (int i = 0; i < 4; ++i) tma_load_3d(adst + i * kABox, &p.a_dz, i * 64, row, 5, abar);
          for (int i = 0; i < 4; ++i) tma_load_3d(adst + (4 + i) * kABox, &p.a_dz, i * 64, row, 0, abar);
        } else {
          for (int i = 0; i < 2; ++i) tma_load_3d(adst + i * kABox, &p.a_dg, i * 64, row, 0, abar);
        }
        if (++ab == (uint32_t)kABufs) { ab = 0; a_ph ^= 1; }
        for (int j = Cfg::kJob0; j < Cfg::kJob0 + Cfg::kNJobs; ++j) {
          int col0, nb, nmma;
          job_of(j, col0, nb, nmma);
          const int boxes = (nmma + 63) / 64;
          const int n_src = j < 2 ? 2 : 1, k_stages = j < 2 ? 4 : 2;
          for (int src = 0; src < n_src; ++src) {
            const CUtensorMap* mb = j < 2 ? (src == 0 ? &p.b_w5 : &p.b_w0) : &p.b_wv;
            for (int ks = 0; ks < k_stages; ++ks) {
              if (!mbar_wait_s(b_empty0 + bs * 8, b_ph, status, 902)) return;
              const uint32_t bar = b_full0 + bs * 8;
              mbar_arrive_expect_tx_s(bar, (uint32_t)boxes * kBBox);
              for (int i = 0; i < boxes; ++i) tma_load_3d(b0 + bs * kBStageBytes + i * kBBox, mb, col0 + i * 64, ks * kKStage, 0, bar);
              if (++bs == (uint32_t)kBStages) { bs = 0; b_ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      uint32_t ab = 0, a_ph = 0, bs = 0, b_ph = 0, acc_ph[2] = {1, 1}, jc = 0;          // accumulators start free
      for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        if (!mbar_wait_s(a_full0 + ab * 8, a_ph, status, 903)) return;
        tc_fence_after_sync();
        const uint32_t abuf = a_s0 + ab * kABytes;
        for (int j = Cfg::kJob0; j < Cfg::kJob0 + Cfg::kNJobs; ++j, ++jc) {
          int col0, nb, nmma;
          job_of(j, col0, nb, nmma);
          const int buf = jc & 1;
          if (!mbar_wait_s(acc_empty0 + buf * 8, acc_ph[buf], status, 904)) return;
          acc_ph[buf] ^= 1;
          tc_fence_after_sync();
          const uint32_t idesc = umma_idesc_bf16(kTile, nmma) | (1u << 16);      // A K-major, B MN-major
          const int n_src = j < 2 ? 2 : 1, k_stages = j < 2 ? 4 : 2;
          bool first = true;
          for (int src = 0; src < n_src; ++src) {
            const uint32_t abase = abuf + (PASS == 0 ? src * 4 * kABox : 0);      // pass 0: dZ_5 then dZ_0
            for (int ks = 0; ks < k_stages; ++ks) {
              if (!mbar_wait_s(b_full0 + bs * 8, b_ph, status, 905)) return;
              tc_fence_after_sync();
              const uint32_t bb = b0 + bs * kBStageBytes;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {          // 16 K per MMA: 32 B inside A's 128-byte rows, 2 KB inside a B box
                umma_bf16(tmem + (uint32_t)buf * 256u, a_desc(abase + ks * kABox + kk * 32), b_desc(bb + kk * 2048), idesc, first ? 0u : 1u);
                first = false;
              }
              umma_commit(&sm.b_empty[bs]);
              if (++bs == (uint32_t)kBStages) { bs = 0; b_ph ^= 1; }
            }
          }
          umma_commit(&sm.acc_full[buf]);
        }
        umma_commit(&sm.a_empty[ab]);
        if (++ab == (uint32_t)kABufs) { ab = 0; a_ph ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM -> bf16 -> HBM
    const int q = warp & 3;
    const int row = q * 32 + lane;
    uint32_t acc_ph[2] = {0, 0}, jc = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const long long grow = t * kTile + row;
      for (int j = Cfg::kJob0; j < Cfg::kJob0 + Cfg::kNJobs; ++j, ++jc) {
        int col0, nb, nmma;
        job_of(j, col0, nb, nmma);
        const int buf = jc & 1;
        if (!mbar_wait_s(acc_full0 + buf * 8, acc_ph[buf], status, 906)) return;
        acc_ph[buf] ^= 1;
        tc_fence_after_sync();
        // row-major: this thread's row; tile-blocked [row / 128][column / 8][128][8]: this row's 16 bytes of the job's first
        // 8-column chunk (the next chunk is 128 rows x 16 B = 1,024 elements further) - a warp's store is then 512
        // contiguous bytes instead of 16 bytes in each of 32 rows (32 lines per instruction)
        __nv_bfloat16* orow = p.tile_blocked
            ? (j < 2 ? p.g_xp + (size_t)t * (432 * kTile) : p.g_d + (size_t)t * (648 * kTile)) + (size_t)(col0 >> 3) * (kTile * 8) + row * 8
            : (j < 2 ? p.g_xp + grow * 432 : p.g_d + grow * 648) + col0;
        const int cstep = p.tile_blocked ? kTile * 8 : 8;      // elements from one 8-column chunk to the next
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * 256u;
        uint32_t v[2][16];
        tmem_ld_32x16(taddr, v[0]);
        for (int c0 = 0, b = 0; c0 < nmma; c0 += 16, ++b) {
          tmem_ld_wait();
          if (c0 + 16 < nmma) tmem_ld_32x16(taddr + (uint32_t)(c0 + 16), v[(b + 1) & 1]);
          const uint32_t* vb = v[b & 1];
          if (grow < p.m || p.tile_blocked) {           // (a tile-blocked buffer holds whole tiles: rows beyond m are zeros)
#pragma unroll
            for (int h = 0; h < 2; ++h)
              if (c0 + 8 * h < nb) {
                uint4 o;
                o.x = pack_bf16x2(__uint_as_float(vb[8 * h]), __uint_as_float(vb[8 * h + 1]));
                o.y = pack_bf16x2(__uint_as_float(vb[8 * h + 2]), __uint_as_float(vb[8 * h + 3]));
                o.z = pack_bf16x2(__uint_as_float(vb[8 * h + 4]), __uint_as_float(vb[8 * h + 5]));
                o.w = pack_bf16x2(__uint_as_float(vb[8 * h + 6]), __uint_as_float(vb[8 * h + 7]));
                *reinterpret_cast<uint4*>(orow + (size_t)((c0 >> 3) + h) * cstep) = o;
              }
          }
        }
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(acc_empty0 + buf * 8) : "memory");
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tmem, 512); }
}

// bf16 copies of the three weight blocks in one buffer: [W_5[:, :432] (256x432) | W_0 (256x432) | W_v[:, 256:904] (128x648)]
__global__ void pgn_pack_input_grad_weights_kernel(const float* __restrict__ w5, const float* __restrict__ w0, const float* __restrict__ wv,
                                                   int view_ld, __nv_bfloat16* __restrict__ out) {
  const int n1 = 256 * 432, n3 = 128 * 648;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n1 + n3; i += gridDim.x * blockDim.x) {
    float v;
    if (i < n1) v = w5[(size_t)(i / 432) * 688 + i % 432];
    else if (i < 2 * n1) v = w0[i - n1];
    else { const int r = i - 2 * n1; v = wv[(size_t)(r / 648) * view_ld + 256 + r % 648]; }
    out[i] = __float2bfloat16_rn(v);
  }
}

cudaError_t make_map(CUtensorMap* map, const void* base, long long cols, long long ld, long long rows, long long layers,
                     long long layer_stride_elems, int box_rows) {
  return pgn_make_map_bf16(map, base, cols, ld, rows, layers, layer_stride_elems, box_rows);
}

}  // namespace

size_t pgn_input_grad_weight_elems() { return (size_t)2 * 256 * 432 + (size_t)128 * 648; }

cudaError_t pgn_launch_pack_input_grad_weights(const float* w5, const float* w0, const float* wv, int view_ld, __nv_bfloat16* out,
                                               cudaStream_t stream) {
  pgn_pack_input_grad_weights_kernel<<<148, 256, 0, stream>>>(w5, w0, wv, view_ld, out);
  return cudaGetLastError();
}

cudaError_t pgn_launch_input_grads(const void* dz, const void* dG, long long m, const __nv_bfloat16* wpack, void* g_xp, void* g_d, int tile_blocked,
                                   int* status, int num_sms, cudaStream_t stream) {
  if (m == 0) return cudaSuccess;
  Params p;
  cudaError_t e;
  if ((e = make_map(&p.a_dz, dz, 256, 256, m, 8, m * 256, kTile)) != cudaSuccess) return e;
  if ((e = make_map(&p.a_dg, dG, 128, 128, m, 1, 0, kTile)) != cudaSuccess) return e;
  if ((e = make_map(&p.b_w5, wpack, 432, 432, 256, 1, 0, kKStage)) != cudaSuccess) return e;
  if ((e = make_map(&p.b_w0, wpack + 256 * 432, 432, 432, 256, 1, 0, kKStage)) != cudaSuccess) return e;
  if ((e = make_map(&p.b_wv, wpack + 2 * 256 * 432, 648, 648, 128, 1, 0, kKStage)) != cudaSuccess) return e;
  p.g_xp = reinterpret_cast<__nv_bfloat16*>(g_xp);
  p.g_d = reinterpret_cast<__nv_bfloat16*>(g_d);
  p.m = m;
  p.tile_blocked = tile_blocked;
  static PgnPerDeviceOnce configured;
  if (configured.need()) {
    e = cudaFuncSetAttribute(pgn_input_grads_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(Smem<0>) + 1024));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(pgn_input_grads_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(Smem<1>) + 1024));
    if (e != cudaSuccess) return e;
    configured.set();
  }
  const long long n_tiles = (m + kTile - 1) / kTile;
  const unsigned grid = (unsigned)(n_tiles < num_sms ? n_tiles : num_sms);
  pgn_input_grads_kernel<0><<<grid, kThreads, sizeof(Smem<0>) + 1024, stream>>>(p, status);
  pgn_input_grads_kernel<1><<<grid, kThreads, sizeof(Smem<1>) + 1024, stream>>>(p, status);
  return cudaGetLastError();
}
