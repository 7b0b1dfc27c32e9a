// Weight gradients of the A-NeRF MLP backward (training step, BASELINE.json configs[3]) as ONE split-K tcgen05 kernel.
//
// What autograd does for every nn.Linear of core/networks/nerf.py:94-148 is dW_l = dZ_l^T h_{l-1}: a [256 x K_in] GEMM
// whose contraction runs over ALL samples of the batch (2-2.5 x 10^5 rows per pass).  Both operands already sit in HBM
// as row-major bf16 matrices with the contraction dimension (rows) outermost: the deltas dZ_l [rows,256] written by
// pgn_delta_chain, the activations h_l [rows,256] dumped by the training forward, the regenerated network input
// [rows,1080].  In UMMA terms both are "MN-major" operands (M/N contiguous, K strided), so no transpose is ever
// materialised:
//
//   * a stage is 32 rows of A (<= 256 columns) and of B (<= 256 columns), moved by the TMA engine as boxes of
//     64 columns x 32 rows (cp.async.bulk.tensor.3d, SWIZZLE_128B; SASS UTMALDG): a box lands as 32 rows of 128 bytes
//     with the 16-byte chunks XOR-swizzled by the row - exactly the canonical MN-major SWIZZLE_128B UMMA operand
//     (LBO = 4 KB to the next 64 columns, SBO = 1 KB to the next 8 rows).  One elected thread issues up to 8 boxes per
//     stage; out-of-range rows / columns are zero-filled by the tensor map.  (A first version staged the operands with
//     per-thread 16-byte cp.async into the SWIZZLE_NONE image: correct, but the L1's outstanding-request tracking capped
//     it at ~12 B/clk/SM, 2.5 TB/s over the chip; bulk tensor copies do not go through it.)
//     The activations h_l come TILE-BLOCKED from the training forward ([row / 128][column / 8][128][8]: the image its
//     epilogue threads hold, written with fully coalesced 16-byte stores): a 128-row tile of a layer is 64 contiguous
//     KB, moved by ONE 1-D cp.async.bulk into a ring of two tiles and used in place as a SWIZZLE_NONE MN-major operand
//     (LBO = 128 B, SBO = 2 KB) by the tile's four 32-row A stages.  Each operand carries its own descriptor layout.
//   * one elected thread issues tcgen05.mma kind::f16 (M = 128, N <= 256, K = 16) with a_major = b_major = MN:
//     per stage 2 K-steps x (1 or 2) M-halves, fp32 accumulators in TMEM (2 x 256 columns = a 256 x 256 output tile).
//   * split-K: the units (one output tile each: an (A matrix, B column range) pair) get a number of CTAs proportional
//     to the bytes they stream; CTA k of a unit takes stages k, k + n, k + 2n, ... so that every unit sweeps the rows at
//     the same rate and operands shared by several units (dZ_5, dG, x_p are each read by 2-4 units) come from the 126 MB
//     L2 instead of HBM a second time (ncu: DRAM bytes = the algorithmic bytes).  At the end every CTA adds its partial
//     tile to the fp32 gradient with red.global.add.v4.f32 (the gradient buffer is zeroed by the launch wrapper).
//
// Roofline: HBM-bound.  Algorithmic bytes per row and pass = 8 x 512 (dZ) + 256 (dG) + 8 x 512 (h) + 2,160 (input)
// = 10.6 KB against 1.72 MFLOP (162 FLOP/B; the machine balance is ~210), i.e. ~0.40 ms per 245,760-row pass at the
// measured 6.5 TB/s; the tensor pipe is <= 45 % busy by construction.
//
// Roles (192 threads): warp 0 = TMA producer (one lane) + TMEM allocation, warp 1 = MMA issuer (one lane), warps 2-5 =
// TMEM -> red.add epilogue.  6 stages x 32 KB in flight per SM.
#include <cuda.h>
#include <cuda_bf16.h>
#include "pgn_common.cuh"
#include "pgn_kernels.h"
#include "pgn_umma.cuh"
#include "pgn_tma.h"

using namespace pgn;

// the producer and the issuer are single latency-critical lanes: they spin on test_wait instead of parking in try_wait
// (wake-up latency of a parked warp; 500 -> 489 us per pass)
#define PGN_LAT_WAIT mbar_spin_s
namespace {

constexpr int kRows = 32;                      // rows (K of the GEMM) per stage
constexpr int kStages = 6;
constexpr int kBoxCols = 64;                   // 128 bytes of bf16: the SWIZZLE_128B span
constexpr int kBoxBytes = kRows * 128;         // one TMA box: 32 rows x 128 B = 4 KB
constexpr int kOpBytes = 4 * kBoxBytes;        // one operand of a stage: up to 4 boxes (256 columns) = 16 KB
constexpr int kTileRows = 128;                 // work is dealt in 128-row tiles = 4 stages (= one tile of the tile-blocked dump)
constexpr int kTbTileBytes = 65536;            // one tile of a tile-blocked [rows,256] matrix: [32 runs][128 rows][8] bf16
constexpr int kTbTiles = 2;                    // tile-blocked B operand: ring of 2 whole tiles
constexpr int kThreads = 192;
constexpr int kMaxUnits = 16;
constexpr int kMaxMaps = 4;
// Soft lock-step of the sweep: operands shared by several units (dZ_5 by 3, dG by 4, x_p by 2 ...) are only served by L2
// if all units pass the same rows within the L2's reach.  The rows are cut into epochs of 2,048 (27 MB of operands); a
// CTA does not start loading epoch e before every CTA has issued its loads of epoch e - kEpochWindow (bounded wait: a
// CTA that is not co-resident can only delay the others by the timeout, never dead-lock them).
constexpr int kEpochRows = 2048;
constexpr int kEpochWindow = 2;

struct Unit {
  int a_map, a_layer, a_col;   // A operand: tensor map, its third coordinate (layer), first column
  int b_map, b_layer, b_col;   // B operand (tensor-map form)
  const uint8_t* b_tb;         // B operand, tile-blocked form (non-NULL): one layer of the training forward's activation dump,
                               // [tile = row / 128][column / 8][128 rows][8] bf16 - a tile is 64 contiguous KB
  float* out;                  // fp32 [Ma, ld_out] tile origin (row = A column, column = B column)
  int ld_out;
  int Ma;                      // 128 | 256
  int Nb;                      // valid B columns (multiple of 8)
  int Nmma;                    // UMMA N: Nb rounded up to 16 (columns beyond Nb are never stored)
  int cta0, ncta;              // CTAs [cta0, cta0 + ncta) work on this unit
};
struct __align__(64) Params {
  CUtensorMap maps[kMaxMaps];  // [layer][row][column] bf16 views of the operands, box = 64 columns x 32 rows
  Unit u[kMaxUnits];
  int n_units;
  long long m;                 // rows
  int* epoch_ctr;              // [ceil(m / kEpochRows)] zeroed per launch: CTAs that have issued every load of an epoch
  int n_ctas;                  // CTAs of the launch
};

struct __align__(1024) Smem {
  uint8_t a[kStages][kOpBytes];
  union {
    uint8_t b[kStages][kOpBytes];              // B through tensor-map boxes: same stages as A
    uint8_t bt[kTbTiles][kTbTileBytes];        // tile-blocked B: whole 128-row tiles, one 64 KB bulk copy each
  };
  uint64_t full[kStages], empty[kStages], bt_full[kTbTiles], bt_empty[kTbTiles], acc_full;
  uint32_t tmem_slot;
};
static_assert(sizeof(Smem) + 1024 <= 232448, "shared memory budget exceeded");

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
// MN-major SWIZZLE_128B descriptor of a stage operand: LBO = 4 KB (next 64 columns = next box), SBO = 1 KB (next 8 rows),
// layout type 2 (SWIZZLE_128B) in bits [61,64)
__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr) { return umma_smem_desc(saddr, kBoxBytes, 1024) | (2ull << 61); }
// MN-major SWIZZLE_NONE descriptor into a tile-blocked tile [column / 8][128 rows][8]: LBO = 128 B (next 8 rows = K),
// SBO = 2 KB (next 8 columns); the start address selects the first row
__device__ __forceinline__ uint64_t mn_desc_tb(uint32_t saddr) { return umma_smem_desc(saddr, 128, 2048); }

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
  const int ncta = u.ncta, Ma = u.Ma, Nb = u.Nb, Nmma = u.Nmma;
  // CTA k of a unit takes the 128-row tiles k, k + ncta, ...; a tile is 4 stages of 32 rows
  const long long n_tiles_total = (p.m + kTileRows - 1) / kTileRows;
  const long long n_mine = 4 * ((active && n_tiles_total > k) ? (n_tiles_total - k + ncta - 1) / ncta : 0);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int s = 0; s < kTbTiles; ++s) { mbar_init(&sm.bt_full[s], 1); mbar_init(&sm.bt_empty[s], 1); }
    mbar_init(&sm.acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc(&sm.tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_slot;
  const int halves = Ma / 128;
  const uint32_t full0 = smem_u32(&sm.full[0]), empty0 = smem_u32(&sm.empty[0]);
  const uint32_t a_s0 = smem_u32(sm.a[0]), b_s0 = smem_u32(sm.b[0]);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0 && n_mine > 0) {
      const CUtensorMap* ma = &p.maps[u.a_map];
      const CUtensorMap* mb = &p.maps[u.b_map];
      const int a_col = u.a_col, a_layer = u.a_layer, b_col = u.b_col, b_layer = u.b_layer;
      const uint8_t* b_tb = u.b_tb;
      const int a_boxes = Ma / kBoxCols, b_boxes = (Nmma + kBoxCols - 1) / kBoxCols;
      const uint32_t bytes = (uint32_t)(a_boxes + (b_tb ? 0 : b_boxes)) * kBoxBytes;
      const uint32_t bt_full0 = smem_u32(&sm.bt_full[0]), bt_empty0 = smem_u32(&sm.bt_empty[0]), bt_s0 = smem_u32(sm.bt[0]);
      int s = 0, bts = 0;
      uint32_t ph = 1, bt_ph = 1;                  // "empty" barriers start released
      int cur_e = 0;
      volatile int* ectr = p.epoch_ctr;
      const int n_all = p.n_ctas;
      for (long long it = 0; it < n_mine; ++it) {
        const long long tile = (long long)k + (it >> 2) * ncta;
        const int sub = (int)(it & 3);
        const int row = (int)(tile * kTileRows) + sub * kRows;
        const int e = row / kEpochRows;
        if (e != cur_e) {
          while (cur_e < e) { atomicAdd(p.epoch_ctr + cur_e, 1); ++cur_e; }
          if (e >= kEpochWindow) {
            const long long t0 = clock64();
            while (ectr[e - kEpochWindow] < n_all && clock64() - t0 < 400000) __nanosleep(200);
          }
        }
        if (b_tb && sub == 0) {                    // the tile's B operand: 64 contiguous KB, one bulk copy
          if (!PGN_LAT_WAIT(bt_empty0 + bts * 8, bt_ph, status, 804)) break;
          mbar_arrive_expect_tx_s(bt_full0 + bts * 8, (uint32_t)kTbTileBytes);
          bulk_g2s_s(bt_s0 + bts * kTbTileBytes, b_tb + (size_t)tile * kTbTileBytes, (uint32_t)kTbTileBytes, bt_full0 + bts * 8);
          if (++bts == kTbTiles) { bts = 0; bt_ph ^= 1; }
        }
        if (!PGN_LAT_WAIT(empty0 + s * 8, ph, status, 801)) break;
        const uint32_t bar = full0 + s * 8;
        mbar_arrive_expect_tx_s(bar, bytes);
        for (int i = 0; i < a_boxes; ++i) tma_load_3d(a_s0 + s * kOpBytes + i * kBoxBytes, ma, a_col + i * kBoxCols, row, a_layer, bar);
        if (!b_tb)
          for (int i = 0; i < b_boxes; ++i) tma_load_3d(b_s0 + s * kOpBytes + i * kBoxBytes, mb, b_col + i * kBoxCols, row, b_layer, bar);
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
      const int n_epochs = (int)((p.m + kEpochRows - 1) / kEpochRows);
      while (cur_e < n_epochs) { atomicAdd(p.epoch_ctr + cur_e, 1); ++cur_e; }
    }
    if (lane == 0 && n_mine == 0) {                // an idle CTA must not hold the others back
      const int n_epochs = (int)((p.m + kEpochRows - 1) / kEpochRows);
      for (int e = 0; e < n_epochs; ++e) atomicAdd(p.epoch_ctr + e, 1);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0 && n_mine > 0) {
      const uint32_t idesc = umma_idesc_bf16(128, Nmma) | (1u << 15) | (1u << 16);      // A and B MN-major
      const bool b_is_tb = u.b_tb != nullptr;
      const uint32_t bt_full0 = smem_u32(&sm.bt_full[0]), bt_s0 = smem_u32(sm.bt[0]);
      bool ok = true;
      int s = 0, bts = 0;
      uint32_t ph = 0, bt_ph = 0;
      for (long long it = 0; it < n_mine; ++it) {
        const int sub = (int)(it & 3);
        if (b_is_tb && sub == 0) {
          if (!PGN_LAT_WAIT(bt_full0 + bts * 8, bt_ph, status, 805)) { ok = false; break; }
        }
        if (!PGN_LAT_WAIT(full0 + s * 8, ph, status, 802)) { ok = false; break; }
        tc_fence_after_sync();
        const uint32_t a0 = a_s0 + s * kOpBytes, b0 = b_s0 + s * kOpBytes;
        const uint32_t bt0 = bt_s0 + bts * kTbTileBytes + sub * (kRows * 16);       // rows [32 sub, 32 sub + 32) of the tile
#pragma unroll
        for (int kk = 0; kk < kRows / 16; ++kk) {                       // 16 rows = 2 KB inside every box / 256 B inside a run
          const uint64_t bd = b_is_tb ? mn_desc_tb(bt0 + kk * 256) : mn_desc(b0 + kk * 2048);
          for (int h = 0; h < halves; ++h)                              // 128 A columns = 2 boxes
            umma_bf16(tmem + (uint32_t)h * 256u, mn_desc(a0 + h * (2 * kBoxBytes) + kk * 2048), bd, idesc, (it > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&sm.empty[s]);
        if (++s == kStages) { s = 0; ph ^= 1; }
        if (b_is_tb && sub == 3) {
          umma_commit(&sm.bt_empty[bts]);
          if (++bts == kTbTiles) { bts = 0; bt_ph ^= 1; }
        }
      }
      if (ok) umma_commit(&sm.acc_full);
    }
  } else {
    // ------------------------------------------------------------ epilogue: partial tile -> fp32 gradient (atomic adds)
    const int q = warp & 3;                                              // TMEM lane quarter this warp may read
    if (n_mine > 0 && mbar_wait(&sm.acc_full, 0, status, 803)) {
      tc_fence_after_sync();
      const int ld_out = u.ld_out;
      float* out0 = u.out;
      for (int h = 0; h < halves; ++h) {
        float* orow = out0 + (size_t)(h * 128 + q * 32 + lane) * ld_out;
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)h * 256u;
        uint32_t v[2][16];
        tmem_ld_32x16(taddr, v[0]);
        for (int c0 = 0, b = 0; c0 < Nmma; c0 += 16, ++b) {
          tmem_ld_wait();
          if (c0 + 16 < Nmma) tmem_ld_32x16(taddr + (uint32_t)(c0 + 16), v[(b + 1) & 1]);
          const uint32_t* vb = v[b & 1];
#pragma unroll
          for (int g = 0; g < 4; ++g)
            if (c0 + 4 * g < Nb)
              red_add_v4(orow + c0 + 4 * g, __uint_as_float(vb[4 * g]), __uint_as_float(vb[4 * g + 1]), __uint_as_float(vb[4 * g + 2]),
                         __uint_as_float(vb[4 * g + 3]));
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tmem, 512); }
}

// alpha_linear.weight gradient: out[c] = sum_r w[r * w_stride] * X[r, c]  (d_sigma^T h7; a [1 x 256] "GEMM") on the
// tile-blocked activation dump X = [tile][column / 8][128 rows][8] bf16: warp w of a block owns the 8-column run
// 8 * blockIdx.y + w; a warp load is 32 consecutive rows of its run = 512 contiguous bytes.
__global__ void __launch_bounds__(256) pgn_weighted_colsum_kernel(const uint8_t* __restrict__ X, long long m, const float* __restrict__ w,
                                                                  int w_stride, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, run = blockIdx.y * 8 + (threadIdx.x >> 5);
  float acc[8] = {};
  const long long n_groups = (m + 31) / 32;                       // groups of 32 rows
  for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
    const long long r = g * 32 + lane;
    if (r < m) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(X + (size_t)(r >> 7) * 65536u + (size_t)run * 2048u + (size_t)(r & 127) * 16u));
      const float wr = __ldg(w + r * w_stride);
      const uint32_t q[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[2 * e] = fmaf(wr, __uint_as_float(q[e] << 16), acc[2 * e]);
        acc[2 * e + 1] = fmaf(wr, __uint_as_float(q[e] & 0xffff0000u), acc[2 * e + 1]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], off);
  }
  if (lane == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) atomicAdd(out + run * 8 + e, acc[e]);
  }
}

// The feature_linear / views_linears.0 block that never exists at batch size (feature_linear has no activation,
// core/networks/nerf.py:125-128): with T = dG^T h7 [128,256] and bv = column sums of dG [128]
//   g_views[:, :256]  = T W_f^T + bv b_f^T      g_feature.weight = W_v[:, :256]^T T      g_feature.bias = W_v[:, :256]^T bv
// fp32 CUDA cores, 33 MFLOP.  grid = 128 + 256 blocks of 256 threads.
__global__ void __launch_bounds__(256) pgn_view_fold_grads_kernel(const float* __restrict__ T, const float* __restrict__ bv,
                                                                  const float* __restrict__ w_f, const float* __restrict__ b_f,
                                                                  const float* __restrict__ w_v, float* __restrict__ g_views,
                                                                  float* __restrict__ g_feat_w, float* __restrict__ g_feat_b, int view_ld) {
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
    g_views[(size_t)n * view_ld + t] = fmaf(bv[n], b_f[t], acc);
  } else {
    // g_feature.weight[j][c] = sum_n W_v[n][j] T[n][c]; this block: row j, thread: column c
    const int j = blockIdx.x - 128;
    __shared__ float wcol[128];
    if (t < 128) wcol[t] = w_v[(size_t)t * view_ld + j];
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

// ---------------------------------------------------------------------------------------------------------------
// host side: tensor maps (pgn_tma.h), boxes of 64 columns x 32 rows
static cudaError_t make_map(CUtensorMap* map, const void* base, long long cols, long long ld, long long rows, long long layers,
                            long long layer_stride_elems) {
  static_assert(kBoxCols == 64, "pgn_make_map_bf16 encodes 64-column boxes");
  return pgn_make_map_bf16(map, base, cols, ld, rows, layers, layer_stride_elems, kRows);
}

// Offsets (floats) of the 12 weight gradients inside the flat buffer: include/posegen_b200.h linear order, nn.Linear layouts
static const int kW_out[12] = {256, 256, 256, 256, 256, 256, 256, 256, 1, 256, 128, 3};
static const int kW_in[12] = {432, 256, 256, 256, 256, 688, 256, 256, 256, 256, 904, 128};

size_t pgn_wgrad_flat_floats_ld(int view_ld) {
  size_t t = 0;
  for (int i = 0; i < 12; ++i) t += (size_t)kW_out[i] * (i == 10 ? view_ld : kW_in[i]);
  return t;
}
size_t pgn_wgrad_flat_floats() { return pgn_wgrad_flat_floats_ld(904); }

static cudaError_t launch_units(Params& p, int n, long long m, int num_sms, int* status, int* epoch_ctr, cudaStream_t stream) {
  p.n_units = n;
  p.m = m;
  p.epoch_ctr = epoch_ctr;
  {
    const long long n_epochs = (m + kEpochRows - 1) / kEpochRows;
    if (n_epochs > PGN_WGRAD_MAX_EPOCHS) return cudaErrorInvalidValue;
    cudaError_t e0 = cudaMemsetAsync(epoch_ctr, 0, (size_t)n_epochs * sizeof(int), stream);
    if (e0 != cudaSuccess) return e0;
  }
  // CTAs per unit proportional to the bytes a stage streams (Ma + Nmma columns): every unit then sweeps the rows at the
  // same rate and shared operands are served by L2.  Largest-remainder rounding, at least one CTA per unit.
  const long long n_stages = (m + kTileRows - 1) / kTileRows;      // CTAs are dealt whole 128-row tiles
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
    cudaError_t e = cudaFuncSetAttribute(pgn_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.set();
  }
  p.n_ctas = c0;
  pgn_wgrad_kernel<<<c0, kThreads, smem, stream>>>(p, status);
  return cudaGetLastError();
}

cudaError_t pgn_launch_weight_grads(const void* dz_, const void* dG_, const void* act_, long long dump_rows, const void* enc_,
                                    long long m, const float* d_raw, const float* bias_v, const float* w_f, const float* b_f,
                                    const float* w_v, int view_ld, float* flat, float* feat_bias, float* tm_scratch, int* epoch_ctr,
                                    int* status, int num_sms, cudaStream_t stream) {
  size_t off[12], o = 0;
  for (int i = 0; i < 12; ++i) { off[i] = o; o += (size_t)kW_out[i] * (i == 10 ? view_ld : kW_in[i]); }
  cudaError_t e = cudaMemsetAsync(flat, 0, o * sizeof(float), stream);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(tm_scratch, 0, 128 * 256 * sizeof(float), stream);
  if (e != cudaSuccess || m == 0) return e;
  Params p;
  enum { kDz = 0, kDg = 1, kAct = 2, kEnc = 3 };
  if ((e = make_map(&p.maps[kDz], dz_, 256, 256, m, 8, m * 256)) != cudaSuccess) return e;            // dZ_l: [8][m][256]
  if ((e = make_map(&p.maps[kDg], dG_, 128, 128, m, 1, 0)) != cudaSuccess) return e;                   // dG:   [m][128]
  p.maps[kAct] = p.maps[kDz];                                                                          // h_l: tile-blocked, no tensor map (b_tb)
  if ((e = make_map(&p.maps[kEnc], enc_, 1080, 1080, m, 1, 0)) != cudaSuccess) return e;               // input: [m][1080]
  int n = 0;
  const uint8_t* act_b = reinterpret_cast<const uint8_t*>(act_);
  auto add = [&](int am, int al, int Ma, int bm, int bl, int bc, int Nb, float* out, int ld_out) {
    Unit& u = p.u[n++];
    u.a_map = am; u.a_layer = al; u.a_col = 0; u.b_map = bm; u.b_layer = bl; u.b_col = bc;
    u.b_tb = bm == kAct ? act_b + (size_t)bl * (size_t)dump_rows * 512u : nullptr;     // activations: tile-blocked dump of layer bl
    u.out = out; u.ld_out = ld_out; u.Ma = Ma; u.Nb = Nb; u.Nmma = (Nb + 15) / 16 * 16; u.cta0 = 0; u.ncta = 0;
  };
  // pts_linears.0: dZ_0^T x_p (432 = 256 + 176 columns)
  add(kDz, 0, 256, kEnc, 0, 0, 256, flat + off[0], 432);
  add(kDz, 0, 256, kEnc, 0, 256, 176, flat + off[0] + 256, 432);
  for (int l = 1; l < 8; ++l) {
    if (l == 5) {       // pts_linears.5 reads [x_p | h4] (skip connection, nerf.py:100-101)
      add(kDz, 5, 256, kEnc, 0, 0, 256, flat + off[5], 688);
      add(kDz, 5, 256, kEnc, 0, 256, 176, flat + off[5] + 256, 688);
      add(kDz, 5, 256, kAct, 4, 0, 256, flat + off[5] + 432, 688);
    } else {
      add(kDz, l, 256, kAct, l - 1, 0, 256, flat + off[l], 256);
    }
  }
  // views_linears.0: dG^T [h7 -> T scratch | d_emb (648 = 256 + 256 + 136 columns)]
  add(kDg, 0, 128, kAct, 7, 0, 256, tm_scratch, 256);
  add(kDg, 0, 128, kEnc, 0, 432, 256, flat + off[10] + 256, view_ld);
  add(kDg, 0, 128, kEnc, 0, 688, 256, flat + off[10] + 512, view_ld);
  add(kDg, 0, 128, kEnc, 0, 944, 136, flat + off[10] + 768, view_ld);      // (frame-code columns 904..919: pgn_framecode_backward)
  if ((e = launch_units(p, n, m, num_sms, status, epoch_ctr, stream)) != cudaSuccess) return e;
  // alpha_linear.weight = d_sigma^T h7
  pgn_weighted_colsum_kernel<<<dim3((unsigned)num_sms, 4), 256, 0, stream>>>(act_b + (size_t)7 * (size_t)dump_rows * 512u, m, d_raw + 3, 4, flat + off[8]);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  pgn_view_fold_grads_kernel<<<128 + 256, 256, 0, stream>>>(tm_scratch, bias_v, w_f, b_f, w_v, flat + off[10], flat + off[9], feat_bias, view_ld);
  return cudaGetLastError();
}

// Generic entry for tests: out[Ma, Nb] (fp32, ld_out) += A[m, :Ma]^T B[m, :Nb]
cudaError_t pgn_launch_wgrad_single(const void* A, int lda, int Ma, const void* B, int ldb, int Nb, long long m, float* out, int ld_out,
                                    int n_ctas, int b_tile_blocked, int* epoch_ctr, int* status, cudaStream_t stream) {
  if (m == 0) return cudaSuccess;
  Params p;
  cudaError_t e;
  if ((e = make_map(&p.maps[0], A, Ma, lda, m, 1, 0)) != cudaSuccess) return e;
  if (b_tile_blocked) p.maps[1] = p.maps[0];
  else if ((e = make_map(&p.maps[1], B, Nb, ldb, m, 1, 0)) != cudaSuccess) return e;
  p.maps[2] = p.maps[0]; p.maps[3] = p.maps[0];
  Unit& u = p.u[0];
  u.a_map = 0; u.a_layer = 0; u.a_col = 0; u.b_map = 1; u.b_layer = 0; u.b_col = 0;
  u.b_tb = b_tile_blocked ? reinterpret_cast<const uint8_t*>(B) : nullptr;
  u.out = out; u.ld_out = ld_out; u.Ma = Ma; u.Nb = Nb; u.Nmma = (Nb + 15) / 16 * 16; u.cta0 = 0; u.ncta = 0;
  return launch_units(p, 1, m, n_ctas, status, epoch_ctr, stream);
}
