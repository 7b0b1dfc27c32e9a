// Fused delta chain of the A-NeRF trunk backward (training step / GAN step) on tcgen05, CTA pairs.
//
// What autograd does per trunk layer (core/networks/nerf.py:94-102 backwards) is a GEMM dL/dh_{l-1} = dZ_l W_l followed
// by a ReLU mask; done layer by layer through HBM that is 5 passes over a [rows,256] matrix per layer.  Here a pair of
// CTAs keeps a 512-row block of deltas in shared memory for the whole chain:
//
//     dG [rows,128]  --W_fold-->  dL/dh7 (+ d_sigma w_alpha)  --mask h7-->  dZ7  --W_7--> ... --W_1--> dZ0
//
// Eight tensor-core layers per block (K = 128 once, K = 256 seven times, N = 256, fp32 accumulators in TMEM); every
// dZ_l is written to HBM exactly once (row-major bf16, the operand of the weight-gradient GEMMs) and nothing is read
// back: the ReLU masks come from the 1-bit-per-activation dump of the forward kernel (32 B per row and layer).
// Algorithmic bytes per row: 8 x 512 B written + 8 x 32 B masks + 256 B dG + 4 B d_sigma = 4.6 KB (HBM-bound:
// 0.47 TFLOP over 2.0 GB for a 3,072-ray batch).
//
// Schedule.  A cluster of two CTAs works on two 256-row tiles (UMMA M = 256, cta_group::2: 128 rows of each tile per
// CTA, the weight operand split along N between the two CTAs' shared memory, so a CTA streams HALF of every weight
// slab).  The leader CTA's issuer runs the layers in the order (j, tile 0), (j, tile 1), (j+1, tile 0), ...: while the
// tensor core works on one tile the other tile's accumulator is drained, masked and written back as the next layer's
// A operand - MMA and drain overlap, which the round-1 kernel (both tiles issued together, cta_group::1) did not do,
// and the MMA reads 8 instead of 12 KB of shared memory per K-step per SM (that kernel's MMAs ran at half rate because
// the drain / store warps competed with them for shared-memory bandwidth).  Both tiles use the same weight slabs: the
// 10 x 8 KB ring (fills of two K-steps) holds one whole layer (8 fills) plus 2 fills of prefetch; a stage is released by
// tile 1's MMAs.  (A CTA's share of the stream is contiguous per fill: one 8 KB bulk copy.  Two 2 KB copies per K-step
// out of the round-1 slab layout starved the issuer - the copy engine's per-request cost, not bytes, was the limit.)
//
// Roles per CTA (768 threads, 1 CTA per SM, persistent over 512-row blocks): warp 0 lane 0 streams this CTA's N half
// of the weights of all eight layers in consumption order with cp.async.bulk (UMMA K-major SWIZZLE_NONE, [2][128][8] bf16
// per K = 16 step); warp 1 issues the MMAs (leader CTA, whole warp convergent, one elected lane per instruction) or
// relays "my half has landed" to the leader's barrier (peer CTA); warps 4-19 drain the two accumulators (8 warps each;
// tcgen05.ld 32x32b), add the sigma head's outer product on the first layer, mask, pack to bf16 and store the next
// layer's A operand into shared memory: K-major SWIZZLE_128B, 4 panels of 64 columns x 128 rows x 128 B - the image a
// TMA box has in shared memory; warps 20-23 then - while the tensor core already runs the next layer on that tile -
// hand it to the TMA engine (one lane: four cp.async.bulk.tensor stores, whole 128-byte lines of the row-major dz,
// rows beyond m clipped by the tensor map) and read it back for the bias gradients (column sums; every column has one
// owner thread, no atomics).  Storing from the accumulator-owning threads directly costs 2x (a thread owns one row: each
// of its stores touches 32 different lines, and the column sums need 128 shuffles); st.global from the four store warps
// took 2x the tensor time per tile.
#include "pgn_common.cuh"
#include "pgn_kernels.h"
#include "pgn_umma.cuh"
#include "pgn_tma.h"

using namespace pgn;

// -DPGN_CHAIN_PROF: per-role wait / work cycle counters of CTA 0 (tools/chain_prof.py reads them); off in the product build
#ifdef PGN_CHAIN_PROF
__device__ unsigned long long g_chain_prof[16];
#define CP_T0() const long long cp_t0 = clock64()
#define CP_ADD(slot) do { cp_acc[slot] += (unsigned long long)(clock64() - cp_t0); } while (0)
#define CP_ARGS , unsigned long long (&cp_acc)[10]
#define CP_PASS , cp_acc
#else
#define CP_T0()
#define CP_ADD(slot)
#define CP_ARGS
#define CP_PASS
#endif

namespace {

constexpr int kTile = 128;                       // rows of a tile held by one CTA (UMMA M = 256 over the pair)
constexpr int kTiles = 2;                        // tiles in flight per CTA pair (MMA of one overlaps the drain of the other)
constexpr int kPairRows = 2 * kTile * kTiles;    // 512 rows per pair-block
constexpr int kPanel = kTile * 128;              // bytes of one 64-column panel of a tile: 128 rows x 128 B (SWIZZLE_128B)
constexpr int kABytes = 4 * kPanel;              // 64 KB: [4 panels][128 rows][64] bf16, 16-byte chunks XOR-swizzled by row & 7
constexpr int kFillBytes = 2 * 2 * 128 * 16;     // one ring stage: two K = 16 steps of this CTA's N half, [2 ks][2][128][8] bf16 (8 KB)
constexpr int kStages = 10;                      // 80 KB: one layer (8 fills, read by both tiles) + 2 fills of prefetch
constexpr int kLayers = 8;                       // fold layer (K = 128) + W_7 .. W_1 (K = 256)
constexpr int kFillsPerBlock = 4 + 7 * 8;        // 60 (x 2 CTAs x 8 KB = the 960 KB weight stream)
constexpr int kThreads = 768;

struct __align__(1024) ChainSmem {
  uint8_t a[kTiles][kABytes];
  uint8_t w[kStages][kFillBytes];
  float colsum[kLayers][256];
  float w_alpha[256];
  uint64_t w_full[kStages], w_empty[kStages], acc_full[kTiles], act_ready[kTiles], cs_ready[kTiles], cs_done[kTiles];
  uint32_t tmem_slot;
};

// byte offset of the 16-byte chunk gc (8 columns 8 gc .. 8 gc + 7) of row r inside an A tile: K-major SWIZZLE_128B, the
// layout a TMA box of 64 columns x 128 rows has in shared memory (so the tile can be stored with cp.async.bulk.tensor)
__device__ __forceinline__ uint32_t a_off(int r, int gc) {
  return (uint32_t)(gc >> 3) * kPanel + (uint32_t)r * 128u + (uint32_t)(((gc & 7) ^ (r & 7)) << 4);
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];\n"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// One layer of the chain for both tiles, issued by the leader CTA's issuer warp (whole warp convergent, one elected
// lane per tcgen05 instruction).  A single warp retires about one instruction per 5-6 cycles and a K-step is 131 cycles
// of tensor time, so the loop has to be lean: ring stages, barrier parities and operand offsets are compile-time
// (120 slabs per block = 6 turns of the 20-stage ring, 8 layers per block: every parity repeats per block), descriptors
// are base + immediate.  (The first version computed stage = slab % 20 and the descriptors per K-step at run time:
// ~390 cycles per K-step, no faster than the round-1 kernel.)
// The issuer, the weight producer and the peer relay are single latency-critical warps: they spin on test_wait.  Parked in
// try_wait (hardware suspend) their wake-up latency sat on the tensor pipe's critical path: 295 -> 253 us per pass.
#define CHAIN_ISSUER_WAIT mbar_spin_s
template <int J>
__device__ __forceinline__ bool chain_issue_layer(uint32_t w_full0, uint32_t w_empty0, uint32_t acc_full0, uint32_t act_ready0,
                                                  uint32_t a_lo0, uint32_t ring_lo, uint32_t tmem, volatile int* status CP_ARGS) {
  constexpr int nfills = J == 0 ? 4 : 8;
  constexpr int fill0 = J == 0 ? 0 : 4 + (J - 1) * 8;
  constexpr uint32_t idesc = umma_idesc_bf16(2 * kTile, 256);
  constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);                 // B: SBO = 128 B, descriptor version 1 (SWIZZLE_NONE)
  constexpr uint32_t kADescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // A: SBO = 1 KB (8 rows), version 1, SWIZZLE_128B
  constexpr uint32_t b_lbo = ((kTile * 16u) >> 4) << 16;                 // B: LBO = 128 rows x 16 B (this CTA's N half)
#pragma unroll
  for (int t = 0; t < kTiles; ++t) {
    { CP_T0(); if (!CHAIN_ISSUER_WAIT(act_ready0 + t * 8, J & 1, status, 702)) return false; CP_ADD(0); }
    tc_fence_after_sync();
#pragma unroll
    for (int f = 0; f < nfills; ++f) {
      const int fl = fill0 + f, st = fl % kStages;
      if (t == 0) {                                                       // (tile 1 re-reads the resident fill)
        CP_T0();
        if (!CHAIN_ISSUER_WAIT(w_full0 + st * 8, (fl / kStages) & 1, status, 703)) return false;
        CP_ADD(1);
        tc_fence_after_sync();
      }
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int ks = 2 * f + g;                                         // K-step = 32 B inside the 128-byte rows of panel ks / 4
        const uint32_t a_lo = a_lo0 + (uint32_t)t * (kABytes >> 4) + (uint32_t)(ks >> 2) * (kPanel >> 4) + (uint32_t)(ks & 3) * 2u;
        const uint32_t b_lo = (ring_lo + (uint32_t)st * (kFillBytes >> 4) + (uint32_t)g * (kFillBytes >> 5)) | b_lbo;
        umma_bf16_2cta_elect(tmem + (uint32_t)t * 256, ((uint64_t)kADescHi << 32) | a_lo, ((uint64_t)kDescHi << 32) | b_lo, idesc,
                             ks > 0 ? 1u : 0u);
      }
      if (t == kTiles - 1) umma_commit_2cta_elect_s(w_empty0 + st * 8);   // both CTAs' ring stages
    }
    umma_commit_2cta_elect_s(acc_full0 + t * 8);
  }
  return true;
}

// A failed (timed-out) wait has already written the status word: every role then leaves through the common tail, so both
// CTAs of the pair still meet at the closing cluster barrier.
#define CHAIN_WAIT(bar, parity, code) do { if (!mbar_wait((bar), (parity), status, (code))) goto tail; } while (0)

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
pgn_delta_chain_kernel(const uint4* __restrict__ dG, const float* __restrict__ d_raw, const uint32_t* __restrict__ mask,
                       long long mask_rows, long long m, const uint8_t* __restrict__ wstream,
                       const float* __restrict__ w_alpha, const __grid_constant__ CUtensorMap dz_map, float* __restrict__ colsum_g,
                       int* __restrict__ status_g, unsigned layer_mask) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  ChainSmem& sm = *reinterpret_cast<ChainSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  volatile int* status = status_g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const long long n_blocks = (m + kPairRows - 1) / kPairRows;
  const long long cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&sm.w_full[i], rank == 0 ? 2 : 1); mbar_init(&sm.w_empty[i], 1); }
    for (int t = 0; t < kTiles; ++t) {
      mbar_init(&sm.acc_full[t], 1);
      mbar_init(&sm.act_ready[t], 16);           // the 8 drain warps of the tile in BOTH CTAs (leader's barrier is the one waited on)
      mbar_init(&sm.cs_ready[t], 8); mbar_init(&sm.cs_done[t], 4);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < kLayers * 256; i += kThreads) (&sm.colsum[0][0])[i] = 0.f;
  for (int i = tid; i < 256; i += kThreads) sm.w_alpha[i] = w_alpha[i];
  if (warp == 0) { tmem_alloc_2cta(&sm.tmem_slot, 512); tmem_relinquish_2cta(); }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_slot;
#ifdef PGN_CHAIN_PROF
  unsigned long long cp_acc[10] = {};
  const long long cp_k0 = clock64();
#endif

  if (warp == 0) {
    // ---------------------------------------------------------------- weight producer (this CTA's N half of every fill)
    if (lane == 0) {
      uint32_t fill = 0;
      for (long long blk = cluster_id; blk < n_blocks; blk += n_clusters) {
        for (int i = 0; i < kFillsPerBlock; ++i, ++fill) {
          const uint32_t st = fill % kStages, use = fill / kStages;
          if (use > 0 && !CHAIN_ISSUER_WAIT(smem_u32(&sm.w_empty[st]), (use - 1) & 1, status, 701)) goto tail;
          mbar_expect_tx(&sm.w_full[st], kFillBytes);
          bulk_g2s_s(smem_u32(sm.w[st]), wstream + ((size_t)i * 2 + rank) * kFillBytes, kFillBytes, smem_u32(&sm.w_full[st]));
        }
      }
    }
  } else if (warp == 1 && rank == 1) {
    // ---------------------------------------------------------------- peer relay: "my half of fill f has landed"
    if (lane == 0) {
      uint32_t fill = 0;
      for (long long blk = cluster_id; blk < n_blocks; blk += n_clusters) {
        for (int i = 0; i < kFillsPerBlock; ++i, ++fill) {
          const uint32_t st = fill % kStages, use = fill / kStages;
          if (!CHAIN_ISSUER_WAIT(smem_u32(&sm.w_full[st]), use & 1, status, 708)) goto tail;
          mbar_arrive_cluster(&sm.w_full[st], 0);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (leader CTA, UMMA M = 256 over the pair)
    static_assert(kFillsPerBlock % kStages == 0 && ((kFillsPerBlock / kStages) & 1) == 0 && (kLayers & 1) == 0,
                  "chain_issue_layer's compile-time ring stages / parities assume whole, even ring turns per block");
    const uint32_t w_full0 = smem_u32(&sm.w_full[0]), w_empty0 = smem_u32(&sm.w_empty[0]);
    const uint32_t acc_full0 = smem_u32(&sm.acc_full[0]), act_ready0 = smem_u32(&sm.act_ready[0]);
    const uint32_t a_lo0 = (smem_u32(sm.a[0]) >> 4) | (1u << 16);                          // A: K-major SWIZZLE_128B (LBO field = 16 B, unused)
    const uint32_t ring_lo = smem_u32(sm.w[0]) >> 4;
    for (long long blk = cluster_id; blk < n_blocks; blk += n_clusters) {
      bool ok = chain_issue_layer<0>(w_full0, w_empty0, acc_full0, act_ready0, a_lo0, ring_lo, tmem, status CP_PASS);
      ok = ok && chain_issue_layer<1>(w_full0, w_empty0, acc_full0, act_ready0, a_lo0, ring_lo, tmem, status CP_PASS);
      ok = ok && chain_issue_layer<2>(w_full0, w_empty0, acc_full0, act_ready0, a_lo0, ring_lo, tmem, status CP_PASS);
      ok = ok && chain_issue_layer<3>(w_full0, w_empty0, acc_full0, act_ready0, a_lo0, ring_lo, tmem, status CP_PASS);
      ok = ok && chain_issue_layer<4>(w_full0, w_empty0, acc_full0, act_ready0, a_lo0, ring_lo, tmem, status CP_PASS);
      ok = ok && chain_issue_layer<5>(w_full0, w_empty0, acc_full0, act_ready0, a_lo0, ring_lo, tmem, status CP_PASS);
      ok = ok && chain_issue_layer<6>(w_full0, w_empty0, acc_full0, act_ready0, a_lo0, ring_lo, tmem, status CP_PASS);
      ok = ok && chain_issue_layer<7>(w_full0, w_empty0, acc_full0, act_ready0, a_lo0, ring_lo, tmem, status CP_PASS);
      if (!ok) goto tail;
    }
  } else if (warp >= 20) {
    // ---------------------------------------------------------------- dZ store + bias-gradient group (128 threads)
    // One lane hands the finished tile to the TMA engine: four tensor stores (one per 64-column panel, box = 64 columns x
    // 128 rows, SWIZZLE_128B: the engine undoes the swizzle and writes whole 128-byte lines of the row-major dz; rows
    // beyond m are clipped by the tensor map).  (Round 1 / the first CTA-pair version stored with st.global from these
    // four warps: ~4.5 k cycles per tile and layer, 2x the tensor time - the whole kernel ran at the store group's pace.)
    // Meanwhile all four warps read the tile for the bias gradients: warp w owns the 8 chunks (64 columns) of panel w,
    // lane = (row & 3) + 4 * chunk; every column has one owner thread, no atomics.
    const int dw = warp - 20, rr = lane & 3, run = dw * 8 + (lane >> 2);
    uint32_t phase = 0;
    for (long long blk = cluster_id; blk < n_blocks; blk += n_clusters) {
      const long long r0 = blk * kPairRows + rank * kTile;
      for (int j = 0; j < kLayers; ++j, ++phase) {
        const int L = 7 - j;
        for (int t = 0; t < kTiles; ++t) {
          { CP_T0(); CHAIN_WAIT(&sm.cs_ready[t], phase & 1, 705); CP_ADD(2); }
          CP_T0();
          if ((layer_mask >> L) & 1u) {                   // layers nobody reads (frozen network) are neither stored nor summed
            // (issuing the four panel stores a quarter of the column-sum loop apart, so that weight fills could slip in
            // between them in the copy engine's queue, was slower: 385 instead of 296 us per pass)
            const uint32_t a_base = smem_u32(sm.a[t]);
            if (warp == 20 && lane == 0) {
              const int grow = (int)(r0 + t * 2 * kTile);
#pragma unroll
              for (int pnl = 0; pnl < 4; ++pnl) tma_store_3d(&dz_map, a_base + pnl * kPanel, pnl * 64, grow, L);
              bulk_commit_group();
            }
            float acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll 8
            for (int i = 0; i < 32; ++i) {
              const uint4 v = lds128(a_base + a_off(rr + 4 * i, run));
              const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                acc[2 * e] += __uint_as_float(w4[e] << 16);
                acc[2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
              }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 1);
              acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 2);
            }
            if (rr == 0) {
#pragma unroll
              for (int e = 0; e < 8; ++e) sm.colsum[L][run * 8 + e] += acc[e];
            }
            if (warp == 20 && lane == 0) bulk_wait_group_read0();      // the engine has read the tile out of shared memory
          }
          __syncwarp();
          CP_ADD(3);
          if (lane == 0) mbar_arrive_local(&sm.cs_done[t]);
        }
      }
    }
    if (warp == 20 && lane == 0) bulk_wait_group0();
    // this thread's columns of the CTA's bias partials -> global
    if (rr == 0) {
      for (int L = 0; L < kLayers; ++L)
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(colsum_g + L * 256 + run * 8 + e, sm.colsum[L][run * 8 + e]);
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue / staging groups (2 x 256 threads)
    // warps 4-11 own tile 0, warps 12-19 tile 1 (this CTA's 128 rows of each)
    const int ew = warp - 4, t = ew >> 3, q = ew & 3, half = (ew >> 2) & 1;
    const int et = (tid - 128) & 255;               // thread of the tile's group
    const int row = q * 32 + lane;                  // TMEM lane = row of this CTA's part of the tile
    const int col0 = half * 128;
    const uint32_t a_base = smem_u32(sm.a[t]);
    const uint32_t taddr = tmem + (uint32_t)t * 256 + ((uint32_t)(q * 32) << 16) + (uint32_t)col0;
    const int gc0 = half * 16;                      // this thread's first 16-byte chunk (8 columns) of its row
    uint32_t acc_phase = 0, csd_phase = 0;          // csd_phase: completed store/sum passes over the A tile waited for so far
    bool first = true;
    for (long long blk = cluster_id; blk < n_blocks; blk += n_clusters) {
      const long long r0 = blk * kPairRows + t * 2 * kTile + rank * kTile;
      // stage in dG: [128 rows x 128 columns] -> the first 16 runs of the A tile (once the store group has read the
      // previous block's last deltas out of it)
      {
        const int srow = et & 127, sh = et >> 7;    // row of the tile, 64-column half
        if (!first) {
          CHAIN_WAIT(&sm.cs_done[t], csd_phase & 1, 706);
          ++csd_phase;
        }
        first = false;
        const long long gr = r0 + srow;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (gr < m) v = __ldg(dG + (size_t)gr * 16 + sh * 8 + i);
          sts128(a_base + a_off(srow, sh * 8 + i), v.x, v.y, v.z, v.w);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&sm.act_ready[t], 0);
      }
      const long long gr = r0 + row;
      const float dsig = gr < m ? __ldg(d_raw + (size_t)gr * 4 + 3) : 0.f;
      for (int j = 0; j < kLayers; ++j, ++acc_phase) {
        const int L = 7 - j;                        // this layer's output is dL/dh_L; its mask is [h_L > 0]
        uint32_t mw[4] = {0u, 0u, 0u, 0u};           // word planes [layer][word][row]: this thread's 128 columns are words 4 half .. + 3
        if (gr < m) {
          const uint32_t* mp = mask + (size_t)(L * 8 + half * 4) * (size_t)mask_rows + (size_t)gr;
#pragma unroll
          for (int i = 0; i < 4; ++i) mw[i] = __ldg(mp + (size_t)i * (size_t)mask_rows);
        }
        { CP_T0(); CHAIN_WAIT(&sm.acc_full[t], acc_phase & 1, 704); CP_ADD(4); }
        tc_fence_after_sync();
        if (j > 0) {                                 // the previous deltas have been stored / summed out of the A tile
          CP_T0();
          CHAIN_WAIT(&sm.cs_done[t], csd_phase & 1, 707);
          CP_ADD(5);
          ++csd_phase;
        }
        CP_T0();
        uint32_t v[2][16];
        tmem_ld_32x16(taddr, v[0]);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          tmem_ld_wait();
          if (b + 1 < 8) tmem_ld_32x16(taddr + (uint32_t)(b + 1) * 16, v[(b + 1) & 1]);
          const uint32_t* vb = v[b & 1];
          const uint32_t bits = mw[b >> 1] >> ((b & 1) * 16);
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float p0 = __uint_as_float(vb[2 * i]), p1 = __uint_as_float(vb[2 * i + 1]);
            if (j == 0) {
              p0 = fmaf(dsig, sm.w_alpha[col0 + b * 16 + 2 * i], p0);
              p1 = fmaf(dsig, sm.w_alpha[col0 + b * 16 + 2 * i + 1], p1);
            }
            if (!(bits & (1u << (2 * i)))) p0 = 0.f;
            if (!(bits & (1u << (2 * i + 1)))) p1 = 0.f;
            pk[i] = pack_bf16x2(p0, p1);
          }
          sts128(a_base + a_off(row, gc0 + 2 * b), pk[0], pk[1], pk[2], pk[3]);
          sts128(a_base + a_off(row, gc0 + 2 * b + 1), pk[4], pk[5], pk[6], pk[7]);
        }
        tc_fence_before_sync();
        fence_proxy_async_smem();                    // the tile is read by the tensor core (next layer) and by the TMA store
        __syncwarp();
        CP_ADD(6);
        if (lane == 0) {
          if (j + 1 < kLayers) mbar_arrive_cluster(&sm.act_ready[t], 0);  // next layer's A operand (and a free accumulator)
          mbar_arrive_local(&sm.cs_ready[t]);                            // dZ_L of this tile is in shared memory
        }
      }
    }
  }
tail:
#ifdef PGN_CHAIN_PROF
  if (blockIdx.x == 0 && (tid == 32 || tid == 128 || tid == 640)) {
    cp_acc[tid == 32 ? 7 : (tid == 128 ? 8 : 9)] = (unsigned long long)(clock64() - cp_k0);
    for (int k = 0; k < 10; ++k) atomicAdd(&g_chain_prof[k], cp_acc[k]);
  }
#endif
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                 // the peer's shared memory / TMEM stay alive until every MMA has retired
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc_2cta(tmem, 512); }
}

// the chain's weight stream from the context's fp32 copies (include/posegen_b200.h: pgn_mlp_delta_chain): slab element
// (j, ks, kc, n, e) = W'_j[n][ks * 16 + kc * 8 + e];  W'_0 = fold^T (fold = W_v[:, :256] W_f, [128][256]),
// W'_j = W_l^T for l = 8 - j (nn.Linear layout [out][in]; the skip layer l = 5 without its first 432 input columns)
struct ChainPackPtrs { const float* w[8]; };       // pts_linears.0 .. 7
__global__ void pgn_pack_chain_kernel(ChainPackPtrs p, const float* __restrict__ fold, __nv_bfloat16* __restrict__ out) {
  const int total = kFillsPerBlock * 2 * (kFillBytes / 2);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int fill = idx >> 13, r = idx & 8191;                       // 8192 elements per fill (both CTA halves)
    const int h = r >> 12, g = (r >> 11) & 1, kc = (r >> 10) & 1, n = h * 128 + ((r >> 3) & 127), e = r & 7;
    float v;
    if (fill < 4) {
      const int k = (2 * fill + g) * 16 + kc * 8 + e;
      v = fold[k * 256 + n];
    } else {
      const int j = 1 + (fill - 4) / 8, ks = 2 * ((fill - 4) % 8) + g, l = 8 - j;
      const int k = ks * 16 + kc * 8 + e;
      v = (l == 5) ? p.w[5][k * 688 + 432 + n] : p.w[l][k * 256 + n];
    }
    out[idx] = __float2bfloat16_rn(v);
  }
}

}  // namespace

cudaError_t pgn_launch_pack_chain_weights(const float* const* w_dev, const float* fold, __nv_bfloat16* out, cudaStream_t stream) {
  ChainPackPtrs p;
  for (int l = 0; l < 8; ++l) p.w[l] = w_dev[l];
  pgn_pack_chain_kernel<<<148 * 2, 256, 0, stream>>>(p, fold, out);
  return cudaGetLastError();
}

cudaError_t pgn_launch_delta_chain(const void* dG, const float* d_raw, const void* mask, long long mask_rows, long long m,
                                   const void* wstream, const float* w_alpha, void* dz, float* colsum, unsigned layer_mask,
                                   int* status, int num_sms, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(colsum, 0, sizeof(float) * kLayers * 256, stream);
  if (e != cudaSuccess || m == 0) return e;
  const size_t smem = sizeof(ChainSmem) + 1024;
  static PgnPerDeviceOnce configured;
  if (configured.need()) {
    e = cudaFuncSetAttribute(pgn_delta_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured.set();
  }
  const long long n_blocks = (m + kPairRows - 1) / kPairRows;
  const long long pairs = num_sms / 2;
  const unsigned grid = 2u * (unsigned)(n_blocks < pairs ? n_blocks : pairs);       // clusters of 2 CTAs
  CUtensorMap dz_map;                 // dz: bf16 [8][m][256] row-major, stored in boxes of 64 columns x 128 rows
  e = pgn_make_map_bf16(&dz_map, dz, 256, 256, m, kLayers, m * 256, kTile);
  if (e != cudaSuccess) return e;
  pgn_delta_chain_kernel<<<grid, kThreads, smem, stream>>>(reinterpret_cast<const uint4*>(dG), d_raw,
                                                           reinterpret_cast<const uint32_t*>(mask), mask_rows, m,
                                                           reinterpret_cast<const uint8_t*>(wstream), w_alpha,
                                                           dz_map, colsum, status, layer_mask);
  return cudaGetLastError();
}

#ifdef PGN_CHAIN_PROF
extern "C" __attribute__((visibility("default"))) int pgn_debug_chain_prof(unsigned long long* out16, int reset) {
  unsigned long long z[16] = {0};
  if (cudaMemcpyFromSymbol(out16, g_chain_prof, sizeof(z)) != cudaSuccess) return -1;
  if (reset && cudaMemcpyToSymbol(g_chain_prof, z, sizeof(z)) != cudaSuccess) return -1;
  return 0;
}
#endif
