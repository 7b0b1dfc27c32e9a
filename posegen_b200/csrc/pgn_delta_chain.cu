// Fused delta chain of the A-NeRF trunk backward (training step / GAN step) on tcgen05.
//
// What autograd does per trunk layer (core/networks/nerf.py:94-102 backwards) is a GEMM dL/dh_{l-1} = dZ_l W_l followed
// by a ReLU mask; done layer by layer through HBM that is 5 passes over a [rows,256] matrix per layer.  Here one CTA
// keeps a 256-row block of deltas in shared memory for the whole chain:
//
//     dG [rows,128]  --W_fold-->  dL/dh7 (+ d_sigma w_alpha)  --mask h7-->  dZ7  --W_7--> ... --W_1--> dZ0
//
// Eight tensor-core layers per block (K = 128 once, K = 256 seven times, N = 256, fp32 accumulators in TMEM); every
// dZ_l is written to HBM exactly once (row-major bf16, the operand of the weight-gradient GEMMs) and nothing is read
// back: the ReLU masks come from the 1-bit-per-activation dump of the forward kernel (32 B per row and layer).
// Algorithmic bytes per row: 8 x 512 B written + 8 x 32 B masks + 256 B dG + 4 B d_sigma = 4.6 KB (HBM-bound:
// 0.47 TFLOP over 2.0 GB for a 3,072-ray batch).
//
// Roles (768 threads, 1 CTA per SM, persistent over row blocks): warp 0 lane 0 streams the weights of all eight layers
// in consumption order through a 10 x 8 KB ring with cp.async.bulk (one K-step slab = [2][256][8] bf16, UMMA K-major
// SWIZZLE_NONE); warp 1 lane 0 issues the MMAs (UMMA M = 128, two row tiles share every weight slab, cta_group::1);
// warps 4-19 drain the two accumulators (8 warps each; tcgen05.ld 32x32b), add the sigma head's outer product on the
// first layer, mask, pack to bf16 and store the next layer's A operand into shared memory ([k/8][row][8]); warps 20-23 then read that
// tile back from shared memory - while the tensor core already runs the next layer on it - and write the dZ rows to
// HBM (4 rows x 128 B per warp store: full lines) and accumulate the bias gradients (column sums; every column has
// one owner thread, no atomics).  Draining and storing from the accumulator-owning threads directly costs 2x: a
// thread owns one row, so each of its stores touches 32 different lines and the column sums need 128 shuffles.
#include "pgn_common.cuh"
#include "pgn_kernels.h"
#include "pgn_umma.cuh"

using namespace pgn;

// -DPGN_CHAIN_PROF: per-role wait / work cycle counters of CTA 0 (tools/chain_prof.py reads them); off in the product build
#ifdef PGN_CHAIN_PROF
__device__ unsigned long long g_chain_prof[16];
#define CP_T0() const long long cp_t0 = clock64()
#define CP_ADD(slot) do { if (blockIdx.x == 0) atomicAdd(&g_chain_prof[slot], (unsigned long long)(clock64() - cp_t0)); } while (0)
#else
#define CP_T0()
#define CP_ADD(slot)
#endif

namespace {

constexpr int kTile = 128;                       // rows per UMMA tile
constexpr int kTiles = 2;                        // row tiles per CTA block (share the weight slabs)
constexpr int kBlockRows = kTile * kTiles;
constexpr int kRun = kTile * 16;                 // bytes of one 8-wide K run of a tile
constexpr int kABytes = 32 * kRun;               // 64 KB: [256/8 runs][128 rows][8] bf16
constexpr int kSlabBytes = 2 * 256 * 16;         // one K = 16 step of a [256 x K] weight: [2][256][8] bf16
constexpr int kStages = 10;                     // 80 KB of weight slabs in flight (a 16-slab layer is 128 KB)
constexpr int kLayers = 8;                       // fold layer (K = 128) + W_7 .. W_1 (K = 256)
constexpr int kSlabsPerBlock = 8 + 7 * 16;       // 120
constexpr int kThreads = 768;
constexpr int kGroup = 4;                       // weight slabs (K-steps) issued per tile before switching accumulators

struct __align__(1024) ChainSmem {
  uint8_t a[kTiles][kABytes];
  uint8_t w[kStages][kSlabBytes];
  float colsum[kLayers][256];
  float w_alpha[256];
  uint64_t w_full[kStages], w_empty[kStages], acc_full, act_ready[kTiles], cs_ready[kTiles], cs_done[kTiles];
  uint32_t tmem_slot;
};

__device__ __forceinline__ void mbar_arrive_local(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__global__ void __launch_bounds__(kThreads, 1)
pgn_delta_chain_kernel(const uint4* __restrict__ dG, const float* __restrict__ d_raw, const uint4* __restrict__ mask,
                       long long mask_rows, long long m, const uint8_t* __restrict__ wstream,
                       const float* __restrict__ w_alpha, uint4* __restrict__ dz, float* __restrict__ colsum_g,
                       int* __restrict__ status_g, unsigned layer_mask) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  ChainSmem& sm = *reinterpret_cast<ChainSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  volatile int* status = status_g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long n_blocks = (m + kBlockRows - 1) / kBlockRows;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&sm.w_full[i], 1); mbar_init(&sm.w_empty[i], 1); }
    mbar_init(&sm.acc_full, 1);
    for (int t = 0; t < kTiles; ++t) { mbar_init(&sm.act_ready[t], 8); mbar_init(&sm.cs_ready[t], 8); mbar_init(&sm.cs_done[t], 4); }
    fence_mbar_init();
  }
  for (int i = tid; i < kLayers * 256; i += kThreads) (&sm.colsum[0][0])[i] = 0.f;
  for (int i = tid; i < 256; i += kThreads) sm.w_alpha[i] = w_alpha[i];
  if (warp == 0) { tmem_alloc(&sm.tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = sm.tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- weight producer
    if (lane == 0) {
      uint32_t slab = 0;
      for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        for (int i = 0; i < kSlabsPerBlock; ++i, ++slab) {
          const uint32_t st = slab % kStages, use = slab / kStages;
          if (use > 0 && !mbar_wait(&sm.w_empty[st], (use - 1) & 1, status, 701)) return;
          mbar_expect_tx(&sm.w_full[st], kSlabBytes);
          bulk_g2s_s(smem_u32(sm.w[st]), wstream + (size_t)i * kSlabBytes, kSlabBytes, smem_u32(&sm.w_full[st]));
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(kTile, 256);
      uint32_t slab = 0, ready_phase = 0;
      for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        for (int j = 0; j < kLayers; ++j, ++ready_phase) {
          { CP_T0();
          for (int t = 0; t < kTiles; ++t)
            if (!mbar_wait(&sm.act_ready[t], ready_phase & 1, status, 702)) return;
          CP_ADD(0); }
          tc_fence_after_sync();
          // K-steps are issued in groups of kGroup slabs per tile: consecutive MMAs that accumulate into the SAME TMEM
          // tile pipeline back to back (131 cycles each); alternating the two accumulators every instruction costs 2x
          const int nks = j == 0 ? 8 : 16;
          for (int ks0 = 0; ks0 < nks; ks0 += kGroup) {
#pragma unroll
            for (int g = 0; g < kGroup; ++g) {
              const uint32_t sl = slab + g, st = sl % kStages, use = sl / kStages;
              CP_T0();
              if (!mbar_wait(&sm.w_full[st], use & 1, status, 703)) return;
              CP_ADD(1);
            }
            tc_fence_after_sync();
#pragma unroll
            for (int t = 0; t < kTiles; ++t) {
#pragma unroll
              for (int g = 0; g < kGroup; ++g) {
                const uint32_t st = (slab + g) % kStages;
                const uint64_t bd = umma_smem_desc(smem_u32(sm.w[st]), 256 * 16, 128);
                const uint64_t ad = umma_smem_desc(smem_u32(sm.a[t]) + (uint32_t)(ks0 + g) * 2 * kRun, kRun, 128);
                umma_bf16(tmem + (uint32_t)t * 256, ad, bd, idesc, (ks0 + g) > 0 ? 1u : 0u);
              }
            }
#pragma unroll
            for (int g = 0; g < kGroup; ++g) umma_commit(&sm.w_empty[(slab + g) % kStages]);
            slab += kGroup;
          }
          umma_commit(&sm.acc_full);
        }
      }
    }
  } else if (warp >= 20) {
    // ---------------------------------------------------------------- dZ store + bias-gradient group (128 threads)
    // warp w owns the 8 runs (64 columns) 8w .. 8w+7 of a tile; lane = (row & 3) + 4 * (run & 7): a warp store covers
    // 4 rows x 128 contiguous bytes (full lines).  (8 rows x 4 runs would make the LDS.128 conflict-free, but its
    // 64-byte row pieces made this group 1.6x slower: it is bound by its global stores, all SMs storing at once.)
    const int dw = warp - 20, rr = lane & 3, run = dw * 8 + (lane >> 2);
    uint32_t phase = 0;
    for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
      const long long r0 = blk * kBlockRows;
      for (int j = 0; j < kLayers; ++j, ++phase) {
        const int L = 7 - j;
        for (int t = 0; t < kTiles; ++t) {
          { CP_T0(); if (!mbar_wait(&sm.cs_ready[t], phase & 1, status, 705)) return; if (warp == 20 && lane == 0) CP_ADD(2); }
          CP_T0();
          const uint32_t src = smem_u32(sm.a[t]) + (uint32_t)run * kRun + rr * 16;
          const long long g0 = r0 + t * kTile + rr;
          uint4* out = dz + ((size_t)L * m + (size_t)g0) * 32 + run;
          float acc[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = 0.f;
          if ((layer_mask >> L) & 1u)                     // layers nobody reads (frozen network) are neither stored nor summed
#pragma unroll 8
          for (int i = 0; i < 32; ++i) {
            const uint4 v = lds128(src + (uint32_t)i * 64);
            if (g0 + 4 * i < m) __stcs(out + (size_t)i * 128, v);      // written once, read by a later kernel: streaming store
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
          __syncwarp();
          if (warp == 20 && lane == 0) CP_ADD(3);
          if (lane == 0) mbar_arrive_local(&sm.cs_done[t]);
        }
      }
    }
    // this thread's columns of the CTA's bias partials -> global
    if (rr == 0) {
      for (int L = 0; L < kLayers; ++L)
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(colsum_g + L * 256 + run * 8 + e, sm.colsum[L][run * 8 + e]);
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue / staging groups (2 x 256 threads)
    // warps 4-11 own tile 0, warps 12-19 tile 1: both accumulators are drained at the same time (a drain is a long
    // dependent chain per thread, so its rate grows with the number of warps on it)
    const int ew = warp - 4, t = ew >> 3, q = ew & 3, half = (ew >> 2) & 1;
    const int et = (tid - 128) & 255;               // thread of the tile's group
    const int row = q * 32 + lane;                  // TMEM lane = row of the tile
    const int col0 = half * 128;
    const uint32_t a_base = smem_u32(sm.a[t]);
    const uint32_t taddr = tmem + (uint32_t)t * 256 + ((uint32_t)(q * 32) << 16) + (uint32_t)col0;
    const uint32_t dst0 = a_base + (uint32_t)(col0 >> 3) * kRun + row * 16;
    uint32_t acc_phase = 0, csd_phase = 0;          // csd_phase: completed store/sum passes over the A tile waited for so far
    bool first = true;
    for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
      const long long r0 = blk * kBlockRows + t * kTile;
      // stage in dG: [128 rows x 128 columns] -> the first 16 runs of the A tile (once the store group has read the
      // previous block's last deltas out of it)
      {
        const int srow = et & 127, sh = et >> 7;    // row of the tile, 64-column half
        if (!first) {
          if (!mbar_wait(&sm.cs_done[t], csd_phase & 1, status, 706)) return;
          ++csd_phase;
        }
        first = false;
        const long long gr = r0 + srow;
        const uint32_t dst = a_base + (uint32_t)(sh * 8) * kRun + srow * 16;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (gr < m) v = __ldg(dG + (size_t)gr * 16 + sh * 8 + i);
          sts128(dst + (uint32_t)i * kRun, v.x, v.y, v.z, v.w);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_local(&sm.act_ready[t]);
      }
      const long long gr = r0 + row;
      const float dsig = gr < m ? __ldg(d_raw + (size_t)gr * 4 + 3) : 0.f;
      for (int j = 0; j < kLayers; ++j, ++acc_phase) {
        const int L = 7 - j;                        // this layer's output is dL/dh_L; its mask is [h_L > 0]
        const uint4 mk = gr < m ? __ldg(mask + ((size_t)L * mask_rows + gr) * 2 + half) : make_uint4(0u, 0u, 0u, 0u);
        { CP_T0(); if (!mbar_wait(&sm.acc_full, acc_phase & 1, status, 704)) return; if (tid == 128) CP_ADD(4); }
        tc_fence_after_sync();
        if (j > 0) {                                 // the previous deltas have been stored / summed out of the A tile
          CP_T0();
          if (!mbar_wait(&sm.cs_done[t], csd_phase & 1, status, 707)) return;
          if (tid == 128) CP_ADD(5);
          ++csd_phase;
        }
        CP_T0();
        const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
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
          sts128(dst0 + (uint32_t)(2 * b) * kRun, pk[0], pk[1], pk[2], pk[3]);
          sts128(dst0 + (uint32_t)(2 * b + 1) * kRun, pk[4], pk[5], pk[6], pk[7]);
        }
        tc_fence_before_sync();
        if (j + 1 < kLayers) fence_proxy_async_smem();
        __syncwarp();
        if (tid == 128) CP_ADD(6);
        if (lane == 0) {
          if (j + 1 < kLayers) mbar_arrive_local(&sm.act_ready[t]);      // next layer's A operand (and a free accumulator)
          mbar_arrive_local(&sm.cs_ready[t]);                            // dZ_L of this tile is in shared memory
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tmem, 512); }
}

// the chain's weight stream from the context's fp32 copies (include/posegen_b200.h: pgn_mlp_delta_chain): slab element
// (j, ks, kc, n, e) = W'_j[n][ks * 16 + kc * 8 + e];  W'_0 = fold^T (fold = W_v[:, :256] W_f, [128][256]),
// W'_j = W_l^T for l = 8 - j (nn.Linear layout [out][in]; the skip layer l = 5 without its first 432 input columns)
struct ChainPackPtrs { const float* w[8]; };       // pts_linears.0 .. 7
__global__ void pgn_pack_chain_kernel(ChainPackPtrs p, const float* __restrict__ fold, __nv_bfloat16* __restrict__ out) {
  const int total = kSlabsPerBlock * 4096;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int slab = idx >> 12, r = idx & 4095;
    const int kc = r >> 11, n = (r >> 3) & 255, e = r & 7;
    float v;
    if (slab < 8) {
      const int k = slab * 16 + kc * 8 + e;
      v = fold[k * 256 + n];
    } else {
      const int j = 1 + (slab - 8) / 16, ks = (slab - 8) % 16, l = 8 - j;
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
  const long long n_blocks = (m + kBlockRows - 1) / kBlockRows;
  const unsigned grid = (unsigned)(n_blocks < num_sms ? n_blocks : num_sms);
  pgn_delta_chain_kernel<<<grid, kThreads, smem, stream>>>(reinterpret_cast<const uint4*>(dG), d_raw,
                                                           reinterpret_cast<const uint4*>(mask), mask_rows, m,
                                                           reinterpret_cast<const uint8_t*>(wstream), w_alpha,
                                                           reinterpret_cast<uint4*>(dz), colsum, status, layer_mask);
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
