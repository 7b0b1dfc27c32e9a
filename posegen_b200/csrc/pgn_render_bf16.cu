// bf16 tensor-core render path (PGN_PRECISION_BF16): fused per-ray-tile pipeline on
// tcgen05 (UMMA M=128, fp32 accumulators in TMEM), weights streamed by the TMA engine
// (cp.async.bulk) as pre-packed K-major slabs, activations resident in shared memory.
//
// One persistent CTA per SM renders groups of 8 rays:
//   coarse: 8x64 samples  = 4 tiles of 128 rows        (network_fn)
//   composite -> inverse-CDF resample -> merge (warp per ray)
//   fine:   8x80 samples  = 5 tiles of 128 rows        (network_fine)
//   composite -> outputs
// Per 128-row tile, 9 tensor-core layers (K-steps of 16):
//   L0  x_p(432)            -> 256  ReLU      A generated on the fly (skeleton-relative
//   L1-4 h(256)             -> 256  ReLU        encoding + cutoff PE, 8 joints per chunk)
//   L5  h(256) | x_p(432)   -> 256  ReLU      (skip concat as two accumulating K ranges)
//   L6-7 h(256)             -> 256  ReLU      (sigma head folded into L7's epilogue, fp32)
//   V   h7(256) | d(672)    -> 128  ReLU      (feature_linear folded into views_linears[0];
//                                              rgb head folded into the epilogue, fp32)
// The samples x joints x embedding tensor only ever exists as 36 KB chunks in shared memory.
//
// Warp roles (320 threads): warps 0-7 encode + epilogue (warp w owns TMEM lanes
// 32*(w%4).., column half w/4), warp 8 = weight producer (one elected lane issues bulk
// copies), warp 9 = MMA issuer (one elected lane) and TMEM allocator.
//
// Reference semantics: core/raycasters.py:361-474, core/encoders.py:8-37,110-122,181-193,
// core/cutoff_embedder.py:111-174, core/networks/nerf.py:94-205, core/utils/ray_utils.py:157-289.
#include "pgn_common.cuh"
#include "pgn_kernels.h"
#include "pgn_umma.cuh"
#include "pgn_bf16_layout.h"

using namespace pgn;

namespace {

constexpr int kComputeThreads = 256;
constexpr int kThreads = 320;
constexpr int kProducerWarp = 8;
constexpr int kIssuerWarp = 9;
constexpr int kRPG = 8;                 // rays per group
constexpr int kTM = 128;                // rows per tile (UMMA M)
constexpr int kRunBytes = kTM * 16;     // one 8-wide K run of all 128 rows
constexpr int kActBytes = 256 / 8 * kRunBytes;          // 65536
constexpr int kStgBytes = 18 * kRunBytes;               // 36864 (144 K)
constexpr int kWStages = 8;             // 8 x 8 KB = one full 256x256 layer half per CTA
constexpr int kWStageBytes = 8192;
constexpr int kTmemCols = 256;
constexpr int kMaxTileRays = 3;

struct __align__(128) Smem {
  uint8_t act[kActBytes];
  uint8_t stg[kStgBytes];
  uint8_t wring[kWStages][kWStageBytes];
  float bias[9 * 256];
  float w_alpha[256];
  float w_rgb[3 * 128];
  float wcache[PGN_J * kTM];            // d-window per (joint,row)
  float dtab[kMaxTileRays][PGN_J * 28]; // PE of joint-frame view dirs per ray of the tile
  float zc[kRPG][PGN_S];
  float zf[kRPG][PGN_T];
  float raw[kRPG][PGN_T * 4];
  float wts[kRPG][PGN_S];
  float scratch[kRPG][128];
  float part[2][kTM][4];                // per column-half partial (rgb, sigma)
  float ray_o[kRPG][3], ray_d[kRPG][3], dnorm[kRPG];
  uint64_t w_full[kWStages], w_empty[kWStages];
  uint64_t stg_full, stg_empty, act_ready, acc_full;
  uint32_t tmem_base;
};

struct TileCtx {
  long long ray0;     // first ray of the group
  int S;              // samples per ray in this pass
  int row0;           // first row of the tile within the group pass
  int total_rows;     // valid rows in the group pass
};

// ------------------------------------------------------------------ encode (compute warps)
// x chunk c (joints 8c..8c+7): thread (row, half) produces joints 8c+4*half..+3 -> 72 values
// = 9 runs of 8 at run index half*9+r.
template <bool kStage>
__device__ __forceinline__ void encode_x_chunk(Smem& sm, const PgnRayRefs& rays, const PgnScalars& sc, const TileCtx& tc,
                                               int chunk, int row, int half, bool write_wcache,
                                               const float* __restrict__ enc_rows, int rows_valid) {
  uint32_t packed[36];
  if (kStage) {
#pragma unroll
    for (int i = 0; i < 36; ++i) {
      const int kp = chunk * 144 + half * 72 + 2 * i;
      float a = 0.f, b = 0.f;
      if (row < rows_valid) {
        a = enc_rows[(size_t)row * PGN_ENC + pgn_xperm_refcol(kp)];
        b = enc_rows[(size_t)row * PGN_ENC + pgn_xperm_refcol(kp + 1)];
      }
      packed[i] = pack_bf16x2(a, b);
    }
  } else {
    const int grow = tc.row0 + row;
    const bool valid = grow < tc.total_rows;
    const int rl = valid ? grow / tc.S : 0;
    const int s = valid ? grow - rl * tc.S : 0;
    const float z = (tc.S == PGN_S) ? sm.zc[rl][s] : sm.zf[rl][s];
    float px, py, pz;
    pgn_sample_point(sm.ray_o[rl], sm.ray_d[rl], z, px, py, pz);
    const float* skt = pgn_ray_skts(rays, tc.ray0 + rl);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = chunk * 8 + half * 4 + jj;
      const float4* m = reinterpret_cast<const float4*>(skt + j * 16);
      PgnJointGeom g = pgn_joint_geom<true>(__ldg(m), __ldg(m + 1), __ldg(m + 2), px, py, pz, sc.tau_v, sc.cutoff_v[j]);
      if (!valid) { g.v = 0.f; g.w = 0.f; g.rx = g.ry = g.rz = 0.f; }
      if (write_wcache) sm.wcache[j * kTM + row] = valid ? pgn_window<true>(g.v, sc.tau_d, sc.cutoff_d[j]) : 0.f;
      float sn, cs;
      __sincosf(g.v, &sn, &cs);
      float vals[18];
      vals[0] = g.v * g.w;
#pragma unroll
      for (int f = 0; f < PGN_LV; ++f) {
        vals[1 + 2 * f] = sn * g.w;
        vals[2 + 2 * f] = cs * g.w;
        const float s2 = 2.f * sn * cs;
        const float c2 = fmaf(cs, cs, -sn * sn);
        sn = s2; cs = c2;
      }
      vals[15] = g.rx; vals[16] = g.ry; vals[17] = g.rz;
#pragma unroll
      for (int i = 0; i < 9; ++i) packed[jj * 9 + i] = pack_bf16x2(vals[2 * i], vals[2 * i + 1]);
    }
  }
  uint8_t* base = sm.stg + (size_t)(half * 9) * kRunBytes + row * 16;
#pragma unroll
  for (int r = 0; r < 9; ++r)
    *reinterpret_cast<uint4*>(base + r * kRunBytes) = make_uint4(packed[4 * r], packed[4 * r + 1], packed[4 * r + 2], packed[4 * r + 3]);
}

// d chunk c (joints 4c..4c+3): thread (row, half) produces joints 4c+2*half, +1 -> 56 values
// = 7 runs at run index half*7+r.
template <bool kStage>
__device__ __forceinline__ void encode_d_chunk(Smem& sm, const TileCtx& tc, int chunk, int row, int half, int tile_ray0,
                                               const float* __restrict__ enc_rows, int rows_valid) {
  uint32_t packed[28];
  if (kStage) {
#pragma unroll
    for (int i = 0; i < 28; ++i) {
      const int q = chunk * 112 + half * 56 + 2 * i;
      const int ca = pgn_dperm_refcol(q), cb = pgn_dperm_refcol(q + 1);
      float a = 0.f, b = 0.f;
      if (row < rows_valid) {
        if (ca >= 0) a = enc_rows[(size_t)row * PGN_ENC + ca];
        if (cb >= 0) b = enc_rows[(size_t)row * PGN_ENC + cb];
      }
      packed[i] = pack_bf16x2(a, b);
    }
  } else {
    const int grow = tc.row0 + row;
    const bool valid = grow < tc.total_rows;
    const int rl = valid ? grow / tc.S : tile_ray0;
    const int tr = min(max(rl - tile_ray0, 0), kMaxTileRays - 1);
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = chunk * 4 + half * 2 + jj;
      const float wd = sm.wcache[j * kTM + row];
      const float4* tab = reinterpret_cast<const float4*>(&sm.dtab[tr][j * 28]);
#pragma unroll
      for (int i = 0; i < 7; ++i) {
        const float4 t = tab[i];
        packed[jj * 14 + 2 * i] = pack_bf16x2(t.x * wd, t.y * wd);
        packed[jj * 14 + 2 * i + 1] = pack_bf16x2(t.z * wd, t.w * wd);
      }
    }
  }
  uint8_t* base = sm.stg + (size_t)(half * 7) * kRunBytes + row * 16;
#pragma unroll
  for (int r = 0; r < 7; ++r)
    *reinterpret_cast<uint4*>(base + r * kRunBytes) = make_uint4(packed[4 * r], packed[4 * r + 1], packed[4 * r + 2], packed[4 * r + 3]);
}

// ------------------------------------------------------------------ epilogue (compute warps)
// MODE 0: hidden layer -> act (bf16, ReLU).  MODE 1: same + sigma partial.  MODE 2: view layer -> rgb partial.
template <int MODE>
__device__ __forceinline__ void epilogue(Smem& sm, uint32_t tmem_base, int layer, int warp, int lane) {
  const int q = warp & 3, half = warp >> 2;
  const int row = q * 32 + lane;
  constexpr int kCols = (MODE == 2) ? 64 : 128;        // columns per thread
  const int col0 = half * kCols;
  const float* bias = sm.bias + layer * 256;
  float sig = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f;
#pragma unroll 1
  for (int b = 0; b < kCols / 32; ++b) {
    uint32_t v[32];
    const int c0 = col0 + b * 32;
    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
    float x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = fmaxf(__uint_as_float(v[i]) + bias[c0 + i], 0.f);
    if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) sig = fmaf(x[i], sm.w_alpha[c0 + i], sig);
    }
    if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        r0 = fmaf(x[i], sm.w_rgb[c0 + i], r0);
        r1 = fmaf(x[i], sm.w_rgb[128 + c0 + i], r1);
        r2 = fmaf(x[i], sm.w_rgb[256 + c0 + i], r2);
      }
    } else {
      uint8_t* dst = sm.act + (size_t)(c0 >> 3) * kRunBytes + row * 16;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4*>(dst + g * kRunBytes) =
            make_uint4(pack_bf16x2(x[8 * g], x[8 * g + 1]), pack_bf16x2(x[8 * g + 2], x[8 * g + 3]),
                       pack_bf16x2(x[8 * g + 4], x[8 * g + 5]), pack_bf16x2(x[8 * g + 6], x[8 * g + 7]));
    }
  }
  if (MODE == 1) sm.part[half][row][3] = sig;
  if (MODE == 2) { sm.part[half][row][0] = r0; sm.part[half][row][1] = r1; sm.part[half][row][2] = r2; }
}

// "my part of the A operand is written / my TMEM reads are done" -> the LEADER CTA's barrier
__device__ __forceinline__ void compute_arrive(uint64_t* bar) {
  tc_fence_before_sync();
  fence_proxy_async_smem();
  mbar_arrive_cluster(bar, 0);
}
__device__ __forceinline__ void compute_bar_sync() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }

// optional phase timers (cycles, one elected thread per role, accumulated per CTA):
//  0 issuer wait w_full | 1 issuer wait stg_full | 2 issuer wait act_ready | 3 issuer total
//  4 producer wait w_empty | 5 producer total
//  6 encode_x | 7 encode_d | 8 epilogue | 9 wait acc_full | 10 wait stg_empty | 12 compute total
//  13 issuer MMA issue | 14 issuer commits
#define PROF_T0() const long long _pt0 = prof ? clock64() : 0
#define PROF_ADD(slot) do { if (prof) pacc[slot] += (unsigned long long)(clock64() - _pt0); } while (0)

// Static per-cluster-iteration schedule: the two CTAs of a pair always run the same tile sequence
// (4 coarse + 5 fine pair-tiles; rows beyond a CTA's rays are masked), so the leader's issuer, both
// weight producers, the peer's relay and both sets of compute warps stay in lock-step by construction.
__device__ __forceinline__ int tiles_in_pass(int pass) { return pass == 0 ? (kRPG * PGN_S) / kTM : (kRPG * PGN_T) / kTM; }

// ------------------------------------------------------------------ the kernel
template <bool kStage>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
pgn_render_bf16_kernel(PgnRayRefs rays, PgnOutputs out, PgnBf16Net net_c, PgnBf16Net net_f,
                       const PgnScalars* __restrict__ scp, const float* __restrict__ near_far,
                       const float* __restrict__ enc_global, long long enc_rows_total, float* __restrict__ raw_global,
                       int* __restrict__ status_g, unsigned long long* __restrict__ prof) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  volatile int* status = status_g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const PgnScalars& sc = *scp;
  const long long n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;

  // work units: a "pair-group" = two ray groups (one per CTA); stage mode: two 128-row tiles
  const long long n_units = kStage ? (enc_rows_total + kTM - 1) / kTM : (rays.n_rays + kRPG - 1) / kRPG;
  const long long n_pairs = (n_units + 1) / 2;
  const int n_pass = kStage ? 1 : 2;

  if (tid == 0) {
    for (int s = 0; s < kWStages; ++s) { mbar_init(&sm.w_full[s], rank == 0 ? 2 : 1); mbar_init(&sm.w_empty[s], 1); }
    mbar_init(&sm.stg_full, 2 * kComputeThreads);
    mbar_init(&sm.stg_empty, 1);
    mbar_init(&sm.act_ready, 2 * kComputeThreads);
    mbar_init(&sm.acc_full, 1);
    fence_mbar_init();
  }
  if (warp == kIssuerWarp) {
    tmem_alloc_2cta(&sm.tmem_base, kTmemCols);
    tmem_relinquish_2cta();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = sm.tmem_base;
  unsigned long long pacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  const long long kernel_t0 = prof ? clock64() : 0;

  if (warp == kProducerWarp) {
    // ===================== weight producer (each CTA streams ITS N-half of every fill) =====================
    if (lane == 0) {
      uint32_t wfill = 0;
      for (long long u = cluster_id; u < n_pairs; u += n_clusters) {
        for (int pass = 0; pass < n_pass; ++pass) {
          const int ntiles = kStage ? 1 : tiles_in_pass(pass);
          const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(pass == 0 ? net_c.wstream : net_f.wstream);
          for (int t = 0; t < ntiles; ++t) {
            size_t off = 0;
            for (int L = 0; L < 9; ++L) {
              const int nh = pgn_layer_n(L) / 2, ks_total = pgn_layer_ksteps(L), kpf = pgn_ks_per_fill(L);
              for (int ks = 0; ks < ks_total; ks += kpf) {
                const int nks = min(kpf, ks_total - ks);
                const uint32_t bytes = (uint32_t)nks * nh * 32u;          // this CTA's half of the fill
                const int stage = wfill % kWStages;
                { PROF_T0(); const bool okw = mbar_wait(&sm.w_empty[stage], ((wfill / kWStages) & 1) ^ 1, status, 101); PROF_ADD(4); if (!okw) goto done; }
                mbar_arrive_expect_tx(&sm.w_full[stage], bytes);
                bulk_g2s(sm.wring[stage], wsrc + off + (size_t)rank * bytes, bytes, &sm.w_full[stage]);
                off += 2u * bytes;
                ++wfill;
              }
            }
          }
        }
      }
    }
  } else if (warp == kIssuerWarp) {
    if (lane == 0 && rank == 1) {
      // ===================== peer relay: "my half of fill f has landed" -> leader's w_full =====================
      uint32_t wfill = 0;
      for (long long u = cluster_id; u < n_pairs; u += n_clusters) {
        for (int pass = 0; pass < n_pass; ++pass) {
          const int ntiles = kStage ? 1 : tiles_in_pass(pass);
          for (int t = 0; t < ntiles; ++t) {
            for (int L = 0; L < 9; ++L) {
              const int ks_total = pgn_layer_ksteps(L), kpf = pgn_ks_per_fill(L);
              for (int ks = 0; ks < ks_total; ks += kpf) {
                const int stage = wfill % kWStages;
                if (!mbar_wait(&sm.w_full[stage], (wfill / kWStages) & 1, status, 401)) goto done;
                mbar_arrive_cluster(&sm.w_full[stage], 0);
                ++wfill;
              }
            }
          }
        }
      }
    } else if (lane == 0) {
      // ===================== MMA issuer (leader CTA): UMMA M=256 over both CTAs =====================
      uint32_t wfill = 0, stg_n = 0, act_n = 0;
      const uint32_t act_addr = smem_u32(sm.act), stg_addr = smem_u32(sm.stg);
      for (long long u = cluster_id; u < n_pairs; u += n_clusters) {
        for (int pass = 0; pass < n_pass; ++pass) {
          const int ntiles = kStage ? 1 : tiles_in_pass(pass);
          for (int t = 0; t < ntiles; ++t) {
            for (int L = 0; L < 9; ++L) {
              const int n = pgn_layer_n(L), nh = n / 2, ks_total = pgn_layer_ksteps(L), kpf = pgn_ks_per_fill(L);
              const int ks_act = pgn_layer_kact(L) / 16;
              const int chunk_ks = (L == 8) ? 7 : 9;
              const uint32_t idesc = umma_idesc_bf16(2 * kTM, n);
              if (ks_act > 0) {
                { PROF_T0(); const bool okw = mbar_wait_cluster(&sm.act_ready, act_n & 1, status, 201); PROF_ADD(2); if (!okw) goto done; }
                ++act_n;
                tc_fence_after_sync();
              }
              int stage = 0;
              for (int ks = 0; ks < ks_total; ++ks) {
                const int kf = ks % kpf;
                if (kf == 0) {
                  stage = wfill % kWStages;
                  { PROF_T0(); const bool okw = mbar_wait_cluster(&sm.w_full[stage], (wfill / kWStages) & 1, status, 202); PROF_ADD(0); if (!okw) goto done; }
                  tc_fence_after_sync();
                }
                uint32_t a_addr;
                bool chunk_end = false;
                if (ks < ks_act) {
                  a_addr = act_addr + (uint32_t)ks * 2 * kRunBytes;
                } else {
                  const int e = ks - ks_act, ce = e % chunk_ks;
                  if (ce == 0) {
                    { PROF_T0(); const bool okw = mbar_wait_cluster(&sm.stg_full, stg_n & 1, status, 203); PROF_ADD(1); if (!okw) goto done; }
                    ++stg_n;
                    tc_fence_after_sync();
                  }
                  a_addr = stg_addr + (uint32_t)ce * 2 * kRunBytes;
                  chunk_end = (ce == chunk_ks - 1);
                }
                const uint64_t adesc = umma_smem_desc(a_addr, kRunBytes, 128);
                const uint64_t bdesc = umma_smem_desc(smem_u32(sm.wring[stage]) + (uint32_t)kf * nh * 32u, (uint32_t)nh * 16u, 128);
                { PROF_T0(); umma_bf16_2cta(tmem_base, adesc, bdesc, idesc, ks > 0 ? 1u : 0u); PROF_ADD(13); }
                { PROF_T0();
                  if (chunk_end) umma_commit_2cta(&sm.stg_empty);
                  if (kf == kpf - 1 || ks == ks_total - 1) { umma_commit_2cta(&sm.w_empty[stage]); ++wfill; }
                  PROF_ADD(14); }
              }
              { PROF_T0(); umma_commit_2cta(&sm.acc_full); PROF_ADD(14); }
            }
          }
        }
      }
    }
  } else {
    // ===================== compute warps (encode, epilogue, composite) =====================
    uint32_t stg_n = 0, acc_n = 0;
    const int row = tid & (kTM - 1), half = tid >> 7;
    for (long long u = cluster_id; u < n_pairs; u += n_clusters) {
      const long long unit = 2 * u + rank;                 // this CTA's ray group (or stage tile)
      const long long ray0 = unit * kRPG;
      const int nr = kStage ? 0 : (int)max(0ll, min((long long)kRPG, rays.n_rays - ray0));
      if (!kStage) {
        compute_bar_sync();
        if (tid < kRPG * 3) {
          const int rl = tid / 3, a = tid % 3;
          if (rl < nr) {
            sm.ray_o[rl][a] = rays.ray_batch[(ray0 + rl) * 11 + a];
            sm.ray_d[rl][a] = rays.ray_batch[(ray0 + rl) * 11 + 3 + a];
          } else { sm.ray_o[rl][a] = 0.f; sm.ray_d[rl][a] = (a == 2) ? 1.f : 0.f; }
        }
        compute_bar_sync();
        if (tid < kRPG) {
          const float* d = sm.ray_d[tid];
          sm.dnorm[tid] = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        }
        for (int i = tid; i < kRPG * PGN_S; i += kComputeThreads) {
          const int rl = i / PGN_S, s = i % PGN_S;
          sm.zc[rl][s] = (rl < nr) ? pgn_coarse_z(near_far[(ray0 + rl) * 2], near_far[(ray0 + rl) * 2 + 1], sc.t_coarse[s]) : 0.f;
        }
      }
      for (int pass = 0; pass < n_pass; ++pass) {
        const int S = pass == 0 ? PGN_S : PGN_T;
        const PgnBf16Net& net = pass == 0 ? net_c : net_f;
        const int total_rows = kStage ? kTM : nr * S;
        const int ntiles = kStage ? 1 : tiles_in_pass(pass);
        // epilogue vectors of this net
        compute_bar_sync();
        for (int i = tid; i < 9 * 256; i += kComputeThreads) sm.bias[i] = net.bias[i];
        for (int i = tid; i < 256; i += kComputeThreads) sm.w_alpha[i] = net.w_alpha[i];
        for (int i = tid; i < 384; i += kComputeThreads) sm.w_rgb[i] = net.w_rgb[i];
        compute_bar_sync();

        for (int t = 0; t < ntiles; ++t) {
          TileCtx tc{ray0, S, t * kTM, total_rows};
          const float* enc_rows = kStage ? enc_global + (size_t)unit * kTM * PGN_ENC : nullptr;
          const int rows_valid = kStage ? (int)max(0ll, min((long long)kTM, enc_rows_total - unit * kTM)) : kTM;
          const int tile_ray0 = kStage ? 0 : min(tc.row0 / S, kRPG - 1);
          if (!kStage) {
            // PE table of the joint-frame view directions for the <=3 rays of this tile
            const int tile_ray1 = min((tc.row0 + kTM - 1) / S, nr - 1);
            for (int i = tid; i < kMaxTileRays * PGN_J; i += kComputeThreads) {
              const int tr = i / PGN_J, j = i % PGN_J;
              const int rl = tile_ray0 + tr;
              if (rl <= tile_ray1) {
                const float4* m = reinterpret_cast<const float4*>(pgn_ray_skts(rays, ray0 + rl) + j * 16);
                float dj[3];
                pgn_joint_dir(__ldg(m), __ldg(m + 1), __ldg(m + 2), sm.ray_d[rl], dj[0], dj[1], dj[2]);
                float* tab = &sm.dtab[tr][j * 28];
                for (int k = 0; k < 1 + 2 * PGN_LD; ++k)
                  for (int a = 0; a < 3; ++a) tab[k * 3 + a] = pgn_pe_term(dj[a], k);
                tab[27] = 0.f;
              }
            }
            compute_bar_sync();
          }
          // ---- L0: x chunks
          for (int c = 0; c < 3; ++c) {
            { PROF_T0(); const bool okw = mbar_wait(&sm.stg_empty, (stg_n & 1) ^ 1, status, 301); PROF_ADD(10); if (!okw) goto done; }
            { PROF_T0(); encode_x_chunk<kStage>(sm, rays, sc, tc, c, row, half, false, enc_rows, rows_valid);
              compute_arrive(&sm.stg_full); PROF_ADD(6); }
            ++stg_n;
          }
          for (int L = 0; L < 8; ++L) {
            if (L == 5) {
              for (int c = 0; c < 3; ++c) {
                { PROF_T0(); const bool okw = mbar_wait(&sm.stg_empty, (stg_n & 1) ^ 1, status, 302); PROF_ADD(10); if (!okw) goto done; }
                { PROF_T0(); encode_x_chunk<kStage>(sm, rays, sc, tc, c, row, half, true, enc_rows, rows_valid);
                  compute_arrive(&sm.stg_full); PROF_ADD(6); }
                ++stg_n;
              }
            }
            { PROF_T0(); const bool okw = mbar_wait(&sm.acc_full, acc_n & 1, status, 303); PROF_ADD(9); if (!okw) goto done; }
            ++acc_n;
            tc_fence_after_sync();
            { PROF_T0();
              if (L == 7) epilogue<1>(sm, tmem_base, L, warp, lane);
              else epilogue<0>(sm, tmem_base, L, warp, lane);
              compute_arrive(&sm.act_ready); PROF_ADD(8); }
          }
          // ---- V: d chunks
          compute_bar_sync();               // wcache (written during L5's encode) visible to all
          for (int c = 0; c < 6; ++c) {
            { PROF_T0(); const bool okw = mbar_wait(&sm.stg_empty, (stg_n & 1) ^ 1, status, 304); PROF_ADD(10); if (!okw) goto done; }
            { PROF_T0(); encode_d_chunk<kStage>(sm, tc, c, row, half, tile_ray0, enc_rows, rows_valid);
              compute_arrive(&sm.stg_full); PROF_ADD(7); }
            ++stg_n;
          }
          { PROF_T0(); const bool okw = mbar_wait(&sm.acc_full, acc_n & 1, status, 305); PROF_ADD(9); if (!okw) goto done; }
          ++acc_n;
          tc_fence_after_sync();
          { PROF_T0(); epilogue<2>(sm, tmem_base, 8, warp, lane); PROF_ADD(8); }
          tc_fence_before_sync();
          compute_bar_sync();
          if (tid < kTM) {
            const float r0 = sm.part[0][tid][0] + sm.part[1][tid][0] + net.b_rgb[0];
            const float r1 = sm.part[0][tid][1] + sm.part[1][tid][1] + net.b_rgb[1];
            const float r2 = sm.part[0][tid][2] + sm.part[1][tid][2] + net.b_rgb[2];
            const float sg = sm.part[0][tid][3] + sm.part[1][tid][3] + net.b_alpha[0];
            if (kStage) {
              if (tid < rows_valid) {
                float* o = raw_global + ((size_t)unit * kTM + tid) * 4;
                o[0] = r0; o[1] = r1; o[2] = r2; o[3] = sg;
              }
            } else {
              const int grow = tc.row0 + tid;
              if (grow < total_rows) {
                const int rl = grow / S, s = grow - rl * S;
                float* o = &sm.raw[rl][s * 4];
                o[0] = r0; o[1] = r1; o[2] = r2; o[3] = sg;
              }
            }
          }
          compute_bar_sync();
        }
        if (kStage) continue;
        // ---- compositing (+ resampling after the coarse pass): one warp per ray
        {
          const int rl = warp;
          if (rl < nr) {
            const long long ri = ray0 + rl;
            float rgb3[3], disp, acc;
            if (pass == 0) {
              float* a0 = out.alpha0 ? out.alpha0 + ri * PGN_S : nullptr;
              pgn_composite_warp<PGN_S>(sm.raw[rl], sm.zc[rl], sm.dnorm[rl], sc.density_scale, sc.rgb_eps, lane,
                                        rgb3, &disp, &acc, sm.wts[rl], a0);
              if (lane == 0) {
                if (out.rgb0) { out.rgb0[ri * 3] = rgb3[0]; out.rgb0[ri * 3 + 1] = rgb3[1]; out.rgb0[ri * 3 + 2] = rgb3[2]; }
                if (out.disp0) out.disp0[ri] = disp;
                if (out.acc0) out.acc0[ri] = acc;
              }
              __syncwarp();
              if (out.weights0) { out.weights0[ri * PGN_S + lane] = sm.wts[rl][lane]; out.weights0[ri * PGN_S + lane + 32] = sm.wts[rl][lane + 32]; }
              if (out.raw0) for (int i = lane; i < PGN_S * 4; i += 32) out.raw0[ri * PGN_S * 4 + i] = sm.raw[rl][i];
              pgn_sample_pdf_warp(sm.zc[rl], sm.wts[rl], sc.u_det, lane, sm.scratch[rl],
                                  out.z_samples ? out.z_samples + ri * PGN_I : nullptr, sm.zf[rl],
                                  out.pdf_inds ? out.pdf_inds + ri * PGN_I : nullptr, nullptr);
              if (out.z_fine) for (int i = lane; i < PGN_T; i += 32) out.z_fine[ri * PGN_T + i] = sm.zf[rl][i];
            } else {
              float* a1 = out.alpha ? out.alpha + ri * PGN_T : nullptr;
              pgn_composite_warp<PGN_T>(sm.raw[rl], sm.zf[rl], sm.dnorm[rl], sc.density_scale, sc.rgb_eps, lane,
                                        rgb3, &disp, &acc, nullptr, a1);
              if (lane == 0) {
                if (out.rgb_map) { out.rgb_map[ri * 3] = rgb3[0]; out.rgb_map[ri * 3 + 1] = rgb3[1]; out.rgb_map[ri * 3 + 2] = rgb3[2]; }
                if (out.disp_map) out.disp_map[ri] = disp;
                if (out.acc_map) out.acc_map[ri] = acc;
              }
              if (out.raw) for (int i = lane; i < PGN_T * 4; i += 32) out.raw[ri * PGN_T * 4 + i] = sm.raw[rl][i];
            }
          }
        }
        compute_bar_sync();
      }
    }
  }
done:
  if (prof) {
    unsigned long long* pp = prof + (size_t)blockIdx.x * 16;
    const unsigned long long total = (unsigned long long)(clock64() - kernel_t0);
    if (warp == kIssuerWarp && lane == 0) { pp[0] = pacc[0]; pp[1] = pacc[1]; pp[2] = pacc[2]; pp[3] = total; pp[13] = pacc[13]; pp[14] = pacc[14]; }
    if (warp == kProducerWarp && lane == 0) { pp[4] = pacc[4]; pp[5] = total; }
    if (tid == 0) { for (int i = 6; i <= 10; ++i) pp[i] = pacc[i]; pp[12] = total; }
  }
  tc_fence_before_sync();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();                 // the peer's TMEM/smem must stay alive until every MMA has retired
  if (warp == kIssuerWarp) {
    tc_fence_after_sync();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------ weight packing
// wsrc[l]: device fp32 nn.Linear weights in the include/posegen_b200.h order.
__global__ void pgn_fold_view_kernel(const float* __restrict__ w_view /*[128][904]*/, const float* __restrict__ w_feat /*[256][256]*/,
                                     const float* __restrict__ b_view, const float* __restrict__ b_feat,
                                     float* __restrict__ fold /*[128][256] + [128]*/) {
  // fold[n][k] = sum_m w_view[n][m] * w_feat[m][k];  fold_b[n] = b_view[n] + sum_m w_view[n][m] * b_feat[m]
  const int n = blockIdx.x, k = threadIdx.x;
  float acc = 0.f;
  for (int m = 0; m < 256; ++m) acc = fmaf(w_view[n * 904 + m], w_feat[m * 256 + k], acc);
  fold[n * 256 + k] = acc;
  if (k == 0) {
    float b = b_view[n];
    for (int m = 0; m < 256; ++m) b = fmaf(w_view[n * 904 + m], b_feat[m], b);
    fold[128 * 256 + n] = b;
  }
}

struct PackPtrs { const float* w[12]; const float* b[12]; };

__global__ void pgn_pack_wstream_kernel(PackPtrs p, const float* __restrict__ fold, __nv_bfloat16* __restrict__ wstream,
                                        float* __restrict__ bias, float* __restrict__ w_alpha, float* __restrict__ w_rgb) {
  const size_t total = pgn_wstream_elems();
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    // locate layer
    size_t off = idx;
    int L = 0;
    for (; L < 9; ++L) {
      const size_t le = (size_t)pgn_layer_n(L) * pgn_layer_ksteps(L) * 16;
      if (off < le) break;
      off -= le;
    }
    const int n_l = pgn_layer_n(L), nh = n_l / 2, kpf = pgn_ks_per_fill(L), ks_total = pgn_layer_ksteps(L);
    // within layer: fills of kpf K-steps; within a fill: [cta rank][kstep][khalf][n_local (N/2)][8]
    const size_t full_fill = (size_t)kpf * n_l * 16;
    const int f = (int)(off / full_fill);
    const int nks = min(kpf, ks_total - f * kpf);
    size_t r = off - (size_t)f * full_fill;
    const size_t per_rank = (size_t)nks * nh * 16;
    const int crank = (int)(r / per_rank);
    r -= (size_t)crank * per_rank;
    const int ks = f * kpf + (int)(r / ((size_t)nh * 16));
    r %= (size_t)nh * 16;
    const int kh = (int)(r / ((size_t)nh * 8));
    r %= (size_t)nh * 8;
    const int n = crank * nh + (int)(r >> 3);
    const int e = (int)(r & 7);
    const int kp = ks * 16 + kh * 8 + e;
    float v = 0.f;
    if (L == 0) v = p.w[0][(size_t)n * 432 + pgn_xperm_refcol(kp)];
    else if (L == 5) v = (kp < 256) ? p.w[5][(size_t)n * 688 + 432 + kp] : p.w[5][(size_t)n * 688 + pgn_xperm_refcol(kp - 256)];
    else if (L == 8) {
      if (kp < 256) v = fold[n * 256 + kp];
      else { const int rc = pgn_dperm_refcol(kp - 256); v = rc >= 0 ? p.w[10][(size_t)n * 904 + 256 + (rc - 432)] : 0.f; }
    } else v = p.w[L][(size_t)n * 256 + kp];
    wstream[idx] = __float2bfloat16_rn(v);
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 9 * 256; i += gridDim.x * blockDim.x) {
    const int L = i / 256, n = i % 256;
    float b = 0.f;
    if (L < 8) b = p.b[L][n];
    else if (n < 128) b = fold[128 * 256 + n];
    bias[i] = b;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 256; i += gridDim.x * blockDim.x) w_alpha[i] = p.w[8][i];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 384; i += gridDim.x * blockDim.x) w_rgb[i] = p.w[11][i];
}

}  // namespace

size_t pgn_bf16_wstream_elems() { return pgn_wstream_elems(); }

cudaError_t pgn_pack_bf16_net(const float* const* w_dev, const float* const* b_dev, __nv_bfloat16* wstream,
                              float* bias, float* w_alpha, float* w_rgb, float* fold_tmp, cudaStream_t stream) {
  PackPtrs p;
  for (int i = 0; i < 12; ++i) { p.w[i] = w_dev[i]; p.b[i] = b_dev[i]; }
  pgn_fold_view_kernel<<<128, 256, 0, stream>>>(w_dev[10], w_dev[9], b_dev[10], b_dev[9], fold_tmp);
  pgn_pack_wstream_kernel<<<148 * 4, 256, 0, stream>>>(p, fold_tmp, wstream, bias, w_alpha, w_rgb);
  return cudaGetLastError();
}

static cudaError_t configure_bf16() {
  static bool done = false;
  if (done) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(pgn_render_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem) + 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(pgn_render_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem) + 1024);
  if (e != cudaSuccess) return e;
  done = true;
  return cudaSuccess;
}

cudaError_t pgn_launch_render_bf16(const PgnRayRefs& rays, const PgnOutputs& out, const PgnBf16Net& nc,
                                   const PgnBf16Net& nf, const PgnScalars* sc_dev, const float* near_far,
                                   int* status, unsigned long long* prof, int num_sms, cudaStream_t stream) {
  cudaError_t e = configure_bf16();
  if (e != cudaSuccess) return e;
  const long long n_groups = (rays.n_rays + kRPG - 1) / kRPG;
  if (n_groups == 0) return cudaSuccess;
  const long long n_pairs = (n_groups + 1) / 2;
  const int grid = 2 * (int)min((long long)(num_sms / 2), n_pairs);      // clusters of 2 CTAs
  pgn_render_bf16_kernel<false><<<grid, kThreads, sizeof(Smem) + 1024, stream>>>(rays, out, nc, nf, sc_dev, near_far,
                                                                                 nullptr, 0, nullptr, status, prof);
  return cudaGetLastError();
}

cudaError_t pgn_launch_mlp_bf16(const PgnBf16Net& net, const float* enc, long long m, float* raw,
                                const PgnScalars* sc_dev, int* status, int num_sms, cudaStream_t stream) {
  cudaError_t e = configure_bf16();
  if (e != cudaSuccess) return e;
  const long long n_tiles = (m + kTM - 1) / kTM;
  if (n_tiles == 0) return cudaSuccess;
  const int grid = 2 * (int)min((long long)(num_sms / 2), (n_tiles + 1) / 2);
  PgnRayRefs rays{};
  PgnOutputs out{};
  pgn_render_bf16_kernel<true><<<grid, kThreads, sizeof(Smem) + 1024, stream>>>(rays, out, net, net, sc_dev, nullptr,
                                                                                enc, m, raw, status, nullptr);
  return cudaGetLastError();
}
