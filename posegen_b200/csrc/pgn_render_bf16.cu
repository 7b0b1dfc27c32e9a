// bf16 tensor-core render path (PGN_PRECISION_BF16): fused per-ray-tile pipeline on tcgen05.
//
// Execution model
//   * CTA pairs (clusters of 2, cta_group::2): one UMMA covers M=256 rows (128 per CTA), fp32
//     accumulators in TMEM; the weight (B) operand is split along N between the two CTAs, so each
//     SM streams only half of every layer from L2 (cp.async.bulk rings).
//   * every CTA runs TWO independent tile pipelines ("slots"): each slot owns 256 TMEM columns, a
//     64 KB activation buffer, a weight ring, a weight-producer thread, an MMA-issuer warp and a
//     compute group of 8 warps that does the slot's encoding, epilogues and compositing.  The two
//     slots only share the tensor core: while one slot is in an epilogue or is generating operand
//     chunks, the tensor core runs the other slot's layer (the hardware interleaves the two issue
//     streams), and the CUDA cores of one slot overlap with the tensor work of the other.
//   * a slot renders ray groups of 8 rays: 4 coarse tiles (2 rays x 64 samples, network_fn) and 5
//     fine tiles (8 x 80 samples cut into 128-row tiles, network_fine), in the order
//     C0 C1 F0 C2 F1 C3 F2 F3 F4.  Compositing / inverse-CDF resampling of a tile is DEFERRED into
//     the shadow of the next tile's hidden layers 1-3, so it never sits on the tensor core's
//     critical path; the order guarantees that a fine tile's z values exist before it is encoded.
//   * per 128-row tile, 9 tensor-core layers (K-steps of 16):
//       L0   x_p(480)            -> 256 ReLU   A generated on the fly (skeleton-relative encoding +
//       L1-4 h(256)              -> 256 ReLU     cutoff PE, 4 joints per 20 KB chunk)
//       L5   h(256) | x_p(480)   -> 256 ReLU   (skip concat = two accumulating K ranges)
//       L6-7 h(256)              -> 256 ReLU   (sigma head folded into L7's epilogue, fp32)
//       V    h7(256) | d(768)    -> 128 ReLU   (feature_linear folded into views_linears[0]; rgb head
//                                               folded into the epilogue, fp32)
//     Generated chunks are staged in a 3-deep ring that lives INSIDE the slot's own activation
//     buffer: the buffer is dead while L0 runs and, for L5/V, ring buffer b is free as soon as the
//     activation K-steps under it have been consumed (act_free[b] barriers), so operand generation
//     runs up to three chunks ahead of the tensor core without any dedicated staging memory.
//   * the samples x joints x embedding tensor only ever exists as 20 KB chunks in shared memory and
//     the per-sample network outputs never leave the SM.
//
// Warp roles (640 threads): warps 0-7 = compute group of slot 0, 8-15 = compute group of slot 1
// (warp w owns TMEM lanes 32*(w%4).., column half (w%8)/4), warps 16/17 = weight producers of slot
// 0/1 (one lane issues bulk copies), warps 18/19 = MMA issuers of slot 0/1 in the leader CTA /
// "my half landed" relays in the peer CTA; warp 18 also owns the TMEM allocation.
//
// Reference semantics: core/raycasters.py:361-474, core/encoders.py:8-37,110-122,181-193,
// core/cutoff_embedder.py:111-174, core/networks/nerf.py:94-205, core/utils/ray_utils.py:157-289.
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include "pgn_common.cuh"
#include "pgn_kernels.h"
#include "pgn_umma.cuh"
#include "pgn_bf16_layout.h"

using namespace pgn;

namespace {

constexpr int kGroupThreads = 256;      // threads of one slot's compute group
constexpr int kGroupWarps = kGroupThreads / 32;
constexpr int kThreads = 640;
constexpr int kProducerWarp0 = 16;      // + slot
constexpr int kIssuerWarp0 = 18;        // + slot
constexpr int kRPG = 8;                 // rays per group (large batches); small batches use groups of 4 (template kG)
constexpr int kTM = 128;                // rows per CTA tile (UMMA M = 256 over the pair)
constexpr int kRunBytes = kTM * 16;     // one 8-wide K run of all 128 rows
constexpr int kActBytes = 256 / 8 * kRunBytes;                 // 65536
constexpr int kStgBytes = (PGN_X_CHUNK_K / 8) * kRunBytes;     // 20480
constexpr int kStgBufs = 3;             // staging ring inside act[slot]
constexpr int kWStages = 3;
constexpr int kWStageBytes = 8192;
constexpr int kTmemCols = 512;          // two 256-column accumulators
constexpr int kMaxTileRays = 3;

static_assert(kStgBufs * kStgBytes <= kActBytes, "staging ring must fit inside the activation buffer");

struct __align__(128) Smem {
  uint8_t act[2][kActBytes];                     // per slot: A operand of the hidden layers; first 60 KB double as the staging ring
  uint8_t wring[2][kWStages][kWStageBytes];      // per slot: weight ring (this CTA's N half)
  __nv_bfloat16 wcache[2][PGN_J * kTM];          // d-window per (joint,row), per slot (bf16)
  __nv_bfloat16 dtab[2][kMaxTileRays][PGN_J * 32];   // PE of joint-frame view dirs per ray of the tile (27 + 5 zeros), bf16
  float jtab[2][kMaxTileRays][PGN_J][8];         // per (tile ray, joint): a = R o + t, b = R d, window offsets (see encode)
  float zf[2][kRPG][PGN_T];                      // merged z of the fine pass of the slot's current ray group
  float carry[2][kRPG][8];                       // incremental compositing state of the fine rays
  float part[2][2][kTM][4];                      // per slot, per column half: partial raw rows (rgb_raw, sigma_raw) of the heads
  uint8_t ones[2 * kRunBytes];                   // constant A operand of the bias K-step: k = 0,1 -> 1.0, else 0
  float cscratch[2][2][256];                     // coarse compositing per slot, per warp: z[64] | weights[64] | sample_pdf scratch[128]
  uint64_t w_full[2][kWStages], w_empty[2][kWStages];
  uint64_t stg_full[2][kStgBufs], stg_empty[2][kStgBufs];
  uint64_t act_ready[2], acc_full[2];
  uint64_t act_free[2][kStgBufs];                // activation K-steps under staging buffer b consumed (skip / view layer)
  uint32_t tmem_base;
};
static_assert(sizeof(Smem) + 1024 <= 232448, "shared memory budget of one CTA exceeded");

// tile k of a ray group.  8 rays: C0 C1 F0 C2 F1 C3 F2 F3 F4 (4 coarse tiles of 2 rays, 640 fine rows = 5 tiles);
// 6 rays: C0 C1 F0 C2 F1 F2 F3 (480 fine rows = 3.75 tiles); 4 rays: C0 C1 F0 F1 F2 (320 fine rows = 2.5 tiles, the last
// one half empty).  A fine tile always comes at least two tiles after the coarse tile of the last ray it touches (that
// tile's compositing / resampling runs in the shadow of the tile after it).  Small batches use the smaller groups: a
// 3,072-ray training batch is 192 pair-units of 8 rays for 148 tile pipelines - two rounds of 9 tiles, the second one with
// most pipelines idle - or 384 pair-units of 4 (three rounds of 5 tiles) or 256 of 6 (two rounds of 7 tiles);
// pgn_bf16_group_rays decides, from the ray count alone.
template <int kG>
__device__ __forceinline__ void tile_of(int k, int& pass, int& t) {
  if (kG == 8) {
    pass = (0x1D4 >> k) & 1;
    t = (int)((0x432312010ull >> (4 * k)) & 0xF);
  } else if (kG == 6) {
    pass = (0x74 >> k) & 1;                          // C C F C F F F
    t = (int)((0x3212010u >> (4 * k)) & 0xF);        // 0 1 0 2 1 2 3
  } else {
    pass = k >= 2 ? 1 : 0;
    t = k >= 2 ? k - 2 : k;
  }
}
template <int kG> constexpr int tiles_per_group() { return kG == 8 ? 9 : (kG == 6 ? 7 : 5); }

struct TileCtx {
  long long ray0;     // first ray of this CTA's group
  long long unit;     // work unit (ray group; stage mode: 128-row tile)
  int nr;             // valid rays in the group (0..8)
  int S;              // samples per ray in this pass
  int pass;
  int t;              // tile index inside the pass
  int row0;           // first row of the tile within the group pass
  int total_rows;     // valid rows in the group pass
  int tile_ray0;      // first ray (local) touched by the tile
};

// ------------------------------------------------------------------ encode (compute group)
// Per tile, PRE(L0) builds two small tables for the <=3 rays the tile touches:
//   jtab[ray][joint] = { a = R_j o + t_j, b = R_j d, -tau_v c_j log2e, -tau_d c_j log2e }   (8 floats)
//       so that a sample's joint-local position is a + z b (3 FMAs) instead of a 3x4 transform of o + z d;
//   dtab[ray][joint] = PE of the normalised joint-frame view direction (27 halfs + 5 zeros).
// Row state (valid, z, tile-ray index) is computed once per tile and reused by L0, L5 and V.
struct RowCtx { bool valid; int tr; float z; };

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ RowCtx make_row_ctx(const Smem& sm, const PgnScalars& sc, const TileCtx& tc,
                                               const float* __restrict__ near_far, int slot, int row,
                                               const float* __restrict__ t_rand = nullptr, int lindisp = 0) {
  RowCtx rc;
  const int grow = tc.row0 + row;
  rc.valid = grow < tc.total_rows;
  rc.tr = 0;
  rc.z = 0.f;
  if (rc.valid) {
    const int rl = grow / tc.S, s = grow - rl * tc.S;
    const long long ri = tc.ray0 + rl;
    rc.tr = rl - tc.tile_ray0;
    if (tc.pass != 0) rc.z = sm.zf[slot][rl][s];
    else if (t_rand) rc.z = pgn_coarse_z_jitter(__ldg(near_far + ri * 2), __ldg(near_far + ri * 2 + 1), sc.t_coarse, s, __ldg(t_rand + ri * PGN_S + s), lindisp);
    else rc.z = pgn_coarse_z(__ldg(near_far + ri * 2), __ldg(near_far + ri * 2 + 1), sc.t_coarse[s], lindisp);
  }
  return rc;
}

// x chunk c (joints 4c..4c+3): thread (row, half) produces joints 4c+2*half, +1 -> 36 values + 4 zeros
// = 5 runs of 8 at run index half*5+r.  The valid path is one basic block so the two joints interleave.
template <bool kWriteW>
__device__ __forceinline__ void encode_x_fast(uint32_t wcache_saddr, uint32_t jtab_saddr, const RowCtx& rc, float tau_v2, float tau_d2,
                                              int chunk, int row, int half, uint32_t (&packed)[20]) {
  const int j0 = chunk * 4 + half * 2;
  if (rc.valid) {
    const uint32_t jt = jtab_saddr + (uint32_t)(rc.tr * PGN_J + j0) * 32u;
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const uint4 qa = lds128(jt + jj * 32), qb = lds128(jt + jj * 32 + 16);
      const float x = fmaf(rc.z, __uint_as_float(qa.w), __uint_as_float(qa.x));
      const float y = fmaf(rc.z, __uint_as_float(qb.x), __uint_as_float(qa.y));
      const float z = fmaf(rc.z, __uint_as_float(qb.y), __uint_as_float(qa.z));
      const float n2 = fmaf(z, z, fmaf(y, y, x * x));
      const float rsq = rsqrtf(fmaxf(n2, 1e-24f));
      const float v = n2 * rsq;
      const float w = __fdividef(1.0f, 1.0f + ex2_approx(fmaf(tau_v2, v, __uint_as_float(qb.z))));   // 1 - sigmoid(tau (v - c))
      if (kWriteW) {
        const float wd = __fdividef(1.0f, 1.0f + ex2_approx(fmaf(tau_d2, v, __uint_as_float(qb.w))));
        sts16(wcache_saddr + (uint32_t)((j0 + jj) * kTM + row) * 2u, __bfloat16_as_ushort(__float2bfloat16_rn(wd)));
      }
      float sn, cs;
      __sincosf(v, &sn, &cs);
      float vals[18];
      vals[0] = v * w;
#pragma unroll
      for (int f = 0; f < PGN_LV; ++f) {
        vals[1 + 2 * f] = sn * w;
        vals[2 + 2 * f] = cs * w;
        const float t2 = cs + cs;                 // double angle: sin 2a = (2 cos a) sin a, cos 2a = (2 cos a) cos a - 1
        sn = t2 * sn;
        cs = fmaf(t2, cs, -1.0f);
      }
      vals[15] = x * rsq; vals[16] = y * rsq; vals[17] = z * rsq;
#pragma unroll
      for (int i = 0; i < 9; ++i) packed[jj * 9 + i] = pack_bf16x2(vals[2 * i], vals[2 * i + 1]);
    }
  } else {
    if (kWriteW) { sts16(wcache_saddr + (uint32_t)(j0 * kTM + row) * 2u, 0); sts16(wcache_saddr + (uint32_t)((j0 + 1) * kTM + row) * 2u, 0); }
#pragma unroll
    for (int i = 0; i < 18; ++i) packed[i] = 0u;
  }
  packed[18] = 0u; packed[19] = 0u;
}
// stage mode (pgn_mlp): the A operand comes from explicit encodings in global memory
__device__ __forceinline__ void encode_x_stage(int chunk, int row, int half, const float* __restrict__ enc_rows, int rows_valid,
                                               uint32_t (&packed)[20]) {
#pragma unroll
  for (int i = 0; i < 20; ++i) {
    const int kp = chunk * PGN_X_CHUNK_K + half * 40 + 2 * i;
    const int ca = pgn_xperm_refcol(kp), cb = pgn_xperm_refcol(kp + 1);
    float a = 0.f, b = 0.f;
    if (row < rows_valid) {
      if (ca >= 0) a = enc_rows[(size_t)row * PGN_ENC + ca];
      if (cb >= 0) b = enc_rows[(size_t)row * PGN_ENC + cb];
    }
    packed[i] = pack_bf16x2(a, b);
  }
}
__device__ __forceinline__ void encode_x_store(uint32_t stg, int row, int half, const uint32_t (&packed)[20]) {
  const uint32_t base = stg + (uint32_t)(half * 5) * kRunBytes + row * 16;
#pragma unroll
  for (int r = 0; r < 5; ++r) sts128(base + r * kRunBytes, packed[4 * r], packed[4 * r + 1], packed[4 * r + 2], packed[4 * r + 3]);
}

// d chunk c (joints 2c, 2c+1): thread (row, half) produces joint 2c+half -> 27 values + 5 zeros
// = 4 runs at run index half*4+r.
__device__ __forceinline__ void encode_d_fast(uint32_t wcache_saddr, uint32_t dtab_saddr, const RowCtx& rc, int chunk, int row, int half,
                                              uint32_t (&packed)[20]) {
  // d_emb[row, (j, t)] = w_d[row, j] * PE(dir_j)[t]: both factors are kept in bf16 and multiplied pairwise
  // (one HMUL2.BF16 per two operand values; the product is the bf16 A-operand element)
  const int j = chunk * 2 + half;
  const uint32_t wbits = rc.valid ? (uint32_t)lds16(wcache_saddr + (uint32_t)(j * kTM + row) * 2u) : 0u;
  const uint32_t w2u = wbits | (wbits << 16);
  const __nv_bfloat162 w2 = *reinterpret_cast<const __nv_bfloat162*>(&w2u);
  const uint32_t tab = dtab_saddr + (uint32_t)(rc.tr * PGN_J + j) * 64u;   // 32 bf16 = 4 x 16 B
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 t = lds128(tab + i * 16);
    const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 p = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]), w2);
      packed[i * 4 + e] = *reinterpret_cast<const uint32_t*>(&p);
    }
  }
}
__device__ __forceinline__ void encode_d_stage(int chunk, int row, int half, const float* __restrict__ enc_rows, int rows_valid,
                                               uint32_t (&packed)[20]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int q = chunk * PGN_D_CHUNK_K + half * 32 + 2 * i;
    const int ca = pgn_dperm_refcol(q), cb = pgn_dperm_refcol(q + 1);
    float a = 0.f, b = 0.f;
    if (row < rows_valid) {
      if (ca >= 0) a = enc_rows[(size_t)row * PGN_ENC + ca];
      if (cb >= 0) b = enc_rows[(size_t)row * PGN_ENC + cb];
    }
    packed[i] = pack_bf16x2(a, b);
  }
}
__device__ __forceinline__ void encode_d_store(uint32_t stg, int row, int half, const uint32_t (&packed)[20]) {
  const uint32_t base = stg + (uint32_t)(half * 4) * kRunBytes + row * 16;
#pragma unroll
  for (int r = 0; r < 4; ++r) sts128(base + r * kRunBytes, packed[4 * r], packed[4 * r + 1], packed[4 * r + 2], packed[4 * r + 3]);
}

// 256-bit global store (sm_100: STG.E.256): one full sector per thread, 32-byte aligned address
__device__ __forceinline__ void stg256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
               "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// 128-bit streaming global store
__device__ __forceinline__ void stg128(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.cs.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ------------------------------------------------------------------ epilogue (compute group)
// The bias is already in the accumulator (bias K-step), so a hidden layer is TMEM -> ReLU+bf16 -> smem.
// MODE 0: hidden layer -> act.  MODE 1: same + sigma head (fp32).  MODE 2: view layer -> rgb head (fp32).
// TMEM loads of the next 32-column batch are issued before the current batch is processed.
template <int MODE>
__device__ __forceinline__ void epilogue(Smem& sm, uint32_t tmem_acc, uint32_t act_saddr, int slot, const float* __restrict__ w_alpha,
                                         const float* __restrict__ w_rgb, int warp, int lane, float& sig_keep,
                                         uint4* dump = nullptr, uint32_t* dump_mask = nullptr, uint32_t* dump_vmask = nullptr,
                                         size_t mask_stride = 0,
                                         const float* __restrict__ fc = nullptr) {
  // fc (view layer only): this row's frame-code term [128] (Optcodes: W_v[:, 904:920] code[cam], fp32), added to the
  // pre-activation before the ReLU; nullptr when the model has no frame codes
  // dump (training forward only).  Trunk layers (MODE 0/1): this thread's row inside its tile of the TILE-BLOCKED dump
  // [rows / 128][256 / 8][128][8] bf16 - the same image the epilogue writes into shared memory - so the 16-byte store of
  // one 8-column run by the 32 lanes of a warp (32 consecutive rows) is 512 contiguous bytes: 4 full lines per store
  // instruction instead of the 32 sectors in 32 lines of a row-major dump (which made the training forward 1.75 instead
  // of 1.0 ms: the epilogue, on each slot's critical path, was bound by the LSU's line throughput).  View layer
  // (MODE 2): row-major [rows,128], 64 consecutive columns of this thread's row.
  const int q = warp & 3, half = warp >> 2;
  const int row = q * 32 + lane;
  constexpr int kCols = (MODE == 2) ? 64 : 128;        // columns per thread
  constexpr int kBatches = kCols / 16;
  const int col0 = half * kCols;
  const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)col0;
  const uint32_t dst0 = act_saddr + (uint32_t)(col0 >> 3) * kRunBytes + row * 16;
  float sig = 0.f, r0 = 0.f, r1 = 0.f, r2 = 0.f;
  uint32_t mask_w = 0u;                        // training forward: [column > 0] bits of the current 32 columns (stored per word:
                                               // one live register instead of four across the drain loop)
  uint32_t v[2][16];
  tmem_ld_32x16(taddr, v[0]);
#pragma unroll
  for (int b = 0; b < kBatches; ++b) {
    tmem_ld_wait();
    if (b + 1 < kBatches) tmem_ld_32x16(taddr + (uint32_t)(b + 1) * 16, v[(b + 1) & 1]);
    const int c0 = col0 + b * 16;
    uint32_t* vb = v[b & 1];
    if (MODE == 2 && fc) {
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(fc + c0) + i4);
        vb[4 * i4] = __float_as_uint(__uint_as_float(vb[4 * i4]) + f.x); vb[4 * i4 + 1] = __float_as_uint(__uint_as_float(vb[4 * i4 + 1]) + f.y);
        vb[4 * i4 + 2] = __float_as_uint(__uint_as_float(vb[4 * i4 + 2]) + f.z); vb[4 * i4 + 3] = __float_as_uint(__uint_as_float(vb[4 * i4 + 3]) + f.w);
      }
    }
    if (MODE == 1) {
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        const float4 wa = __ldg(reinterpret_cast<const float4*>(w_alpha + c0) + i4);
        sig = fmaf(fmaxf(__uint_as_float(vb[4 * i4]), 0.f), wa.x, sig);
        sig = fmaf(fmaxf(__uint_as_float(vb[4 * i4 + 1]), 0.f), wa.y, sig);
        sig = fmaf(fmaxf(__uint_as_float(vb[4 * i4 + 2]), 0.f), wa.z, sig);
        sig = fmaf(fmaxf(__uint_as_float(vb[4 * i4 + 3]), 0.f), wa.w, sig);
      }
    }
    if (MODE == 2) {
#pragma unroll
      for (int i4 = 0; i4 < 4; ++i4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(w_rgb + c0) + i4);
        const float4 g = __ldg(reinterpret_cast<const float4*>(w_rgb + 128 + c0) + i4);
        const float4 c = __ldg(reinterpret_cast<const float4*>(w_rgb + 256 + c0) + i4);
        const float x0 = fmaxf(__uint_as_float(vb[4 * i4]), 0.f), x1 = fmaxf(__uint_as_float(vb[4 * i4 + 1]), 0.f);
        const float x2 = fmaxf(__uint_as_float(vb[4 * i4 + 2]), 0.f), x3 = fmaxf(__uint_as_float(vb[4 * i4 + 3]), 0.f);
        r0 = fmaf(x0, a.x, r0); r0 = fmaf(x1, a.y, r0); r0 = fmaf(x2, a.z, r0); r0 = fmaf(x3, a.w, r0);
        r1 = fmaf(x0, g.x, r1); r1 = fmaf(x1, g.y, r1); r1 = fmaf(x2, g.z, r1); r1 = fmaf(x3, g.w, r1);
        r2 = fmaf(x0, c.x, r2); r2 = fmaf(x1, c.y, r2); r2 = fmaf(x2, c.z, r2); r2 = fmaf(x3, c.w, r2);
      }
    } else {
      uint32_t pk[8];
#pragma unroll
      for (int g = 0; g < 2; ++g) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          pk[4 * g + i] = pack_relu_bf16x2(__uint_as_float(vb[8 * g + 2 * i]), __uint_as_float(vb[8 * g + 2 * i + 1]));
        sts128(dst0 + (uint32_t)(2 * b + g) * kRunBytes, pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
      }
#ifndef PGN_EXP_NOSTG      // (-DPGN_EXP_NOSTG: timing experiment, the trunk activations are not stored)
      if (dump) {                                                // runs (col0 / 8 + 2 b) and (+ 1): 128 rows x 16 B = 2 KB apart
        uint4* d = dump + (size_t)((col0 >> 3) + 2 * b) * 128;
        stg128(d, pk[0], pk[1], pk[2], pk[3]);
        stg128(d + 128, pk[4], pk[5], pk[6], pk[7]);
      }
#endif
#ifndef PGN_EXP_NOMASK     // (-DPGN_EXP_NOMASK: timing experiment, no mask bits)
      if (dump_mask) {
        // ReLU mask of the 16 columns (what the fused delta chain of the backward reads instead of the activations):
        // one funnel shift per column collects the accumulators' sign bits (last column first, so column 0 ends up in
        // bit 0); active = not negative (a positive fp32 stays positive in bf16; an exact +0 counts as active, its
        // activation and hence its delta's effect are zero anyway)
        uint32_t mb = 0;
#pragma unroll
        for (int i = 15; i >= 0; --i) mb = __funnelshift_l(vb[i], mb, 1);
        // word planes ([layer][word][row], include/posegen_b200.h): the 32 lanes of a warp (32 consecutive rows) write one
        // 128-byte line per word instead of 4 bytes in each of 32 sectors
        if (b & 1) dump_mask[(size_t)(b >> 1) * mask_stride] = mask_w | ((~mb & 0xffffu) << 16);
        else mask_w = ~mb & 0xffffu;
      }
#endif
    }
    if (MODE == 2 && dump) {      // view layer: relu(g) of this thread's 16 columns
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pk[i] = pack_relu_bf16x2(__uint_as_float(vb[2 * i]), __uint_as_float(vb[2 * i + 1]));
      stg256(dump + (col0 >> 3) + 2 * b, pk);
    }
    if (MODE == 2 && dump_vmask) {   // masks-only dump: [g > 0] of this thread's 64 view-layer columns as bits
      uint32_t mb = 0;
#pragma unroll
      for (int i = 15; i >= 0; --i) mb = __funnelshift_l(vb[i], mb, 1);
      if (b & 1) dump_vmask[(size_t)(b >> 1) * mask_stride] = mask_w | ((~mb & 0xffffu) << 16);
      else mask_w = ~mb & 0xffffu;
    }
  }
  // heads: this thread's column half of (rgb_raw, sigma_raw) -> one conflict-free 16-byte store per tile
  if (MODE == 1) sig_keep = sig;
  if (MODE == 2) sts128(smem_u32(&sm.part[slot][half][row][0]), __float_as_uint(r0), __float_as_uint(r1), __float_as_uint(r2), __float_as_uint(sig_keep));
}

// "my part of the A operand is written / my TMEM reads are done" -> the LEADER CTA's barrier
// (one elected arrive per warp: every lane fences its own writes, __syncwarp orders them before lane 0's release)
__device__ __forceinline__ void compute_arrive(uint32_t bar, int lane) {
  tc_fence_before_sync();
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster_s(bar, 0);
}
// chunk hand-off: plain shared-memory writes only (no tcgen05 operation of this thread to order)
__device__ __forceinline__ void chunk_arrive(uint32_t bar, int lane) {
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) mbar_arrive_cluster_s(bar, 0);
}
__device__ __forceinline__ void group_bar_sync(int slot) { asm volatile("bar.sync %0, 256;\n" ::"r"(slot + 1) : "memory"); }

// ------------------------------------------------------------------ MMA issue (one warp per slot, leader CTA)
// Layer kinds and their static structure (pgn_bf16_layout.h): K-steps, activation K-steps, chunk size, fill size,
// and the weight-ring stage the layer starts on (fills per tile: L0 16 | H 9 | L5 24 | V 17; 111 = 0 mod 3, so every
// layer kind always starts on the same ring stage; chunks per layer are multiples of 3, so chunk c always
// uses staging buffer c % 3).
constexpr int kKindL0 = 0, kKindH = 1, kKindL5 = 2, kKindV = 3;
static_assert(kWStages == 3 && kStgBufs == 3, "issue_layer's static ring schedule assumes 3-deep rings");
static_assert((pgn_layer_ksteps(0) + 1) / 2 == 16 && (pgn_layer_ksteps(1) + 1) / 2 == 9 && (pgn_layer_ksteps(5) + 1) / 2 == 24 &&
              (pgn_layer_ksteps(8) + 3) / 4 == 17, "fills per layer changed: recompute the static ring stages");

struct IssuerCtx {
  uint32_t w_full0, w_empty0, stg_full0, stg_empty0, acc_full, act_free;
  uint32_t act_lo, ones_lo, ring_lo, tmem_acc;
  uint32_t wph, sph;            // per-stage / per-buffer phase bits
  volatile int* status;
};

template <int KIND>
__device__ __forceinline__ bool issue_layer(IssuerCtx& ic) {
  constexpr int L = KIND == kKindL0 ? 0 : (KIND == kKindH ? 1 : (KIND == kKindL5 ? 5 : 8));
  constexpr int n = pgn_layer_n(L), nh = n / 2;
  constexpr int ks_total = pgn_layer_ksteps(L), kpf = pgn_ks_per_fill(L);
  constexpr int ks_act = pgn_layer_kact(L) / 16;
  constexpr int chunk_ks = pgn_layer_chunk_ks(L);
  constexpr bool has_chunks = pgn_layer_chunks(L) != 0;
  constexpr int stage0 = KIND == kKindL0 ? 0 : 1;
  constexpr uint32_t idesc = umma_idesc_bf16(2 * kTM, n);
  constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);                 // SBO = 128 B, descriptor version 1
  constexpr uint32_t b_lbo = ((nh * 16u) >> 4) << 16;                    // B: LBO = (N/2)*16 B
  constexpr uint32_t b_step = (nh * 32u) >> 4;                           // one K-step of this CTA's B half
  constexpr uint32_t kAStep = (2 * kRunBytes) >> 4;                      // one K-step of A (two runs)
#pragma unroll
  for (int ks = 0; ks < ks_total; ++ks) {
    const int f = ks / kpf, st = (stage0 + f) % kWStages;
    if (ks % kpf == 0) {                                                  // ---- next weight fill
      if (!mbar_wait_s(ic.w_full0 + st * 8, (ic.wph >> st) & 1u, ic.status, 202)) return false;
      ic.wph ^= 1u << st;
      tc_fence_after_sync();
    }
    uint32_t a_lo;
    bool chunk_end = false;
    int sb = 0;
    if (ks == ks_total - 1) {
      a_lo = ic.ones_lo;                                                  // bias K-step
    } else if (ks < ks_act) {
      a_lo = ic.act_lo + (uint32_t)ks * kAStep;
    } else {
      const int cs = ks - ks_act, c = cs / chunk_ks, ce = cs % chunk_ks;
      sb = c % kStgBufs;
      if (ce == 0) {
        if (!mbar_wait_s(ic.stg_full0 + sb * 8, (ic.sph >> sb) & 1u, ic.status, 203)) return false;
        ic.sph ^= 1u << sb;
        tc_fence_after_sync();
      }
      a_lo = ic.act_lo + (uint32_t)sb * (uint32_t)(kStgBytes >> 4) + (uint32_t)ce * kAStep;
      chunk_end = ce == chunk_ks - 1;
    }
    const uint32_t b_lo = (ic.ring_lo + (uint32_t)st * (kWStageBytes >> 4) + (uint32_t)(ks % kpf) * b_step) | b_lbo;
    umma_bf16_2cta_elect(ic.tmem_acc, ((uint64_t)kDescHi << 32) | a_lo, ((uint64_t)kDescHi << 32) | b_lo, idesc, ks > 0 ? 1u : 0u);
    if (chunk_end) umma_commit_2cta_elect_s(ic.stg_empty0 + sb * 8);
    // activation K-steps 0-4 / 5-9 / 10-14 consumed -> staging buffer 0 / 1 / 2 (20 KB = 5 K-steps each) may be overwritten
    if (has_chunks && ks_act > 0 && (ks == 4 || ks == 9 || ks == 14)) umma_commit_2cta_elect_s(ic.act_free + (ks / 5) * 8);
    if (ks % kpf == kpf - 1 || ks == ks_total - 1) umma_commit_2cta_elect_s(ic.w_empty0 + st * 8);
  }
  umma_commit_2cta_elect_s(ic.acc_full);
  return true;
}

// optional phase timers (cycles, one elected thread per role of SLOT 0, accumulated per CTA):
//  3 issuer total | 4 producer wait w_empty | 5 producer total
//  6 encode_x | 7 encode_d | 8 epilogue | 9 wait acc_full | 10 wait stg_empty | 11 composite | 12 compute total
//  13 wait act_free | 14 per-tile tables | 15 chunk store + arrive
#define PROF_T0() const long long _pt0 = kProf ? clock64() : 0
#define PROF_ADD(slot) do { if (kProf) pacc[slot] += (unsigned long long)(clock64() - _pt0); } while (0)

// ------------------------------------------------------------------ the kernel
// kDump: 0 = inference, 1 = training forward (activation + mask dump, training-time randomness), 2 = masks only
// (deterministic sampling; the fine pass's ReLU masks for the pose gradient through a frozen network)
// kFC: the model has Optcodes frame codes (the view layer's epilogue adds the ray's code term).  A separate instantiation
// (compiled in its own translation unit, pgn_render_bf16_fc.cu) because the hooks cost 2.4 % of the frame-code-free
// kernel when they are only switched off at run time (4.38 vs 4.49 M rays/s, A/B on one box).
template <bool kStage, bool kProf, int kDump, int kG = 8, bool kFC = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
pgn_render_bf16_kernel(PgnRayRefs rays, PgnOutputs out, PgnBf16Net net_c, PgnBf16Net net_f,
                       const PgnScalars* __restrict__ scp, const float* __restrict__ near_far,
                       const float* __restrict__ enc_global, long long enc_rows_total, float* __restrict__ raw_global,
                       int* __restrict__ status_g, unsigned long long* __restrict__ prof, PgnActDump dump) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  volatile int* status = status_g;
  // volatile reads: keeps tid / lane in registers instead of re-reading the special registers (S2R) in hot loops
  int tid, lane;
  asm volatile("mov.u32 %0, %%tid.x;\n" : "=r"(tid));
  asm volatile("mov.u32 %0, %%laneid;\n" : "=r"(lane));
  const int warp = tid >> 5;
  const uint32_t sm_base = smem_u32(&sm);
#define SADDR(field) (sm_base + (uint32_t)offsetof(Smem, field))
  const uint32_t rank = cluster_ctarank();
  const PgnScalars& sc = *scp;
  const long long n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;

  // work units: ray groups of 8 (stage mode: 128-row tiles); a pair-unit = one unit per CTA of the pair.
  // Pair-units are dealt round-robin to clusters; inside a cluster, alternately to the two slots.
  const long long n_units = kStage ? (enc_rows_total + kTM - 1) / kTM : (rays.n_rays + kG - 1) / kG;
  const long long n_pairs = (n_units + 1) / 2;
  const int n_local = (int)((n_pairs > cluster_id) ? (n_pairs - cluster_id + n_clusters - 1) / n_clusters : 0);
  constexpr int kTiles = kStage ? 1 : tiles_per_group<kG>();

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      for (int i = 0; i < kWStages; ++i) { mbar_init(&sm.w_full[s][i], rank == 0 ? 2 : 1); mbar_init(&sm.w_empty[s][i], 1); }
      for (int i = 0; i < kStgBufs; ++i) { mbar_init(&sm.stg_full[s][i], 2 * kGroupWarps); mbar_init(&sm.stg_empty[s][i], 1); }
      mbar_init(&sm.act_ready[s], 2 * kGroupWarps);
      mbar_init(&sm.acc_full[s], 1);
      for (int i = 0; i < kStgBufs; ++i) mbar_init(&sm.act_free[s][i], 1);
    }
    fence_mbar_init();
  }
  if (warp == kIssuerWarp0) {
    tmem_alloc_2cta(&sm.tmem_base, kTmemCols);
    tmem_relinquish_2cta();
  }
  if (tid < kTM) {     // constant A operand of the bias K-step
    *reinterpret_cast<uint4*>(sm.ones + tid * 16) = make_uint4(0x3F803F80u, 0u, 0u, 0u);      // bf16 (1.0, 1.0, 0...)
    *reinterpret_cast<uint4*>(sm.ones + kRunBytes + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = sm.tmem_base;
  unsigned long long pacc[32] = {};
  const long long kernel_t0 = kProf ? clock64() : 0;

  if (warp >= kProducerWarp0 && warp < kProducerWarp0 + 2) {
    // ===================== weight producer of slot s (each CTA streams ITS N-half of every fill) =====================
    const int s = warp - kProducerWarp0;
    const int n_slot = (n_local + 1 - s) / 2;
    if (lane == 0) {
      const uint32_t w_full0 = SADDR(w_full) + s * kWStages * 8, w_empty0 = SADDR(w_empty) + s * kWStages * 8;
      const uint32_t ring0 = SADDR(wring) + s * kWStages * kWStageBytes;
      uint32_t stage = 0, wphase = 1;          // "empty" barriers start released
      for (int i = 0; i < n_slot; ++i) {
        for (int k = 0; k < kTiles; ++k) {
          int pass, t;
          tile_of<kG>(k, pass, t);
          const uint8_t* wbase = reinterpret_cast<const uint8_t*>((kStage || pass == 0) ? net_c.wstream : net_f.wstream);
          const uint8_t* src = wbase;          // layers are contiguous in consumption order
          for (int L = 0; L < 9; ++L) {
            const int nh = pgn_layer_n(L) / 2, ks_total = pgn_layer_ksteps(L), kpf = pgn_ks_per_fill(L);
            for (int ks = 0; ks < ks_total; ks += kpf) {
              const int nks = min(kpf, ks_total - ks);
              const uint32_t bytes = (uint32_t)nks * nh * 32u;          // this CTA's half of the fill
              { PROF_T0(); const bool okw = mbar_wait_s(w_empty0 + stage * 8, wphase, status, 101); PROF_ADD(4); if (!okw) goto done; }
              mbar_arrive_expect_tx_s(w_full0 + stage * 8, bytes);
              bulk_g2s_s(ring0 + stage * kWStageBytes, src + (size_t)rank * bytes, bytes, w_full0 + stage * 8);
              src += 2u * bytes;
              if (++stage == kWStages) { stage = 0; wphase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp >= kIssuerWarp0) {
    const int s = warp - kIssuerWarp0;
    const int n_slot = (n_local + 1 - s) / 2;
    if (rank == 1) {
      // ===================== peer relay: "my half of fill f has landed" -> leader's w_full =====================
      if (lane == 0) {
        const uint32_t w_full0 = SADDR(w_full) + s * kWStages * 8;
        uint32_t stage = 0, wphase = 0;
        for (int i = 0; i < n_slot; ++i)
          for (int k = 0; k < kTiles; ++k)
            for (int L = 0; L < 9; ++L) {
              const int ks_total = pgn_layer_ksteps(L), kpf = pgn_ks_per_fill(L);
              for (int ks = 0; ks < ks_total; ks += kpf) {
                if (!mbar_wait_s(w_full0 + stage * 8, wphase, status, 401)) goto done;
                mbar_arrive_cluster_s(w_full0 + stage * 8, 0);
                if (++stage == kWStages) { stage = 0; wphase ^= 1; }
              }
            }
      }
    } else {
      // ===================== MMA issuer of slot s (leader CTA): UMMA M=256 over both CTAs =====================
      // The whole warp runs this code convergently; one elected lane issues each tcgen05 instruction.
      // A single warp retires roughly one instruction per 5-6 cycles, so a K-step (131 cycles of tensor
      // time at N=256) leaves room for ~20 instructions: every layer kind is fully unrolled
      // (issue_layer<KIND>) so that operand descriptors and barrier addresses are base + immediate.
      IssuerCtx ic;
      ic.w_full0 = SADDR(w_full) + s * kWStages * 8; ic.w_empty0 = SADDR(w_empty) + s * kWStages * 8;
      ic.stg_full0 = SADDR(stg_full) + s * kStgBufs * 8; ic.stg_empty0 = SADDR(stg_empty) + s * kStgBufs * 8;
      ic.acc_full = SADDR(acc_full) + s * 8; ic.act_free = SADDR(act_free) + s * kStgBufs * 8;
      const uint32_t a_lbo = (uint32_t)(kRunBytes >> 4) << 16;           // A: LBO = 2048 B
      ic.act_lo = ((SADDR(act) + s * kActBytes) >> 4) | a_lbo;
      ic.ones_lo = (SADDR(ones) >> 4) | a_lbo;
      ic.ring_lo = (SADDR(wring) + s * kWStages * kWStageBytes) >> 4;
      ic.tmem_acc = tmem_base + (uint32_t)s * 256u;
      ic.wph = 0; ic.sph = 0;
      ic.status = status;
      const uint32_t act_ready_a = SADDR(act_ready) + s * 8;
      uint32_t jobs = 0;                     // jobs already issued (act_ready phase)
      for (int i = 0; i < n_slot; ++i) {
        for (int k = 0; k < kTiles; ++k) {
#pragma unroll 1
          for (int L = 0; L < 9; ++L) {
            // the previous job of this slot must have been drained (its epilogue wrote act[s] / freed the accumulator)
            if (jobs > 0) {
              PROF_T0();
              if (!mbar_wait_s(act_ready_a, (jobs - 1) & 1, status, 201)) goto done;
              tc_fence_after_sync();
              PROF_ADD(2);
            }
            ++jobs;
            const long long t_job0 = kProf ? clock64() : 0;
            bool ok;
            if (L == 0) ok = issue_layer<kKindL0>(ic);
            else if (L == 5) ok = issue_layer<kKindL5>(ic);
            else if (L == 8) ok = issue_layer<kKindV>(ic);
            else ok = issue_layer<kKindH>(ic);
            if (!ok) goto done;
            if (kProf && s == 0 && L != 0 && L != 5 && L != 8) pacc[28] += (unsigned long long)(clock64() - t_job0);   // hidden layers: issue time
          }
        }
      }
    }
  } else {
    // ===================== compute group of slot s =====================
    // Runs the slot's tiles in order; per layer job: PRE (generated A-operand chunks), then POST
    // (accumulator drain -> next layer's A operand / heads).  The compositing of tile n is cut into three
    // stages that run between PRE and POST of layers 1..3 of tile n+1, i.e. while the tensor core works on
    // that tile's hidden layers.
    const int s = warp >> 3;
    const int n_slot = (n_local + 1 - s) / 2;
    const int gtid = tid - s * kGroupThreads;     // thread index inside the group
    const int gwarp = gtid >> 5;
    const int row = gtid & (kTM - 1), half = gtid >> 7;
    const uint32_t tmem_acc = tmem_base + (uint32_t)s * 256u;
    const uint32_t act_saddr = SADDR(act) + s * kActBytes;
    const uint32_t jtab_saddr = SADDR(jtab) + s * (uint32_t)sizeof(sm.jtab[0]);
    const uint32_t dtab_saddr = SADDR(dtab) + s * (uint32_t)sizeof(sm.dtab[0]);
    const uint32_t wcache_saddr = SADDR(wcache) + s * (uint32_t)sizeof(sm.wcache[0]);
    const uint32_t stg_full0 = SADDR(stg_full) + s * kStgBufs * 8, stg_empty0 = SADDR(stg_empty) + s * kStgBufs * 8;
    const uint32_t act_ready_a = SADDR(act_ready) + s * 8, acc_full_a = SADDR(acc_full) + s * 8, act_free_a = SADDR(act_free) + s * kStgBufs * 8;
    const bool timed = kProf && s == 0;
    const float kLog2e = 1.4426950408889634f;
    const float tau_v2 = sc.tau_v * kLog2e, tau_d2 = sc.tau_d * kLog2e;
    uint32_t accs = 0, afree = 0;
    uint32_t sbuf = 0, sphase = 1;                // staging ring cursor ("empty" barriers start released)

    auto make_ctx = [&](int i, int k, TileCtx& tc) {
      const long long q = 2ll * i + s;
      const long long u = cluster_id + q * n_clusters;
      tc.unit = 2 * u + rank;
      tc.ray0 = tc.unit * kG;
      tc.nr = kStage ? 0 : (int)max(0ll, min((long long)kG, rays.n_rays - tc.ray0));
      if (kStage) { tc.pass = 0; tc.t = 0; } else tile_of<kG>(k, tc.pass, tc.t);
      tc.S = tc.pass == 0 ? PGN_S : PGN_T;
      tc.row0 = tc.t * kTM;
      tc.total_rows = kStage ? kTM : tc.nr * tc.S;
      tc.tile_ray0 = min(tc.row0 / tc.S, kG - 1);
    };

    // ---- per-tile tables (jtab, dtab) for the <=3 rays of the tile: one (ray, joint, axis) item per thread
    auto build_tables = [&](const TileCtx& tc) {
      const int tile_ray1 = min((tc.row0 + kTM - 1) / tc.S, tc.nr - 1);
      if (gtid < kMaxTileRays * PGN_J * 3) {
        const int tr = gtid / (PGN_J * 3), rem = gtid - tr * (PGN_J * 3);
        const int jn = rem / 3, axis = rem - jn * 3;
        const int rl = tc.tile_ray0 + tr;
        if (rl <= tile_ray1) {
          const long long ri = tc.ray0 + rl;
          const float* rb = rays.ray_batch + ri * 11;
          const float4* m = reinterpret_cast<const float4*>(pgn_ray_skts(rays, ri) + jn * 16);
          const float4 m0 = __ldg(m), m1 = __ldg(m + 1), m2 = __ldg(m + 2);
          const float o0 = __ldg(rb), o1 = __ldg(rb + 1), o2 = __ldg(rb + 2);
          const float dd[3] = {__ldg(rb + 3), __ldg(rb + 4), __ldg(rb + 5)};
          const float b0 = fmaf(m0.z, dd[2], fmaf(m0.y, dd[1], m0.x * dd[0]));
          const float b1 = fmaf(m1.z, dd[2], fmaf(m1.y, dd[1], m1.x * dd[0]));
          const float b2 = fmaf(m2.z, dd[2], fmaf(m2.y, dd[1], m2.x * dd[0]));
          if (axis == 0) {
            const float a0 = fmaf(m0.z, o2, fmaf(m0.y, o1, m0.x * o0)) + m0.w;
            const float a1 = fmaf(m1.z, o2, fmaf(m1.y, o1, m1.x * o0)) + m1.w;
            const float a2 = fmaf(m2.z, o2, fmaf(m2.y, o1, m2.x * o0)) + m2.w;
            const uint32_t jt = jtab_saddr + (uint32_t)(tr * PGN_J + jn) * 32u;
            sts128(jt, __float_as_uint(a0), __float_as_uint(a1), __float_as_uint(a2), __float_as_uint(b0));
            sts128(jt + 16, __float_as_uint(b1), __float_as_uint(b2), __float_as_uint(-tau_v2 * sc.cutoff_v[jn]),
                   __float_as_uint(-tau_d2 * sc.cutoff_d[jn]));
          }
          // normalised joint-frame view direction (core/encoders.py:25-37,181-193), component `axis`
          const float den = fmaxf(sqrtf(fmaf(b2, b2, fmaf(b1, b1, b0 * b0))), 1e-12f);
          const float x = (axis == 0 ? b0 : (axis == 1 ? b1 : b2)) / den;
          __nv_bfloat16* tab = &sm.dtab[s][tr][jn * 32];
          float sn, cs;
          __sincosf(x, &sn, &cs);            // |x| <= 1
          tab[axis] = __float2bfloat16_rn(x);
#pragma unroll
          for (int f = 0; f < PGN_LD; ++f) {
            tab[(1 + 2 * f) * 3 + axis] = __float2bfloat16_rn(sn);
            tab[(2 + 2 * f) * 3 + axis] = __float2bfloat16_rn(cs);
            const float s2 = 2.f * sn * cs;
            const float c2 = fmaf(cs, cs, -sn * sn);
            sn = s2; cs = c2;
          }
          if (axis == 0) for (int e = 27; e < 32; ++e) tab[e] = __float2bfloat16_rn(0.f);
        }
      }
    };

    // ---- deferred compositing of a finished tile (raw rows of the tile are in part[s]), in three stages
    auto composite_stage = [&](const TileCtx& tc, int stage) {
      const PgnBf16Net& net = tc.pass == 0 ? net_c : net_f;
      if (stage == 1) {
        group_bar_sync(s);                                   // every head partial of the tile has been stored
        const float br = net.b_rgb[0], bg = net.b_rgb[1], bb = net.b_rgb[2], ba = net.b_alpha[0];
        if (tc.pass == 0) {
          // coarse tile = 2 whole rays: head biases, composite, outputs, weights for the resampling
          const int rl = 2 * tc.t + gwarp;
          if (gwarp < 2 && rl < tc.nr) {
            const long long ri = tc.ray0 + rl;
            float* zc = sm.cscratch[s][gwarp];
            float* wts = zc + 64;
            float* rawrows = &sm.part[s][0][gwarp * PGN_S][0];           // summed in place into the half-0 rows
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              float4* p = reinterpret_cast<float4*>(rawrows + (lane + 32 * h) * 4);
              const float4 u = *reinterpret_cast<const float4*>(&sm.part[s][1][gwarp * PGN_S + lane + 32 * h][0]);
              float4 v = *p;
              v.x += u.x + br; v.y += u.y + bg; v.z += u.z + bb; v.w += u.w + ba;
              *p = v;
            }
            const float nn = __ldg(near_far + ri * 2), ff = __ldg(near_far + ri * 2 + 1);
            if (kDump == 1 && dump.t_rand) {
              zc[lane] = pgn_coarse_z_jitter(nn, ff, sc.t_coarse, lane, __ldg(dump.t_rand + ri * PGN_S + lane), rays.lindisp);
              zc[lane + 32] = pgn_coarse_z_jitter(nn, ff, sc.t_coarse, lane + 32, __ldg(dump.t_rand + ri * PGN_S + lane + 32), rays.lindisp);
            } else {
              zc[lane] = pgn_coarse_z(nn, ff, sc.t_coarse[lane], rays.lindisp);
              zc[lane + 32] = pgn_coarse_z(nn, ff, sc.t_coarse[lane + 32], rays.lindisp);
            }
            const float d0 = __ldg(rays.ray_batch + ri * 11 + 3), d1 = __ldg(rays.ray_batch + ri * 11 + 4), d2 = __ldg(rays.ray_batch + ri * 11 + 5);
            const float dn = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
            float* cr = sm.carry[s][rl];
            if (lane < 8) cr[lane] = (lane == 0) ? 1.f : 0.f;
            __syncwarp();
            pgn_composite_segment_warp<PGN_S, true, 2>(rawrows, zc, 0, PGN_S, dn, sc.density_scale, sc.rgb_eps, lane, cr, wts,
                                                       out.alpha0 ? out.alpha0 + ri * PGN_S : nullptr,
                                                       (kDump == 1 && dump.noise0) ? dump.noise0 + ri * PGN_S : nullptr);
            if (lane == 0) {
              float rgb3[3], disp, acc;
              pgn_composite_finalize(cr, rgb3, &disp, &acc);
              if (out.rgb0) { out.rgb0[ri * 3] = rgb3[0]; out.rgb0[ri * 3 + 1] = rgb3[1]; out.rgb0[ri * 3 + 2] = rgb3[2]; }
              if (out.disp0) out.disp0[ri] = disp;
              if (out.acc0) out.acc0[ri] = acc;
            }
            __syncwarp();
            if (lane < 8) cr[lane] = (lane == 0) ? 1.f : 0.f;      // carry now belongs to the fine pass of this ray
            if (out.weights0) { out.weights0[ri * PGN_S + lane] = wts[lane]; out.weights0[ri * PGN_S + lane + 32] = wts[lane + 32]; }
            if (out.raw0) for (int i = lane; i < PGN_S * 4; i += 32) out.raw0[ri * PGN_S * 4 + i] = rawrows[i];
            __syncwarp();
          }
        } else {
          // fine tile: rows [row0, row0+128) cut up to 3 rays; continue each ray's compositing
          const int rl = tc.tile_ray0 + gwarp;
          if (gwarp < kMaxTileRays && rl < tc.nr && rl * PGN_T < tc.row0 + kTM) {
            const long long ri = tc.ray0 + rl;
            const int s0 = max(0, tc.row0 - rl * PGN_T), s1 = min(PGN_T, tc.row0 + kTM - rl * PGN_T);
            float* rawrows = &sm.part[s][0][rl * PGN_T + s0 - tc.row0][0];
            for (int i = lane; i < s1 - s0; i += 32) {
              float4* p = reinterpret_cast<float4*>(rawrows + i * 4);
              const float4 u = *reinterpret_cast<const float4*>(&sm.part[s][1][rl * PGN_T + s0 - tc.row0 + i][0]);
              float4 v = *p;
              v.x += u.x + br; v.y += u.y + bg; v.z += u.z + bb; v.w += u.w + ba;
              *p = v;
            }
            __syncwarp();
            const float d0 = __ldg(rays.ray_batch + ri * 11 + 3), d1 = __ldg(rays.ray_batch + ri * 11 + 4), d2 = __ldg(rays.ray_batch + ri * 11 + 5);
            const float dn = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
            float* cr = sm.carry[s][rl];
            pgn_composite_segment_warp<PGN_T, true, 3>(rawrows, sm.zf[s][rl], s0, s1, dn, sc.density_scale, sc.rgb_eps, lane, cr,
                                                       nullptr, out.alpha ? out.alpha + ri * PGN_T : nullptr,
                                                       (kDump == 1 && dump.noise) ? dump.noise + ri * PGN_T : nullptr);
            if (out.raw) for (int i = lane; i < (s1 - s0) * 4; i += 32) out.raw[(ri * PGN_T + s0) * 4 + i] = rawrows[i];
            if (s1 == PGN_T && lane == 0) {
              float rgb3[3], disp, acc;
              pgn_composite_finalize(cr, rgb3, &disp, &acc);
              if (out.rgb_map) { out.rgb_map[ri * 3] = rgb3[0]; out.rgb_map[ri * 3 + 1] = rgb3[1]; out.rgb_map[ri * 3 + 2] = rgb3[2]; }
              if (out.disp_map) out.disp_map[ri] = disp;
              if (out.acc_map) out.acc_map[ri] = acc;
            }
          }
        }
      } else if (stage == 2 || stage == 3) {
        // coarse tile: inverse-CDF resampling of the two rays (cdf, then draw + merge -> zf)
        const int rl = 2 * tc.t + gwarp;
        if (tc.pass == 0 && gwarp < 2 && rl < tc.nr) {
          const long long ri = tc.ray0 + rl;
          float* zc = sm.cscratch[s][gwarp];
          float* wts = zc + 64;
          float* scr = zc + 128;
          if (stage == 2) {
            pgn_sample_pdf_cdf_warp_fast(zc, wts, lane, scr);
          } else {
            if (kDump == 1 && dump.u_is)
              pgn_sample_pdf_draw_warp_rand(zc, dump.u_is + ri * PGN_I, lane, scr, out.z_samples ? out.z_samples + ri * PGN_I : nullptr,
                                            sm.zf[s][rl], out.pdf_inds ? out.pdf_inds + ri * PGN_I : nullptr);
            else
              pgn_sample_pdf_draw_warp_fast(zc, sc.u_det, lane, scr, out.z_samples ? out.z_samples + ri * PGN_I : nullptr,
                                            sm.zf[s][rl], out.pdf_inds ? out.pdf_inds + ri * PGN_I : nullptr);
            if (out.z_fine) for (int i = lane; i < PGN_T; i += 32) out.z_fine[ri * PGN_T + i] = sm.zf[s][rl][i];
          }
        }
      }
    };

    // ---- POST(job): drain the accumulator (epilogue)
    float sig_keep = 0.f;             // this thread's sigma-head partial, carried from L7's epilogue to V's
    int fc_row = rays.n_codes;        // frame-code table row of this thread's ray (set before the next tile's row context replaces rc)
    auto post = [&](int L, const TileCtx& tc) -> bool {
      const PgnBf16Net& net = (kStage || tc.pass == 0) ? net_c : net_f;
      { PROF_T0(); const bool okw = mbar_wait_s(acc_full_a, accs & 1, status, 303); if (timed) { PROF_ADD(9); PROF_ADD(16 + L); } if (!okw) return false; }
      ++accs;
      tc_fence_after_sync();
      uint4* dptr = nullptr;
      uint32_t* mptr = nullptr;          // mask words of this thread's (row, column half): word planes [layer][word 0..7][row]
      uint32_t* vptr = nullptr;          // view layer: [word 0..3][row]
      size_t mstride = 0;
      if (kDump == 2 && tc.pass == 1) {
        // masks only (frozen network, pose gradient): trunk masks of the fine pass as word planes [layer 0..7][word 0..7][row]
        // (word w = [column 32 w + bit > 0]), then the view layer's [word 0..3][row]
        const long long m = dump.rows_f;
        const long long grow = tc.unit * (long long)(kG * tc.S) + tc.row0 + ((gwarp & 3) * 32 + lane);
        uint4* base = reinterpret_cast<uint4*>(dump.f);
        uint32_t* words = reinterpret_cast<uint32_t*>(base);
        mstride = (size_t)m;
        if (L < 8) mptr = words + (size_t)(L * 8 + (gwarp >> 2) * 4) * (size_t)m + (size_t)grow;
        else vptr = words + (size_t)(64 + (gwarp >> 2) * 2) * (size_t)m + (size_t)grow;
      }
      // (groups of 4: the third fine tile holds 64 rows of the group; its upper half would land in the next group's rows)
      if (kDump == 1 && (tc.pass == 0 ? dump.c : dump.f) != nullptr && tc.row0 + ((gwarp & 3) * 32 + lane) < kG * tc.S) {   // a pass without a buffer is not dumped
        // training forward: post-ReLU activations of every layer, bf16, per pass [layer][row][256] row-major (view
        // layer: [row][128], after the eight trunk layers); rows in (ray, sample) order:
        // row = unit * rows_per_group + tile * 128 + row_in_tile (units padded to pairs)
        const long long m = tc.pass == 0 ? dump.rows_c : dump.rows_f;
        const long long grow = tc.unit * (long long)(kG * tc.S) + tc.row0 + ((gwarp & 3) * 32 + lane);
        uint4* base = reinterpret_cast<uint4*>(tc.pass == 0 ? dump.c : dump.f);
        // trunk layer L: tile-blocked, tile grow / 128 (64 KB = 4096 uint4), this row's 16 bytes of run 0; view layer: row-major
        dptr = L == 8 ? base + (size_t)8 * 32 * (size_t)m + (size_t)grow * 16
                      : base + (size_t)L * 32 * (size_t)m + (size_t)(grow >> 7) * 4096 + (size_t)(grow & 127);
        // ReLU masks of the trunk layers behind the activations: [layer 0..7][row][column half] x 128 bits
        mstride = (size_t)m;
        if (L < 8) mptr = reinterpret_cast<uint32_t*>(base + (size_t)272 * (size_t)m) + (size_t)(L * 8 + (gwarp >> 2) * 4) * (size_t)m + (size_t)grow;
      }
      { PROF_T0();
        if (L == 8) epilogue<2>(sm, tmem_acc, act_saddr, s, net.w_alpha, net.w_rgb, gwarp, lane, sig_keep, dptr, mptr, vptr, mstride,
                                kFC ? net.fc_table + (size_t)fc_row * 128 : nullptr);
        else if (L == 7) epilogue<1>(sm, tmem_acc, act_saddr, s, net.w_alpha, net.w_rgb, gwarp, lane, sig_keep, dptr, mptr, vptr, mstride);
        else epilogue<0>(sm, tmem_acc, act_saddr, s, net.w_alpha, net.w_rgb, gwarp, lane, sig_keep, dptr, mptr, vptr, mstride);
        compute_arrive(act_ready_a, lane); if (timed) PROF_ADD(8); }
      if (kStage && L == 8) {
        group_bar_sync(s);
        if (gtid < kTM) {
          const float* p0 = sm.part[s][0][gtid];
          const float* p1 = sm.part[s][1][gtid];
          const int rows_valid = (int)max(0ll, min((long long)kTM, enc_rows_total - tc.unit * kTM));
          if (gtid < rows_valid) {
            float* o = raw_global + ((size_t)tc.unit * kTM + gtid) * 4;
            o[0] = p0[0] + p1[0] + net.b_rgb[0]; o[1] = p0[1] + p1[1] + net.b_rgb[1];
            o[2] = p0[2] + p1[2] + net.b_rgb[2]; o[3] = p0[3] + p1[3] + net.b_alpha[0];
          }
        }
        group_bar_sync(s);
      }
      return true;
    };

    // ---- PRE(job): everything the tensor core needs from the CUDA cores before/while it runs the layer
    RowCtx rc;
    rc.valid = false; rc.tr = 0; rc.z = 0.f;
    bool tables_ready = false;
    auto pre = [&](int L, const TileCtx& tc) -> bool {
      const int nchunks = pgn_layer_chunks(L);
      if (nchunks == 0) return true;
      const float* enc_rows = kStage ? enc_global + (size_t)tc.unit * kTM * PGN_ENC : nullptr;
      const int rows_valid = kStage ? (int)max(0ll, min((long long)kTM, enc_rows_total - tc.unit * kTM)) : kTM;
      if (!kStage && L == 0 && !tables_ready) {      // (normally built ahead, in the shadow of the previous tile's view layer)
        PROF_T0();
        build_tables(tc);
        rc = make_row_ctx(sm, sc, tc, near_far, s, row, kDump == 1 ? dump.t_rand : nullptr, rays.lindisp);
        group_bar_sync(s);                 // tables visible to every thread of the group
        if (timed) PROF_ADD(14);
      }
      if (L == 0) tables_ready = false;
      // software pipeline: the values of chunk c+1 are computed while chunk c travels through the
      // staging ring / tensor core; only the 16-byte stores wait for a ring buffer to be released
      uint32_t packed[20];
      auto compute_chunk = [&](int c) {
        if (L == 8) {
          PROF_T0();
          if (kStage) encode_d_stage(c, row, half, enc_rows, rows_valid, packed);
          else encode_d_fast(wcache_saddr, dtab_saddr, rc, c, row, half, packed);
          if (timed) PROF_ADD(7);
        } else {
          PROF_T0();
          if (kStage) encode_x_stage(c, row, half, enc_rows, rows_valid, packed);
          else if (L == 5) encode_x_fast<true>(wcache_saddr, jtab_saddr, rc, tau_v2, tau_d2, c, row, half, packed);
          else encode_x_fast<false>(wcache_saddr, jtab_saddr, rc, tau_v2, tau_d2, c, row, half, packed);
          if (timed) PROF_ADD(6);
        }
      };
      if (L == 8) group_bar_sync(s);       // wcache (written by L5's encode) visible to every thread
      compute_chunk(0);
      for (int c = 0; c < nchunks; ++c) {
        if (L != 0 && c < kStgBufs) {   // the activation K-steps of this layer that lie under ring buffer c must have been consumed
          PROF_T0(); const bool okw = mbar_wait_s(act_free_a + c * 8, afree & 1, status, 302); if (timed) PROF_ADD(13); if (!okw) return false;
          if (c == kStgBufs - 1) ++afree;
        }
        { PROF_T0(); const bool okw = mbar_wait_s(stg_empty0 + sbuf * 8, sphase, status, 301); if (timed) PROF_ADD(10); if (!okw) return false; }
        { PROF_T0();
          const uint32_t stg = act_saddr + sbuf * (uint32_t)kStgBytes;
          if (L == 8) encode_d_store(stg, row, half, packed); else encode_x_store(stg, row, half, packed);
          chunk_arrive(stg_full0 + sbuf * 8, lane); if (timed) PROF_ADD(15); }
        if (++sbuf == kStgBufs) { sbuf = 0; sphase ^= 1; }
        if (c + 1 < nchunks) compute_chunk(c + 1);
      }
      return true;
    };

    TileCtx tc, prev;
    bool pending = false;
    for (int i = 0; i < n_slot; ++i) {
      for (int k = 0; k < kTiles; ++k) {
        make_ctx(i, k, tc);
        for (int L = 0; L < 9; ++L) {
          { PROF_T0(); const bool okp = pre(L, tc); if (timed && (L == 0 || L == 5 || L == 8)) PROF_ADD(L == 0 ? 25 : (L == 5 ? 26 : 27)); if (!okp) goto done; }
          if (!kStage && pending && L >= 1 && L <= 3) { PROF_T0(); composite_stage(prev, L); if (timed) PROF_ADD(11); }
          if (kFC && !kStage && L == 8)                 // this row's ray -> frame-code row (while rc still describes THIS tile)
            fc_row = (tc.nr > 0) ? pgn_ray_code_row(rays, tc.ray0 + min(tc.tile_ray0 + rc.tr, tc.nr - 1)) : rays.n_codes;
          if (!kStage && L == 8 && (k + 1 < kTiles || i + 1 < n_slot)) {
            // the view layer's last chunks are still in the tensor core: build the next tile's tables now
            PROF_T0();
            TileCtx nx;
            if (k + 1 < kTiles) make_ctx(i, k + 1, nx); else make_ctx(i + 1, 0, nx);
            group_bar_sync(s);               // every thread is done reading this tile's tables
            build_tables(nx);
            rc = make_row_ctx(sm, sc, nx, near_far, s, row, kDump == 1 ? dump.t_rand : nullptr, rays.lindisp);
            group_bar_sync(s);               // tables visible to every thread of the group
            tables_ready = true;
            if (timed) PROF_ADD(14);
          }
          if (!post(L, tc)) goto done;
        }
        if (!kStage) { prev = tc; pending = true; }
      }
    }
    if (pending) for (int st = 1; st <= 3; ++st) composite_stage(prev, st);
  }
done:
  if (kProf) {
    unsigned long long* pp = prof + (size_t)blockIdx.x * 32;
    const unsigned long long total = (unsigned long long)(clock64() - kernel_t0);
    if (warp == kIssuerWarp0 && lane == 0 && rank == 0) { pp[0] = pacc[0]; pp[1] = pacc[1]; pp[2] = pacc[2]; pp[3] = total; pp[28] = pacc[28]; pp[29] = pacc[29]; }
    if (warp == kIssuerWarp0 && lane == 0 && rank == 1) { pp[0] = 0; pp[1] = 0; pp[2] = 0; pp[3] = total; pp[28] = 0; pp[29] = 0; }
    if (warp == kProducerWarp0 && lane == 0) { pp[4] = pacc[4]; pp[5] = total; }
    if (tid == 0) { pp[6] = pacc[6]; pp[7] = pacc[7]; pp[8] = pacc[8]; pp[9] = pacc[9]; pp[10] = pacc[10]; pp[11] = pacc[11];
                    pp[12] = total; pp[13] = pacc[13]; pp[14] = pacc[14]; pp[15] = pacc[15];
                    for (int k = 16; k < 28; ++k) pp[k] = pacc[k]; }
  }
  tc_fence_before_sync();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();                 // the peer's TMEM/smem must stay alive until every MMA has retired
  if (warp == kIssuerWarp0) {
    tc_fence_after_sync();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

#ifndef PGN_RENDER_FC_UNIT
// ------------------------------------------------------------------ weight packing
// wsrc[l]: device fp32 nn.Linear weights in the include/posegen_b200.h order.
__global__ void pgn_fold_view_kernel(const float* __restrict__ w_view /*[128][view_ld]*/, const float* __restrict__ w_feat /*[256][256]*/,
                                     const float* __restrict__ b_view, const float* __restrict__ b_feat,
                                     float* __restrict__ fold /*[128][256] + [128]*/, int view_ld) {
  // fold[n][k] = sum_m w_view[n][m] * w_feat[m][k];  fold_b[n] = b_view[n] + sum_m w_view[n][m] * b_feat[m]
  // (runs on every weight upload, i.e. once per training step: four independent accumulation chains per thread and a
  // block reduction for the bias instead of 256 serial FMAs + a serial bias loop in thread 0)
  __shared__ float red[8];
  const int n = blockIdx.x, k = threadIdx.x;
  const float* wv = w_view + (size_t)n * view_ld;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 16
  for (int m = 0; m < 256; m += 4) {                 // (64 loads in flight per thread: the kernel is L2-latency bound)
    a0 = fmaf(__ldg(wv + m), __ldg(w_feat + m * 256 + k), a0);
    a1 = fmaf(__ldg(wv + m + 1), __ldg(w_feat + (m + 1) * 256 + k), a1);
    a2 = fmaf(__ldg(wv + m + 2), __ldg(w_feat + (m + 2) * 256 + k), a2);
    a3 = fmaf(__ldg(wv + m + 3), __ldg(w_feat + (m + 3) * 256 + k), a3);
  }
  fold[n * 256 + k] = (a0 + a1) + (a2 + a3);
  float b = __ldg(wv + k) * __ldg(b_feat + k);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
  if ((k & 31) == 0) red[k >> 5] = b;
  __syncthreads();
  if (k == 0) {
    float t = b_view[n];
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    fold[128 * 256 + n] = t;
  }
}

struct PackPtrs { const float* w[12]; const float* b[12]; };

__global__ void pgn_pack_wstream_kernel(PackPtrs p, const float* __restrict__ fold, __nv_bfloat16* __restrict__ wstream,
                                        float* __restrict__ bias, float* __restrict__ w_alpha, float* __restrict__ w_rgb, int view_ld) {
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
    if (ks == ks_total - 1) {
      // bias K-step: k = 0 -> bf16 hi part, k = 1 -> residual (the tensor core adds both in fp32)
      const float b = (L < 8) ? p.b[L][n] : fold[128 * 256 + n];
      const float hi = __bfloat162float(__float2bfloat16_rn(b));
      v = (kh == 0 && e == 0) ? hi : ((kh == 0 && e == 1) ? b - hi : 0.f);
    } else
    if (L == 0) { const int rc = pgn_xperm_refcol(kp); v = rc >= 0 ? p.w[0][(size_t)n * 432 + rc] : 0.f; }
    else if (L == 5) {
      if (kp < 256) v = p.w[5][(size_t)n * 688 + 432 + kp];
      else { const int rc = pgn_xperm_refcol(kp - 256); v = rc >= 0 ? p.w[5][(size_t)n * 688 + rc] : 0.f; }
    }
    else if (L == 8) {
      if (kp < 256) v = fold[n * 256 + kp];
      else { const int rc = pgn_dperm_refcol(kp - 256); v = rc >= 0 ? p.w[10][(size_t)n * view_ld + 256 + (rc - 432)] : 0.f; }
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

// Rays per group of a launch over n_rays rays: 8, or 6 / 4 when that shortens the critical path of the 148 tile pipelines
// (2 per CTA, pair-units dealt round-robin): rounds of pair-units x tiles per group (9 / 7 / 5), the larger group on a tie.  A pure function of the ray count (a 148-SM B200 is assumed), so that the
// dump-size queries and the launch agree.  The masks-only (GAN) forward always uses 8.
int pgn_bf16_group_rays(long long n_rays, int masks_only) {
  if (masks_only || n_rays <= 0) return kRPG;
  const long long slots = 148;                       // 74 CTA pairs x 2 pipelines
  int best = kRPG;
  long long best_tiles = 0;
  for (int g = 8; g >= 4; g -= 2) {                  // critical path in tiles: rounds of pair-units x tiles per group
    const long long pairs = ((n_rays + g - 1) / g + 1) / 2;
    const long long tiles = ((pairs + slots - 1) / slots) * (g == 8 ? 9 : (g == 6 ? 7 : 5));
    if (g == 8 || tiles < best_tiles) { best = g; best_tiles = tiles; }
  }
  return best;
}

// rows of the activation dump of one pass (units are padded to CTA pairs): samples_per_ray = 64 | 80
long long pgn_bf16_dump_rows(long long n_rays, int samples_per_ray, int masks_only) {
  const int g = pgn_bf16_group_rays(n_rays, masks_only);
  const long long n_groups = (n_rays + g - 1) / g;
  const long long rows = ((n_groups + 1) / 2) * 2 * (long long)g * samples_per_ray;
  return (rows + 127) / 128 * 128;
}

cudaError_t pgn_pack_bf16_net(const float* const* w_dev, const float* const* b_dev, __nv_bfloat16* wstream,
                              float* bias, float* w_alpha, float* w_rgb, float* fold_tmp, int view_ld, cudaStream_t stream) {
  PackPtrs p;
  for (int i = 0; i < 12; ++i) { p.w[i] = w_dev[i]; p.b[i] = b_dev[i]; }
  pgn_fold_view_kernel<<<128, 256, 0, stream>>>(w_dev[10], w_dev[9], b_dev[10], b_dev[9], fold_tmp, view_ld);
  pgn_pack_wstream_kernel<<<148 * 4, 256, 0, stream>>>(p, fold_tmp, wstream, bias, w_alpha, w_rgb, view_ld);
  return cudaGetLastError();
}

#else
}  // namespace
#endif  // !PGN_RENDER_FC_UNIT  (weight packing, dump geometry: main unit only)

// ------------------------------------------------------------------ launch (compiled twice: kFC = false here, kFC = true in
// pgn_render_bf16_fc.cu, which defines PGN_RENDER_FC_UNIT and includes this file)
#ifdef PGN_RENDER_FC_UNIT
#define PGN_FC true
#define PGN_LAUNCH_RENDER pgn_launch_render_bf16_fc
#define PGN_LAUNCH_MLP pgn_launch_mlp_bf16_fc
#else
#define PGN_FC false
#define PGN_LAUNCH_RENDER pgn_launch_render_bf16
#define PGN_LAUNCH_MLP pgn_launch_mlp_bf16
cudaError_t pgn_launch_render_bf16_fc(const PgnRayRefs& rays, const PgnOutputs& out, const PgnBf16Net& nc, const PgnBf16Net& nf,
                                      const PgnScalars* sc_dev, const float* near_far, int* status, unsigned long long* prof,
                                      const PgnActDump* dump, int num_sms, cudaStream_t stream);
cudaError_t pgn_launch_mlp_bf16_fc(const PgnBf16Net& net, const float* enc, long long m, float* raw, const PgnScalars* sc_dev,
                                   int* status, int num_sms, int n_codes, cudaStream_t stream);
#endif

static cudaError_t configure_bf16() {
  static PgnPerDeviceOnce done;
  if (!done.need()) return cudaSuccess;
  const int bytes = (int)sizeof(Smem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(pgn_render_bf16_kernel<false, false, 0, 8, PGN_FC>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
#ifndef PGN_RENDER_FC_UNIT
  e = cudaFuncSetAttribute(pgn_render_bf16_kernel<false, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
#endif
  e = cudaFuncSetAttribute(pgn_render_bf16_kernel<true, false, 0, 8, PGN_FC>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(pgn_render_bf16_kernel<false, false, 1, 8, PGN_FC>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(pgn_render_bf16_kernel<false, false, 2, 8, PGN_FC>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(pgn_render_bf16_kernel<false, false, 0, 4, PGN_FC>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(pgn_render_bf16_kernel<false, false, 1, 4, PGN_FC>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(pgn_render_bf16_kernel<false, false, 0, 6, PGN_FC>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(pgn_render_bf16_kernel<false, false, 1, 6, PGN_FC>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  done.set();
  return cudaSuccess;
}

cudaError_t PGN_LAUNCH_RENDER(const PgnRayRefs& rays, const PgnOutputs& out, const PgnBf16Net& nc,
                              const PgnBf16Net& nf, const PgnScalars* sc_dev, const float* near_far,
                              int* status, unsigned long long* prof, const PgnActDump* dump, int num_sms, cudaStream_t stream) {
#ifndef PGN_RENDER_FC_UNIT
  if (nc.fc_table) return pgn_launch_render_bf16_fc(rays, out, nc, nf, sc_dev, near_far, status, prof, dump, num_sms, stream);
#endif
  cudaError_t e = configure_bf16();
  if (e != cudaSuccess) return e;
  const bool use_prof = prof != nullptr && !PGN_FC;          // the phase timers exist for the frame-code-free inference kernel
  const int g = use_prof ? kRPG : pgn_bf16_group_rays(rays.n_rays, dump && dump->masks_only);
  const long long n_groups = (rays.n_rays + g - 1) / g;
  if (n_groups == 0) return cudaSuccess;
  const long long n_pairs = (n_groups + 1) / 2;
  const int grid = 2 * (int)min((long long)(num_sms / 2), n_pairs);      // clusters of 2 CTAs
  const size_t smem = sizeof(Smem) + 1024;
  const PgnActDump nodump{};
  if (dump && dump->masks_only)
    pgn_render_bf16_kernel<false, false, 2, 8, PGN_FC><<<grid, kThreads, smem, stream>>>(rays, out, nc, nf, sc_dev, near_far, nullptr, 0, nullptr, status, nullptr, *dump);
  else if (dump && g == 4)
    pgn_render_bf16_kernel<false, false, 1, 4, PGN_FC><<<grid, kThreads, smem, stream>>>(rays, out, nc, nf, sc_dev, near_far, nullptr, 0, nullptr, status, nullptr, *dump);
  else if (dump && g == 6)
    pgn_render_bf16_kernel<false, false, 1, 6, PGN_FC><<<grid, kThreads, smem, stream>>>(rays, out, nc, nf, sc_dev, near_far, nullptr, 0, nullptr, status, nullptr, *dump);
  else if (dump)
    pgn_render_bf16_kernel<false, false, 1, 8, PGN_FC><<<grid, kThreads, smem, stream>>>(rays, out, nc, nf, sc_dev, near_far, nullptr, 0, nullptr, status, nullptr, *dump);
  else if (g == 4)
    pgn_render_bf16_kernel<false, false, 0, 4, PGN_FC><<<grid, kThreads, smem, stream>>>(rays, out, nc, nf, sc_dev, near_far, nullptr, 0, nullptr, status, nullptr, nodump);
  else if (g == 6)
    pgn_render_bf16_kernel<false, false, 0, 6, PGN_FC><<<grid, kThreads, smem, stream>>>(rays, out, nc, nf, sc_dev, near_far, nullptr, 0, nullptr, status, nullptr, nodump);
#ifndef PGN_RENDER_FC_UNIT
  else if (use_prof)
    pgn_render_bf16_kernel<false, true, 0><<<grid, kThreads, smem, stream>>>(rays, out, nc, nf, sc_dev, near_far, nullptr, 0, nullptr, status, prof, nodump);
#endif
  else
    pgn_render_bf16_kernel<false, false, 0, 8, PGN_FC><<<grid, kThreads, smem, stream>>>(rays, out, nc, nf, sc_dev, near_far, nullptr, 0, nullptr, status, nullptr, nodump);
  return cudaGetLastError();
}

cudaError_t PGN_LAUNCH_MLP(const PgnBf16Net& net, const float* enc, long long m, float* raw,
                           const PgnScalars* sc_dev, int* status, int num_sms, int n_codes, cudaStream_t stream) {
#ifndef PGN_RENDER_FC_UNIT
  if (net.fc_table) return pgn_launch_mlp_bf16_fc(net, enc, m, raw, sc_dev, status, num_sms, n_codes, stream);
#endif
  cudaError_t e = configure_bf16();
  if (e != cudaSuccess) return e;
  const long long n_tiles = (m + kTM - 1) / kTM;
  if (n_tiles == 0) return cudaSuccess;
  const int grid = 2 * (int)min((long long)(num_sms / 2), (n_tiles + 1) / 2);
  PgnRayRefs rays{};
  rays.n_codes = n_codes;          // explicit encodings carry no camera index: a frame-code model uses its mean code
  PgnOutputs out{};
  pgn_render_bf16_kernel<true, false, 0, 8, PGN_FC><<<grid, kThreads, sizeof(Smem) + 1024, stream>>>(rays, out, net, net, sc_dev, nullptr,
                                                                                                   enc, m, raw, status, nullptr, PgnActDump{});
  return cudaGetLastError();
}
