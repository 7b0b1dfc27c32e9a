// Static schedule of the bf16 tensor-core MLP: layer shapes, K ordering and the packed
// weight-stream layout shared by the packer, the weight producers and the MMA issuer.
//
// MMA layers (index L):          N    K (act part | generated part)
//   0      pts_linears.0        256   0   | 480 x_p  (6 chunks of 4 joints; 18 values per joint + 8 zero pad per chunk)
//   1..4   pts_linears.1-4      256   256 | 0
//   5      pts_linears.5        256   256 | 480 x_p  (reference order is [x_p | h]; K is permuted)
//   6,7    pts_linears.6-7      256   256 | 0
//   8      feature∘views        128   256 | 768 d    (12 chunks of 2 joints; 27 values per joint + 5 zero pad)
// The zero pads make every per-thread run of generated values a whole number of 16-byte
// shared-memory stores (8 bf16) and every chunk a whole number of K=16 steps while keeping the
// chunk small enough (20 KB) for two tiles to be resident per CTA; they cost 4.6 % extra MMA work.
//
// The kernel runs on CTA pairs (cta_group::2, UMMA M=256): B is split along N between the two CTAs.
// Stream = layers in order; per layer fills of `ks_per_fill` K-steps; per fill
// [cta rank][kstep][khalf][n_local (N/2)][8] bf16, i.e. for one CTA each K=16 step is two "runs"
// (8 consecutive k for its N/2 rows) -> UMMA K-major SWIZZLE_NONE with LBO = (N/2)*16 B, SBO = 128 B.
// Each CTA moves its half of a fill (8 KB) with one bulk copy into a ring deep enough to keep
// several copies in flight (the stream is L2-latency bound at shallow depth).
#pragma once
#include <stddef.h>

#ifndef __host__
#define __host__
#define __device__
#endif

#define PGN_X_CHUNK_K 80      // 4 joints x 18 + 8 pad
#define PGN_X_CHUNKS 6
#define PGN_D_CHUNK_K 64      // 2 joints x (27 + 5 pad)
#define PGN_D_CHUNKS 12

__host__ __device__ constexpr int pgn_layer_n(int L) { return L == 8 ? 128 : 256; }
__host__ __device__ constexpr int pgn_layer_kact(int L) { return L == 0 ? 0 : 256; }
__host__ __device__ constexpr int pgn_layer_kenc(int L) {
  return (L == 0 || L == 5) ? PGN_X_CHUNK_K * PGN_X_CHUNKS : (L == 8 ? PGN_D_CHUNK_K * PGN_D_CHUNKS : 0);
}
// +1: the last K-step of every layer multiplies a constant "ones" block (columns 0,1 = 1) with
// [bf16 hi(bias) ; bf16 lo(bias)], i.e. the bias is added by the tensor core in fp32 and the epilogue
// is reduced to ReLU + convert + store.
__host__ __device__ constexpr int pgn_layer_ksteps(int L) { return (pgn_layer_kact(L) + pgn_layer_kenc(L)) / 16 + 1; }
__host__ __device__ constexpr int pgn_layer_chunk_ks(int L) { return L == 8 ? PGN_D_CHUNK_K / 16 : PGN_X_CHUNK_K / 16; }
__host__ __device__ constexpr int pgn_layer_chunks(int L) { return (L == 0 || L == 5) ? PGN_X_CHUNKS : (L == 8 ? PGN_D_CHUNKS : 0); }
__host__ __device__ constexpr int pgn_ks_per_fill(int L) { return L == 8 ? 4 : 2; }   // 8 KB per CTA per bulk copy
__host__ __device__ constexpr size_t pgn_wstream_elems() {
  size_t t = 0;
  for (int L = 0; L < 9; ++L) t += (size_t)pgn_layer_n(L) * pgn_layer_ksteps(L) * 16;
  return t;
}
// position inside the 480-wide x part -> reference column of the density-net input (or -1: zero pad).
// chunk c = kp/80 holds joints 4c..4c+3; thread half h = (kp%80)/40 owns joints 4c+2h, 4c+2h+1.
__host__ __device__ inline int pgn_xperm_refcol(int kp) {
  const int c = kp / PGN_X_CHUNK_K, rr = kp - c * PGN_X_CHUNK_K;
  const int h = rr / 40, r = rr - h * 40;
  if (r >= 36) return -1;
  const int j = 4 * c + 2 * h + r / 18, t = r % 18;
  return t < 15 ? t * 24 + j : 360 + j * 3 + (t - 15);
}
// position inside the 768-wide d part -> reference column of the 1080 vector (or -1: zero pad)
__host__ __device__ inline int pgn_dperm_refcol(int q) {
  const int j = q / 32, t = q - j * 32;
  if (t >= 27) return -1;
  const int k = t / 3, a = t - k * 3;
  return 432 + k * 72 + j * 3 + a;
}
