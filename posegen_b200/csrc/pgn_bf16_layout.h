// Static schedule of the bf16 tensor-core MLP: layer shapes, K ordering and the packed
// weight-stream layout shared by the packer, the weight producer and the MMA issuer.
//
// MMA layers (index L):          N    K (act part | generated part)
//   0      pts_linears.0        256   0   | 432 x_p  (joint-major, 18 per joint)
//   1..4   pts_linears.1-4      256   256 | 0
//   5      pts_linears.5        256   256 | 432 x_p  (reference order is [x_p | h]; K is permuted)
//   6,7    pts_linears.6-7      256   256 | 0
//   8      feature∘views        128   256 | 672 d    (joint-major, 27 + 1 zero pad per joint)
// The kernel runs on CTA pairs (cta_group::2, UMMA M=256): B is split along N between the two CTAs.
// Stream = layers in order; per layer fills of `ks_per_fill` K-steps; per fill
// [cta rank][kstep][khalf][n_local (N/2)][8] bf16, i.e. for one CTA each K=16 step is two "runs"
// (8 consecutive k for its N/2 rows) -> UMMA K-major SWIZZLE_NONE with LBO = (N/2)*16 B, SBO = 128 B.
// The producer moves `ks_per_fill` K-steps per bulk copy (8 KB) into a deep ring so that many
// copies are in flight (the stream is L2-latency bound, not bandwidth bound, at shallow depth).
#pragma once
#include <stddef.h>

#ifndef __host__
#define __host__
#define __device__
#endif

__host__ __device__ constexpr int pgn_layer_n(int L) { return L == 8 ? 128 : 256; }
__host__ __device__ constexpr int pgn_layer_kact(int L) { return L == 0 ? 0 : 256; }
__host__ __device__ constexpr int pgn_layer_kenc(int L) { return (L == 0 || L == 5) ? 432 : (L == 8 ? 672 : 0); }
__host__ __device__ constexpr int pgn_layer_ksteps(int L) { return (pgn_layer_kact(L) + pgn_layer_kenc(L)) / 16; }
__host__ __device__ constexpr int pgn_ks_per_fill(int L) { return L == 8 ? 4 : 2; }   // 8 KB per CTA per bulk copy
__host__ __device__ constexpr size_t pgn_wstream_elems() {
  size_t t = 0;
  for (int L = 0; L < 9; ++L) t += (size_t)pgn_layer_n(L) * pgn_layer_ksteps(L) * 16;
  return t;
}
// position inside the 432-wide x part -> reference column of the density-net input
__host__ __device__ inline int pgn_xperm_refcol(int kp) {
  const int j = kp / 18, t = kp - j * 18;
  return t < 15 ? t * 24 + j : 360 + j * 3 + (t - 15);
}
// position inside the 672-wide d part -> reference column of the 1080 vector (or -1: zero pad)
__host__ __device__ inline int pgn_dperm_refcol(int q) {
  const int j = q / 28, t = q - j * 28;
  if (t == 27) return -1;
  const int k = t / 3, a = t - k * 3;
  return 432 + k * 72 + j * 3 + a;
}
