// Bring-up probe for the tcgen05 plumbing used by pgn_render_bf16.cu: one CTA computes
// D[128][N] = A[128][K] * B[N][K]^T (bf16 inputs, fp32 accumulate) with the same shared
// memory layout ([k/8][row][8], SWIZZLE_NONE K-major), descriptors, commit/mbarrier and
// TMEM load path as the render kernel.  `variant` bit 0 swaps the LBO/SBO fields so a
// single GPU run settles the descriptor convention.
#include "pgn_umma.cuh"
#include "pgn_kernels.h"

using namespace pgn;

__global__ void __launch_bounds__(128, 1)
pgn_probe_umma_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K, int N,
                      int variant, int* __restrict__ status_g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                          // K/8 runs of 128 rows x 16 B
  uint8_t* sB = smem + (size_t)(K / 8) * 2048; // K/8 runs of N rows x 16 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  volatile int* status = status_g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 128 * K; i += 128) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sA + (size_t)(k / 8) * 2048 + r * 16 + (k % 8) * 2) = __float2bfloat16_rn(A[i]);
  }
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sB + (size_t)(k / 8) * (N * 16) + n * 16 + (k % 8) * 2) = __float2bfloat16_rn(B[i]);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 256); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const int reps = (variant & 4) ? 8 : 1;     // bit 2: timing mode (same K-steps issued 8x back to back)
  long long t_issue0 = 0, t_issue1 = 0;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    t_issue0 = clock64();
    for (int rep = 0; rep < reps; ++rep)
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t a_addr = smem_u32(sA) + ks * 2 * 2048;
      const uint32_t b_addr = smem_u32(sB) + ks * 2 * (N * 16);
      uint64_t ad, bd;
      if (variant & 1) { ad = umma_smem_desc(a_addr, 128, 2048); bd = umma_smem_desc(b_addr, 128, N * 16); }
      else             { ad = umma_smem_desc(a_addr, 2048, 128); bd = umma_smem_desc(b_addr, N * 16, 128); }
      umma_bf16(tmem, ad, bd, idesc, (ks > 0 || rep > 0) ? 1u : 0u);
    }
    umma_commit(&bar);
    t_issue1 = clock64();
  }
  __syncwarp();
  mbar_wait(&bar, 0, status, 901);
  if (tid == 0 && (variant & 4)) { D[(size_t)128 * N] = (float)(t_issue1 - t_issue0); D[(size_t)128 * N + 1] = (float)(clock64() - t_issue0); }
  tc_fence_after_sync();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[(size_t)row * N + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tmem, 256); }
}

// CTA-pair variant: D[256][N] = A[256][K] * B[N][K]^T on a 2-CTA cluster with cta_group::2
// (UMMA M=256).  CTA r stages A rows [128r, 128r+128) and B rows [N/2*r, N/2*(r+1)); the leader
// issues the MMAs and multicasts the commit; each CTA drains its own 128 TMEM lanes.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pgn_probe_umma2_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K, int N,
                       int timing, int* __restrict__ status_g) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int NH = N / 2;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)(K / 8) * 2048;
  __shared__ uint64_t bar_done, bar_ready;
  __shared__ uint32_t tmem_slot;
  volatile int* status = status_g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  for (int i = tid; i < 128 * K; i += 128) {
    const int r = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sA + (size_t)(k / 8) * 2048 + r * 16 + (k % 8) * 2) = __float2bfloat16_rn(A[(size_t)(rank * 128 + r) * K + k]);
  }
  for (int i = tid; i < NH * K; i += 128) {
    const int n = i / K, k = i % K;
    *reinterpret_cast<__nv_bfloat16*>(sB + (size_t)(k / 8) * (NH * 16) + n * 16 + (k % 8) * 2) = __float2bfloat16_rn(B[(size_t)(rank * NH + n) * K + k]);
  }
  if (tid == 0) { mbar_init(&bar_done, 1); mbar_init(&bar_ready, 256); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc_2cta(&tmem_slot, 256); tmem_relinquish_2cta(); }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  // operands of both CTAs ready -> leader's barrier
  fence_proxy_async_smem();
  mbar_arrive_cluster(&bar_ready, 0);
  const int reps = timing ? 8 : 1;
  long long t_issue0 = 0, t_issue1 = 0;
  if (rank == 0 && tid == 0) {
    mbar_wait_cluster(&bar_ready, 0, status, 911);
    tc_fence_after_sync();
    const uint32_t idesc = umma_idesc_bf16(256, N);
    t_issue0 = clock64();
    for (int rep = 0; rep < reps; ++rep)
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t ad = umma_smem_desc(smem_u32(sA) + ks * 2 * 2048, 2048, 128);
      const uint64_t bd = umma_smem_desc(smem_u32(sB) + ks * 2 * (NH * 16), NH * 16, 128);
      umma_bf16_2cta(tmem, ad, bd, idesc, (ks > 0 || rep > 0) ? 1u : 0u);
    }
    umma_commit_2cta(&bar_done);
    t_issue1 = clock64();
  }
  __syncwarp();
  mbar_wait(&bar_done, 0, status, 912);
  if (rank == 0 && tid == 0 && timing) { D[(size_t)256 * N] = (float)(t_issue1 - t_issue0); D[(size_t)256 * N + 1] = (float)(clock64() - t_issue0); }
  tc_fence_after_sync();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[(size_t)(rank * 128 + row) * N + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc_2cta(tmem, 256); }
}

cudaError_t pgn_launch_probe_umma(const float* A, const float* B, float* D, int K, int N, int variant, int* status,
                                  cudaStream_t stream) {
  if (variant & 2) {
    const size_t smem2 = (size_t)(K / 8) * 2048 + (size_t)(K / 8) * (N / 2) * 16 + 1024;
    cudaError_t e2 = cudaFuncSetAttribute(pgn_probe_umma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e2 != cudaSuccess) return e2;
    pgn_probe_umma2_kernel<<<2, 128, smem2, stream>>>(A, B, D, K, N, (variant & 4) ? 1 : 0, status);
    return cudaGetLastError();
  }
  const size_t smem = (size_t)(K / 8) * 2048 + (size_t)(K / 8) * N * 16 + 1024;
  cudaError_t e = cudaFuncSetAttribute(pgn_probe_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  pgn_probe_umma_kernel<<<1, 128, smem, stream>>>(A, B, D, K, N, variant, status);
  return cudaGetLastError();
}
