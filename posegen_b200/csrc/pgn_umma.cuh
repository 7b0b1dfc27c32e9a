// Thin inline-PTX wrappers for the sm_100a features the bf16 render kernel uses:
// mbarrier, bulk async copies (UBLKCP), tcgen05 MMA / TMEM (UTCHMMA, LDTM), fences.
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" tables.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pgn {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: gives up (and latches *status) after ~2^31 cycles so a pipeline bug
// surfaces as PGN_E_KERNEL instead of a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, volatile int* status, int code) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  int spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 63) == 0) {
      if (*status != 0) return false;
      if (clock64() - t0 > (1ll << 31)) { *status = code; return false; }
    }
  }
  return true;
}

// CTA-pair variants: arrive on the barrier at the same offset in CTA `cta` of the cluster.
// Default (.release.cta / .acquire.cta) semantics on purpose, like CUTLASS' ClusterBarrier: a
// .cluster-scope release/acquire compiles to MEMBAR.ALL.GPU / CCTL.IVALL (L1 invalidate) per call;
// operand hand-off to the tensor core is ordered by fence.proxy.async + the mbarrier itself.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile("{\n .reg .b32 ra;\n mapa.shared::cluster.u32 ra, %0, %1;\n"
               " mbarrier.arrive.shared::cluster.b64 _, [ra];\n}\n" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_cluster(uint64_t* bar, uint32_t parity, volatile int* status, int code) {
  if (mbar_try_wait_cluster(bar, parity)) return true;
  const long long t0 = clock64();
  int spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++spins & 63) == 0) {
      if (*status != 0) return false;
      if (clock64() - t0 > (1ll << 31)) { *status = code; return false; }
    }
  }
  return true;
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// ---- the same primitives on 32-bit shared-window addresses (no generic->shared conversion per call;
// the render kernel computes every barrier address once as smem base + offset)
__device__ __forceinline__ void mbar_arrive_expect_tx_s(uint32_t bar, uint32_t bytes) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the waiting warp is parked by the hardware until the phase completes
// (or the hint expires) instead of spinning through the issue slots -- the render kernel has up to 20 warps
// waiting on barriers at any time and runs at the board's power cap.
__device__ __forceinline__ bool mbar_try_wait_s(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_s(uint32_t bar, uint32_t parity, volatile int* status, int code) {
  if (mbar_try_wait_s(bar, parity)) return true;
  const long long t0 = clock64();
  int spins = 0;
  while (!mbar_try_wait_s(bar, parity)) {
    if ((++spins & 255) == 0) {
      if (*status != 0) return false;
      if (clock64() - t0 > (1ll << 31)) { *status = code; return false; }
    }
  }
  return true;
}
// pure spin on test_wait (no hardware suspend): for a single latency-critical warp (an MMA issuer) whose wake-up latency
// from a parked try_wait would sit on the tensor pipe's critical path
__device__ __forceinline__ bool mbar_test_wait_s(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_spin_s(uint32_t bar, uint32_t parity, volatile int* status, int code) {
  if (mbar_test_wait_s(bar, parity)) return true;
  const long long t0 = clock64();
  int spins = 0;
  while (!mbar_test_wait_s(bar, parity)) {
    if ((++spins & 1023) == 0) {
      if (*status != 0) return false;
      if (clock64() - t0 > (1ll << 31)) { *status = code; return false; }
    }
  }
  return true;
}
__device__ __forceinline__ void mbar_arrive_cluster_s(uint32_t bar, uint32_t cta) {
  asm volatile("{\n .reg .b32 ra;\n mapa.shared::cluster.u32 ra, %0, %1;\n"
               " mbarrier.arrive.shared::cluster.b64 _, [ra];\n}\n" ::"r"(bar), "r"(cta) : "memory");
}

// ------------------------------------------------------------ proxies/fences
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// ------------------------------------------------------------ bulk copy (TMA engine, 1-D)

__device__ __forceinline__ void bulk_g2s_s(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(smem_dst), "l"(gmem_src), "r"(bytes), "r"(bar) : "memory");
}

// ------------------------------------------------------------ TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives lane (base_lane + t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}

// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr) : "memory");
}

// ------------------------------------------------------------ explicit shared-space vector accesses
// (a 16-byte store through a generic pointer may be split into scalar ST.E by the register allocator,
//  which turns a conflict-free 128-bit store into four 4-way-conflicted 32-bit stores)
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t saddr, uint16_t v) {
  asm volatile("st.shared.b16 [%0], %1;\n" ::"r"(saddr), "h"(v) : "memory");
}
__device__ __forceinline__ uint16_t lds16(uint32_t saddr) {
  uint16_t v;
  asm volatile("ld.shared.b16 %0, [%1];\n" : "=h"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr));
  return r;
}

// ------------------------------------------------------------ UMMA descriptors
// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleaved" core matrices):
//   core matrix = 8 rows x 16 bytes, stored as 128 contiguous bytes;
//   LBO = byte distance between the two core matrices adjacent in K inside one K=16 step;
//   SBO = byte distance between core matrices adjacent in M/N (8-row groups).
// bits [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=0
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major.
//   [4,6) c_format=1(f32) | [7,10) a_format=1(bf16) | [10,13) b_format=1 | 15 a_major=0 | 16 b_major=0
//   [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// ---- CTA-pair (cta_group::2) variants: TMEM alloc by the same warp of BOTH CTAs, MMA issued by the
// leader CTA only (UMMA M=256: 128 rows per CTA, B split along N between the two CTAs' smem),
// commit multicast to the barrier at the same offset in both CTAs.
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// relu(lo), relu(hi) -> packed bf16x2 in one instruction
// Warp-uniform issue: the WHOLE warp executes the call convergently and one elected lane issues the
// instruction (operands are warp-uniform, so they live in uniform registers; no per-lane broadcast loops).
__device__ __forceinline__ void umma_bf16_2cta_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n .reg .pred p, e;\n elect.sync _|e, 0xffffffff;\n setp.ne.b32 p, %4, 0;\n"
               " @e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit_2cta_elect_s(uint32_t bar) {
  asm volatile("{\n .reg .pred e;\n elect.sync _|e, 0xffffffff;\n"
               " @e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n}\n"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;\n" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

}  // namespace pgn
