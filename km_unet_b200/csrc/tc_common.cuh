// tc_common.cuh -- inline-PTX helpers shared by the tcgen05 kernels (mbarrier, bulk TMA copy, TMEM, UMMA descriptors)
// and the closed-form Phi evaluators of the KAN producers.
#pragma once
#include <cuda_bf16.h>

#include <cstdio>

#include "common.cuh"

namespace kmu {
namespace tcx {

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded waits: a protocol bug must fault (trap) within a second or two instead of hanging the GPU.
// mbar_wait     -- producers / epilogue / loader warps: back off with nanosleep between polls so that waiting warps do not
//                  take issue slots from the warps doing the work on the same scheduler (the first profile of the forward
//                  kernel showed 40 % of all stall samples on the spin loops' branches).
// mbar_wait_hot -- the single MMA-issuing thread (critical path): tight poll, no clock reads in the loop.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(32);
    if (++spins > (1u << 25)) {
      printf("kmu tcgen05 kernel: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait_hot(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 28)) {
      printf("kmu tcgen05 kernel: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// UMMA shared-memory descriptor, no swizzle, K-major: 8 rows x 16 B core matrices; LBO = byte step between the two
// K-halves of one K=16 instruction, SBO = byte step between consecutive 8-row groups.  (cute::UMMA::SmemDescriptor:
// start[0,14) lbo[16,30) sbo[32,46) version=1 at [46,48) layout_type=0 at [61,64).)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// Same descriptor with the start address moved by byte_off (a multiple of 16).  Shared-memory addresses are < 256 KB, so the
// sum never carries out of the 14-bit field: one 32-bit add per MMA instead of re-deriving the descriptor through a chain
// of dependent uniform-datapath instructions (measured: ~64 cycles per small-N MMA with the chain, issue-bound).
__device__ __forceinline__ uint64_t desc_advance(uint64_t desc, uint32_t byte_off) {
  const uint32_t lo = (uint32_t)desc + (byte_off >> 4);
  return (desc & 0xFFFFFFFF00000000ull) | (uint64_t)lo;
}
// Instruction descriptor for kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), both K-major, N>>3 at
// [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Same, with both operands MN-major (bits 15 / 16): the shared-memory core matrix is then 8 K-rows x 16 B of 8
// MN-contiguous elements; in the descriptor SBO stays the byte step between 8-element MN groups and LBO the byte
// step between the two 8-row K groups of one K=16 instruction (cute make_umma_desc<Major::MN>, INTERLEAVE).
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int M, int N) {
  return make_idesc_bf16(M, N) | (1u << 15) | (1u << 16);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r[0]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ Phi producers
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// Eight cubic B-spline values of x on uniform knots t_j = t0 + j h (j = 0..11) as 8 bf16: only basis i-3..i are non-zero
// for x in knot span i; they are computed in closed form and shifted into place.  Zero outside [t0, t11).
__device__ __forceinline__ uint4 spline_group(float x, float t0, float inv_h) {
  float s = (x - t0) * inv_h;
  unsigned long long lo = 0ull, hi = 0ull;
  if (s >= 0.f && s < 11.f) {
    float fi = floorf(s);
    float u = s - fi;
    int i = (int)fi;
    float u2 = u * u, u3 = u2 * u, om = 1.f - u;
    const float k6 = 1.0f / 6.0f;
    float w0 = om * om * om * k6;
    float w1 = (3.f * u3 - 6.f * u2 + 4.f) * k6;
    float w2 = (-3.f * u3 + 3.f * u2 + 3.f * u + 1.f) * k6;
    float w3 = u3 * k6;
    unsigned long long v = ((unsigned long long)pack_bf16x2(w2, w3) << 32) | (unsigned long long)pack_bf16x2(w0, w1);
    int a = 16 * (i - 3);  // bit position of basis i-3 inside the 128-bit group, may be negative
    lo = a >= 0 ? (a < 64 ? (v << a) : 0ull) : (v >> (-a));
    int b = a - 64;
    hi = b >= 0 ? (v << b) : ((-b) < 64 ? (v >> (-b)) : 0ull);
  }
  return make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
}


}  // namespace tcx
}  // namespace kmu
