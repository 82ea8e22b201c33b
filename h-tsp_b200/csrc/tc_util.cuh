// tcgen05 / TMEM / mbarrier PTX helpers for the Blackwell-native rollout kernel (decode_tc.cu).
// Descriptor formats follow cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor); the
// shared-memory operand layout is K-major SWIZZLE_128B throughout (rows of 128 bytes, 8-row atoms of
// 1024 bytes, the 16-byte chunk index of a row XORed with row%8), the same layout gemm_tc.cu gets
// from TMA tensor maps -- here the tiles are written pre-swizzled by generic stores / bulk copies.
#pragma once
#include "common.cuh"

namespace htsp {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
#ifdef HTSP_DEBUG_WAIT
// lab builds (-DHTSP_DEBUG_WAIT): every warp records which barrier it is waiting for; a wait that times out prints the
// records of its CTA's warps and then lets the kernel run to its end
__device__ int g_wait_abort = 0;
__device__ int g_wait_victim = -1;
__device__ int g_prog[512 * 32];
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t tag = 0) {
  uint32_t done = 0;
  const int me = (int)(blockIdx.x & 511) * 32 + ((int)threadIdx.x >> 5);
  if ((threadIdx.x & 31) == 0) *(volatile int*)&g_prog[me] = 0x10000 | ((int)(bar - 116544u) / 8 << 4) | (int)parity;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if ((spin & 0xfffu) == 0xfffu && *(volatile int*)&g_wait_abort) return;
    if (spin == (1u << 17)) {
      if (atomicAdd(&g_wait_victim, 1) < 6) {
        printf("[wait timeout] block %d warp %d lane %d barrier %d parity %u tag %u\n", (int)blockIdx.x, (int)threadIdx.x >> 5,
               (int)threadIdx.x & 31, (int)(bar - 116544u) / 8, parity, tag);
        for (int w = 0; w < 16; ++w) {
          const int v = *(volatile int*)&g_prog[(int)(blockIdx.x & 511) * 32 + w];
          printf("   block %d warp %d: state 0x%x (waiting %d, barrier %d, parity %d)\n", (int)blockIdx.x, w, v, v >> 16, (v >> 4) & 0xfff, v & 1);
        }
      }
    }
    if (spin > (1u << 20)) {
      atomicAdd(&g_wait_abort, 1);
      return;
    }
  }
  if ((threadIdx.x & 31) == 0) *(volatile int*)&g_prog[me] = 0x20000 | ((int)(bar - 116544u) / 8 << 4) | (int)parity;   // passed
}
#else
// bounded wait: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU.  The polling loops of the
// waiting warps are a quarter of all issued instructions of the rollout kernels (ncu, profiles/r02_decode_tc2_*) and
// compete with the softmax warps for issue slots; probing MORE often per counter update makes it worse (measured on
// decode_tc2, 1184 instances: 1 probe per iteration 46.4 ms, 4 probes 46.85 ms, 8 probes 47.8 ms), a __nanosleep
// back-off or a suspend-time hint adds wake-up latency to the decode chain (-0.4 % / -1 %): one probe it stays.
__device__ __forceinline__ bool mbar_probe(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t tag = 0) {
  (void)tag;       // (lab builds print it)
  for (uint32_t spin = 0;; ++spin) {
#ifndef HTSP_WAIT_PROBES
#define HTSP_WAIT_PROBES 1
#endif
#pragma unroll
    for (int i = 0; i < HTSP_WAIT_PROBES; ++i)
      if (mbar_probe(bar, parity)) return;
    if (spin > (1u << 21) / HTSP_WAIT_PROBES) __trap();
  }
}
#endif
// one lane of a converged warp (the others get false)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n.reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(pred));
  return pred != 0;
}
// non-blocking probe of a phase (used by the producer that serves two rings)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void fence_async_smem() {   // generic-proxy smem writes -> async proxy (MMA)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_128B matrix descriptor: start>>4 | SBO = 1024 B (8 rows x 128 B) | version 1 | layout 2
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor, kind::f16: D = f32, A = B = f16, both K-major, M = 128
__host__ __device__ constexpr uint32_t idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T   (A: lane = row, one 32-bit column = two consecutive k)
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Cheap-to-issue forms: the descriptors are passed as their low words (start address >> 4; the high
// word of a K-major SWIZZLE_128B descriptor is the constant kDescHi), so stepping through k-slices
// or key blocks is one 32-bit add per operand.
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ void mma_ss_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
      : "memory");
}
__device__ __forceinline__ void mma_ts_lo(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 db;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
      : "memory");
}
// kind::tf32 (A, B = fp32 words of which the hardware reads the top 19 bits -- measured, tools/lab_tc.cu part A)
__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// high descriptor word of a K-major SWIZZLE_64B operand (rows of 64 bytes, 8-row atoms of 512 bytes)
constexpr uint32_t kDescHi64 = (512u >> 4) | (1u << 14) | (4u << 29);
// A: K-major SWIZZLE_128B, B: K-major SWIZZLE_64B (the per-head key blocks of decode_tc2.cu)
__device__ __forceinline__ void mma_ss_lo_b64(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 da, db;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 da, {%1, %5};\n"
      "mov.b64 db, {%2, %6};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi), "r"(kDescHi64)
      : "memory");
}
// D[tmem] (+)= A[tmem, fp32 words] * B[smem, tf32 K-major SWIZZLE_128B]^T
__device__ __forceinline__ void mma_ts_tf32_lo(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 db;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "mov.b64 db, {%2, %5};\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
      : "memory");
}
__host__ __device__ __forceinline__ uint32_t sw64_off(uint32_t row, uint32_t chunk) {
  return row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4);
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

// TMEM <-> registers, 32 lanes x 32 bit, one row (lane) per thread.  The loads are asynchronous:
// ld16_wait() makes the registers of every earlier load valid AND ties them to the wait through
// "+r" operands so the compiler cannot move a use above it.
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ld16_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                 "+r"(r[15])::"memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// byte offset of element (row, 16-byte chunk) inside a K-major SWIZZLE_128B block (rows of 128 bytes)
__host__ __device__ __forceinline__ uint32_t sw128_off(uint32_t row, uint32_t chunk) {
  return row * 128u + ((chunk ^ (row & 7u)) << 4);
}

}  // namespace tc
}  // namespace htsp
