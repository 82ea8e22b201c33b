// Helpers shared by the tcgen05 rollout kernels (decode_tc.cu, decode_tc2.cu): fp16 hi/lo splits, the visited-mask
// word select, TMEM <-> register transfers of 8/16/32 columns, Philox4x32-10 and Gumbel noise for sampling rollouts.
#pragma once
#include "common.cuh"
#include "mma_util.cuh"
#include "tc_util.cuh"

namespace htsp {

__device__ __forceinline__ void tcd_split8(const float* v, uint4& hi, uint4& lo) {
  __half2 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    float2 f = __half22float2(h[i]);
    l[i] = __floats2half2_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
  }
  hi = *reinterpret_cast<uint4*>(h);
  lo = *reinterpret_cast<uint4*>(l);
}

// sm_100 mixed-precision add (one FHADD: fp16 operand read straight from a packed register half)
__device__ __forceinline__ float tcd_sub_h_f(uint32_t pair, int hi_half, float x) {   // half - x
  float r;
  const unsigned short h = hi_half ? (unsigned short)(pair >> 16) : (unsigned short)(pair & 0xffffu);
  asm("sub.rn.f32.f16 %0, %1, %2;" : "=f"(r) : "h"(h), "f"(x));
  return r;
}
__device__ __forceinline__ float tcd_add_h_f(uint32_t pair, int hi_half, float acc) {   // half + acc
  float r;
  const unsigned short h = hi_half ? (unsigned short)(pair >> 16) : (unsigned short)(pair & 0xffffu);
  asm("add.rn.f32.f16 %0, %1, %2;" : "=f"(r) : "h"(h), "f"(acc));
  return r;
}
// x = hi + lo for a pair: hi = RN_fp16(x); x - hi is exact in fp32 (<= 13 significant bits), lo = RN_fp16(x - hi).
// The residual is one packed FFMA2 after unpacking hi (2 HADD2.F32); the 2-FHADD form (tcd_sub_h_f) has
// fewer instructions but measured 1.8 % slower on the whole rollout.
__device__ __forceinline__ void tcd_split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 hh = __floats2half2_rn(x0, x1);
  hi = *reinterpret_cast<const uint32_t*>(&hh);
  const float2 r = __ffma2_rn(__half22float2(hh), make_float2(-1.f, -1.f), make_float2(x0, x1));
  const __half2 ll = __floats2half2_rn(r.x, r.y);
  lo = *reinterpret_cast<const uint32_t*>(&ll);
}
__device__ __forceinline__ void tcd_split8_packed(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) tcd_split_pair(v[2 * i], v[2 * i + 1], h[i], l[i]);
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ unsigned tcd_word(const unsigned (&vis)[7], int w) {
  unsigned r = vis[0];
#pragma unroll
  for (int i = 1; i < 7; ++i) r = (w == i) ? vis[i] : r;
  return r;
}

// TMEM <-> registers for W = 16 or 32 columns into / from the first W entries of a 32-register array
template <int W>
__device__ __forceinline__ void tld(uint32_t taddr, uint32_t (&r)[32]) {
  if (W == 32) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
  } else {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
  }
}
// completes every earlier tcgen05.ld of this thread; the "+r" operands tie the registers to the wait
__device__ __forceinline__ void tld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                 "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]),
                 "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]),
                 "+r"(r[29]), "+r"(r[30]), "+r"(r[31])::"memory");
}
__device__ __forceinline__ void tld_wait16(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]),
                 "+r"(r[15])::"memory");
}
template <int W>
__device__ __forceinline__ void tst(uint32_t taddr, const uint32_t (&r)[32]) {
  if (W == 8) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
  } else if (W == 32) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
        "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
  } else {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
  }
}

// Philox4x32-10 (counter-based: the noise of (instance, row, step, key quad) does not depend on the launch geometry)
__device__ __forceinline__ uint4 tcd_philox(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
// standard Gumbel noise -ln(-ln u), u uniform in (0,1) from 24 random bits
__device__ __forceinline__ float tcd_gumbel(uint32_t bits) {
  const float u = ((float)(bits >> 8) + 0.5f) * 5.9604644775390625e-08f;
  return -__logf(-__logf(u));
}

// 2^x for a pair on the FMA pipe (offloads MUFU.EX2): Cody-Waite split around the nearest integer, degree-4 minimax
// polynomial on [-0.5, 0.5] (max relative error 2.7e-6), exponent inserted with an integer shift-add; inputs below
// -125 (including -inf) return 2^-125
__device__ __forceinline__ float2 ex2_poly2(float x0, float x1) {
  constexpr float kMagic = 12582912.f;   // 1.5 * 2^23
  const float2 x = make_float2(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const float2 t = __fadd2_rn(x, make_float2(kMagic, kMagic));
  const float2 nf = __fadd2_rn(t, make_float2(-kMagic, -kMagic));
  const float2 f = __ffma2_rn(nf, make_float2(-1.f, -1.f), x);
  float2 p = make_float2(0.009570099413394928f, 0.009570099413394928f);
  p = __ffma2_rn(p, f, make_float2(0.05591785907745361f, 0.05591785907745361f));
  p = __ffma2_rn(p, f, make_float2(0.240247443318367f, 0.240247443318367f));
  p = __ffma2_rn(p, f, make_float2(0.6931217908859253f, 0.6931217908859253f));
  p = __ffma2_rn(p, f, make_float2(0.9999992847442627f, 0.9999992847442627f));
  float2 r;
  r.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return r;
}

}  // namespace htsp
