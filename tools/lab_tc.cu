// Laboratory for the rollout kernel's design decisions (tools only; not part of libhtsp_b200):
//   part A  tcgen05.mma semantics the kernel relies on, checked against the CPU:
//           (1) kind::tf32 with A = fp32 words left in place in TMEM (hardware reads the top 19 bits),
//           (2) kind::f16 with A = bf16 (hi, lo) pairs in one TMEM word and B = fp16 (mixed formats),
//           (3) the same with B = bf16;
//   part B  issue / MUFU throughput of candidate softmax inner loops with the production warp layout
//           (16 softmax warps, four per scheduler, one row per lane, 16-key chunks through tcgen05.ld/st),
//           no MMAs and no barriers: cycles per 16-key chunk and warp for each variant.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o lab_tc tools/lab_tc.cu
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__host__ __device__ __forceinline__ uint32_t sw128_off(uint32_t row, uint32_t chunk) { return row * 128u + ((chunk ^ (row & 7u)) << 4); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// a_fmt / b_fmt: 0 = f16, 1 = bf16, 2 = tf32; D = f32, both K-major, M = 128
__host__ __device__ constexpr uint32_t idesc(int n, int a_fmt, int b_fmt) {
  return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ts_f16(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t id, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(bdesc), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_tf32(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t id, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(bdesc), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void tld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tld16_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
}
__device__ __forceinline__ void tst16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// =====================================================================================================
// part A
// =====================================================================================================
// test: 0 = tf32 (A fp32 words), 1 = A bf16 pairs / B fp16, 2 = A bf16 pairs / B bf16.
// A: [128 rows x 32 words], B image: 4 KB pre-swizzled K-major SWIZZLE_128B block [32 rows x 128 B], D: [128 x 32]
__global__ void __launch_bounds__(128) lab_mma_kernel(const uint32_t* __restrict__ A, const unsigned char* __restrict__ Bimg, float* __restrict__ D, int test) {
  __shared__ __align__(1024) unsigned char bsm[4096];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 4096 / 16; i += 128) ((uint4*)bsm)[i] = ((const uint4*)Bimg)[i];
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tb = tmem_slot;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  uint32_t a[16];
  for (int c = 0; c < 32; c += 16) {
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = A[tid * 32 + c + j];
    tst16(tb + lane_addr + c, a);
  }
  st_wait();
  fence_before();
  __syncthreads();
  fence_after();
  if (tid == 0) {
    const uint32_t baddr = smem_u32(bsm);
    if (test == 0) {
      const uint32_t id = idesc(32, 2, 2);
      for (int ks = 0; ks < 4; ++ks) mma_ts_tf32(tb + 64, tb + 8 * ks, desc_sw128(baddr + 32 * ks), id, ks != 0);
    } else {
      const uint32_t id = idesc(32, 1, test == 1 ? 0 : 1);
      for (int ks = 0; ks < 4; ++ks) mma_ts_f16(tb + 64, tb + 8 * ks, desc_sw128(baddr + 32 * ks), id, ks != 0);
    }
    commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  fence_after();
  for (int c = 0; c < 32; c += 16) {
    tld16(tb + lane_addr + 64 + c, a);
    tld16_wait(a);
#pragma unroll
    for (int j = 0; j < 16; ++j) D[tid * 32 + c + j] = __uint_as_float(a[j]);
  }
  fence_before();
  __syncthreads();
  if (warp == 0) { fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(128) : "memory"); }
}

static float trunc_bits(float x, int drop) { uint32_t u; memcpy(&u, &x, 4); u &= ~((1u << drop) - 1u); float y; memcpy(&y, &u, 4); return y; }
static float rn_bits(float x, int drop) { uint32_t u; memcpy(&u, &x, 4); u = (u + (1u << (drop - 1))) & ~((1u << drop) - 1u); float y; memcpy(&y, &u, 4); return y; }
static float frand() { return (float)rand() / (float)RAND_MAX; }

static void part_a() {
  uint32_t* dA; unsigned char* dB; float* dD;
  CK(cudaMalloc(&dA, 128 * 32 * 4)); CK(cudaMalloc(&dB, 4096)); CK(cudaMalloc(&dD, 128 * 32 * 4));
  std::vector<float> D(128 * 32);
  for (int test = 0; test < 3; ++test) {
    if (test == 1) { printf("{\"part\": \"A\", \"test\": \"kind::f16 A=bf16 pair, B=fp16\", \"result\": \"illegal instruction on sm_100a (measured): A and B formats must agree\"}\n"); continue; }
    std::vector<uint32_t> A(128 * 32);
    std::vector<unsigned char> B(4096, 0);
    std::vector<double> ref_t(128 * 32, 0.0), ref_r(128 * 32, 0.0), ref_f(128 * 32, 0.0);
    std::vector<float> pa(128 * 32), vb(32 * 32);
    for (auto& v : pa) v = expf(4.f * frand() - 2.f);
    for (auto& v : vb) v = 2.f * frand() - 1.f;
    if (test == 0) {
      for (int i = 0; i < 128 * 32; ++i) memcpy(&A[i], &pa[i], 4);
      for (int n = 0; n < 32; ++n) for (int k = 0; k < 32; ++k) memcpy(&B[sw128_off(n, k >> 2) + (k & 3) * 4], &vb[n * 32 + k], 4);
      for (int r = 0; r < 128; ++r) for (int n = 0; n < 32; ++n) for (int k = 0; k < 32; ++k) {
        ref_t[r * 32 + n] += (double)trunc_bits(pa[r * 32 + k], 13) * (double)trunc_bits(vb[n * 32 + k], 13);
        ref_r[r * 32 + n] += (double)rn_bits(pa[r * 32 + k], 13) * (double)rn_bits(vb[n * 32 + k], 13);
        ref_f[r * 32 + n] += (double)pa[r * 32 + k] * (double)vb[n * 32 + k];
      }
    } else {
      // A word j of a row: key j as (lo bf16 | hi bf16 << 16); B row n: k slots 2j, 2j+1 both hold v(n, j)
      for (int i = 0; i < 128 * 32; ++i) {
        const float hi = trunc_bits(pa[i], 16), lo = trunc_bits(pa[i] - hi, 16);
        uint32_t uh, ul; memcpy(&uh, &hi, 4); memcpy(&ul, &lo, 4);
        A[i] = (uh & 0xffff0000u) | (ul >> 16);
        pa[i] = hi + lo;
      }
      for (int n = 0; n < 32; ++n) for (int j = 0; j < 32; ++j) {
        uint16_t h;
        float back;
        if (test == 1) { __half hh = __float2half_rn(vb[n * 32 + j]); memcpy(&h, &hh, 2); back = __half2float(hh); }
        else { __nv_bfloat16 hh = __float2bfloat16_rn(vb[n * 32 + j]); memcpy(&h, &hh, 2); back = __bfloat162float(hh); }
        vb[n * 32 + j] = back;
        for (int e = 0; e < 2; ++e) { const int k = 2 * j + e; memcpy(&B[sw128_off(n, k >> 3) + (k & 7) * 2], &h, 2); }
      }
      for (int r = 0; r < 128; ++r) for (int n = 0; n < 32; ++n) for (int k = 0; k < 32; ++k)
        ref_f[r * 32 + n] += (double)pa[r * 32 + k] * (double)vb[n * 32 + k];
      ref_t = ref_f; ref_r = ref_f;
    }
    CK(cudaMemcpy(dA, A.data(), 128 * 32 * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), 4096, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, 128 * 32 * 4));
    lab_mma_kernel<<<1, 128>>>(dA, dB, dD, test);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"part\": \"A\", \"test\": %d, \"error\": \"%s\"}\n", test, cudaGetErrorString(e)); exit(1); }
    CK(cudaMemcpy(D.data(), dD, 128 * 32 * 4, cudaMemcpyDeviceToHost));
    double et = 0, er = 0, ef = 0, mag = 0;
    for (int i = 0; i < 128 * 32; ++i) {
      et = fmax(et, fabs(D[i] - ref_t[i])); er = fmax(er, fabs(D[i] - ref_r[i])); ef = fmax(ef, fabs(D[i] - ref_f[i]));
      mag = fmax(mag, fabs(ref_f[i]));
    }
    const char* names[3] = {"tf32 A=fp32 words in TMEM", "kind::f16 A=bf16 pair, B=fp16", "kind::f16 A=bf16 pair, B=bf16"};
    printf("{\"part\": \"A\", \"test\": \"%s\", \"max_abs_diff_vs_truncated_inputs\": %.3e, \"vs_rn_inputs\": %.3e, \"vs_full_inputs\": %.3e, \"max_abs_ref\": %.3f}\n",
           names[test], et, er, ef, mag);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

// =====================================================================================================
// part B
// =====================================================================================================
constexpr int NP = 208;
constexpr float kMagic = 12582912.f;   // 1.5 * 2^23
// 2^x for a pair on the FMA pipe: Cody-Waite split around the nearest integer, degree-4 minimax on [-0.5, 0.5]
// (max relative error 2.7e-6), exponent inserted with an integer shift-add.  x >= -125 after the clamp.
__device__ __forceinline__ float2 ex2_poly2(float x0, float x1) {
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
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 hh = __floats2half2_rn(x0, x1);
  hi = *reinterpret_cast<const uint32_t*>(&hh);
  const float2 r = __ffma2_rn(__half22float2(hh), make_float2(-1.f, -1.f), make_float2(x0, x1));
  const __half2 ll = __floats2half2_rn(r.x, r.y);
  lo = *reinterpret_cast<const uint32_t*>(&ll);
}
__device__ __forceinline__ unsigned vis_word(const unsigned (&vis)[7], int w) {
  unsigned r = vis[0];
#pragma unroll
  for (int i = 1; i < 7; ++i) r = (w == i) ? vis[i] : r;
  return r;
}

// VARIANT 0: production x3 chunk (shift, ex2, row sum, fp16 hi/lo split)
//         1: tf32 in place: mask, ex2, store                     (KPOLY of the 16 elements on the FMA pipe)
//         2: bf16 (hi, lo) pair in place: mask, ex2, split 3 instr, store   (KPOLY as above)
//         3: as 1 but the mask is applied after the exponential (FSEL to zero)
template <int VARIANT, int KPOLY>
__device__ __forceinline__ float chunk_body(uint32_t tin, uint32_t tout, unsigned bits, float m) {
  uint32_t s[16], out[16];
  tld16(tin, s);
  tld16_wait(s);
  float csum = 0.f;
  if (VARIANT == 0) {
    float tv[16], p[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) tv[j] = ((bits >> j) & 1u) ? -INFINITY : __uint_as_float(s[j]);
    const float2 nm = make_float2(-m, -m);
    float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float2 d0 = __fadd2_rn(make_float2(tv[j], tv[j + 1]), nm);
      const float2 d1 = __fadd2_rn(make_float2(tv[j + 2], tv[j + 3]), nm);
      p[j] = ex2_approx(d0.x); p[j + 1] = ex2_approx(d0.y); p[j + 2] = ex2_approx(d1.x); p[j + 3] = ex2_approx(d1.y);
      acc0 = __fadd2_rn(acc0, make_float2(p[j], p[j + 1]));
      acc1 = __fadd2_rn(acc1, make_float2(p[j + 2], p[j + 3]));
    }
    acc0 = __fadd2_rn(acc0, acc1);
    csum = acc0.x + acc0.y;
#pragma unroll
    for (int j = 0; j < 16; j += 2) split_pair(p[j], p[j + 1], out[j >> 1], out[8 + (j >> 1)]);
  } else {
    float p[16];
    // the first KPOLY elements (in pairs) go through the polynomial, the others through MUFU; interleaved by pair
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const bool poly = (j < KPOLY);
      float x0 = __uint_as_float(s[j]), x1 = __uint_as_float(s[j + 1]);
      if (VARIANT != 3 && !poly) {
        x0 = ((bits >> j) & 1u) ? -INFINITY : x0;
        x1 = ((bits >> (j + 1)) & 1u) ? -INFINITY : x1;
      }
      if (poly) {
        const float2 r = ex2_poly2(x0, x1);
        p[j] = ((bits >> j) & 1u) ? 0.f : r.x;
        p[j + 1] = ((bits >> (j + 1)) & 1u) ? 0.f : r.y;
      } else {
        p[j] = ex2_approx(x0);
        p[j + 1] = ex2_approx(x1);
        if (VARIANT == 3) {
          p[j] = ((bits >> j) & 1u) ? 0.f : p[j];
          p[j + 1] = ((bits >> (j + 1)) & 1u) ? 0.f : p[j + 1];
        }
      }
    }
    if (VARIANT == 2) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t u = __float_as_uint(p[j]);
        const float hi = __uint_as_float(u & 0xffff0000u);
        const float r = p[j] - hi;
        out[j] = __byte_perm(__float_as_uint(r), u, 0x7632);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) out[j] = __float_as_uint(p[j]);
    }
  }
  tst16(tout, out);
  return csum;
}


// software-pipelined variants: the tcgen05.ld of chunk c+1 is in flight while chunk c is processed (two register
// buffers).  VARIANT 4: tf32 in place (mask, ex2, store); VARIANT 5: the x3 production arithmetic.
template <int VARIANT>
__device__ __forceinline__ void process16(uint32_t (&s)[16], uint32_t tout, unsigned bits, float m, float& l) {
  uint32_t out[16];
  if (VARIANT == 4) {
#pragma unroll
    for (int j = 0; j < 16; ++j) out[j] = __float_as_uint(ex2_approx(((bits >> j) & 1u) ? -INFINITY : __uint_as_float(s[j])));
  } else {
    float tv[16], p[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) tv[j] = ((bits >> j) & 1u) ? -INFINITY : __uint_as_float(s[j]);
    const float2 nm = make_float2(-m, -m);
    float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float2 d0 = __fadd2_rn(make_float2(tv[j], tv[j + 1]), nm);
      const float2 d1 = __fadd2_rn(make_float2(tv[j + 2], tv[j + 3]), nm);
      p[j] = ex2_approx(d0.x); p[j + 1] = ex2_approx(d0.y); p[j + 2] = ex2_approx(d1.x); p[j + 3] = ex2_approx(d1.y);
      acc0 = __fadd2_rn(acc0, make_float2(p[j], p[j + 1]));
      acc1 = __fadd2_rn(acc1, make_float2(p[j + 2], p[j + 3]));
    }
    acc0 = __fadd2_rn(acc0, acc1);
    l += acc0.x + acc0.y;
#pragma unroll
    for (int j = 0; j < 16; j += 2) split_pair(p[j], p[j + 1], out[j >> 1], out[8 + (j >> 1)]);
  }
  tst16(tout, out);
}
template <int VARIANT>
__device__ __forceinline__ float prefetch_loop(uint32_t tS, int c_lo, int c_hi, const unsigned (&vis)[7], int reps) {
  float l = 0.f;
  uint32_t a[16], b[16];
  for (int r = 0; r < reps; ++r) {
#pragma unroll 1
    for (int hd = 0; hd < 8; ++hd) {
      const float m = (float)(3 + (hd & 1));
      tld16(tS + 16 * c_lo, a);
#pragma unroll 1
      for (int c = c_lo; c < c_hi; c += 2) {
        const unsigned w0 = vis_word(vis, c >> 1) >> ((c & 1) * 16), w1 = vis_word(vis, (c + 1) >> 1) >> (((c + 1) & 1) * 16);
        tld16_wait(a);
        if (c + 1 < c_hi) tld16(tS + 16 * (c + 1), b);
        process16<VARIANT>(a, tS + 16 * c, w0 & 0xffffu, m, l);
        if (c + 1 < c_hi) {
          tld16_wait(b);
          if (c + 2 < c_hi) tld16(tS + 16 * (c + 2), a);
          process16<VARIANT>(b, tS + 16 * (c + 1), w1 & 0xffffu, m, l);
        }
      }
    }
  }
  st_wait();
  return l;
}

template <int VARIANT, int KPOLY>
__global__ void __launch_bounds__(640, 1) lab_softmax_kernel(long long* __restrict__ cycles, float* __restrict__ sink, int reps, unsigned seed, unsigned warp_mask) {
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tb = tmem_slot;
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;" ::: "memory");
    const int sw = warp - 4;
    const int t = sw >> 3, half = (sw >> 2) & 1, q = warp & 3;
    const bool active = !(t == 1 && q == 3) && ((warp_mask >> sw) & 1u);            // 200 rows: tile 1 has three lane quarters
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t tS = tb + lane_addr + (uint32_t)(t * NP);
    const int nch = NP >> 4, n0 = nch >> 1;
    const int c_lo = half ? n0 : 0, c_hi = half ? nch : n0;
    // scores in [-3, 3], written once; visited bits ~50 %
    unsigned h = seed ^ (unsigned)(blockIdx.x * 977 + tid * 31);
    auto rnd = [&]() { h = h * 1664525u + 1013904223u; return h; };
    if (active) {
      for (int c = c_lo; c < c_hi; ++c) {
        uint32_t v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(6.f * ((float)(rnd() >> 8) * (1.f / 16777216.f)) - 3.f);
        tst16(tS + 16 * c, v);
      }
      st_wait();
    }
    unsigned vis[7];
#pragma unroll
    for (int w = 0; w < 7; ++w) vis[w] = rnd() ^ (rnd() << 7);
    float acc = 0.f;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    const long long t0 = clock64();
    if (VARIANT == 4 || VARIANT == 5) {
      if (active) acc = prefetch_loop<VARIANT>(tS, c_lo, c_hi, vis, reps);
    } else
    if (active) {
      for (int r = 0; r < reps; ++r) {
#pragma unroll 1
        for (int hd = 0; hd < 8; ++hd) {
          float l = 0.f;
          const float m = (float)(3 + (hd & 1));
#pragma unroll 1
          for (int w = c_lo >> 1; 2 * w < c_hi; ++w) {
            const unsigned word = vis_word(vis, w);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int c = 2 * w + e;
              if (c >= c_lo && c < c_hi)
                l += chunk_body<VARIANT, KPOLY>(tS + 16 * c, tS + 16 * c, e ? (word >> 16) : (word & 0xffffu), m);
            }
          }
          acc += l;
          // in-place results are not valid scores for the next pass: rewrite nothing, but keep the values finite
          // by clearing the exponent's top bit on reload is not needed for timing (MUFU / FMA latency is data independent)
        }
      }
      st_wait();
    }
    const long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x * 16 + sw] = active ? (t1 - t0) : 0;
    if (acc == 123.456f) sink[0] = acc;
  }
  fence_before();
  __syncthreads();
  if (warp == 1) { fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory"); }
}


// =====================================================================================================
// part C: dispatch rate of small tcgen05.mma (one issuing thread, back to back), with the A operand in TMEM
// (".ts": P V) or in shared memory (".ss": S, logits), alone and while 16 softmax warps stream tcgen05.ld
// =====================================================================================================
__device__ __forceinline__ void mma_ss_f16(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t id, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(id), "r"(acc) : "memory");
}
// kind: 0 = f16 .ts, 1 = tf32 .ts, 2 = f16 .ss;  ldtm: 0 = tensor pipe alone, 1 = 16 warps loop over tcgen05.ld.x16
__global__ void __launch_bounds__(640, 1) lab_issue_kernel(long long* __restrict__ out, float* __restrict__ sink, int kind, int n, int n_mma, int ldtm, int n_acc) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  for (int i = tid; i < 65536 / 16; i += 640) ((uint4*)sm)[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tb = tmem_slot;
  if (warp >= 4) {     // finite values in the A columns
    const uint32_t la = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 0x3c003c00u;
    if (warp < 8) for (int c = 0; c < 416; c += 16) tst16(tb + la + c, v);
    st_wait();
  }
  fence_before();
  __syncthreads();
  fence_after();
  if (warp == 3 || (warp == 2 && n_acc == 3)) {
    if (lane == 0) {
      const uint32_t id = idesc(n, kind == 1 ? 2 : 0, kind == 1 ? 2 : 0);
      const uint64_t bdesc = desc_sw128(smem_u32(sm));
      const uint64_t adesc = desc_sw128(smem_u32(sm) + 32768);
      const long long t0 = clock64();
      const uint32_t d0 = tb + 128 + (warp == 2 ? 208u : 0u);
      uint64_t* mybar = &bar + 0;
      const uint32_t dstep = (n_acc > 1) ? (uint32_t)n : 0u;
      // tight issue loop: 8 MMAs per iteration over 8 A slices, up to two accumulators, no divisions / branches on kind
      if (kind == 0) {
        for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) mma_ts_f16(d0 + (u & 1) * dstep, tb + 8 * u, bdesc, id, 1);
        }
      } else if (kind == 1) {
        for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) mma_ts_tf32(d0 + (u & 1) * dstep, tb + 8 * u, bdesc, id, 1);
        }
      } else {
        for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) mma_ss_f16(d0 + (u & 1) * dstep, adesc, bdesc, id, 1);
        }
      }
      const long long t1 = clock64();
      commit(smem_u32(warp == 3 ? &bar : &bar2));
      mbar_wait(smem_u32(warp == 3 ? &bar : &bar2), 0);
      const long long t2 = clock64();
      if (warp == 3) stop = 1;
      if (blockIdx.x == 0 && warp == 3) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  } else if (warp >= 4 && ldtm) {
    const uint32_t la = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    long long n_ld = 0;
    const long long t0 = clock64();
    while (!stop) {
      for (int c = 0; c < 13; ++c) {
        uint32_t v[16];
        tld16(tb + la + (uint32_t)(((warp >> 2) & 1) * 208 + 16 * c), v);
        tld16_wait(v);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc ^= v[j];
        ++n_ld;
      }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) { out[2 + 2 * (warp - 4)] = n_ld; out[3 + 2 * (warp - 4)] = t1 - t0; }
    if (acc == 0x12345u) sink[0] = 1.f;
  }
  fence_before();
  __syncthreads();
  if (warp == 1) { fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory"); }
}

static void part_c() {
  long long* dO; float* dS;
  CK(cudaMalloc(&dO, 64 * sizeof(long long))); CK(cudaMalloc(&dS, 4));
  CK(cudaFuncSetAttribute(lab_issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const char* kinds[3] = {"f16 A in TMEM", "tf32 A in TMEM", "f16 A in smem"};
  const int ns[] = {16, 32, 48, 112, 208};
  for (int ldtm = 0; ldtm < 2; ++ldtm)
    for (int kind = 0; kind < 3; ++kind)
      for (int n : ns)
       for (int n_acc = 1; n_acc <= 3; n_acc += 2) {
        if (n > 176) continue;
        const int n_mma = 2048;
        CK(cudaMemset(dO, 0, 64 * sizeof(long long)));
        for (int it = 0; it < 2; ++it) {
          lab_issue_kernel<<<148, 640, 65536>>>(dO, dS, kind, n, n_mma, ldtm, n_acc);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("{\"part\": \"C\", \"error\": \"%s\"}\n", cudaGetErrorString(e)); exit(1); }
        }
        long long o[64];
        CK(cudaMemcpy(o, dO, sizeof(o), cudaMemcpyDeviceToHost));
        double ld_bytes_per_clk = 0;
        if (ldtm) for (int w = 0; w < 16; ++w) if (o[3 + 2 * w] > 0) ld_bytes_per_clk += (double)o[2 + 2 * w] * 2048.0 / (double)o[3 + 2 * w];
        printf("{\"part\": \"C\", \"kind\": \"%s\", \"N\": %d, \"issuers\": %d, \"concurrent_ldtm_warps\": %d, \"issue_clk_per_mma\": %.1f, \"clk_per_mma_to_completion\": %.1f, \"ldtm_bytes_per_clk_sm\": %.1f}\n",
               kinds[kind], n, n_acc == 3 ? 2 : 1, ldtm ? 16 : 0, (double)o[0] / n_mma, (double)o[1] / n_mma, ld_bytes_per_clk);
      }
  cudaFree(dO); cudaFree(dS);
}

// =====================================================================================================
// part D: kind::f16 with A in a K-major SWIZZLE_128B tile and B in a K-major SWIZZLE_64B tile (rows of 64 bytes:
// 16-byte chunk index XOR ((row >> 1) & 3), 8-row atoms of 512 bytes) -- the layout of the per-head key blocks
// =====================================================================================================
__host__ __device__ __forceinline__ uint32_t sw64_off(uint32_t row, uint32_t chunk) { return row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4); }
__global__ void __launch_bounds__(128) lab_sw64_kernel(const unsigned char* __restrict__ Aimg, const unsigned char* __restrict__ Bimg, float* __restrict__ D) {
  __shared__ __align__(1024) unsigned char asm_[16384];
  __shared__ __align__(1024) unsigned char bsm[4096];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 16384 / 16; i += 128) ((uint4*)asm_)[i] = ((const uint4*)Aimg)[i];
  for (int i = tid; i < 4096 / 16; i += 128) ((uint4*)bsm)[i] = ((const uint4*)Bimg)[i];
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tb = tmem_slot;
  if (tid == 0) {
    const uint32_t id = idesc(64, 0, 0);
    for (int ks = 0; ks < 2; ++ks) {
      const uint64_t ad = desc_sw128(smem_u32(asm_) + 32 * ks);
      uint64_t bd = 0;
      bd |= (uint64_t)(((smem_u32(bsm) + 32 * ks) >> 4) & 0x3FFF);
      bd |= (uint64_t)((512 >> 4) & 0x3FFF) << 32;
      bd |= (uint64_t)1 << 46;
      bd |= (uint64_t)4 << 61;
      mma_ss_f16(tb, ad, bd, id, ks != 0);
    }
    commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  fence_after();
  uint32_t a[16];
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  for (int c = 0; c < 64; c += 16) {
    tld16(tb + lane_addr + c, a);
    tld16_wait(a);
#pragma unroll
    for (int j = 0; j < 16; ++j) D[tid * 64 + c + j] = __uint_as_float(a[j]);
  }
  fence_before();
  __syncthreads();
  if (warp == 0) { fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(128) : "memory"); }
}
static void part_d() {
  std::vector<unsigned char> A(16384, 0), B(4096, 0);
  std::vector<float> fa(128 * 32), fb(64 * 32);
  for (auto& v : fa) v = __half2float(__float2half_rn(2.f * frand() - 1.f));
  for (auto& v : fb) v = __half2float(__float2half_rn(2.f * frand() - 1.f));
  for (int r = 0; r < 128; ++r) for (int k = 0; k < 32; ++k) { __half h = __float2half_rn(fa[r * 32 + k]); memcpy(&A[sw128_off(r, k >> 3) + (k & 7) * 2], &h, 2); }
  for (int n = 0; n < 64; ++n) for (int k = 0; k < 32; ++k) { __half h = __float2half_rn(fb[n * 32 + k]); memcpy(&B[sw64_off(n, k >> 3) + (k & 7) * 2], &h, 2); }
  unsigned char *dA, *dB; float* dD;
  CK(cudaMalloc(&dA, 16384)); CK(cudaMalloc(&dB, 4096)); CK(cudaMalloc(&dD, 128 * 64 * 4));
  CK(cudaMemcpy(dA, A.data(), 16384, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), 4096, cudaMemcpyHostToDevice));
  lab_sw64_kernel<<<1, 128>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("{\"part\": \"D\", \"error\": \"%s\"}\n", cudaGetErrorString(e)); exit(1); }
  std::vector<float> D(128 * 64);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double err = 0;
  for (int r = 0; r < 128; ++r) for (int n = 0; n < 64; ++n) {
    double ref = 0;
    for (int k = 0; k < 32; ++k) ref += (double)fa[r * 32 + k] * (double)fb[n * 32 + k];
    err = fmax(err, fabs(ref - D[r * 64 + n]));
  }
  printf("{\"part\": \"D\", \"test\": \"A SWIZZLE_128B x B SWIZZLE_64B (64 keys x 32 halfs), two k-steps\", \"max_abs_diff\": %.3e}\n", err);
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

// =====================================================================================================
// part E: do tcgen05.mma instructions issued CONCURRENTLY by several threads into the SAME accumulator lose updates?
// Three issuing threads (warps 0-2) each add n_mma products of small integers (exact in fp32 in any order) to one
// 32-column accumulator; a fourth configuration uses one thread for the same work.  Repeated `reps` times.
// =====================================================================================================
__global__ void __launch_bounds__(128) lab_shared_acc_kernel(int n_issuers, int n_mma, int reps, int* __restrict__ mismatches, float* __restrict__ last) {
  __shared__ __align__(1024) unsigned char bsm[4096];
  __shared__ uint64_t bar_go, bar_done;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // B [32 rows n x 32 keys] tf32: b(n, k) = (n + k) % 3 (0, 1, 2); A [128 x 32]: a(r, k) = 1 + ((r + k) & 1)
  for (int i = tid; i < 32 * 32; i += 128) {
    const int n = i >> 5, k = i & 31;
    *(float*)(bsm + sw128_off(n, k >> 2) + (k & 3) * 4) = (float)((n + k) % 3);
  }
  if (tid == 0) { mbar_init(smem_u32(&bar_go), 128); mbar_init(smem_u32(&bar_done), n_issuers); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_async_smem();
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tb = tmem_slot;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  uint32_t a[16];
  for (int c = 0; c < 32; c += 16) {
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = __float_as_uint((float)(1 + ((tid + c + j) & 1)));
    tst16(tb + lane_addr + c, a);
  }
  st_wait();
  int bad = 0;
  for (int r = 0; r < reps; ++r) {
    // zero the accumulator (columns 64-95), then release the issuers together
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = 0u;
    tst16(tb + lane_addr + 64, a);
    tst16(tb + lane_addr + 80, a);
    st_wait();
    fence_before();
    __syncthreads();
    fence_after();
    if (lane == 0 && warp < n_issuers) {
      const uint32_t id = idesc(32, 2, 2);
      const int per = n_mma / n_issuers;
      for (int i = 0; i < per; ++i) {
        const int ks = (i + warp) & 3;
        mma_ts_tf32(tb + 64, tb + 8 * ks, desc_sw128(smem_u32(bsm) + 32 * ks), id, 1);
      }
      commit(smem_u32(&bar_done));
    }
    mbar_wait(smem_u32(&bar_done), r & 1);
    fence_after();
    // expected: sum over issued k-steps of sum_{k in step} a(r,k) b(n,k)
    for (int c = 0; c < 32; c += 16) {
      tld16(tb + lane_addr + 64 + c, a);
      tld16_wait(a);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = c + j;
        float ref = 0.f;
        for (int w = 0; w < n_issuers; ++w)
          for (int i = 0; i < n_mma / n_issuers; ++i) {
            const int ks = (i + w) & 3;
            for (int k = 8 * ks; k < 8 * ks + 8; ++k) ref += (float)(1 + ((tid + k) & 1)) * (float)((n + k) % 3);
          }
        if (__uint_as_float(a[j]) != ref) ++bad;
        if (r == reps - 1 && tid == 5 && n == 3) { last[0] = __uint_as_float(a[j]); last[1] = ref; }
      }
    }
    fence_before();
    __syncthreads();
  }
  atomicAdd(mismatches, bad);
  if (warp == 0) { fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(128) : "memory"); }
}
static void part_e() {
  int* dM; float* dL;
  CK(cudaMalloc(&dM, 4)); CK(cudaMalloc(&dL, 8));
  for (int n_issuers = 1; n_issuers <= 3; n_issuers += 2) {
    CK(cudaMemset(dM, 0, 4));
    lab_shared_acc_kernel<<<148, 128>>>(n_issuers, 48, 200, dM, dL);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"part\": \"E\", \"error\": \"%s\"}\n", cudaGetErrorString(e)); exit(1); }
    int m; float l[2];
    CK(cudaMemcpy(&m, dM, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(l, dL, 8, cudaMemcpyDeviceToHost));
    printf("{\"part\": \"E\", \"test\": \"%d thread(s) issue 48 accumulating tf32 MMAs into one accumulator, 200 rounds x 148 CTAs\", \"mismatching_elements\": %d, \"sample_value\": %.1f, \"sample_expected\": %.1f}\n",
           n_issuers, m, l[0], l[1]);
  }
  cudaFree(dM); cudaFree(dL);
}

template <int VARIANT, int KPOLY>
static void run_variant(const char* name, int reps, unsigned warp_mask = 0xffffu) {
  const int grid = 148;
  long long* dC; float* dS;
  CK(cudaMalloc(&dC, grid * 16 * sizeof(long long))); CK(cudaMalloc(&dS, 4));
  for (int it = 0; it < 2; ++it) {
    lab_softmax_kernel<VARIANT, KPOLY><<<grid, 640>>>(dC, dS, reps, 1234u + it, warp_mask);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"part\": \"B\", \"variant\": \"%s\", \"error\": \"%s\"}\n", name, cudaGetErrorString(e)); exit(1); }
  }
  std::vector<long long> c(grid * 16);
  CK(cudaMemcpy(c.data(), dC, c.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  double mean_max = 0, mean_all = 0; int n_all = 0;
  for (int b = 0; b < grid; ++b) {
    long long mx = 0;
    for (int w = 0; w < 16; ++w) { mx = c[b * 16 + w] > mx ? c[b * 16 + w] : mx; if (c[b * 16 + w] > 0) { mean_all += (double)c[b * 16 + w]; ++n_all; } }
    mean_max += (double)mx;
  }
  mean_max /= grid; mean_all /= n_all;
  // a scheduler with four softmax warps runs 2 x (6 + 7) = 26 chunks per head
  printf("{\"part\": \"B\", \"variant\": \"%s\", \"warp_mask\": \"0x%04x\", \"clk_per_head_pass\": %.0f, \"clk_per_chunk_on_a_4_warp_scheduler\": %.1f, \"mean_warp_clk_per_head\": %.0f}\n",
         name, warp_mask, mean_max / (reps * 8.0), mean_max / (reps * 8.0) / 26.0, mean_all / (reps * 8.0));
  cudaFree(dC); cudaFree(dS);
}

int main(int argc, char** argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 200;
  part_a();
  part_d();
  part_e();
  if (argc == 5) return 0;
  if (argc <= 3) part_c();
  if (argc == 3) return 0;
  if (argc > 3) {     // warp-count study: how fast is a softmax warp alone / two per scheduler / four per scheduler?
    const unsigned masks[4] = {0x0001u, 0x0011u, 0x00ffu, 0xffffu};   // one warp; t0 halves A+B of quarter 0; tile 0; both tiles
    for (unsigned m : masks) {
      run_variant<0, 0>("x3 production chunk", reps, m);
      run_variant<5, 0>("x3 chunk, next tcgen05.ld in flight", reps, m);
      run_variant<1, 0>("tf32 in place, all MUFU", reps, m);
      run_variant<4, 0>("tf32 in place, next tcgen05.ld in flight", reps, m);
      run_variant<1, 4>("tf32 in place, 4/16 polynomial", reps, m);
    }
    return 0;
  }
  run_variant<0, 0>("x3 production chunk", reps);
  run_variant<1, 0>("tf32 in place, all MUFU", reps);
  run_variant<1, 4>("tf32 in place, 4/16 polynomial", reps);
  run_variant<1, 6>("tf32 in place, 6/16 polynomial", reps);
  run_variant<1, 8>("tf32 in place, 8/16 polynomial", reps);
  run_variant<1, 16>("tf32 in place, all polynomial", reps);
  run_variant<3, 0>("tf32 in place, mask after exp, all MUFU", reps);
  run_variant<2, 0>("bf16 pair in place, all MUFU", reps);
  run_variant<2, 4>("bf16 pair in place, 4/16 polynomial", reps);
  run_variant<2, 6>("bf16 pair in place, 6/16 polynomial", reps);
  return 0;
}
