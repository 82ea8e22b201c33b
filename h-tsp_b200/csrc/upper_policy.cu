// K8: upper-level policy forward: HTSP_PPO.forward (h_tsp.py:1272-1276)
// = ActorPPO.forward (rl_models.py:329-336) o IMPALAEncoder.forward (rl_models.py:187-273)
// o IMPALACNN.forward (rl_models.py:117-133), state [E,Ntot,8] -> action [E,2], all fp32.
//
//   pillar stage   every city falls into one of 100 x 100 pixels ("pillars") of its instance; its 12
//                  features are [x, y, offset to the pixel centre, offset to the pillar mean, the 6 extra
//                  state columns]; 1x1 Conv1d (12 -> C0, no bias), InstanceNorm1d over the cities of the
//                  instance, ReLU, per-pillar max -> pseudo image [E,C0,100,100] (zeros elsewhere).
//                  The reference builds pillars with torch.unique + pytorch-scatter; here the pixel id
//                  addresses them directly: pillar means from order-independent fixed-point atomics
//                  (deterministic), InstanceNorm statistics from the 12-vector first and 12x12 second
//                  moments of the features (y = W f is linear, so mean and variance of every channel
//                  follow from 90 numbers per instance, accumulated in fp64 in a fixed order), per-pillar
//                  max by integer atomicMax on the non-negative floats (order independent).
//   CNN            3 x [conv3x3, maxpool(3,2,1), 2 residual blocks (ReLU, conv, ReLU, conv, + input)]
//                  as shared-memory tiled direct convolutions (bias / input ReLU / residual fused),
//                  then ReLU(flatten) -> fc -> ReLU and the 4-layer actor MLP (ReLU, ReLU, Hardswish,
//                  tanh) in one kernel per instance.
// 325 MFLOP per instance; at E = 16 environments this is ~2 % of an environment step, so the kernels are
// plain fp32 CUDA-core code (same arithmetic type as the reference) rather than tensor-core GEMMs.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"

namespace htsp {

constexpr int PG = 100;                        // bev grid (rl_models.py:155-158)
constexpr int PS = PG * PG;
constexpr int PIN = 12;                        // input_dim + 4 (rl_models.py:171)
constexpr int PMOM = PIN + PIN * (PIN + 1) / 2;   // 12 first + 78 second moments
constexpr int PFLAT = 32 * 13 * 13;            // flatten dim of the default depths [16,32,32]
constexpr int PCONVS = 15;
constexpr float kPixel = 0.01f;
constexpr double kFix = 1099511627776.0;       // 2^40: fixed-point scale of the pillar sums
constexpr double kInEps = 1e-3;                // InstanceNorm1d eps (rl_models.py:178-180)

struct PolGeom {
  int E, mid, adim, C0;
  int cin[PCONVS], cout[PCONVS];
};
static PolGeom pol_geom(int embed, int mid, int adim) {
  PolGeom g;
  g.E = embed; g.mid = mid; g.adim = adim; g.C0 = embed / 2;
  const int depth[3] = {16, 32, 32};
  int cin = g.C0;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 5; ++j) { g.cin[5 * i + j] = (j == 0) ? cin : depth[i]; g.cout[5 * i + j] = depth[i]; }
    cin = depth[i];
  }
  return g;
}
struct PolOff {
  size_t w_in, gamma, beta, conv_w[PCONVS], conv_b[PCONVS], fc_w, fc_b, a_w[4], a_b[4], end;
};
static PolOff pol_off(const PolGeom& g) {
  PolOff o;
  size_t p = 0;
  auto take = [&](size_t n) { const size_t r = p; p += (n + 3) / 4 * 4; return r; };   // 16-byte aligned sections
  o.w_in = take((size_t)PIN * g.C0);       // [12][C0]
  o.gamma = take(g.C0);
  o.beta = take(g.C0);
  for (int j = 0; j < PCONVS; ++j) { o.conv_w[j] = take((size_t)g.cin[j] * 9 * g.cout[j]); o.conv_b[j] = take(g.cout[j]); }
  o.fc_w = take((size_t)g.E * PFLAT);      // [E][5408]
  o.fc_b = take(g.E);
  const int ain[4] = {g.E, g.mid, g.mid, g.mid}, aout[4] = {g.mid, g.mid, g.mid, g.adim};
  for (int l = 0; l < 4; ++l) { o.a_w[l] = take((size_t)ain[l] * aout[l]); o.a_b[l] = take(aout[l]); }   // [in][out]
  o.end = p;
  return o;
}
static bool pol_dims_ok(int embed, int mid, int adim) {
  return embed >= 16 && embed <= 256 && embed % 16 == 0 && mid >= 1 && mid <= 256 && adim >= 1 && adim <= 16;
}

// ---- workspace ----------------------------------------------------------------------------------------
struct PolWs {
  size_t sums, cnt, img, zero_bytes, pix, part, coef, t0, xa, xb, xc, end;
  int nblk;
};
static PolWs pol_ws(int B, int N, int C0) {
  PolWs w;
  size_t p = 0;
  auto take = [&](size_t bytes) { const size_t r = p; p += align_up(bytes, 256); return r; };
  w.sums = take((size_t)2 * B * PS * 8);
  w.cnt = take((size_t)B * PS * 4);
  w.img = take((size_t)B * C0 * PS * 4);
  w.zero_bytes = p;                                   // [sums | cnt | img] are cleared by one memset per call
  w.pix = take((size_t)B * N * 4);
  w.nblk = (N + 255) / 256;
  w.part = take((size_t)B * w.nblk * PMOM * 8);
  w.coef = take((size_t)B * 2 * C0 * 4);
  w.t0 = take((size_t)B * 16 * PS * 4);               // conv output before the pool: max(16*100^2, 32*50^2, 32*25^2)
  w.xa = take((size_t)B * 16 * 2500 * 4);             // activations after the pool: max(16*50^2, 32*25^2, 32*13^2)
  w.xb = take((size_t)B * 16 * 2500 * 4);
  w.xc = take((size_t)B * 16 * 2500 * 4);
  w.end = p;
  return w;
}

// ---- pillar stage -------------------------------------------------------------------------------------
__device__ __forceinline__ void pol_pixel_of(float x, float y, int& cx, int& cy) {
  // floor((xy - bev_range[0:2]) / bev_pixel_size).int()  (rl_models.py:196-197); range origin is 0
  cx = (int)floorf(__fdiv_rn(x, kPixel));
  cy = (int)floorf(__fdiv_rn(y, kPixel));
}

// pix[i] = b*10^4 + cx*100 + cy (the reference's bev_merge_coords, rl_models.py:207-211), -1 when it leaves
// [0, B*10^4); fixed-point sums of x, y and the count of every pillar
__global__ void __launch_bounds__(256)
pol_pixel_kernel(const float* __restrict__ state, int B, int N, long long* __restrict__ sums,
                 int* __restrict__ cnt, int* __restrict__ pix) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * N) return;
  const int b = (int)(i / N);
  const float2 xy = *(const float2*)(state + i * 8);
  int cx, cy;
  pol_pixel_of(xy.x, xy.y, cx, cy);
  long long u = (long long)b * PS + (long long)cx * PG + cy;
  if (u < 0 || u >= (long long)B * PS) u = -1;
  pix[i] = (int)u;
  if (u >= 0) {
    atomicAdd((unsigned long long*)&sums[u], (unsigned long long)__double2ll_rn((double)xy.x * kFix));
    atomicAdd((unsigned long long*)&sums[(size_t)B * PS + u], (unsigned long long)__double2ll_rn((double)xy.y * kFix));
    atomicAdd(&cnt[u], 1);
  }
}

// mvf_input row of one city (rl_models.py:213-222): [xy, xy - pixel centre, xy - pillar mean, extra(6)]
__device__ __forceinline__ void pol_features(const float* __restrict__ st, int u, const long long* __restrict__ sums,
                                             const int* __restrict__ cnt, size_t BS, float (&f)[PIN]) {
  const float4 s0 = *(const float4*)st, s1 = *(const float4*)(st + 4);
  const float x = s0.x, y = s0.y;
  int cx, cy;
  pol_pixel_of(x, y, cx, cy);
  f[0] = x;
  f[1] = y;
  f[2] = __fsub_rn(x, __fmul_rn(__fadd_rn((float)cx, 0.5f), kPixel));
  f[3] = __fsub_rn(y, __fmul_rn(__fadd_rn((float)cy, 0.5f), kPixel));
  float mx = x, my = y;                         // a dropped pillar (u < 0) keeps a zero cluster offset
  if (u >= 0) {
    const double inv = 1.0 / ((double)cnt[u] * kFix);
    mx = (float)((double)sums[u] * inv);
    my = (float)((double)sums[BS + u] * inv);
  }
  f[4] = __fsub_rn(x, mx);
  f[5] = __fsub_rn(y, my);
  f[6] = s0.z; f[7] = s0.w; f[8] = s1.x; f[9] = s1.y; f[10] = s1.z; f[11] = s1.w;
}

// first and second moments of the feature rows of one instance, one partial per 256 cities
__global__ void __launch_bounds__(256)
pol_moments_kernel(const float* __restrict__ state, const int* __restrict__ pix, const long long* __restrict__ sums,
                   const int* __restrict__ cnt, int B, int N, double* __restrict__ part) {
  __shared__ double red[8][PMOM];
  const int b = blockIdx.y, n = blockIdx.x * 256 + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float f[PIN];
#pragma unroll
  for (int k = 0; k < PIN; ++k) f[k] = 0.f;
  if (n < N) pol_features(state + ((size_t)b * N + n) * 8, pix[(size_t)b * N + n], sums, cnt, (size_t)B * PS, f);
  int m = 0;
#pragma unroll
  for (int i = 0; i < PIN; ++i) {
#pragma unroll
    for (int j = i - 1; j < PIN; ++j) {         // j == i-1 stands for the first moment of f[i]
      double v = (j < i) ? (double)f[i] : (double)f[i] * (double)f[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][m] = v;
      ++m;
    }
  }
  __syncthreads();
  if (threadIdx.x < PMOM) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    part[((size_t)b * gridDim.x + blockIdx.x) * PMOM + threadIdx.x] = v;
  }
}

// InstanceNorm1d as a per-(instance, channel) scale and shift: z = y * a + s
__global__ void __launch_bounds__(128)
pol_coef_kernel(const double* __restrict__ part, int nblk, int N, const float* __restrict__ w_in,
                const float* __restrict__ gamma, const float* __restrict__ beta, int C0, float* __restrict__ coef) {
  __shared__ double mom[PMOM];
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < PMOM; k += blockDim.x) {
    double v = 0.0;
    for (int i = 0; i < nblk; ++i) v += part[((size_t)b * nblk + i) * PMOM + k];
    mom[k] = v / (double)N;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C0; c += blockDim.x) {
    double w[PIN];
#pragma unroll
    for (int k = 0; k < PIN; ++k) w[k] = (double)w_in[k * C0 + c];
    double mean = 0.0, ey2 = 0.0;
    int m = 0;
#pragma unroll
    for (int i = 0; i < PIN; ++i) {
#pragma unroll
      for (int j = i - 1; j < PIN; ++j) {
        if (j < i) mean += w[i] * mom[m];
        else ey2 += ((j == i) ? 1.0 : 2.0) * w[i] * w[j] * mom[m];
        ++m;
      }
    }
    const double var = fmax(ey2 - mean * mean, 0.0);          // biased variance (InstanceNorm)
    const double a = (double)gamma[c] / sqrt(var + kInEps);
    coef[((size_t)b * 2 + 0) * C0 + c] = (float)a;
    coef[((size_t)b * 2 + 1) * C0 + c] = (float)((double)beta[c] - mean * a);
  }
}

// 1x1 conv + norm + ReLU per city, per-pillar max into the image; block = C0 channels x 4 cities, 32
// cities per block
__global__ void __launch_bounds__(512)
pol_scatter_kernel(const float* __restrict__ state, const int* __restrict__ pix, const long long* __restrict__ sums,
                   const int* __restrict__ cnt, int B, int N, const float* __restrict__ w_in,
                   const float* __restrict__ coef, int C0, float* __restrict__ img) {
  __shared__ float sf[32][PIN];
  __shared__ int su[32];
  const int b = blockIdx.y, n0 = blockIdx.x * 32;
  const int c = threadIdx.x, py = threadIdx.y, t = py * C0 + c;
  if (t < 32) {
    float f[PIN];
    int u = -1;
    if (n0 + t < N) {
      u = pix[(size_t)b * N + n0 + t];
      pol_features(state + ((size_t)b * N + n0 + t) * 8, u, sums, cnt, (size_t)B * PS, f);
#pragma unroll
      for (int k = 0; k < PIN; ++k) sf[t][k] = f[k];
    }
    su[t] = u;
  }
  float w[PIN];
#pragma unroll
  for (int k = 0; k < PIN; ++k) w[k] = w_in[k * C0 + c];
  const float a = coef[((size_t)b * 2 + 0) * C0 + c], s = coef[((size_t)b * 2 + 1) * C0 + c];
  __syncthreads();
  for (int p = py; p < 32; p += blockDim.y) {
    const int u = su[p];
    if (u < 0) continue;
    float y = 0.f;
#pragma unroll
    for (int k = 0; k < PIN; ++k) y = fmaf(sf[p][k], w[k], y);
    const float z = fmaf(y, a, s);
    if (z > 0.f) {
      const int ub = u / PS, rem = u - ub * PS;      // a coordinate outside [0,1) aliases into another pillar, as in the reference
      atomicMax((int*)&img[((size_t)ub * C0 + c) * PS + rem], __float_as_int(z));
    }
  }
}

// ---- CNN ----------------------------------------------------------------------------------------------
// direct 3x3 convolution, padding 1.  One CTA = TYN whole rows of output pixels x CT output channels; a
// thread owns PW horizontally adjacent pixels x CT channels.  Shared memory delivers one weight per clock
// and SM (a broadcast LDS.128 occupies the pipe for 4 clocks), so a weight vector must feed >= 32 FMAs
// per lane to keep the FMA pipes busy: PW = 10 / 5 / 5 / 7 pixels for 100 / 50 / 25 / 13-pixel rows with
// CT = 4 channels (40 / 20 / 20 / 28 accumulators per thread).  Rows are whole, so the left/right halo columns
// and the rows outside the image are always zero: shared memory is cleared once and only the interior is
// copied, one VEC*4-byte cp.async per lane and row (16 bytes for the 100-pixel rows); every copy of a
// chunk of CK input channels is in flight before the first wait.  Row layout in shared memory:
// [3 pad | 0 | Ww pixels | 0 | pad], the interior 16-byte aligned.  Weights are pre-arranged [Cin][9][Cout].
template <int PW, int CT, int VEC>
__global__ void __launch_bounds__(256, 2)
pol_conv3x3_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                   const float* __restrict__ res, float* __restrict__ out, int Cin, int Cout, int Hh, int Ww,
                   int relu_in, int TXN, int TYN, int CK) {
  extern __shared__ __align__(16) float pol_smem[];
  const int SW = (Ww + 8 + PW + 3) & ~3, SH = TYN + 2;                // smem row pitch: multiple of 4 floats
  float* s_in = pol_smem;                                             // [CK][SH][SW]
  float* s_w = pol_smem + (size_t)CK * SH * SW;                       // [CK][9][CT]
  const int ty0 = blockIdx.x * TYN;
  const int co0 = blockIdx.y * CT, b = blockIdx.z;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int tx = tid % TXN, ty = tid / TXN;
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;
  float acc[PW][CT];
#pragma unroll
  for (int j = 0; j < CT; ++j) {
    const float bj = bias[co0 + j];
#pragma unroll
    for (int px = 0; px < PW; ++px) acc[px][j] = bj;
  }
  const int per_ch = SH * SW;
  for (int i = tid; i < CK * per_ch / 4; i += nthr) ((float4*)s_in)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* inb = in + (size_t)b * Cin * Hh * Ww;
  const uint32_t s_in_addr = (uint32_t)__cvta_generic_to_shared(s_in);
  const int nvec = Ww / VEC;                                           // copy lanes per row (<= 32)
  for (int c0 = 0; c0 < Cin; c0 += CK) {
    __syncthreads();
    if (lane < nvec) {
      for (int ci = warp; ci < CK; ci += nwarps) {                 // one warp per input channel of the chunk
        const float* src = inb + ((ptrdiff_t)(c0 + ci) * Hh + (ty0 - 1)) * Ww + lane * VEC;   // row yy = 0
        const uint32_t dst = s_in_addr + 4u * (uint32_t)(ci * per_ch + 4 + lane * VEC);
#pragma unroll 4
        for (int yy = 0; yy < SH; ++yy) {
          const int gy = ty0 + yy - 1;
          if (gy >= 0 && gy < Hh) {
            const float* sp = src + (ptrdiff_t)yy * Ww;
            const uint32_t dp = dst + 4u * (uint32_t)(yy * SW);
            if (VEC == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dp), "l"(sp) : "memory");
            else if (VEC == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dp), "l"(sp) : "memory");
            else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dp), "l"(sp) : "memory");
          }
        }
      }
    }
    for (int i = tid; i < CK * 9 * (CT / 4); i += nthr) {        // 16-byte weight copies: [ci][tap][CT]
      const int q = i % (CT / 4), ct = i / (CT / 4);             // ct = ci * 9 + tap
      const float* src = w + ((size_t)c0 * 9 + ct) * Cout + co0 + 4 * q;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(s_w + 4 * i)),
                   "l"(src)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (ty < TYN) {
#pragma unroll 1
      for (int ci = 0; ci < CK; ++ci) {
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const float* row = s_in + (size_t)ci * per_ch + (ty + dy) * SW + 3 + tx * PW;   // pixel tx*PW - 1 of the row
          float seg[PW + 2];
#pragma unroll
          for (int k = 0; k < PW + 2; ++k) seg[k] = relu_in ? fmaxf(row[k], 0.f) : row[k];
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const float4* wp = (const float4*)(s_w + ((size_t)ci * 9 + dy * 3 + dx) * CT);
#pragma unroll
            for (int j4 = 0; j4 < CT / 4; ++j4) {
              const float4 ww = wp[j4];
#pragma unroll
              for (int px = 0; px < PW; ++px) {
                const float v = seg[px + dx];
                acc[px][4 * j4 + 0] = fmaf(v, ww.x, acc[px][4 * j4 + 0]);
                acc[px][4 * j4 + 1] = fmaf(v, ww.y, acc[px][4 * j4 + 1]);
                acc[px][4 * j4 + 2] = fmaf(v, ww.z, acc[px][4 * j4 + 2]);
                acc[px][4 * j4 + 3] = fmaf(v, ww.w, acc[px][4 * j4 + 3]);
              }
            }
          }
        }
      }
    }
  }
  const int gy = ty0 + ty;
  if (ty < TYN && gy < Hh) {
#pragma unroll
    for (int j = 0; j < CT; ++j) {
      const size_t o = (((size_t)b * Cout + co0 + j) * Hh + gy) * Ww + tx * PW;
#pragma unroll
      for (int px = 0; px < PW; ++px)
        if (tx * PW + px < Ww) out[o + px] = res ? acc[px][j] + res[o + px] : acc[px][j];
    }
  }
}

// MaxPool2d(kernel 3, stride 2, padding 1)
__global__ void __launch_bounds__(256)
pol_maxpool_kernel(const float* __restrict__ in, float* __restrict__ out, size_t planes, int Hh, int Ww, int Ho, int Wo) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= planes * Ho * Wo) return;
  const int ox = (int)(i % Wo), oy = (int)((i / Wo) % Ho);
  const size_t pl = i / ((size_t)Wo * Ho);
  const float* p = in + pl * Hh * Ww;
  float m = -INFINITY;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int y = 2 * oy + dy;
    if (y < 0 || y >= Hh) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int x = 2 * ox + dx;
      if (x >= 0 && x < Ww) m = fmaxf(m, p[(size_t)y * Ww + x]);
    }
  }
  out[i] = m;
}

// ReLU(flatten) -> fc -> ReLU = graph_feat (rl_models.py:130-131): one warp per (instance, output), the
// 5408-long dot product with eight 16-byte weight loads in flight per lane; grid (E/8, instances)
__global__ void __launch_bounds__(256)
pol_fc_kernel(const float* __restrict__ x3, const float* __restrict__ fc_w, const float* __restrict__ fc_b, int E,
              float* __restrict__ feat) {
  const int b = blockIdx.y, lane = threadIdx.x & 31, o = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (o >= E) return;
  const float4* wr = (const float4*)(fc_w + (size_t)o * PFLAT);
  const float4* xr = (const float4*)(x3 + (size_t)b * PFLAT);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  constexpr int NV = PFLAT / 4;                       // 1352 float4 per row
  for (int k0 = 0; k0 < NV; k0 += 256) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = k0 + u * 32 + lane;
      if (k < NV) {
        const float4 ww = __ldg(wr + k), xx = __ldg(xr + k);
        acc[u] = fmaf(fmaxf(xx.x, 0.f), ww.x, acc[u]);
        acc[u] = fmaf(fmaxf(xx.y, 0.f), ww.y, acc[u]);
        acc[u] = fmaf(fmaxf(xx.z, 0.f), ww.z, acc[u]);
        acc[u] = fmaf(fmaxf(xx.w, 0.f), ww.w, acc[u]);
      }
    }
  }
  float a = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
  if (lane == 0) feat[(size_t)b * E + o] = fmaxf(a + fc_b[o], 0.f);
}

// ActorPPO.net and tanh (rl_models.py:287-301, 329-336); one CTA per instance, weights [in][out]
__global__ void __launch_bounds__(256)
pol_actor_kernel(const float* __restrict__ feat, const float* __restrict__ w1, const float* __restrict__ b1,
                 const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ w3,
                 const float* __restrict__ b3, const float* __restrict__ w4, const float* __restrict__ b4, int E,
                 int mid, int adim, float* __restrict__ action) {
  __shared__ float h0[256], h1[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < E) h0[tid] = feat[(size_t)b * E + tid];
  __syncthreads();
  auto dense = [&](const float* hin, int nin, const float* wt, const float* bb, int nout, int j) {
    float acc = bb[j];
#pragma unroll 16
    for (int i = 0; i < nin; ++i) acc = fmaf(hin[i], __ldg(wt + (size_t)i * nout + j), acc);
    return acc;
  };
  if (tid < mid) h1[tid] = fmaxf(dense(h0, E, w1, b1, mid, tid), 0.f);
  __syncthreads();
  if (tid < mid) h0[tid] = fmaxf(dense(h1, mid, w2, b2, mid, tid), 0.f);
  __syncthreads();
  if (tid < mid) {
    const float v = dense(h0, mid, w3, b3, mid, tid);
    h1[tid] = v * fminf(fmaxf(v + 3.f, 0.f), 6.f) / 6.f;          // Hardswish
  }
  __syncthreads();
  if (tid < adim) action[(size_t)b * adim + tid] = tanhf(dense(h1, mid, w4, b4, adim, tid));
}

}  // namespace htsp

using namespace htsp;

extern "C" int htsp_policy_n_tensors(void) { return 3 + 2 * PCONVS + 2 + 8; }

extern "C" size_t htsp_policy_blob_bytes(int embed, int mid, int adim) {
  if (!pol_dims_ok(embed, mid, adim)) return 0;
  return pol_off(pol_geom(embed, mid, adim)).end * sizeof(float);
}

extern "C" int htsp_policy_pack_weights(const float* const* t, int n_tensors, int embed, int mid, int adim,
                                        void* dev_blob, void* stream) {
  HTSP_REQUIRE(t && dev_blob, "null pointer");
  HTSP_REQUIRE(pol_dims_ok(embed, mid, adim), "policy dimensions");
  HTSP_REQUIRE(n_tensors == htsp_policy_n_tensors(), "tensor count");
  for (int i = 0; i < n_tensors; ++i) HTSP_REQUIRE(t[i] != nullptr, "null tensor");
  const PolGeom g = pol_geom(embed, mid, adim);
  const PolOff o = pol_off(g);
  std::vector<float> blob(o.end, 0.f);
  float* p = blob.data();
  int k = 0;
  const float* w_in = t[k++];                                   // [C0,12,1] -> [12][C0]
  for (int c = 0; c < g.C0; ++c)
    for (int i = 0; i < PIN; ++i) p[o.w_in + (size_t)i * g.C0 + c] = w_in[(size_t)c * PIN + i];
  memcpy(p + o.gamma, t[k++], sizeof(float) * g.C0);
  memcpy(p + o.beta, t[k++], sizeof(float) * g.C0);
  for (int j = 0; j < PCONVS; ++j) {                            // [Cout,Cin,3,3] -> [Cin][9][Cout]
    const float* w = t[k++];
    const int ci_n = g.cin[j], co_n = g.cout[j];
    for (int co = 0; co < co_n; ++co)
      for (int ci = 0; ci < ci_n; ++ci)
        for (int tap = 0; tap < 9; ++tap)
          p[o.conv_w[j] + ((size_t)ci * 9 + tap) * co_n + co] = w[((size_t)co * ci_n + ci) * 9 + tap];
    memcpy(p + o.conv_b[j], t[k++], sizeof(float) * co_n);
  }
  memcpy(p + o.fc_w, t[k++], sizeof(float) * (size_t)g.E * PFLAT);
  memcpy(p + o.fc_b, t[k++], sizeof(float) * g.E);
  const int ain[4] = {g.E, g.mid, g.mid, g.mid}, aout[4] = {g.mid, g.mid, g.mid, g.adim};
  for (int l = 0; l < 4; ++l) {                                 // [out,in] -> [in][out]
    const float* w = t[k++];
    for (int oo = 0; oo < aout[l]; ++oo)
      for (int i = 0; i < ain[l]; ++i) p[o.a_w[l] + (size_t)i * aout[l] + oo] = w[(size_t)oo * ain[l] + i];
    memcpy(p + o.a_b[l], t[k++], sizeof(float) * aout[l]);
  }
  HTSP_CHECK_CUDA(cudaMemcpyAsync(dev_blob, p, o.end * sizeof(float), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  HTSP_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));   // the staging vector dies at return
  return HTSP_OK;
}

extern "C" size_t htsp_policy_workspace_bytes(int E_envs, int Ntot, int embed) {
  if (E_envs <= 0 || Ntot <= 0 || embed < 16) return 0;
  return pol_ws(E_envs, Ntot, embed / 2).end + 256;
}

extern "C" int htsp_policy_forward(const void* blob, int embed, int mid, int adim, const float* state, int B,
                                   int N, float* feat, float* action, float* image_out, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  HTSP_REQUIRE(blob && state && feat && action && workspace, "null pointer");
  HTSP_REQUIRE(pol_dims_ok(embed, mid, adim), "policy dimensions");
  HTSP_REQUIRE(B >= 1 && B <= 65535 && N >= 1, "batch / instance size");
  HTSP_REQUIRE((long long)B * PS < (1ll << 24), "pillar ids must stay exact in fp32 (rl_models.py:207-211)");
  HTSP_REQUIRE(workspace_bytes >= htsp_policy_workspace_bytes(B, N, embed), "workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const PolGeom g = pol_geom(embed, mid, adim);
  const PolOff o = pol_off(g);
  const PolWs w = pol_ws(B, N, g.C0);
  const float* wb = (const float*)blob;
  char* base = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  long long* sums = (long long*)(base + w.sums);
  int* cnt = (int*)(base + w.cnt);
  float* img = (float*)(base + w.img);
  int* pix = (int*)(base + w.pix);
  double* part = (double*)(base + w.part);
  float* coef = (float*)(base + w.coef);
  float* t0 = (float*)(base + w.t0);
  float* xa = (float*)(base + w.xa);
  float* xb = (float*)(base + w.xb);
  float* xc = (float*)(base + w.xc);

  HTSP_CHECK_CUDA(cudaMemsetAsync(base, 0, w.zero_bytes, s));
  const size_t pts = (size_t)B * N;
  pol_pixel_kernel<<<(unsigned)((pts + 255) / 256), 256, 0, s>>>(state, B, N, sums, cnt, pix);
  HTSP_CHECK_LAUNCH();
  pol_moments_kernel<<<dim3(w.nblk, B), 256, 0, s>>>(state, pix, sums, cnt, B, N, part);
  HTSP_CHECK_LAUNCH();
  pol_coef_kernel<<<B, 128, 0, s>>>(part, w.nblk, N, wb + o.w_in, wb + o.gamma, wb + o.beta, g.C0, coef);
  HTSP_CHECK_LAUNCH();
  {
    const int py = g.C0 >= 128 ? 4 : (512 / g.C0 < 32 ? 512 / g.C0 : 32);
    HTSP_REQUIRE(g.C0 * py >= 32, "block too small for the feature stage");
    pol_scatter_kernel<<<dim3((N + 31) / 32, B), dim3(g.C0, py), 0, s>>>(state, pix, sums, cnt, B, N, wb + o.w_in,
                                                                         coef, g.C0, img);
    HTSP_CHECK_LAUNCH();
  }
  if (image_out)
    HTSP_CHECK_CUDA(cudaMemcpyAsync(image_out, img, (size_t)B * g.C0 * PS * 4, cudaMemcpyDeviceToDevice, s));

  // per layer size: pixels per thread PW, copy vector VEC, rows per CTA TYN, input channels per chunk CK
  auto conv = [&](int j, const float* in, const float* res, float* out, int hw, int relu_in) -> int {
    const int pw = hw == 100 ? 10 : (hw == 13 ? 7 : 5);
    HTSP_REQUIRE(hw == 100 || hw == 50 || hw == 25 || hw == 13, "image size of the default IMPALA geometry");
    const int txn = (hw + pw - 1) / pw;
    int tyn = 256 / txn;
    if (tyn > 25) tyn = 25;
    if (tyn > hw) tyn = hw;
    const int cin = g.cin[j], ck = hw == 100 ? 8 : (hw == 50 ? 16 : cin);
    HTSP_REQUIRE(cin % ck == 0, "input channels per chunk");
    const int sw = (hw + 8 + pw + 3) & ~3, ct = 4;
    const size_t smem = ((size_t)ck * (tyn + 2) * sw + (size_t)ck * 9 * ct) * sizeof(float);
    const int tiles = (hw + tyn - 1) / tyn;
    const dim3 grid(tiles, g.cout[j] / ct, B);
    int threads = (txn * tyn + 31) / 32 * 32;
    if (threads < 128) threads = 128;              // the copy phase wants a few warps
#define POL_LAUNCH_CONV(PWV, VECV)                                                                                   \
    do {                                                                                                             \
      HTSP_CHECK_CUDA(cudaFuncSetAttribute(pol_conv3x3_kernel<PWV, 4, VECV>,                                          \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
      pol_conv3x3_kernel<PWV, 4, VECV><<<grid, threads, smem, s>>>(in, wb + o.conv_w[j], wb + o.conv_b[j], res, out,  \
                                                                  cin, g.cout[j], hw, hw, relu_in, txn, tyn, ck);   \
    } while (0)
    if (hw == 100) POL_LAUNCH_CONV(10, 4);
    else if (hw == 50) POL_LAUNCH_CONV(5, 2);
    else if (hw == 25) POL_LAUNCH_CONV(5, 1);
    else POL_LAUNCH_CONV(7, 1);
#undef POL_LAUNCH_CONV
    HTSP_CHECK_LAUNCH();
    return HTSP_OK;
  };
  const float* x = img;
  int hw = PG;
  for (int i = 0; i < 3; ++i) {
    int rc = conv(5 * i, x, nullptr, t0, hw, 0);
    if (rc != HTSP_OK) return rc;
    const int ho = (hw - 1) / 2 + 1;
    const size_t planes = (size_t)B * g.cout[5 * i];
    pol_maxpool_kernel<<<(unsigned)((planes * ho * ho + 255) / 256), 256, 0, s>>>(t0, xa, planes, hw, hw, ho, ho);
    HTSP_CHECK_LAUNCH();
    hw = ho;
    // residual blocks: xa -> (xb) -> xc = xa + conv(relu(conv(relu(xa)))); then xc -> (xb) -> xa
    if ((rc = conv(5 * i + 1, xa, nullptr, xb, hw, 1)) != HTSP_OK) return rc;
    if ((rc = conv(5 * i + 2, xb, xa, xc, hw, 1)) != HTSP_OK) return rc;
    if ((rc = conv(5 * i + 3, xc, nullptr, xb, hw, 1)) != HTSP_OK) return rc;
    if ((rc = conv(5 * i + 4, xb, xc, xa, hw, 1)) != HTSP_OK) return rc;
    x = xa;
  }
  HTSP_REQUIRE(hw * hw * 32 == PFLAT, "flatten dimension");
  pol_fc_kernel<<<dim3((g.E + 7) / 8, B), 256, 0, s>>>(xa, wb + o.fc_w, wb + o.fc_b, g.E, feat);
  HTSP_CHECK_LAUNCH();
  pol_actor_kernel<<<B, 256, 0, s>>>(feat, wb + o.a_w[0], wb + o.a_b[0], wb + o.a_w[1], wb + o.a_b[1], wb + o.a_w[2],
                                     wb + o.a_b[2], wb + o.a_w[3], wb + o.a_b[3], g.E, g.mid, g.adim, action);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}
