// K1: subproblem gather + min/max normalisation (+ 8-fold augmentation), HBM-bound.
// Replaces h_tsp.py:176-191 (torch.gather, 4 min/max reductions, maximum, 2 strided writes)
// and rl4cop/utils/utils.py:37-56 (9 torch.cat) with one launch.
// Algorithmic bytes per subproblem: N*(8 idx + 8 gathered coords + 8 x_new) + 4 (scale)
// [+ n_aug*N*8 when the augmented copies are materialised].
#include "common.cuh"

namespace htsp {

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// The 8 isometries of the unit square in the reference's block order (utils.py:44-51).
__device__ __forceinline__ float2 isometry(int k, float x, float y) {
  switch (k) {
    case 0: return make_float2(x, y);
    case 1: return make_float2(1.f - x, y);
    case 2: return make_float2(x, 1.f - y);
    case 3: return make_float2(1.f - x, 1.f - y);
    case 4: return make_float2(y, x);
    case 5: return make_float2(1.f - y, x);
    case 6: return make_float2(y, 1.f - x);
    default: return make_float2(1.f - y, 1.f - x);
  }
}

// One CTA per subproblem; each thread keeps its gathered coordinates in registers between the
// reduction and the write so HBM is touched once per element.
constexpr int kExtractThreads = 256;
constexpr int kExtractPerThread = 4;  // supports N <= 1024

__global__ void __launch_bounds__(kExtractThreads)
extract_normalize_kernel(const float2* __restrict__ data, const long long* __restrict__ fragment,
                         int B, int Ntot, int N, int n_aug, float2* __restrict__ x_new,
                         float* __restrict__ scale, float2* __restrict__ x_aug) {
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  __shared__ float red[4][kExtractThreads / 32];
  float2 v[kExtractPerThread];
  float mnx = INFINITY, mny = INFINITY, mxx = -INFINITY, mxy = -INFINITY;
  bool bad = false;
#pragma unroll
  for (int i = 0; i < kExtractPerThread; ++i) {
    int j = tid + i * kExtractThreads;
    if (j < N) {
      const long long idx = __ldg(&fragment[(size_t)b * N + j]);
      bad = bad || idx < 0 || idx >= Ntot;       // the reference's torch.gather raises; here the subproblem returns NaN
      v[i] = __ldg(&data[(size_t)b * Ntot + clamp_index(idx, Ntot)]);
      mnx = fminf(mnx, v[i].x); mxx = fmaxf(mxx, v[i].x);
      mny = fminf(mny, v[i].y); mxy = fmaxf(mxy, v[i].y);
    }
  }
  mnx = warp_min(mnx); mny = warp_min(mny); mxx = warp_max(mxx); mxy = warp_max(mxy);
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = mnx; red[1][tid >> 5] = mny; red[2][tid >> 5] = mxx; red[3][tid >> 5] = mxy;
  }
  __syncthreads();
  mnx = red[0][0]; mny = red[1][0]; mxx = red[2][0]; mxy = red[3][0];
#pragma unroll
  for (int w = 1; w < kExtractThreads / 32; ++w) {
    mnx = fminf(mnx, red[0][w]); mny = fminf(mny, red[1][w]);
    mxx = fmaxf(mxx, red[2][w]); mxy = fmaxf(mxy, red[3][w]);
  }
  // s = 0.9 / max(dx, dy): torch evaluates `scalar / tensor` as reciprocal(tensor) * scalar (two
  // roundings); x' = s*(x - min) + 0.05 as separate ops, no FMA contraction (h_tsp.py:188-191)
  float s = __fmul_rn(__frcp_rn(fmaxf(__fsub_rn(mxx, mnx), __fsub_rn(mxy, mny))), 0.9f);
  if (__syncthreads_or(bad)) s = __int_as_float(0x7fc00000);    // out-of-range fragment ids: scale, coordinates and length are NaN
  if (tid == 0) scale[b] = s;
#pragma unroll
  for (int i = 0; i < kExtractPerThread; ++i) {
    int j = tid + i * kExtractThreads;
    if (j < N) {
      float nx = __fadd_rn(__fmul_rn(s, __fsub_rn(v[i].x, mnx)), 0.05f);
      float ny = __fadd_rn(__fmul_rn(s, __fsub_rn(v[i].y, mny)), 0.05f);
      x_new[(size_t)b * N + j] = make_float2(nx, ny);
      if (x_aug != nullptr) {
        if (n_aug == 8) {
#pragma unroll
          for (int k = 0; k < 8; ++k) x_aug[((size_t)k * B + b) * N + j] = isometry(k, nx, ny);
        } else {
          x_aug[(size_t)b * N + j] = make_float2(nx, ny);
        }
      }
    }
  }
}

__global__ void augment8_kernel(const float2* __restrict__ x, int B, int N,
                                float2* __restrict__ x_aug) {
  size_t total = (size_t)B * N;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    float2 v = __ldg(&x[i]);
#pragma unroll
    for (int k = 0; k < 8; ++k) x_aug[(size_t)k * total + i] = isometry(k, v.x, v.y);
  }
}

}  // namespace htsp

using namespace htsp;

extern "C" int htsp_extract_normalize(const float* data, const int64_t* fragment, int B, int Ntot,
                                      int N, int n_aug, float* x_new, float* scale, float* x_aug,
                                      void* stream) {
  HTSP_REQUIRE(data && fragment && x_new && scale, "null pointer");
  HTSP_REQUIRE(B >= 0 && N >= 1 && Ntot >= 1, "bad shape");
  HTSP_REQUIRE(N <= kExtractThreads * kExtractPerThread, "N > 1024 unsupported");
  HTSP_REQUIRE(n_aug == 1 || n_aug == 8, "n_aug must be 1 or 8");
  if (B == 0) return HTSP_OK;
  extract_normalize_kernel<<<B, kExtractThreads, 0, (cudaStream_t)stream>>>(
      (const float2*)data, (const long long*)fragment, B, Ntot, N, n_aug, (float2*)x_new, scale,
      (float2*)x_aug);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

extern "C" int htsp_augment8(const float* x, int B, int N, float* x_aug, void* stream) {
  HTSP_REQUIRE(x && x_aug, "null pointer");
  HTSP_REQUIRE(B >= 0 && N >= 1, "bad shape");
  if (B == 0) return HTSP_OK;
  size_t total = (size_t)B * N;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  augment8_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float2*)x, B, N,
                                                             (float2*)x_aug);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}
