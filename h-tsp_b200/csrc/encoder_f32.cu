// K2 (fp32 variant): MHAEncoder.forward in eval mode on CUDA cores, reference op order.
// rl4cop/models/models.py:30-60 and rl4cop/utils/utils.py:12-34.  This is the parity anchor
// (HTSP_MODE_FP32); every Linear is one tiled SGEMM with the bias / ReLU / residual /
// eval-BatchNorm epilogue fused, attention never materialises the [Bp,8,N,N] score tensor.
#include "common.cuh"

namespace htsp {

// ---------------------------------------------------------------------------------------
// Y[M,Nout] = epi(X[M,K] * W[Nout,K]^T): 128x128 tile, BK=8, 256 threads, 8x8 per thread.
// epi: (+bias) -> (ReLU) -> (residual + .) -> (. * alpha + beta)
// ---------------------------------------------------------------------------------------
constexpr int LBM = 128, LBN = 128, LBK = 8;

template <bool RELU>
__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ X, const float* __restrict__ W,
                  const float* __restrict__ bias, const float* __restrict__ residual,
                  const float* __restrict__ bn_alpha, const float* __restrict__ bn_beta,
                  float* __restrict__ Y, int M, int Nout, int K) {
  __shared__ __align__(16) float As[2][LBK][LBM];
  __shared__ __align__(16) float Bs[2][LBK][LBN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * LBM, n0 = blockIdx.x * LBN;
  const int tx = tid & 15, ty = tid >> 4;
  const int lrow = tid >> 1, lk = (tid & 1) * 4;  // loader: one float4 of A and one of W

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const bool a_ok = (m0 + lrow) < M;
  const float* a_ptr = X + (size_t)(m0 + lrow) * K + lk;
  const float* b_ptr = W + (size_t)(n0 + lrow) * K + lk;
  float4 a_reg = a_ok ? __ldg((const float4*)a_ptr) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 b_reg = __ldg((const float4*)b_ptr);
  const int nk = K / LBK;
  int buf = 0;
  As[0][lk + 0][lrow] = a_reg.x; As[0][lk + 1][lrow] = a_reg.y;
  As[0][lk + 2][lrow] = a_reg.z; As[0][lk + 3][lrow] = a_reg.w;
  Bs[0][lk + 0][lrow] = b_reg.x; Bs[0][lk + 1][lrow] = b_reg.y;
  Bs[0][lk + 2][lrow] = b_reg.z; Bs[0][lk + 3][lrow] = b_reg.w;
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) {
      a_reg = a_ok ? __ldg((const float4*)(a_ptr + (kt + 1) * LBK)) : make_float4(0.f, 0.f, 0.f, 0.f);
      b_reg = __ldg((const float4*)(b_ptr + (kt + 1) * LBK));
    }
#pragma unroll
    for (int k = 0; k < LBK; ++k) {
      float4 a0 = *(const float4*)&As[buf][k][ty * 4];
      float4 a1 = *(const float4*)&As[buf][k][64 + ty * 4];
      float4 b0 = *(const float4*)&Bs[buf][k][tx * 4];
      float4 b1 = *(const float4*)&Bs[buf][k][64 + tx * 4];
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      buf ^= 1;
      As[buf][lk + 0][lrow] = a_reg.x; As[buf][lk + 1][lrow] = a_reg.y;
      As[buf][lk + 2][lrow] = a_reg.z; As[buf][lk + 3][lrow] = a_reg.w;
      Bs[buf][lk + 0][lrow] = b_reg.x; Bs[buf][lk + 1][lrow] = b_reg.y;
      Bs[buf][lk + 2][lrow] = b_reg.z; Bs[buf][lk + 3][lrow] = b_reg.w;
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      if (bias != nullptr) {
        float4 bb = __ldg((const float4*)(bias + n));
        v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
      }
      if (RELU) {
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = fmaxf(v[q], 0.f);
      }
      if (residual != nullptr) {
        float4 r = __ldg((const float4*)(residual + (size_t)m * Nout + n));
        v[0] = r.x + v[0]; v[1] = r.y + v[1]; v[2] = r.z + v[2]; v[3] = r.w + v[3];
      }
      if (bn_alpha != nullptr) {
        float4 al = __ldg((const float4*)(bn_alpha + n));
        float4 be = __ldg((const float4*)(bn_beta + n));
        v[0] = fmaf(v[0], al.x, be.x); v[1] = fmaf(v[1], al.y, be.y);
        v[2] = fmaf(v[2], al.z, be.z); v[3] = fmaf(v[3], al.w, be.w);
      }
      *(float4*)(Y + (size_t)m * Nout + n) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

int launch_linear_f32(const float* X, const float* W, const float* bias, const float* residual,
                      const float* bn_alpha, const float* bn_beta, bool relu, float* Y, int M,
                      int Nout, int K, cudaStream_t s) {
  HTSP_REQUIRE(Nout % LBN == 0 && K % LBK == 0 && M > 0, "linear_f32 shape");
  dim3 grid(Nout / LBN, (M + LBM - 1) / LBM);
  if (relu)
    linear_f32_kernel<true><<<grid, 256, 0, s>>>(X, W, bias, residual, bn_alpha, bn_beta, Y, M, Nout, K);
  else
    linear_f32_kernel<false><<<grid, 256, 0, s>>>(X, W, bias, residual, bn_alpha, bn_beta, Y, M, Nout, K);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

// init_projection_layer: Linear(2,128)  (models.py:47,57)
__global__ void init_proj_kernel(const float2* __restrict__ x, const float* __restrict__ W,
                                 const float* __restrict__ b, float* __restrict__ out, size_t M) {
  size_t total = M * H;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    size_t m = i / H;
    int c = (int)(i % H);
    float2 p = __ldg(&x[m]);
    float w0 = __ldg(&W[c * 2]), w1 = __ldg(&W[c * 2 + 1]);
    out[i] = fmaf(p.y, w1, p.x * w0) + __ldg(&b[c]);
  }
}

// Encoder self-attention for one (instance, head): softmax(q k^T / 4) v, fp32, two passes over
// the keys (max, then exp/accumulate) with K_h and V_h staged in shared memory; one query per
// thread.  qkv row layout: [q(128) | k(128) | v(128)].           (utils.py:12-29)
__global__ void __launch_bounds__(256)
enc_attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                         __half* __restrict__ ohi, __half* __restrict__ olo, int N) {
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;
  float* Vs = sm + (size_t)N * DK;
  const int h = blockIdx.x, b = blockIdx.y;
  const float* base = qkv + (size_t)b * N * (3 * H);
  for (int i = threadIdx.x; i < N * (DK / 4); i += blockDim.x) {
    int j = i / (DK / 4), c = (i % (DK / 4)) * 4;
    *(float4*)&Ks[j * DK + c] = __ldg((const float4*)(base + (size_t)j * 3 * H + H + h * DK + c));
    *(float4*)&Vs[j * DK + c] = __ldg((const float4*)(base + (size_t)j * 3 * H + 2 * H + h * DK + c));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float q[DK];
#pragma unroll
    for (int c = 0; c < DK; c += 4) {
      float4 t = __ldg((const float4*)(base + (size_t)i * 3 * H + h * DK + c));
      q[c] = t.x; q[c + 1] = t.y; q[c + 2] = t.z; q[c + 3] = t.w;
    }
    float m = -INFINITY;
    for (int j = 0; j < N; ++j) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < DK; ++c) s = fmaf(q[c], Ks[j * DK + c], s);
      m = fmaxf(m, s * 0.25f);
    }
    float l = 0.f, o[DK];
#pragma unroll
    for (int c = 0; c < DK; ++c) o[c] = 0.f;
    for (int j = 0; j < N; ++j) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < DK; ++c) s = fmaf(q[c], Ks[j * DK + c], s);
      float p = expf(s * 0.25f - m);
      l += p;
#pragma unroll
      for (int c = 0; c < DK; ++c) o[c] = fmaf(p, Vs[j * DK + c], o[c]);
    }
    float inv = 1.f / l;
    const size_t doff = ((size_t)b * N + i) * H + h * DK;
#pragma unroll
    for (int c = 0; c < DK; c += 4) {
      const float v0 = o[c] * inv, v1 = o[c + 1] * inv, v2 = o[c + 2] * inv, v3 = o[c + 3] * inv;
      if (out != nullptr) *(float4*)(out + doff + c) = make_float4(v0, v1, v2, v3);
      if (ohi != nullptr) {   // fp16 hi/lo split for the tensor-core combine GEMM
        __half2 h0 = __floats2half2_rn(v0, v1), h1 = __floats2half2_rn(v2, v3);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        __half2 l0 = __floats2half2_rn(v0 - f0.x, v1 - f0.y), l1 = __floats2half2_rn(v2 - f1.x, v3 - f1.y);
        *(uint2*)(ohi + doff + c) = make_uint2(*(uint32_t*)&h0, *(uint32_t*)&h1);
        *(uint2*)(olo + doff + c) = make_uint2(*(uint32_t*)&l0, *(uint32_t*)&l1);
      }
    }
  }
}

size_t encode_f32_ws_floats(int Bp, int N) {
  size_t M = (size_t)Bp * N;
  return M * (size_t)(3 * H + H + H + FF);
}

int launch_init_proj(const float* blob, const float* x, float* emb, size_t M, cudaStream_t s) {
  size_t total = M * H;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  init_proj_kernel<<<blocks, 256, 0, s>>>((const float2*)x, blob + kEncInitW, blob + kEncInitB, emb, M);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

int launch_enc_attention(const float* qkv, float* out, __half* ohi, __half* olo, int Bp, int N,
                         cudaStream_t s) {
  const size_t att_smem = (size_t)N * DK * 2 * sizeof(float);
  if (att_smem > 48 * 1024)
    HTSP_CHECK_CUDA(cudaFuncSetAttribute(enc_attention_f32_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_smem));
  enc_attention_f32_kernel<<<dim3(NH, Bp), 256, att_smem, s>>>(qkv, out, ohi, olo, N);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

int launch_encode_f32(const float* blob, int n_layers, const float* x, int Bp, int N, float* emb,
                      float* ws, cudaStream_t s) {
  const size_t M = (size_t)Bp * N;
  float* qkv = ws;
  float* att = qkv + M * 3 * H;
  float* h1 = att + M * H;
  float* hid = h1 + M * H;
  {
    int rc = launch_init_proj(blob, x, emb, M, s);
    if (rc) return rc;
  }
  for (int l = 0; l < n_layers; ++l) {
    EncLayerOff o = enc_layer_off(l);
    int rc;
    rc = launch_linear_f32(emb, blob + o.wqkv, nullptr, nullptr, nullptr, nullptr, false, qkv, (int)M, 3 * H, H, s);
    if (rc) return rc;
    rc = launch_enc_attention(qkv, att, nullptr, nullptr, Bp, N, s);
    if (rc) return rc;
    rc = launch_linear_f32(att, blob + o.wo, blob + o.bo, emb, blob + o.bn1a, blob + o.bn1b, false, h1, (int)M, H, H, s);
    if (rc) return rc;
    rc = launch_linear_f32(h1, blob + o.w1, blob + o.b1, nullptr, nullptr, nullptr, true, hid, (int)M, FF, H, s);
    if (rc) return rc;
    rc = launch_linear_f32(hid, blob + o.w2, blob + o.b2, h1, blob + o.bn2a, blob + o.bn2b, false, emb, (int)M, H, FF, s);
    if (rc) return rc;
  }
  return HTSP_OK;
}

}  // namespace htsp
