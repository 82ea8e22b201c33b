// PathDecoder.reset (rl4cop/models/models.py:360-402) hoisted into per-instance tables:
//   proj[m, 0:128]   = Wk  emb_m        (glimpse keys)
//   proj[m,128:256]  = Wv  emb_m        (glimpse values)
//   proj[m,256:384]  = Wq_last  emb_m   (models.py:409-411 applied once per node, not per row-step)
//   proj[m,384:512]  = Wq_first emb_m   (models.py:393-397)
//   proj[m,512:640]  = Wc^T emb_m       (multi_head_combine folded into the pointer keys,
//                                        used by the tensor-core modes only)
//   qfix[b]  = (Wq_graph mean_n(emb) + Wq_source emb[src]) + Wq_target emb[tgt]   (models.py:378-392)
//   c0[b,n]  = bc . emb[b,n]            (bias term of the folded pointer logits)
#include <algorithm>
#include "common.cuh"

namespace htsp {

// One CTA (4 warps) per instance.  Pass 1 walks the embedding rows once: warp w takes rows w, w+4, ...,
// lane l holds columns 4l..4l+3 (one 512-byte row per load instruction), accumulates the column sums for
// the graph mean and reduces bc . emb[n] (c0) with shuffles.  Pass 2: thread t computes output t of the
// three 128 x 128 mat-vecs from float4 weight-row reads (L1-resident across the CTAs of an SM).
__global__ void __launch_bounds__(128)
decode_prep_kernel(const float* __restrict__ emb, const float* __restrict__ Wg,
                   const float* __restrict__ Ws, const float* __restrict__ Wt,
                   const float* __restrict__ bc, const int32_t* __restrict__ src,
                   const int32_t* __restrict__ tgt, int N, float* __restrict__ qfix,
                   float* __restrict__ c0, float amax_limit, int32_t* __restrict__ flag) {
  __shared__ __align__(16) float part[4][H];
  __shared__ __align__(16) float mean[H], es[H], et[H];
  const int b = blockIdx.x, t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const float* e = emb + (size_t)b * N * H;
  const float4 bc4 = __ldg((const float4*)bc + lane);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float amax = 0.f;                      // largest |embedding| this lane has seen (NaNs show up in the sums instead)
#pragma unroll 4
  for (int n = warp; n < N; n += 4) {
    const float4 v = __ldg((const float4*)(e + (size_t)n * H) + lane);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    amax = fmaxf(fmaxf(amax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    float d = fmaf(v.w, bc4.w, fmaf(v.z, bc4.z, fmaf(v.y, bc4.y, v.x * bc4.x)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) c0[(size_t)b * N + n] = d;
  }
  *(float4*)&part[warp][4 * lane] = acc;
  es[t] = __ldg(&e[(size_t)clamp_index(src[b], N) * H + t]);
  et[t] = __ldg(&e[(size_t)clamp_index(tgt[b], N) * H + t]);
  __syncthreads();
  mean[t] = ((part[0][t] + part[1][t]) + (part[2][t] + part[3][t])) / (float)N;
  // Embeddings the arithmetic of this call cannot represent -- non-finite ones (an overflow upstream), or, in the
  // tensor-core modes, |v| >= 65504 (the operands are split into fp16 hi + lo) -- flag the call: its rewards come
  // back NaN (decode_poison_kernel) instead of tours decided by garbage logits.
  const bool unrepresentable = !(fabsf(mean[t]) <= 3.0e38f) || !(amax < amax_limit);
  if (__syncthreads_or(unrepresentable) && t == 0) atomicOr(flag, 2);
  const float4* wg = (const float4*)(Wg + (size_t)t * H);
  const float4* ws = (const float4*)(Ws + (size_t)t * H);
  const float4* wt = (const float4*)(Wt + (size_t)t * H);
  float g = 0.f, s = 0.f, q = 0.f;
#pragma unroll 8
  for (int i = 0; i < H / 4; ++i) {
    const float4 a = __ldg(wg + i), c = __ldg(ws + i), d = __ldg(wt + i);
    const float4 m4 = *(const float4*)&mean[4 * i], s4 = *(const float4*)&es[4 * i], t4 = *(const float4*)&et[4 * i];
    g = fmaf(m4.w, a.w, fmaf(m4.z, a.z, fmaf(m4.y, a.y, fmaf(m4.x, a.x, g))));
    s = fmaf(s4.w, c.w, fmaf(s4.z, c.z, fmaf(s4.y, c.y, fmaf(s4.x, c.x, s))));
    q = fmaf(t4.w, d.w, fmaf(t4.z, d.z, fmaf(t4.y, d.y, fmaf(t4.x, d.x, q))));
  }
  qfix[(size_t)b * H + t] = (g + s) + q;
}

size_t decode_prep_tc_ws_bytes(int Bp, int N) {
  return align_up((size_t)Bp * N * H * 2 * 2, 256) + 256;   // emb hi/lo
}

int launch_decode_prep(const void* blobv, int n_layers, const float* emb, const int32_t* src,
                       const int32_t* tgt, int Bp, int N, float* proj, float* qfix, float* c0,
                       bool tensor_cores, void* tc_ws, int32_t* flag, cudaStream_t s) {
  const float* blob = (const float*)blobv;
  DecOff d = dec_off(n_layers);
  int rc;
  if (tensor_cores) {
    const size_t M = (size_t)Bp * N;
    __half* e_hi = (__half*)(((uintptr_t)tc_ws + 255) & ~(uintptr_t)255);
    __half* e_lo = e_hi + M * H;
    const __half* hb = blob_halfs(blobv, n_layers);
    if ((rc = launch_split_f16(emb, e_hi, e_lo, M * H, s))) return rc;
    rc = launch_linear_tc(e_hi, e_lo, hb + wproj_hi_off_h(n_layers), hb + wproj_lo_off_h(n_layers),
                          nullptr, nullptr, nullptr, nullptr, false, proj, nullptr, nullptr, (int)M,
                          PROJ, H, s);
  } else {
    rc = launch_linear_f32(emb, blob + d.wproj, nullptr, nullptr, nullptr, nullptr, false, proj,
                           Bp * N, PROJ, H, s);
  }
  if (rc) return rc;
  decode_prep_kernel<<<Bp, 128, 0, s>>>(emb, blob + d.wg, blob + d.ws, blob + d.wt, blob + d.bc,
                                        src, tgt, N, qfix, c0, tensor_cores ? 65504.f : 3.0e38f, flag);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

// Range check of everything the rollout kernels use as an index (they clamp; a flagged call returns NaN rewards)
__global__ void decode_validate_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ tgt,
                                       const int32_t* __restrict__ first, const int32_t* __restrict__ forced_pi,
                                       int Bp, int N, int G, int32_t* __restrict__ flag) {
  const size_t n_forced = forced_pi ? (size_t)Bp * G * N : 0;
  const size_t total = (size_t)2 * Bp + G + n_forced;
  bool bad = false;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int v;
    if (i < (size_t)Bp) v = src[i];
    else if (i < (size_t)2 * Bp) v = tgt[i - Bp];
    else if (i < (size_t)2 * Bp + G) v = first[i - 2 * (size_t)Bp];
    else v = forced_pi[i - 2 * (size_t)Bp - G];
    bad = bad || v < 0 || v >= N;
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1);
}
__global__ void decode_poison_kernel(float* __restrict__ reward, size_t n, const int32_t* __restrict__ flag) {
  if (*flag == 0) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    reward[i] = __int_as_float(0x7fc00000);
}
int launch_decode_validate(const int32_t* src, const int32_t* tgt, const int32_t* first, const int32_t* forced_pi,
                           int Bp, int N, int G, int32_t* flag, cudaStream_t s) {
  HTSP_CHECK_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t), s));
  const size_t total = (size_t)2 * Bp + G + (forced_pi ? (size_t)Bp * G * N : 0);
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 2048);
  decode_validate_kernel<<<blocks, 256, 0, s>>>(src, tgt, first, forced_pi, Bp, N, G, flag);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}
int launch_decode_poison(float* reward, size_t n, const int32_t* flag, cudaStream_t s) {
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 1024);
  decode_poison_kernel<<<blocks, 256, 0, s>>>(reward, n, flag);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

}  // namespace htsp
