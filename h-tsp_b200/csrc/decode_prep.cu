// PathDecoder.reset (rl4cop/models/models.py:360-402) hoisted into per-instance tables:
//   proj[m, 0:128]   = Wk  emb_m        (glimpse keys)
//   proj[m,128:256]  = Wv  emb_m        (glimpse values)
//   proj[m,256:384]  = Wq_last  emb_m   (models.py:409-411 applied once per node, not per row-step)
//   proj[m,384:512]  = Wq_first emb_m   (models.py:393-397)
//   proj[m,512:640]  = Wc^T emb_m       (multi_head_combine folded into the pointer keys,
//                                        used by the tensor-core modes only)
//   qfix[b]  = (Wq_graph mean_n(emb) + Wq_source emb[src]) + Wq_target emb[tgt]   (models.py:378-392)
//   c0[b,n]  = bc . emb[b,n]            (bias term of the folded pointer logits)
#include "common.cuh"

namespace htsp {

__global__ void __launch_bounds__(128)
decode_prep_kernel(const float* __restrict__ emb, const float* __restrict__ Wg,
                   const float* __restrict__ Ws, const float* __restrict__ Wt,
                   const float* __restrict__ bc, const int32_t* __restrict__ src,
                   const int32_t* __restrict__ tgt, int N, float* __restrict__ qfix,
                   float* __restrict__ c0) {
  __shared__ float mean[H], es[H], et[H], bcs[H];
  const int b = blockIdx.x, t = threadIdx.x;
  const float* e = emb + (size_t)b * N * H;
  float acc = 0.f;
  for (int n = 0; n < N; ++n) acc += __ldg(&e[(size_t)n * H + t]);
  mean[t] = acc / (float)N;
  es[t] = __ldg(&e[(size_t)src[b] * H + t]);
  et[t] = __ldg(&e[(size_t)tgt[b] * H + t]);
  bcs[t] = __ldg(&bc[t]);
  __syncthreads();
  float g = 0.f, s = 0.f, q = 0.f;
  for (int i = 0; i < H; ++i) {
    g = fmaf(mean[i], __ldg(&Wg[t * H + i]), g);
    s = fmaf(es[i], __ldg(&Ws[t * H + i]), s);
    q = fmaf(et[i], __ldg(&Wt[t * H + i]), q);
  }
  qfix[(size_t)b * H + t] = (g + s) + q;
  for (int n = t; n < N; n += blockDim.x) {
    float d = 0.f;
    for (int i = 0; i < H; ++i) d = fmaf(bcs[i], __ldg(&e[(size_t)n * H + i]), d);
    c0[(size_t)b * N + n] = d;
  }
}

size_t decode_prep_tc_ws_bytes(int Bp, int N) {
  return align_up((size_t)Bp * N * H * 2 * 2, 256) + 256;   // emb hi/lo
}

int launch_decode_prep(const void* blobv, int n_layers, const float* emb, const int32_t* src,
                       const int32_t* tgt, int Bp, int N, float* proj, float* qfix, float* c0,
                       bool tensor_cores, void* tc_ws, cudaStream_t s) {
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
                                        src, tgt, N, qfix, c0);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

}  // namespace htsp
