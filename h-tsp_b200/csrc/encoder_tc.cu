// K2 (tensor-core variant): MHAEncoder.forward (rl4cop/models/models.py:30-60) with every Linear on
// tcgen05 (gemm_tc.cu, split-fp16 x3) and the bias / ReLU / residual / eval-BatchNorm epilogues fused.
// Activations travel between layers as fp32 (residual stream) plus the fp16 hi/lo split the next
// GEMM's TMA loads consume.  The [N x N] self-attention runs on mma.sync tensor cores (enc_attention_mma.cu) for N <= 1024.
#include <stdlib.h>
#include "common.cuh"

namespace htsp {

// per row (Bp*N rows): h hi/lo 2*256, qkv f32 1536, att hi/lo 2*256, h1 f32 512 + hi/lo 2*256,
// hidden hi/lo 2*1024
size_t encode_tc_ws_bytes(int Bp, int N) {
  const size_t M = (size_t)Bp * N;
  return align_up(M * (512 + 1536 + 512 + 512 + 512 + 2048), 256) + 1024;
}

// train mode: + z [M,128] fp32 (the pre-normalisation sums) and the statistics partials
size_t encode_tc_train_ws_bytes(int Bp, int N) {
  const size_t M = (size_t)Bp * N;
  return encode_tc_ws_bytes(Bp, N) + align_up(M * H * 4, 256) + bn_train_ws_bytes(M) + 256;
}

int launch_encode_tc(const void* blob, int n_layers, const float* x, int Bp, int N, float* emb,
                     void* ws, cudaStream_t s, const float* bn_gb, float* bn_stats) {
  const size_t M = (size_t)Bp * N;
  const float* fb = (const float*)blob;
  const __half* hb = blob_halfs(blob, n_layers);
  char* w = (char*)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
  __half* h_hi = (__half*)w;   w += M * H * 2;
  __half* h_lo = (__half*)w;   w += M * H * 2;
  float* qkv = (float*)w;      w += M * 3 * H * 4;   // fp32 [M,384] or, aliased, fp16 hi | lo [M,384] each
  __half* q_hi = (__half*)qkv;
  __half* q_lo = q_hi + M * 3 * H;
  __half* a_hi = (__half*)w;   w += M * H * 2;
  __half* a_lo = (__half*)w;   w += M * H * 2;
  float* h1 = (float*)w;       w += M * H * 4;
  __half* h1_hi = (__half*)w;  w += M * H * 2;
  __half* h1_lo = (__half*)w;  w += M * H * 2;
  __half* f_hi = (__half*)w;   w += M * FF * 2;
  __half* f_lo = (__half*)w;   w += M * FF * 2;
  const bool train = bn_gb != nullptr;        // batch statistics (PathSolver.training_step): see bn_train.cu
  HTSP_REQUIRE(train == (bn_stats != nullptr), "train-mode BatchNorm needs both the affine parameters and the statistics output");
  float* z = nullptr;
  void* bn_ws = nullptr;
  if (train) {
    w = (char*)(((uintptr_t)w + 255) & ~(uintptr_t)255);
    z = (float*)w;             w += align_up(M * H * 4, 256);
    bn_ws = w;
  }
  int rc;
  const char* ffe = getenv("HTSP_FF_FUSED");
  const bool ff_fused = !(ffe != nullptr && atoi(ffe) == 0) && !train;
  if ((rc = launch_init_proj(fb, x, emb, M, s))) return rc;
  if ((rc = launch_split_f16(emb, h_hi, h_lo, M * H, s))) return rc;
  for (int l = 0; l < n_layers; ++l) {
    const EncLayerOff o = enc_layer_off(l);
    const EncLayerOffH oh = enc_layer_off_h(l);
    // q,k,v projections (models.py:31-33), no bias
    const bool tc_att = N <= 1024;    // tensor-core attention kernel stages one head's K/V tile (128 bytes per key)
    if ((rc = launch_linear_tc(h_hi, h_lo, hb + oh.wqkv_hi, hb + oh.wqkv_lo, nullptr, nullptr, nullptr,
                               nullptr, false, tc_att ? nullptr : qkv, tc_att ? q_hi : nullptr,
                               tc_att ? q_lo : nullptr, (int)M, 3 * H, H, s))) return rc;
    // softmax(q k^T / 4) v  (utils.py:12-29)
    if (tc_att) {
      if ((rc = launch_enc_attention_mma(q_hi, q_lo, a_hi, a_lo, Bp, N, s))) return rc;
    } else {
      if ((rc = launch_enc_attention(qkv, nullptr, a_hi, a_lo, Bp, N, s))) return rc;
    }
    // x = norm1(x + combine(attn))  (models.py:34-35)
    if (train) {
      if ((rc = launch_linear_tc(a_hi, a_lo, hb + oh.wo_hi, hb + oh.wo_lo, fb + o.bo, emb, nullptr, nullptr, false, z,
                                 nullptr, nullptr, (int)M, H, H, s))) return rc;
      if ((rc = launch_bn_train(z, M, bn_gb + ((size_t)l * 4 + 0) * H, bn_gb + ((size_t)l * 4 + 1) * H,
                                bn_stats + ((size_t)l * 4 + 0) * H, h1, h1_hi, h1_lo, bn_ws, s))) return rc;
    } else
    if ((rc = launch_linear_tc(a_hi, a_lo, hb + oh.wo_hi, hb + oh.wo_lo, fb + o.bo, emb, fb + o.bn1a,
                               fb + o.bn1b, false, h1, h1_hi, h1_lo, (int)M, H, H, s))) return rc;
    // feed_forward: Linear -> ReLU -> Linear, x = norm2(x + ff)  (models.py:22-26,36-37) in one kernel: the hidden
    // activation stays on the SM (HTSP_FF_FUSED=0 keeps the two GEMM launches: same bits, A/B comparisons)
    if (ff_fused) {
      if ((rc = launch_ff_fused_tc(h1_hi, h1_lo, h1, hb + oh.w1_hi, hb + oh.w1_lo, fb + o.b1, hb + oh.w2_hi, hb + oh.w2_lo,
                                   fb + o.b2, fb + o.bn2a, fb + o.bn2b, emb, h_hi, h_lo, (int)M, s))) return rc;
      continue;
    }
    // feed_forward: Linear -> ReLU -> Linear  (models.py:22-26,36)
    if ((rc = launch_linear_tc(h1_hi, h1_lo, hb + oh.w1_hi, hb + oh.w1_lo, fb + o.b1, nullptr, nullptr,
                               nullptr, true, nullptr, f_hi, f_lo, (int)M, FF, H, s))) return rc;
    // x = norm2(x + ff)  (models.py:36-37)
    if (train) {
      if ((rc = launch_linear_tc(f_hi, f_lo, hb + oh.w2_hi, hb + oh.w2_lo, fb + o.b2, h1, nullptr, nullptr, false, z,
                                 nullptr, nullptr, (int)M, H, FF, s))) return rc;
      if ((rc = launch_bn_train(z, M, bn_gb + ((size_t)l * 4 + 2) * H, bn_gb + ((size_t)l * 4 + 3) * H,
                                bn_stats + ((size_t)l * 4 + 2) * H, emb, h_hi, h_lo, bn_ws, s))) return rc;
      continue;
    }
    if ((rc = launch_linear_tc(f_hi, f_lo, hb + oh.w2_hi, hb + oh.w2_lo, fb + o.b2, h1, fb + o.bn2a,
                               fb + o.bn2b, false, emb, h_hi, h_lo, (int)M, H, FF, s))) return rc;
  }
  return HTSP_OK;
}

}  // namespace htsp
