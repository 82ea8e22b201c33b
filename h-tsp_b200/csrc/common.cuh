// Shared internals of libhtsp_b200: packed-weights layout, error plumbing, launch counter.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/htsp_b200.h"

namespace htsp {

constexpr int H = HTSP_EMBED_DIM;   // 128
constexpr int NH = HTSP_N_HEADS;    // 8
constexpr int DK = H / NH;          // 16
constexpr int FF = HTSP_FF_DIM;     // 512
constexpr int PROJ = 5 * H;         // decoder projections: K | V | QL | QF | K2  (640)

// ---- packed weight blob (fp32, offsets in floats) -------------------------------------
struct EncLayerOff {
  size_t wqkv, wo, bo, bn1a, bn1b, w1, b1, w2, b2, bn2a, bn2b;
};
constexpr size_t kEncInitW = 0;                  // [128,2]
constexpr size_t kEncInitB = kEncInitW + H * 2;  // [128]
constexpr size_t kEncLayer0 = kEncInitB + H;
constexpr size_t kEncLayerFloats =
    (size_t)3 * H * H + H * H + H + H + H + (size_t)FF * H + FF + (size_t)H * FF + H + H + H;

__host__ __device__ inline EncLayerOff enc_layer_off(int l) {
  EncLayerOff o;
  size_t p = kEncLayer0 + (size_t)l * kEncLayerFloats;
  o.wqkv = p; p += (size_t)3 * H * H;   // rows: Wq | Wk | Wv, each [128,128] (out,in)
  o.wo = p;   p += (size_t)H * H;
  o.bo = p;   p += H;
  o.bn1a = p; p += H;
  o.bn1b = p; p += H;
  o.w1 = p;   p += (size_t)FF * H;      // [512,128]
  o.b1 = p;   p += FF;
  o.w2 = p;   p += (size_t)H * FF;      // [128,512]
  o.b2 = p;   p += H;
  o.bn2a = p; p += H;
  o.bn2b = p; p += H;
  return o;
}
struct DecOff {
  size_t wg, ws, wt, wproj, wc, bc, end;
};
__host__ __device__ inline DecOff dec_off(int n_layers) {
  DecOff d;
  size_t p = kEncLayer0 + (size_t)n_layers * kEncLayerFloats;
  d.wg = p;    p += (size_t)H * H;
  d.ws = p;    p += (size_t)H * H;
  d.wt = p;    p += (size_t)H * H;
  d.wproj = p; p += (size_t)PROJ * H;   // rows: Wk | Wv | Wq_last | Wq_first | Wc^T
  d.wc = p;    p += (size_t)H * H;
  d.bc = p;    p += H;
  d.end = p;
  return d;
}

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void kernel_timer_begin(cudaStream_t s);  // records the armed start event, if any
void kernel_timer_end(cudaStream_t s);    // records the armed stop event and disarms

#define HTSP_CHECK_CUDA(expr)                                                        \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      ::htsp::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                \
                        cudaGetErrorString(_e));                                     \
      return HTSP_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

#define HTSP_CHECK_LAUNCH()                                                          \
  do {                                                                               \
    ::htsp::count_launch();                                                          \
    HTSP_CHECK_CUDA(cudaGetLastError());                                             \
  } while (0)

#define HTSP_REQUIRE(cond, msg)                                                      \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      ::htsp::set_error("%s:%d: invalid argument: %s (%s)", __FILE__, __LINE__, msg, \
                        #cond);                                                      \
      return HTSP_ERR_INVALID_ARG;                                                   \
    }                                                                                \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- internal launchers (one per .cu) --------------------------------------------------
// encoder_f32.cu
int launch_linear_f32(const float* X, const float* W, const float* bias, const float* residual,
                      const float* bn_alpha, const float* bn_beta, bool relu, float* Y, int M,
                      int Nout, int K, cudaStream_t s);
int launch_encode_f32(const float* blob, int n_layers, const float* x, int Bp, int N, float* emb,
                      float* ws, cudaStream_t s);
size_t encode_f32_ws_floats(int Bp, int N);
// decode_prep.cu
int launch_decode_prep(const float* blob, int n_layers, const float* emb, const int32_t* src,
                       const int32_t* tgt, int Bp, int N, float* proj, float* qfix, float* c0,
                       cudaStream_t s);
// decode_f32.cu
int launch_decode_f32(const float* blob, int n_layers, const float* x, const float* emb,
                      const float* proj, const float* qfix, const int32_t* src,
                      const int32_t* tgt, const int32_t* first, int Bp, int N, int G,
                      const int32_t* forced_pi, int32_t* pi, float* reward, float* logits_out,
                      cudaStream_t s);
// decode_mma.cu
size_t decode_mma_pack_bytes(int Bp, int N);
int launch_decode_mma(const float* x, const float* proj, const float* qfix, const float* c0,
                      const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp, int N,
                      int G, int mode, const int32_t* forced_pi, int32_t* pi, float* reward,
                      float* logits_out, void* pack_ws, cudaStream_t s);

}  // namespace htsp
