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

// ---- fp16 hi/lo copies of the GEMM weights (tensor-core modes), offsets in halfs from the start of
// the fp16 section that follows the fp32 section (256-byte aligned)
struct EncLayerOffH {
  size_t wqkv_hi, wqkv_lo, wo_hi, wo_lo, w1_hi, w1_lo, w2_hi, w2_lo;
};
constexpr size_t kEncLayerHalfs = 2 * ((size_t)3 * H * H + (size_t)H * H + (size_t)FF * H + (size_t)H * FF);
__host__ __device__ inline EncLayerOffH enc_layer_off_h(int l) {
  EncLayerOffH o;
  size_t p = (size_t)l * kEncLayerHalfs;
  o.wqkv_hi = p; p += (size_t)3 * H * H;
  o.wqkv_lo = p; p += (size_t)3 * H * H;
  o.wo_hi = p;   p += (size_t)H * H;
  o.wo_lo = p;   p += (size_t)H * H;
  o.w1_hi = p;   p += (size_t)FF * H;
  o.w1_lo = p;   p += (size_t)FF * H;
  o.w2_hi = p;   p += (size_t)H * FF;
  o.w2_lo = p;   p += (size_t)H * FF;
  return o;
}
__host__ __device__ inline size_t wproj_hi_off_h(int n_layers) { return (size_t)n_layers * kEncLayerHalfs; }
__host__ __device__ inline size_t wproj_lo_off_h(int n_layers) { return wproj_hi_off_h(n_layers) + (size_t)PROJ * H; }
__host__ __device__ inline size_t half_section_halfs(int n_layers) { return wproj_lo_off_h(n_layers) + (size_t)PROJ * H; }
__host__ __device__ inline size_t half_section_byte_off(int n_layers) {
  return (dec_off(n_layers).end * sizeof(float) + 255) / 256 * 256;
}
inline const __half* blob_halfs(const void* blob, int n_layers) {
  return (const __half*)((const char*)blob + half_section_byte_off(n_layers));
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
// caller-supplied indices never address memory unchecked: kernels clamp them, decode_validate_kernel / K1 flag them
// (the results of a flagged call are NaN), see ADVICE round 1
__host__ __device__ inline int clamp_index(long long i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : (int)i); }

// ---- internal launchers (one per .cu) --------------------------------------------------
// encoder_f32.cu
int launch_linear_f32(const float* X, const float* W, const float* bias, const float* residual,
                      const float* bn_alpha, const float* bn_beta, bool relu, float* Y, int M,
                      int Nout, int K, cudaStream_t s);
int launch_encode_f32(const float* blob, int n_layers, const float* x, int Bp, int N, float* emb,
                      float* ws, cudaStream_t s);
size_t encode_f32_ws_floats(int Bp, int N);
// gemm_tc.cu (tcgen05 + TMA)
int launch_linear_tc(const __half* xhi, const __half* xlo, const __half* whi, const __half* wlo,
                     const float* bias, const float* residual, const float* bn_alpha,
                     const float* bn_beta, bool relu, float* y32, __half* yhi, __half* ylo, int M,
                     int Nout, int K, cudaStream_t s);
int launch_split_f16(const float* x, __half* hi, __half* lo, size_t n, cudaStream_t s);
int launch_ff_fused_tc(const __half* xhi, const __half* xlo, const float* x32, const __half* w1hi, const __half* w1lo,
                       const float* b1, const __half* w2hi, const __half* w2lo, const float* b2, const float* bn_alpha,
                       const float* bn_beta, float* y32, __half* yhi, __half* ylo, int M, cudaStream_t s);
// encoder_tc.cu
size_t encode_tc_ws_bytes(int Bp, int N);
size_t encode_tc_train_ws_bytes(int Bp, int N);
size_t bn_train_ws_bytes(size_t M);
int launch_bn_train(const float* z, size_t M, const float* gamma, const float* beta, float* stats, float* y32,
                    __half* yhi, __half* ylo, void* ws, cudaStream_t s);
// bn_gb / bn_stats: NULL = eval-mode BatchNorm (running statistics folded into the blob); else train mode: device
// [L][4][128] gamma1 | beta1 | gamma2 | beta2 in, batch mean1 | biased var1 | mean2 | biased var2 out
int launch_encode_tc(const void* blob, int n_layers, const float* x, int Bp, int N, float* emb,
                     void* ws, cudaStream_t s, const float* bn_gb = nullptr, float* bn_stats = nullptr);
int launch_init_proj(const float* blob, const float* x, float* emb, size_t M, cudaStream_t s);
int launch_enc_attention(const float* qkv, float* out, __half* ohi, __half* olo, int Bp, int N,
                         cudaStream_t s);
// enc_attention_mma.cu
int launch_enc_attention_mma(const __half* qkv_hi, const __half* qkv_lo, __half* out_hi, __half* out_lo,
                             int Bp, int N, cudaStream_t s);
// decode_prep.cu
size_t decode_prep_tc_ws_bytes(int Bp, int N);
int launch_decode_prep(const void* blob, int n_layers, const float* emb, const int32_t* src,
                       const int32_t* tgt, int Bp, int N, float* proj, float* qfix, float* c0,
                       bool tensor_cores, void* tc_ws, int32_t* flag, cudaStream_t s);
// flag[0] |= 1 if any src / tgt / first / forced_pi entry is outside [0, N); poison: reward := NaN when flagged
int launch_decode_validate(const int32_t* src, const int32_t* tgt, const int32_t* first, const int32_t* forced_pi,
                           int Bp, int N, int G, int32_t* flag, cudaStream_t s);
int launch_decode_poison(float* reward, size_t n, const int32_t* flag, cudaStream_t s);
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

// decode_tc.cu (tcgen05 + TMEM rollout kernel, N <= 224)
bool decode_tc_supported(int N, int G);
size_t decode_tc_pack_bytes(int Bp, int N);
size_t decode_tc_dbg_floats(int N);
int launch_decode_tc(const float* x, const float* proj, const float* qfix, const float* c0,
                     const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp, int N,
                     int G, int mode, const int32_t* forced_pi, int32_t* pi, float* reward,
                     float* logits_out, void* pack_ws, float* dbg, int dbg_step, const uint64_t* sample_seed,
                     cudaStream_t s);

// decode_tc2.cu (second tcgen05 rollout kernel: tf32 probabilities in place, three softmax warps per row quarter)
bool decode_tc2_supported(int N, int G);
size_t decode_tc2_pack_bytes(int Bp, int N);
size_t decode_tc2_dbg_floats(int N);
int decode_tc2_ctas_per_sm(int Bp, int N, int G);
bool decode_tc2_geometry(int N, int32_t out[5]);
bool decode_tc2_stage(int N, int i, uint32_t out[2]);
int launch_decode_tc2(const float* x, const float* proj, const float* qfix, const float* c0,
                      const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp, int N,
                      int G, const int32_t* forced_pi, int32_t* pi, float* reward,
                      float* logits_out, void* pack_ws, float* dbg, int dbg_step, const uint64_t* sample_seed,
                      float* logp, cudaStream_t s);

}  // namespace htsp
