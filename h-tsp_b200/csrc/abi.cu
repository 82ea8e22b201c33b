// extern "C" entry points of libhtsp_b200 that are not defined next to their kernels:
// library bookkeeping, weight packing, encode and decode dispatch.  See include/htsp_b200.h.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <vector>
#include "common.cuh"

namespace htsp {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static thread_local cudaEvent_t g_ev_start = nullptr, g_ev_stop = nullptr;
void kernel_timer_begin(cudaStream_t s) {
  if (g_ev_start && g_ev_stop) cudaEventRecord(g_ev_start, s);
}
void kernel_timer_end(cudaStream_t s) {
  if (g_ev_start && g_ev_stop) cudaEventRecord(g_ev_stop, s);
  g_ev_start = g_ev_stop = nullptr;
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

}  // namespace htsp

using namespace htsp;

extern "C" int htsp_abi_version(void) { return HTSP_ABI_VERSION; }
extern "C" const char* htsp_last_error(void) { return g_err; }
extern "C" uint64_t htsp_launch_count(void) { return g_launches.load(); }

extern "C" int htsp_set_kernel_timer(void* start_event, void* stop_event) {
  g_ev_start = (cudaEvent_t)start_event;
  g_ev_stop = (cudaEvent_t)stop_event;
  return HTSP_OK;
}

extern "C" size_t htsp_weights_blob_bytes(int n_layers) {
  if (n_layers < 0) return 0;
  return half_section_byte_off(n_layers) + half_section_halfs(n_layers) * sizeof(__half);
}
extern "C" int htsp_n_weight_tensors(int n_layers) { return 2 + 17 * n_layers + 9; }

extern "C" int htsp_pack_weights(const float* const* t, int n_tensors, int n_layers,
                                 void* dev_blob, void* stream) {
  HTSP_REQUIRE(t && dev_blob, "null pointer");
  HTSP_REQUIRE(n_layers >= 0 && n_tensors == htsp_n_weight_tensors(n_layers), "tensor count");
  for (int i = 0; i < n_tensors; ++i) HTSP_REQUIRE(t[i] != nullptr, "null tensor");
  const DecOff d = dec_off(n_layers);
  const size_t total_bytes = htsp_weights_blob_bytes(n_layers);
  std::vector<float> blob((total_bytes + 3) / 4, 0.f);
  float* o = blob.data();
  int k = 0;
  memcpy(o + kEncInitW, t[k++], sizeof(float) * H * 2);
  memcpy(o + kEncInitB, t[k++], sizeof(float) * H);
  auto fold_bn = [&](const float* w, const float* b, const float* rm, const float* rv, float* alpha,
                     float* beta) {
    for (int c = 0; c < H; ++c) {  // eval BatchNorm1d, eps 1e-5 (models.py:35,37)
      const float invstd = 1.0f / sqrtf(rv[c] + 1e-5f);
      alpha[c] = invstd * w[c];
      beta[c] = b[c] - rm[c] * alpha[c];
    }
  };
  for (int l = 0; l < n_layers; ++l) {
    const EncLayerOff e = enc_layer_off(l);
    memcpy(o + e.wqkv, t[k++], sizeof(float) * H * H);                  // Wq
    memcpy(o + e.wqkv + (size_t)H * H, t[k++], sizeof(float) * H * H);  // Wk
    memcpy(o + e.wqkv + (size_t)2 * H * H, t[k++], sizeof(float) * H * H);  // Wv
    memcpy(o + e.wo, t[k++], sizeof(float) * H * H);
    memcpy(o + e.bo, t[k++], sizeof(float) * H);
    memcpy(o + e.w1, t[k++], sizeof(float) * FF * H);
    memcpy(o + e.b1, t[k++], sizeof(float) * FF);
    memcpy(o + e.w2, t[k++], sizeof(float) * H * FF);
    memcpy(o + e.b2, t[k++], sizeof(float) * H);
    fold_bn(t[k], t[k + 1], t[k + 2], t[k + 3], o + e.bn1a, o + e.bn1b); k += 4;
    fold_bn(t[k], t[k + 1], t[k + 2], t[k + 3], o + e.bn2a, o + e.bn2b); k += 4;
  }
  const float* wq_graph = t[k++]; const float* wq_source = t[k++]; const float* wq_target = t[k++];
  const float* wq_first = t[k++]; const float* wq_last = t[k++];
  const float* wk = t[k++]; const float* wv = t[k++]; const float* wc = t[k++]; const float* bc = t[k++];
  memcpy(o + d.wg, wq_graph, sizeof(float) * H * H);
  memcpy(o + d.ws, wq_source, sizeof(float) * H * H);
  memcpy(o + d.wt, wq_target, sizeof(float) * H * H);
  memcpy(o + d.wproj + (size_t)0 * H * H, wk, sizeof(float) * H * H);
  memcpy(o + d.wproj + (size_t)1 * H * H, wv, sizeof(float) * H * H);
  memcpy(o + d.wproj + (size_t)2 * H * H, wq_last, sizeof(float) * H * H);
  memcpy(o + d.wproj + (size_t)3 * H * H, wq_first, sizeof(float) * H * H);
  for (int i = 0; i < H; ++i)
    for (int c = 0; c < H; ++c) o[d.wproj + (size_t)4 * H * H + (size_t)i * H + c] = wc[(size_t)c * H + i];
  memcpy(o + d.wc, wc, sizeof(float) * H * H);
  memcpy(o + d.bc, bc, sizeof(float) * H);
  // fp16 hi/lo copies of the GEMM weights for the tensor-core layers (w = hi + lo)
  {
    __half* hs = (__half*)((char*)o + half_section_byte_off(n_layers));
    auto split = [&](const float* w, size_t n, size_t hi_off, size_t lo_off) {
      for (size_t i = 0; i < n; ++i) {
        const __half h = __float2half_rn(w[i]);
        hs[hi_off + i] = h;
        hs[lo_off + i] = __float2half_rn(w[i] - __half2float(h));
      }
    };
    for (int l = 0; l < n_layers; ++l) {
      const EncLayerOff e = enc_layer_off(l);
      const EncLayerOffH eh = enc_layer_off_h(l);
      split(o + e.wqkv, (size_t)3 * H * H, eh.wqkv_hi, eh.wqkv_lo);
      split(o + e.wo, (size_t)H * H, eh.wo_hi, eh.wo_lo);
      split(o + e.w1, (size_t)FF * H, eh.w1_hi, eh.w1_lo);
      split(o + e.w2, (size_t)H * FF, eh.w2_hi, eh.w2_lo);
    }
    split(o + d.wproj, (size_t)PROJ * H, wproj_hi_off_h(n_layers), wproj_lo_off_h(n_layers));
  }
  // synchronous copy: the staging vector dies at return
  HTSP_CHECK_CUDA(cudaMemcpyAsync(dev_blob, o, total_bytes, cudaMemcpyHostToDevice,
                                  (cudaStream_t)stream));
  HTSP_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return HTSP_OK;
}

extern "C" size_t htsp_encode_workspace_bytes(int Bp, int N, int mode) {
  if (Bp <= 0 || N <= 0) return 0;
  if (mode != HTSP_MODE_FP32) return encode_tc_ws_bytes(Bp, N);
  return encode_f32_ws_floats(Bp, N) * sizeof(float);
}

extern "C" int htsp_encode(const void* blob, int n_layers, const float* x, int Bp, int N, int mode,
                           float* emb, void* workspace, size_t workspace_bytes, void* stream) {
  HTSP_REQUIRE(blob && x && emb, "null pointer");
  HTSP_REQUIRE(Bp >= 0 && N >= 1 && n_layers >= 0, "bad shape");
  HTSP_REQUIRE(mode >= HTSP_MODE_FP32 && mode <= HTSP_MODE_TC_FAST, "bad mode");
  if (Bp == 0) return HTSP_OK;
  if (workspace == nullptr || workspace_bytes < htsp_encode_workspace_bytes(Bp, N, mode)) {
    set_error("htsp_encode: workspace too small (%zu < %zu)", workspace_bytes,
              htsp_encode_workspace_bytes(Bp, N, mode));
    return HTSP_ERR_WORKSPACE;
  }
  if (mode != HTSP_MODE_FP32)
    return launch_encode_tc(blob, n_layers, x, Bp, N, emb, workspace, (cudaStream_t)stream);
  return launch_encode_f32((const float*)blob, n_layers, x, Bp, N, emb, (float*)workspace,
                           (cudaStream_t)stream);
}

extern "C" size_t htsp_encode_train_workspace_bytes(int Bp, int N, int mode) {
  if (Bp <= 0 || N <= 0 || mode == HTSP_MODE_FP32) return 0;
  return encode_tc_train_ws_bytes(Bp, N);
}

extern "C" int htsp_encode_train(const void* blob, int n_layers, const float* bn_gamma_beta, const float* x, int Bp, int N,
                                 int mode, float* emb, float* bn_batch_stats, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  HTSP_REQUIRE(blob && x && emb && bn_gamma_beta && bn_batch_stats, "null pointer");
  HTSP_REQUIRE(Bp >= 1 && N >= 1 && (size_t)Bp * N >= 2 && n_layers >= 0, "bad shape");
  if (mode != HTSP_MODE_TC_X3 && mode != HTSP_MODE_TC_FAST) {
    set_error("htsp_encode_train: train-mode BatchNorm is served by the tensor-core encoder (mode x3 / fast), got mode %d", mode);
    return HTSP_ERR_UNSUPPORTED;
  }
  if (workspace == nullptr || workspace_bytes < htsp_encode_train_workspace_bytes(Bp, N, mode)) {
    set_error("htsp_encode_train: workspace too small (%zu < %zu)", workspace_bytes,
              htsp_encode_train_workspace_bytes(Bp, N, mode));
    return HTSP_ERR_WORKSPACE;
  }
  return launch_encode_tc(blob, n_layers, x, Bp, N, emb, workspace, (cudaStream_t)stream, bn_gamma_beta, bn_batch_stats);
}

static size_t dec_proj_bytes(int Bp, int N) { return align_up((size_t)Bp * N * PROJ * 4, 256); }
static size_t dec_qfix_bytes(int Bp) { return align_up((size_t)Bp * H * 4, 256); }
static size_t dec_c0_bytes(int Bp, int N) { return align_up((size_t)Bp * N * 4, 256); }

static size_t dec_pack_bytes(int Bp, int N, int G) {
  size_t a = decode_mma_pack_bytes(Bp, N);
  if (decode_tc_supported(N, G)) a = std::max(a, decode_tc_pack_bytes(Bp, N));
  if (decode_tc2_supported(N, G)) a = std::max(a, decode_tc2_pack_bytes(Bp, N));
  return a;
}
// HTSP_DECODE_IMPL=mma forces the mma.sync rollout kernel (A/B comparisons); default is tcgen05 when it fits
static bool use_tc_kernel(int N, int G) {
  const char* e = getenv("HTSP_DECODE_IMPL");
  if (e != nullptr && strcmp(e, "mma") == 0) return false;
  return decode_tc_supported(N, G);
}
// HTSP_MODE_TC_FAST runs on the second tcgen05 kernel (decode_tc2.cu) when the shape fits; HTSP_DECODE_IMPL=tc1
// keeps the first kernel's single-fp16 variant (A/B comparisons)
static bool use_tc2_kernel(int N, int G, int mode) {
  const char* e = getenv("HTSP_DECODE_IMPL");
  if (e != nullptr && (strcmp(e, "mma") == 0 || strcmp(e, "tc1") == 0)) return false;
  return mode == HTSP_MODE_TC_FAST && decode_tc2_supported(N, G);
}

extern "C" size_t htsp_decode_workspace_bytes(int Bp, int N, int G, int mode) {
  (void)G;
  if (Bp <= 0 || N <= 0) return 0;
  size_t b = dec_proj_bytes(Bp, N) + dec_qfix_bytes(Bp) + dec_c0_bytes(Bp, N) + 256 /* index-range flag */;
  if (mode != HTSP_MODE_FP32) b += dec_pack_bytes(Bp, N, G) + decode_prep_tc_ws_bytes(Bp, N);
  return b;
}

static int decode_dispatch(const void* blob, int n_layers, const float* x, const float* emb, float* proj, float* qfix,
                           float* c0, const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp, int N, int G,
                           int mode, const int32_t* forced_pi, int32_t* pi, float* reward, float* logits_out,
                           void* pack_ws, float* dbg, int dbg_step, const uint64_t* sample_seed, float* logp,
                           cudaStream_t s);

static int decode_impl(const void* blob, int n_layers, const float* x, const float* emb,
                       const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp,
                       int N, int G, int mode, const int32_t* forced_pi, int32_t* pi,
                       float* reward, float* logits_out, void* workspace,
                       size_t workspace_bytes, float* dbg, int dbg_step, void* stream,
                       const uint64_t* sample_seed = nullptr, float* logp = nullptr) {
  HTSP_REQUIRE(blob && x && emb && src && tgt && first && pi && reward, "null pointer");
  HTSP_REQUIRE(Bp >= 0 && N >= 3 && G >= 1 && G <= N && n_layers >= 0, "bad shape");
  HTSP_REQUIRE(mode >= HTSP_MODE_FP32 && mode <= HTSP_MODE_TC_FAST, "bad mode");
  if (Bp == 0) return HTSP_OK;
  if (workspace == nullptr || workspace_bytes < htsp_decode_workspace_bytes(Bp, N, G, mode)) {
    set_error("htsp_decode: workspace too small (%zu < %zu)", workspace_bytes,
              htsp_decode_workspace_bytes(Bp, N, G, mode));
    return HTSP_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  char* w = (char*)workspace;
  float* proj = (float*)w; w += dec_proj_bytes(Bp, N);
  float* qfix = (float*)w; w += dec_qfix_bytes(Bp);
  float* c0 = (float*)w;   w += dec_c0_bytes(Bp, N);
  int32_t* idx_flag = (int32_t*)w; w += 256;
  const bool tc = mode != HTSP_MODE_FP32;
  void* pack_ws = w;
  void* prep_ws = tc ? (void*)(w + dec_pack_bytes(Bp, N, G)) : nullptr;
  // src / tgt / first / forced_pi are used as indices: the kernels clamp them, this flags the call (NaN rewards)
  int rc = launch_decode_validate(src, tgt, first, forced_pi, Bp, N, G, idx_flag, s);
  if (rc) return rc;
  rc = launch_decode_prep(blob, n_layers, emb, src, tgt, Bp, N, proj, qfix, c0, tc, prep_ws, idx_flag, s);
  if (rc) return rc;
  rc = decode_dispatch(blob, n_layers, x, emb, proj, qfix, c0, src, tgt, first, Bp, N, G, mode, forced_pi, pi, reward,
                       logits_out, pack_ws, dbg, dbg_step, sample_seed, logp, s);
  if (rc) return rc;
  if (logp != nullptr) {
    rc = launch_decode_poison(logp, (size_t)Bp * G, idx_flag, s);
    if (rc) return rc;
  }
  return launch_decode_poison(reward, (size_t)Bp * G, idx_flag, s);
}

static int decode_dispatch(const void* blob, int n_layers, const float* x, const float* emb, float* proj, float* qfix,
                           float* c0, const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp, int N, int G,
                           int mode, const int32_t* forced_pi, int32_t* pi, float* reward, float* logits_out,
                           void* pack_ws, float* dbg, int dbg_step, const uint64_t* sample_seed, float* logp,
                           cudaStream_t s) {
  if (logp != nullptr && !use_tc2_kernel(N, G, mode)) {
    set_error("htsp_decode_sample_logp: log-probabilities are returned by the second tcgen05 kernel only (mode fast, "
              "33 <= N <= 208 and G <= 256, or 208 < N <= 576 in key blocks); got mode %d, N %d, G %d", mode, N, G);
    return HTSP_ERR_UNSUPPORTED;
  }
  if (sample_seed != nullptr && (mode == HTSP_MODE_FP32 || !(use_tc_kernel(N, G) || use_tc2_kernel(N, G, mode)))) {
    set_error("htsp_decode_sample: sampling rollouts are served by the tcgen05 kernels only (tensor-core mode, "
              "17 <= N <= 208, G <= 256; mode fast: up to N = 576 in key blocks); got mode %d, N %d, G %d", mode, N, G);
    return HTSP_ERR_UNSUPPORTED;
  }
  if (mode == HTSP_MODE_FP32)
    return launch_decode_f32((const float*)blob, n_layers, x, emb, proj, qfix, src, tgt, first, Bp,
                             N, G, forced_pi, pi, reward, logits_out, s);
  if (use_tc2_kernel(N, G, mode))
    return launch_decode_tc2(x, proj, qfix, c0, src, tgt, first, Bp, N, G, forced_pi, pi, reward, logits_out, pack_ws,
                             dbg, dbg_step, sample_seed, logp, s);
  if (use_tc_kernel(N, G))
    return launch_decode_tc(x, proj, qfix, c0, src, tgt, first, Bp, N, G, mode, forced_pi, pi, reward,
                            logits_out, pack_ws, dbg, dbg_step, sample_seed, s);
  HTSP_REQUIRE(dbg == nullptr, "debug dump needs the tcgen05 kernel");
  return launch_decode_mma(x, proj, qfix, c0, src, tgt, first, Bp, N, G, mode, forced_pi, pi, reward,
                           logits_out, pack_ws, s);
}

extern "C" int htsp_decode(const void* blob, int n_layers, const float* x, const float* emb,
                           const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp,
                           int N, int G, int mode, const int32_t* forced_pi, int32_t* pi,
                           float* reward, float* logits_out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  return decode_impl(blob, n_layers, x, emb, src, tgt, first, Bp, N, G, mode, forced_pi, pi, reward,
                     logits_out, workspace, workspace_bytes, nullptr, 0, stream);
}

extern "C" int htsp_decode_sample(const void* blob, int n_layers, const float* x, const float* emb,
                                  const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp,
                                  int N, int G, int mode, uint64_t seed, int32_t* pi, float* reward,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  return decode_impl(blob, n_layers, x, emb, src, tgt, first, Bp, N, G, mode, nullptr, pi, reward, nullptr,
                     workspace, workspace_bytes, nullptr, 0, stream, &seed);
}

extern "C" int htsp_decode_sample_logp(const void* blob, int n_layers, const float* x, const float* emb,
                                       const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp,
                                       int N, int G, int mode, uint64_t seed, int32_t* pi, float* reward, float* logp,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  HTSP_REQUIRE(logp != nullptr, "null pointer");
  return decode_impl(blob, n_layers, x, emb, src, tgt, first, Bp, N, G, mode, nullptr, pi, reward, nullptr,
                     workspace, workspace_bytes, nullptr, 0, stream, &seed, logp);
}

extern "C" int htsp_decode_ctas_per_sm(int Bp, int N, int G, int mode) {
  if (Bp <= 0 || N < 3 || G < 1) return 0;
  return use_tc2_kernel(N, G, mode) ? decode_tc2_ctas_per_sm(Bp, N, G) : 1;
}
extern "C" int htsp_decode_key_blocks(int N, int32_t* out5) {
  HTSP_REQUIRE(out5 != nullptr, "null pointer");
  if (!decode_tc2_geometry(N, out5)) {
    set_error("htsp_decode_key_blocks: N=%d is not served by the second tcgen05 kernel (33 <= N <= 576)", N);
    return HTSP_ERR_UNSUPPORTED;
  }
  return HTSP_OK;
}
extern "C" int htsp_decode_key_block_stage(int N, int i, uint32_t* out2) {
  HTSP_REQUIRE(out2 != nullptr, "null pointer");
  if (!decode_tc2_stage(N, i, out2)) {
    set_error("htsp_decode_key_block_stage: no stage %d for N=%d", i, N);
    return HTSP_ERR_UNSUPPORTED;
  }
  return HTSP_OK;
}
extern "C" size_t htsp_decode_debug_floats(int N) { return std::max(decode_tc_dbg_floats(N), decode_tc2_dbg_floats(N)); }
extern "C" int htsp_decode_debug(const void* blob, int n_layers, const float* x, const float* emb,
                                 const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp,
                                 int N, int G, int mode, const int32_t* forced_pi, int32_t* pi,
                                 float* reward, void* workspace, size_t workspace_bytes, float* dbg,
                                 int dbg_step, void* stream) {
  HTSP_REQUIRE(dbg != nullptr && mode != HTSP_MODE_FP32, "debug dump: tensor-core modes only");
  return decode_impl(blob, n_layers, x, emb, src, tgt, first, Bp, N, G, mode, forced_pi, pi, reward,
                     nullptr, workspace, workspace_bytes, dbg, dbg_step, stream);
}
