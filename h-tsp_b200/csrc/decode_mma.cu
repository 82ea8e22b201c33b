#include "common.cuh"
namespace htsp {
size_t decode_mma_pack_bytes(int Bp, int N) { return 0; }
int launch_decode_mma(const float* x, const float* proj, const float* qfix, const float* c0,
                      const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp, int N,
                      int G, int mode, const int32_t* forced_pi, int32_t* pi, float* reward,
                      float* logits_out, void* pack_ws, cudaStream_t s) {
  set_error("tensor-core decode not built yet");
  return HTSP_ERR_UNSUPPORTED;
}
}  // namespace htsp
