// Train-mode BatchNorm1d of the encoder (rl4cop/models/models.py:35,37 with the module in train() mode, as
// PathSolver.training_step runs it, rl4cop/train_path_solver.py:352-375): the statistics are those of the batch --
// mean and biased variance of every channel over all M = B * N rows -- instead of the running ones, so the affine
// cannot be folded into a GEMM epilogue.  The GEMM writes z = x + sublayer(x) (fp32), then
//   bn_stats_partial_kernel  per-channel sum and sum of squares of a row block, fp64, fixed order
//   bn_stats_final_kernel    partials -> mean, biased variance (what torch normalises with; the caller turns the
//                            variance into the unbiased one for the running-statistics update)
//   bn_apply_kernel          y = (z - mean) / sqrt(var + 1e-5) * gamma + beta  -> fp32 row + the fp16 hi/lo split the
//                            next GEMM's TMA loads consume
// HBM-bound passes over [M,128]; deterministic (no atomics): the same batch gives the same bits.
#include "common.cuh"
#include "mma_util.cuh"

namespace htsp {

constexpr int BN_ROWS_PER_BLOCK = 256;

__global__ void __launch_bounds__(256)
bn_stats_partial_kernel(const float* __restrict__ z, size_t M, double* __restrict__ partial) {
  __shared__ double sh[2][2][H];
  const int c = threadIdx.x & (H - 1), half = threadIdx.x >> 7;        // two row streams per block
  const size_t r0 = (size_t)blockIdx.x * BN_ROWS_PER_BLOCK;
  const size_t r1 = r0 + BN_ROWS_PER_BLOCK < M ? r0 + BN_ROWS_PER_BLOCK : M;
  double s = 0.0, ss = 0.0;
  for (size_t r = r0 + half; r < r1; r += 2) {
    const double v = (double)__ldg(z + r * H + c);
    s += v;
    ss += v * v;
  }
  sh[half][0][c] = s;
  sh[half][1][c] = ss;
  __syncthreads();
  if (half == 0) {
    partial[((size_t)blockIdx.x * 2 + 0) * H + c] = sh[0][0][c] + sh[1][0][c];
    partial[((size_t)blockIdx.x * 2 + 1) * H + c] = sh[0][1][c] + sh[1][1][c];
  }
}

__global__ void __launch_bounds__(H)
bn_stats_final_kernel(const double* __restrict__ partial, int nb, size_t M, float* __restrict__ stats) {
  const int c = threadIdx.x;
  double s = 0.0, ss = 0.0;
  for (int b = 0; b < nb; ++b) {
    s += partial[((size_t)b * 2 + 0) * H + c];
    ss += partial[((size_t)b * 2 + 1) * H + c];
  }
  const double mean = s / (double)M;
  double var = ss / (double)M - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[c] = (float)mean;
  stats[H + c] = (float)var;
}

__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ z, const float* __restrict__ stats, const float* __restrict__ gamma,
                const float* __restrict__ beta, size_t n4, float* __restrict__ y32, __half* __restrict__ yhi,
                __half* __restrict__ ylo) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i & (H / 4 - 1)) * 4;
    const float4 v = __ldg((const float4*)z + i);
    const float4 mu = __ldg((const float4*)(stats + c)), va = __ldg((const float4*)(stats + H + c));
    const float4 g = __ldg((const float4*)(gamma + c)), b = __ldg((const float4*)(beta + c));
    // torch: alpha = invstd * weight, y = x * alpha + (bias - mean * alpha)
    const float a0 = g.x / sqrtf(va.x + 1e-5f), a1 = g.y / sqrtf(va.y + 1e-5f);
    const float a2 = g.z / sqrtf(va.z + 1e-5f), a3 = g.w / sqrtf(va.w + 1e-5f);
    const float y0 = v.x * a0 + (b.x - mu.x * a0), y1 = v.y * a1 + (b.y - mu.y * a1);
    const float y2 = v.z * a2 + (b.z - mu.z * a2), y3 = v.w * a3 + (b.w - mu.w * a3);
    ((float4*)y32)[i] = make_float4(y0, y1, y2, y3);
    uint32_t h0, l0, h1, l1;
    split2p(y0, y1, h0, l0);
    split2p(y2, y3, h1, l1);
    ((uint2*)yhi)[i] = make_uint2(h0, h1);
    ((uint2*)ylo)[i] = make_uint2(l0, l1);
  }
}

size_t bn_train_ws_bytes(size_t M) {
  const size_t nb = (M + BN_ROWS_PER_BLOCK - 1) / BN_ROWS_PER_BLOCK;
  return align_up(nb * 2 * H * sizeof(double), 256);
}

// z [M,128] -> y32 / yhi / ylo [M,128]; stats [2][128] receives the batch mean and biased variance
int launch_bn_train(const float* z, size_t M, const float* gamma, const float* beta, float* stats, float* y32,
                    __half* yhi, __half* ylo, void* ws, cudaStream_t s) {
  HTSP_REQUIRE(z && gamma && beta && stats && y32 && yhi && ylo && ws && M > 1, "bn_train arguments");
  const int nb = (int)((M + BN_ROWS_PER_BLOCK - 1) / BN_ROWS_PER_BLOCK);
  double* partial = (double*)ws;
  bn_stats_partial_kernel<<<nb, 256, 0, s>>>(z, M, partial);
  HTSP_CHECK_LAUNCH();
  bn_stats_final_kernel<<<1, H, 0, s>>>(partial, nb, M, stats);
  HTSP_CHECK_LAUNCH();
  const size_t n4 = M * H / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  bn_apply_kernel<<<blocks, 256, 0, s>>>(z, stats, gamma, beta, n4, y32, yhi, ylo);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

}  // namespace htsp
