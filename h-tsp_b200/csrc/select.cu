// K4: best-of-(A*G) selection (rl4cop/train_path_solver.py:290-306) and the path orientation /
// global-id mapping of the hook (h_tsp.py:206-226); K5: closed tour length from coordinates
// (rl_utils.py:87-102 over dist_matrix of h_tsp.py:449-453) without the Ntot x Ntot matrix.
// All three are HBM/latency-bound index kernels: one warp per subproblem (selection), one CTA per
// subproblem (orientation) / tour (length).
#include "common.cuh"

namespace htsp {

// reward [A*B, G] (instance a*B+b), pi [A*B, G, N] i32.  torch semantics: max over g inside each
// augmentation block, then max over blocks, first maximum wins both times == the smallest
// linear index a*G+g among the global maxima.
// One WARP per subproblem (8 per CTA): every lane scans its share of the A*G rewards (coalesced, all loads
// in flight), one shuffle reduction with the first-index rule, then the 32 lanes copy the winning tour.
__global__ void __launch_bounds__(256)
select_best_kernel(const float* __restrict__ reward, const int32_t* __restrict__ pi, int A, int B,
                   int G, int N, float* __restrict__ best_len, int32_t* __restrict__ best_row,
                   long long* __restrict__ best_pi) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int a = 0; a < A; ++a) {
    const float* rp = reward + ((size_t)a * B + b) * G;
#pragma unroll 4
    for (int g = lane; g < G; g += 32) {
      const float r = __ldg(&rp[g]);
      if (r > best || bi == 0x7fffffff) { best = r; bi = a * G + g; }   // ascending index: first maximum stays
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, off);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
    if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > best || (ob == best && oi < bi))) { best = ob; bi = oi; }
  }
  const int a = bi / G, g = bi - a * G;
  if (lane == 0) { best_len[b] = -best; best_row[b] = bi; }
  const int32_t* srcp = pi + (((size_t)a * B + b) * G + g) * N;
#pragma unroll 4
  for (int i = lane; i < N; i += 32) best_pi[(size_t)b * N + i] = (long long)__ldg(&srcp[i]);
}

// h_tsp.py:206-226.  status bits: 1 = local node 0 missing, 2 = path does not end at N-1 after
// orientation, 4 = not a permutation of 0..N-1.
__global__ void __launch_bounds__(128)
finalize_paths_kernel(const long long* __restrict__ best_pi, const long long* __restrict__ fragment,
                      const float* __restrict__ best_len, const float* __restrict__ scale, int N,
                      long long* __restrict__ out_paths, float* __restrict__ out_len,
                      int32_t* __restrict__ status) {
  extern __shared__ unsigned seen[];
  const int b = blockIdx.x, tid = threadIdx.x;
  __shared__ int zero_pos, bad;
  const long long* p = best_pi + (size_t)b * N;
  if (tid == 0) { zero_pos = -1; bad = 0; }
  for (int i = tid; i < (N + 31) / 32; i += blockDim.x) seen[i] = 0u;
  __syncthreads();
  for (int i = tid; i < N; i += blockDim.x) {
    const long long v = __ldg(&p[i]);
    if (v == 0) zero_pos = i;
    if (v < 0 || v >= N) atomicOr(&bad, 4);
    else if (atomicOr(&seen[v >> 5], 1u << (v & 31)) & (1u << (v & 31))) atomicOr(&bad, 4);
  }
  __syncthreads();
  const int z = zero_pos;
  if (z < 0) {
    if (tid == 0) { status[b] = bad | 1; out_len[b] = best_len[b] / scale[b]; }
    return;
  }
  // rotated path r[i] = p[(z+i)%N]; if r[N-1] != N-1 use flip(roll(r,-1)) : r'[i] = r[(N-i)%N]
  const bool flip = __ldg(&p[(z + N - 1) % N]) != (long long)(N - 1);
  for (int i = tid; i < N; i += blockDim.x) {
    const int k = flip ? (N - i) % N : i;
    const long long local = __ldg(&p[(z + k) % N]);
    if (i == N - 1 && local != (long long)(N - 1)) atomicOr(&bad, 2);
    const long long lc = local < 0 ? 0 : (local >= N ? N - 1 : local);
    out_paths[(size_t)b * N + i] = __ldg(&fragment[(size_t)b * N + lc]);
  }
  __syncthreads();
  if (tid == 0) { status[b] = bad; out_len[b] = best_len[b] / scale[b]; }
}

// closed tour length: sum_i fp32( |1000 x[t_i] - 1000 x[t_{i+1}]|_fp64 ), fp32 sum, / 1000.
// One CTA of 1024 threads per tour: a thread gathers ONE coordinate per edge (the edge's other end comes from
// the next lane by shuffle; lane 31 fetches it) and keeps four edges in flight.
constexpr int kTourThreads = 1024;
__global__ void __launch_bounds__(kTourThreads)
tour_length_kernel(const float2* __restrict__ coords, const long long* __restrict__ tours,
                   const int32_t* __restrict__ tour_len, int Ntot, int Tmax,
                   float* __restrict__ out) {
  const int e = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int T = tour_len[e];
  const float2* xy = coords + (size_t)e * Ntot;
  const long long* t = tours + (size_t)e * Tmax;
  float acc = 0.f;
  for (int base = 0; base < T; base += 4 * kTourThreads) {
    float2 pa[4], pn[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * kTourThreads + tid;
      pa[u] = make_float2(0.f, 0.f);
      if (i < T) pa[u] = __ldg(&xy[__ldg(&t[i])]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * kTourThreads + tid;
      pn[u].x = __shfl_down_sync(0xffffffffu, pa[u].x, 1);
      pn[u].y = __shfl_down_sync(0xffffffffu, pa[u].y, 1);
      if (i < T && (lane == 31 || i + 1 == T)) pn[u] = __ldg(&xy[__ldg(&t[i + 1 == T ? 0 : i + 1])]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * kTourThreads + tid;
      if (i < T) {
        const double dx = (double)pa[u].x * 1000.0 - (double)pn[u].x * 1000.0;
        const double dy = (double)pa[u].y * 1000.0 - (double)pn[u].y * 1000.0;
        acc += (float)sqrt(dx * dx + dy * dy);
      }
    }
  }
  __shared__ float red[kTourThreads / 32];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) red[tid >> 5] = acc;
  __syncthreads();
  if (tid < 32) {
    float s = red[tid];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (tid == 0) out[e] = T > 0 ? s / 1000.0f : 0.f;
  }
}

}  // namespace htsp

using namespace htsp;

extern "C" int htsp_select_best(const float* reward, const int32_t* pi, int n_aug, int B, int G,
                                int N, float* best_len, int32_t* best_row, int64_t* best_pi,
                                void* stream) {
  HTSP_REQUIRE(reward && pi && best_len && best_row && best_pi, "null pointer");
  HTSP_REQUIRE((n_aug == 1 || n_aug == 8) && B >= 0 && G >= 1 && N >= 1, "bad shape");
  if (B == 0) return HTSP_OK;
  select_best_kernel<<<(B + 7) / 8, 256, 0, (cudaStream_t)stream>>>(reward, pi, n_aug, B, G, N, best_len,
                                                          best_row, (long long*)best_pi);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

extern "C" int htsp_finalize_paths(const int64_t* best_pi, const int64_t* fragment,
                                   const float* best_len, const float* scale, int B, int N,
                                   int64_t* out_paths, float* out_len, int32_t* status,
                                   void* stream) {
  HTSP_REQUIRE(best_pi && fragment && best_len && scale && out_paths && out_len && status, "null pointer");
  HTSP_REQUIRE(B >= 0 && N >= 2, "bad shape");
  if (B == 0) return HTSP_OK;
  size_t smem = (size_t)((N + 31) / 32) * 4;
  finalize_paths_kernel<<<B, 128, smem, (cudaStream_t)stream>>>(
      (const long long*)best_pi, (const long long*)fragment, best_len, scale, N,
      (long long*)out_paths, out_len, status);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

extern "C" int htsp_tour_length(const float* coords, const int64_t* tours, const int32_t* tour_len,
                                int E, int Ntot, int Tmax, float* out, void* stream) {
  HTSP_REQUIRE(coords && tours && tour_len && out, "null pointer");
  HTSP_REQUIRE(E >= 0 && Ntot >= 1 && Tmax >= 1, "bad shape");
  if (E == 0) return HTSP_OK;
  tour_length_kernel<<<E, kTourThreads, 0, (cudaStream_t)stream>>>((const float2*)coords,
                                                          (const long long*)tours, tour_len, Ntot,
                                                          Tmax, out);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}
