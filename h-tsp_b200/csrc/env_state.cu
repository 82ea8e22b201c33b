// K6: device-resident state of the large-TSP environment, batched
// over E environments, one CTA per environment:
//   htsp_env_splice  -- LargeState.move_to (h_tsp.py:288-320): splice the solved path into the
//                       current tour (both branches), rebuild the city->position index, the
//                       reference's asserts as status codes
//   htsp_env_update  -- rl_utils.update_neighbor_coord_ (rl_utils.py:135-159), LargeState.mask
//                       (h_tsp.py:330-340), get_tour_distance over fp32(fp64 cdist x 1000)
//                       (rl_utils.py:87-102, h_tsp.py:449-453,586-589) WITHOUT the Ntot x Ntot
//                       matrix, the reward of Env._step (h_tsp.py:599) and LargeState.to_tensor
//                       (h_tsp.py:396-408)
// All of it is HBM-bound integer/byte work: coalesced grid-stride loops, no tensor cores.
// Algorithmic bytes per environment step: splice T*(8+8) + T*4*2 (index), update
// T*(8+8+16) (neighbour rows) + A*k*(8+1) (A available cities, k-NN lists) + T*(8+8) (length)
// + Ntot*(8+1+1+16+32) (state tensor).
#include "common.cuh"

namespace htsp {

constexpr int ENV_THREADS = 1024;

enum { ENV_OK = 0, ENV_ERR_ENDPOINT = 1, ENV_ERR_SAME_ENDPOINT = 2, ENV_ERR_DUPLICATE = 3,
       ENV_ERR_NO_GROWTH = 4, ENV_ERR_RANGE = 5 };

__global__ void __launch_bounds__(ENV_THREADS)
env_splice_kernel(const long long* __restrict__ new_paths, const int32_t* __restrict__ path_len, int P,
                  int Ntot, const long long* __restrict__ tour_in, long long* __restrict__ tour_out,
                  int32_t* __restrict__ tour_len, int32_t* __restrict__ pos, int32_t* __restrict__ status) {
  const int e = blockIdx.x, tid = threadIdx.x;
  const long long* path = new_paths + (size_t)e * P;
  const long long* tin = tour_in + (size_t)e * Ntot;
  long long* tout = tour_out + (size_t)e * Ntot;
  int32_t* ps = pos + (size_t)e * Ntot;
  const int T = tour_len[e], L = path_len[e];
  __shared__ int s_bad;
  if (L == 0) {                       // inactive (finished) environment: the tour carries over unchanged
    for (int i = tid; i < T; i += ENV_THREADS) tout[i] = tin[i];
    if (tid == 0) status[e] = ENV_OK;
    return;
  }
  if (tid == 0) s_bad = ENV_OK;
  __syncthreads();
  // range check of the path (everything below indexes with it)
  for (int i = tid; i < L; i += ENV_THREADS)
    if (path[i] < 0 || path[i] >= Ntot) s_bad = ENV_ERR_RANGE;
  __syncthreads();
  // a failed splice leaves the environment's tour as it was (the caller flips the double buffer for all environments)
  auto keep_old_tour = [&]() {
    for (int i = tid; i < T; i += ENV_THREADS) tout[i] = tin[i];
  };
  if (s_bad != ENV_OK || L < 1 || L > P) {
    keep_old_tour();
    if (tid == 0) status[e] = ENV_ERR_RANGE;
    return;
  }
  int head_n, head_off, tail_n, tail_off;      // out = in[head_off : +head_n] + path + in[tail_off : +tail_n]
  if (T == 0) {
    head_n = head_off = tail_n = tail_off = 0;
  } else {
    const int s = ps[path[0]], en = ps[path[L - 1]];
    if (s < 0 || en < 0 || s == en) {           // h_tsp.py:293-295
      keep_old_tour();
      if (tid == 0) status[e] = (s == en && s >= 0) ? ENV_ERR_SAME_ENDPOINT : ENV_ERR_ENDPOINT;
      return;
    }
    if (en > s) { head_off = 0; head_n = s; tail_off = en + 1; tail_n = T - en - 1; }      // :298-303
    else { head_off = en + 1; head_n = s - en - 1; tail_off = 0; tail_n = 0; }             // :304-307
  }
  const int newT = head_n + L + tail_n;
  if (newT > Ntot) {
    keep_old_tour();
    if (tid == 0) status[e] = ENV_ERR_DUPLICATE;
    return;
  }
  for (int i = tid; i < head_n; i += ENV_THREADS) tout[i] = tin[head_off + i];
  for (int i = tid; i < L; i += ENV_THREADS) tout[head_n + i] = path[i];
  for (int i = tid; i < tail_n; i += ENV_THREADS) tout[head_n + L + i] = tin[tail_off + i];
  // rebuild the city -> position index; a city that lands twice is the reference's duplicate assert (:309-311)
  for (int i = tid; i < T; i += ENV_THREADS) ps[tin[i]] = -1;
  __syncthreads();
  for (int i = tid; i < newT; i += ENV_THREADS)
    if (atomicExch(&ps[tout[i]], i) != -1) s_bad = ENV_ERR_DUPLICATE;
  __syncthreads();
  if (tid == 0) {
    int st = s_bad;
    if (st == ENV_OK && newT < Ntot && newT <= T) st = ENV_ERR_NO_GROWTH;                  // :313-320
    status[e] = st;
    tour_len[e] = newT;
  }
}

__global__ void __launch_bounds__(ENV_THREADS)
env_update_kernel(const float2* __restrict__ x, const long long* __restrict__ tour,
                  const int32_t* __restrict__ tour_len, const long long* __restrict__ new_paths,
                  const int32_t* __restrict__ path_len, int P, const long long* __restrict__ knn, int k,
                  int Ntot, unsigned char* __restrict__ selected, unsigned char* __restrict__ available,
                  float4* __restrict__ neighbor, float* __restrict__ state, float* __restrict__ cur_len,
                  float* __restrict__ reward) {
  const int e = blockIdx.x, tid = threadIdx.x;
  const float2* xy = x + (size_t)e * Ntot;
  const long long* t = tour + (size_t)e * Ntot;
  const long long* path = new_paths + (size_t)e * P;
  const long long* nn = knn + (size_t)e * Ntot * k;
  unsigned char* sel = selected + (size_t)e * Ntot;
  unsigned char* av = available + (size_t)e * Ntot;
  float4* nb = neighbor + (size_t)e * Ntot;
  const int T = tour_len[e], L = path_len[e];
  // neighbour coordinates of every city of the tour: [x[next], x[previous]] (rl_utils.py:143-159);
  // the closed-tour length rides along: fp32( |1000 x_i - 1000 x_{i+1}| in fp64 ), fp32 sum
  float acc = 0.f;
  for (int i = tid; i < T; i += ENV_THREADS) {
    const long long c = t[i], nx = t[i + 1 == T ? 0 : i + 1], pv = t[i == 0 ? T - 1 : i - 1];
    const float2 pc = xy[c], pn = xy[nx], pp = xy[pv];
    nb[c] = make_float4(pn.x, pn.y, pp.x, pp.y);
    const double dx = (double)pc.x * 1000.0 - (double)pn.x * 1000.0;
    const double dy = (double)pc.y * 1000.0 - (double)pn.y * 1000.0;
    acc += (float)sqrt(dx * dx + dy * dy);
  }
  // masks (h_tsp.py:330-340): mark the path, then every available city stays available iff not all
  // of its k nearest neighbours are selected
  for (int i = tid; i < L; i += ENV_THREADS) { sel[path[i]] = 1; av[path[i]] = 1; }
  __syncthreads();
  for (int c = tid; c < Ntot; c += ENV_THREADS) {
    if (av[c]) {
      bool all_sel = true;
      for (int j = 0; j < k; ++j) all_sel = all_sel && (sel[nn[(size_t)c * k + j]] != 0);
      av[c] = all_sel ? 0 : 1;
    }
  }
  __shared__ float red[ENV_THREADS / 32];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((tid & 31) == 0) red[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < ENV_THREADS / 32; ++w) s += red[w];
    const float new_len = T > 0 ? s / 1000.0f : 0.f;
    reward[e] = cur_len[e] - new_len;            // h_tsp.py:599
    cur_len[e] = new_len;
  }
  // LargeState.to_tensor (h_tsp.py:396-408): [x, selected, available, neighbour coords]
  if (state != nullptr) {
    float4* sp = (float4*)(state + (size_t)e * Ntot * 8);
    for (int c = tid; c < Ntot; c += ENV_THREADS) {
      const float2 pc = xy[c];
      sp[2 * c] = make_float4(pc.x, pc.y, sel[c] ? 1.f : 0.f, av[c] ? 1.f : 0.f);
      sp[2 * c + 1] = nb[c];
    }
  }
}

}  // namespace htsp

using namespace htsp;

extern "C" int htsp_env_splice(const int64_t* new_paths, const int32_t* path_len, int E, int P, int Ntot,
                               const int64_t* tour_in, int64_t* tour_out, int32_t* tour_len,
                               int32_t* pos, int32_t* status, void* stream) {
  HTSP_REQUIRE(new_paths && path_len && tour_in && tour_out && tour_len && pos && status, "null pointer");
  HTSP_REQUIRE(E >= 0 && P >= 1 && Ntot >= 1 && tour_in != tour_out, "bad shape");
  if (E == 0) return HTSP_OK;
  env_splice_kernel<<<E, ENV_THREADS, 0, (cudaStream_t)stream>>>((const long long*)new_paths, path_len, P, Ntot,
                                                                 (const long long*)tour_in, (long long*)tour_out,
                                                                 tour_len, pos, status);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

extern "C" int htsp_env_update(const float* x, const int64_t* tour, const int32_t* tour_len,
                               const int64_t* new_paths, const int32_t* path_len, const int64_t* knn, int k,
                               int E, int P, int Ntot, uint8_t* selected, uint8_t* available,
                               float* neighbor_coord, float* state, float* cur_len, float* reward,
                               void* stream) {
  HTSP_REQUIRE(x && tour && tour_len && new_paths && path_len && knn && selected && available &&
               neighbor_coord && cur_len && reward, "null pointer");
  HTSP_REQUIRE(E >= 0 && P >= 1 && Ntot >= 1 && k >= 1, "bad shape");
  if (E == 0) return HTSP_OK;
  env_update_kernel<<<E, ENV_THREADS, 0, (cudaStream_t)stream>>>(
      (const float2*)x, (const long long*)tour, tour_len, (const long long*)new_paths, path_len, P,
      (const long long*)knn, k, Ntot, selected, available, (float4*)neighbor_coord, state, cur_len, reward);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}
