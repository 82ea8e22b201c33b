// K7: fragment selection of the large-TSP environment on the device:
//   htsp_env_knn       -- Env.reset (h_tsp.py:443-455): k nearest neighbours of every city under
//                         fp32(fp64 euclid x 1000) with an infinite diagonal, ascending, WITHOUT the
//                         Ntot x Ntot matrix (400 MB fp32 + 800 MB fp64 per TSP-10000 instance there)
//   htsp_env_fragment  -- Env.get_fragment_knn (h_tsp.py:482-542): rl_utils.get_nearest_city_idx x 3
//                         (rl_utils.py:46-62), the FIFO expansion over the k-NN lists (numba
//                         append_selected, h_tsp.py:472-480, 512-524) and Env._extend_fragment
//                         (h_tsp.py:638-661), one CTA per environment, no host round trips (the
//                         reference does three .item() syncs and a mask D2H per environment and step)
// Integer / byte work, HBM- and latency-bound; results are exact integers (city ids).
#include "common.cuh"

namespace htsp {

constexpr int FRAG_THREADS = 256;

// lexicographic (distance, city) minimum: "first index among equal distances" of torch.argmin / topk
struct DistIdx {
  float d;
  int i;
};
__device__ __forceinline__ DistIdx di_min(DistIdx a, DistIdx b) {
  return (b.d < a.d || (b.d == a.d && b.i < a.i)) ? b : a;
}
// warp arg-min by two redux instructions: the distances are non-negative floats (or +inf, or the NaN 0x7fc00000 of a
// retired entry), whose bit patterns order like the numbers (NaN last)
__device__ __forceinline__ DistIdx warp_min(unsigned dbits, int i) {
  const unsigned m = __reduce_min_sync(0xffffffffu, dbits);
  const unsigned mi = __reduce_min_sync(0xffffffffu, dbits == m ? (unsigned)i : 0xffffffffu);
  return DistIdx{__uint_as_float(m), (int)mi};
}
__device__ __forceinline__ DistIdx block_min(DistIdx v, DistIdx* red) {
  v = warp_min(__float_as_uint(v.d), v.i);
  const int tid = threadIdx.x, lane = tid & 31;
  __syncthreads();                       // red may still be read from a previous call
  if (lane == 0) red[tid >> 5] = v;
  __syncthreads();
  const bool have = lane < (int)(blockDim.x >> 5);
  const DistIdx r = have ? red[lane] : DistIdx{0.f, 0x7fffffff};
  return warp_min(have ? __float_as_uint(r.d) : 0xffffffffu, r.i);
}
// float32( |a - b| in float64 ) with a scale applied in float64 first (x1000 for the k-NN matrix, h_tsp.py:449-452)
__device__ __forceinline__ float dist32(float2 a, double bx, double by, double scale) {
  const double dx = (double)a.x * scale - bx, dy = (double)a.y * scale - by;
  return (float)sqrt(dx * dx + dy * dy);
}

// ---- k-NN: one CTA per (environment, city); distances of the row live in shared memory -------------
__global__ void __launch_bounds__(FRAG_THREADS)
env_knn_kernel(const float2* __restrict__ x, int Ntot, int k, long long* __restrict__ knn) {
  extern __shared__ float drow[];        // [Ntot]
  __shared__ DistIdx red[FRAG_THREADS / 32];
  const int e = blockIdx.y, c = blockIdx.x, tid = threadIdx.x;
  const float2* xy = x + (size_t)e * Ntot;
  const float2 pc = xy[c];
  const double bx = (double)pc.x * 1000.0, by = (double)pc.y * 1000.0;
  // every thread owns the entries tid, tid + 256, ... of the row and is the only one that reads or retires them
  // (no barrier guards drow); it keeps the minimum of its entries in registers and rescans them only after one of
  // them has won a round -- a round costs the other threads the block arg-min alone
  auto scan_mine = [&]() {
    DistIdx m = {INFINITY, 0x7fffffff};
    for (int j = tid; j < Ntot; j += FRAG_THREADS) m = di_min(m, DistIdx{drow[j], j});
    return m;
  };
  for (int j = tid; j < Ntot; j += FRAG_THREADS) drow[j] = (j == c) ? INFINITY : dist32(xy[j], bx, by, 1000.0);
  DistIdx mine = scan_mine();
  long long* out = knn + ((size_t)e * Ntot + c) * k;
  for (int r = 0; r < k; ++r) {          // k selection rounds (k = 40): block arg-min, then retire the winner
    const DistIdx best = block_min(mine, red);
    if (tid == 0) out[r] = best.i;
    if ((unsigned)best.i < (unsigned)Ntot && best.i % FRAG_THREADS == tid) {
      // a retired entry must lose against every remaining one, including other infinities (the diagonal)
      drow[best.i] = __int_as_float(0x7fc00000);     // NaN: never '<' and never '==' anything
      mine = scan_mine();
    }
  }
}

struct FragArgs {
  const float2* x;                 // [E,Ntot]
  const unsigned char* selected;   // [E,Ntot]
  const long long* knn;            // [E,Ntot,k]
  const long long* tour;           // [E,Ntot]
  const int32_t* tour_len;         // [E]
  const int32_t* pos;              // [E,Ntot]
  const float2* action;            // [E] scaled to [0,1]^2 (Env.scale_action applied)
  const unsigned char* active;     // [E] or null (all active)
  long long* fragment;             // [E,frag_len]
  long long* new_cities;           // [E,frag_len]  (-1 padded)
  int32_t* n_new;                  // [E]
  int32_t* start_city;             // [E]  (-1 when the tour is empty)
  int32_t* status;                 // [E]
  int Ntot, k, frag_len, max_new_nodes, qcap;
};

// nearest city to (px,py) among cities whose selected flag equals `want`
__device__ __forceinline__ int nearest_city(const float2* xy, const unsigned char* sel, int Ntot, float2 p,
                                            unsigned char want, DistIdx* red) {
  DistIdx best = {INFINITY, 0x7fffffff};
  const double bx = (double)p.x, by = (double)p.y;
  for (int j = threadIdx.x; j < Ntot; j += FRAG_THREADS)
    if ((sel[j] != 0) == (want != 0)) best = di_min(best, DistIdx{dist32(xy[j], bx, by, 1.0), j});
  return block_min(best, red).i;
}

__global__ void __launch_bounds__(FRAG_THREADS) env_fragment_kernel(const FragArgs A) {
  extern __shared__ unsigned frag_smem[];
  __shared__ DistIdx red[FRAG_THREADS / 32];
  __shared__ int s_nnew;
  const int e = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int Ntot = A.Ntot, k = A.k, FL = A.frag_len;
  if (A.active != nullptr && !A.active[e]) {
    if (tid == 0) { A.n_new[e] = 0; A.start_city[e] = -1; A.status[e] = 0; }
    return;
  }
  const float2* xy = A.x + (size_t)e * Ntot;
  const unsigned char* sel = A.selected + (size_t)e * Ntot;
  const long long* nn = A.knn + (size_t)e * Ntot * k;
  const long long* tour = A.tour + (size_t)e * Ntot;
  const int32_t* pos = A.pos + (size_t)e * Ntot;
  long long* frag = A.fragment + (size_t)e * FL;
  long long* newc = A.new_cities + (size_t)e * FL;
  const int T = A.tour_len[e];
  const int nwords = (Ntot + 31) / 32;
  unsigned* seen = frag_smem;                    // [nwords] bitmap of the FIFO's selected_set
  int* queue = (int*)(frag_smem + nwords);       // [qcap]
  // the bitmap starts as the selected flags: "not yet seen and not selected" (h_tsp.py:472-480) is one bit test
  for (int i0 = (tid >> 5) * 32; i0 < Ntot; i0 += FRAG_THREADS) {
    const int i = i0 + lane;
    const unsigned m = __ballot_sync(0xffffffffu, i < Ntot && sel[i] != 0);
    if (lane == 0) seen[i0 >> 5] = m;
  }
  for (int i = tid; i < FL; i += FRAG_THREADS) newc[i] = -1;
  int start_city = -1, n_new = 0;
  const float2 act = A.action[e];
  if (T < Ntot) {
    // choose the starting city from the predicted coordinate (h_tsp.py:496-506)
    const int nearest_new = nearest_city(xy, sel, Ntot, act, 0, red);
    int new_start = nearest_new;
    if (T != 0) {
      start_city = nearest_city(xy, sel, Ntot, xy[nearest_new], 1, red);
      new_start = nearest_city(xy, sel, Ntot, xy[start_city], 0, red);
    }
    const int max_cities = max(A.max_new_nodes, FL - T);
    __syncthreads();
    // FIFO expansion over the k-NN lists by one warp (h_tsp.py:512-524): pop, record, append the
    // not-yet-seen unselected neighbours in list order (ballot + prefix count keeps the order).  The pops form a
    // chain of dependent L2 reads (190 of them per step): the k-NN rows of the next two queue entries are requested
    // while the current one is processed (k <= 64: a row is two registers per lane).
    if (tid < 32) {
      int head = 0, tail = 1;
      if (lane == 0) { queue[0] = new_start; seen[new_start >> 5] |= 1u << (new_start & 31); }
      __syncwarp();
      int ra[3], rb[3];
      bool have[3] = {false, false, false};
      auto request = [&](int slot, int qpos) {          // row of queue[qpos] -> registers of `slot` (if it is queued yet)
        if (have[slot] || qpos >= tail || k > 64) return;
        const long long* row = nn + (size_t)queue[qpos] * k;
        ra[slot] = lane < k ? (int)row[lane] : -1;
        rb[slot] = 32 + lane < k ? (int)row[32 + lane] : -1;
        have[slot] = true;
      };
      while (head < tail && n_new < max_cities && n_new < FL) {
        const int node = queue[head];
        request(0, head);
        request(1, head + 1);
        request(2, head + 2);
        ++head;
        if (lane == 0) newc[n_new] = node;
        ++n_new;
        for (int j0 = 0; j0 < k; j0 += 32) {
          const int j = j0 + lane;
          int nb = -1;
          if (j < k) nb = have[0] ? (j0 == 0 ? ra[0] : (j0 == 32 ? rb[0] : (int)nn[(size_t)node * k + j])) : (int)nn[(size_t)node * k + j];
          const bool ok = nb >= 0 && !((seen[nb >> 5] >> (nb & 31)) & 1u);
          const unsigned m = __ballot_sync(0xffffffffu, ok);
          if (ok) {
            const int slot = tail + __popc(m & ((1u << lane) - 1u));
            if (slot < A.qcap) queue[slot] = nb;
            atomicOr(&seen[nb >> 5], 1u << (nb & 31));
          }
          tail = min(tail + __popc(m), A.qcap);
          __syncwarp();
        }
        ra[0] = ra[1]; rb[0] = rb[1]; have[0] = have[1];
        ra[1] = ra[2]; rb[1] = rb[2]; have[1] = have[2];
        have[2] = false;
      }
      if (lane == 0) s_nnew = n_new;
    }
    __syncthreads();
    n_new = s_nnew;
  } else {
    // refine an old fragment: start at the visited city nearest to the prediction (h_tsp.py:525-530)
    start_city = nearest_city(xy, sel, Ntot, act, 1, red);
    __syncthreads();
  }
  // Env._extend_fragment (h_tsp.py:638-661)
  int st = 0;
  if (T != 0) {
    const int total = FL - n_new;
    if (total <= 0 || total > T) {
      st = 1;                                     // the reference's length asserts (:641, :656-658)
    } else {
      const int half = total / 2;
      const int offset = pos[start_city] - half;
      // reorder = tour[offset:] + tour[:offset] with Python slice semantics
      const int rot = (offset >= 0) ? offset : ((-offset <= T) ? T + offset : 0);
      for (int i = tid; i < FL; i += FRAG_THREADS) {
        long long v;
        if (i < half) v = tour[(rot + i) % T];
        else if (i < half + n_new) v = newc[i - half];
        else v = tour[(rot + (i - n_new)) % T];
        frag[i] = v;
      }
    }
  } else {
    if (n_new != FL) st = 1;
    for (int i = tid; i < FL; i += FRAG_THREADS) frag[i] = (i < n_new) ? newc[i] : -1;
  }
  if (tid == 0) { A.n_new[e] = n_new; A.start_city[e] = start_city; A.status[e] = st; }
}

}  // namespace htsp

using namespace htsp;

extern "C" int htsp_env_knn(const float* x, int E, int Ntot, int k, int64_t* knn, void* stream) {
  HTSP_REQUIRE(x && knn, "null pointer");
  HTSP_REQUIRE(E >= 0 && Ntot >= 2 && k >= 1 && k < Ntot, "bad shape");
  if (E == 0) return HTSP_OK;
  const size_t smem = (size_t)Ntot * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("htsp_env_knn: Ntot=%d does not fit one distance row in shared memory", Ntot);
    return HTSP_ERR_UNSUPPORTED;
  }
  HTSP_CHECK_CUDA(cudaFuncSetAttribute(env_knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  env_knn_kernel<<<dim3(Ntot, E), FRAG_THREADS, smem, (cudaStream_t)stream>>>((const float2*)x, Ntot, k,
                                                                             (long long*)knn);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

extern "C" int htsp_env_fragment(const float* x, const uint8_t* selected, const int64_t* knn, int k,
                                 const int64_t* tour, const int32_t* tour_len, const int32_t* pos,
                                 const float* action, const uint8_t* active, int E, int Ntot, int frag_len,
                                 int max_new_nodes, int64_t* fragment, int64_t* new_cities, int32_t* n_new,
                                 int32_t* start_city, int32_t* status, void* stream) {
  HTSP_REQUIRE(x && selected && knn && tour && tour_len && pos && action && fragment && new_cities &&
               n_new && start_city && status, "null pointer");
  HTSP_REQUIRE(E >= 0 && Ntot >= 2 && k >= 1 && frag_len >= 2 && max_new_nodes >= 1, "bad shape");
  if (E == 0) return HTSP_OK;
  FragArgs a;
  a.x = (const float2*)x; a.selected = selected; a.knn = (const long long*)knn; a.tour = (const long long*)tour;
  a.tour_len = tour_len; a.pos = pos; a.action = (const float2*)action; a.active = active;
  a.fragment = (long long*)fragment; a.new_cities = (long long*)new_cities; a.n_new = n_new;
  a.start_city = start_city; a.status = status;
  a.Ntot = Ntot; a.k = k; a.frag_len = frag_len; a.max_new_nodes = max_new_nodes;
  a.qcap = frag_len * k + 1;                      // every recorded city appends at most k entries
  const size_t smem = (size_t)((Ntot + 31) / 32) * 4 + (size_t)a.qcap * 4;
  if (smem > 200 * 1024) {
    set_error("htsp_env_fragment: Ntot=%d, frag_len=%d, k=%d need %zu bytes of shared memory", Ntot, frag_len, k, smem);
    return HTSP_ERR_UNSUPPORTED;
  }
  HTSP_CHECK_CUDA(cudaFuncSetAttribute(env_fragment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  env_fragment_kernel<<<E, FRAG_THREADS, smem, (cudaStream_t)stream>>>(a);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}
