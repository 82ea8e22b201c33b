// K3 (fp32 variant, HTSP_MODE_FP32): the whole POMO rollout of RB rows (starts) of one instance
// in one persistent CTA -- all N-2 decode steps, the visited mask, the source<->target coupling,
// argmax, tour and reward -- with fp32 FMAs in the reference's op order (no algebraic folds
// except the per-node Wq_last/Wq_first tables).  This is the parity anchor for the tensor-core
// kernel in decode_mma.cu, not the fast path.
//   PathDecoder.forward            rl4cop/models/models.py:404-446
//   multi_head_attention           rl4cop/utils/utils.py:12-29
//   GroupState.move_to / mask      rl4cop/train_path_solver.py:45-72
//   Env._get_path_distance         rl4cop/train_path_solver.py:109-149
#include "common.cuh"

namespace htsp {

constexpr int DF_THREADS = 256;

template <int RB>
__global__ void __launch_bounds__(DF_THREADS)
decode_f32_kernel(const float* __restrict__ x, const float* __restrict__ emb,
                  const float* __restrict__ proj, const float* __restrict__ qfix,
                  const float* __restrict__ WcT, const float* __restrict__ bc,
                  const int32_t* __restrict__ src_a, const int32_t* __restrict__ tgt_a,
                  const int32_t* __restrict__ first, int N, int G,
                  const int32_t* __restrict__ forced_pi, int32_t* __restrict__ pi,
                  float* __restrict__ reward, float* __restrict__ logits_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.y;
  const int g0 = blockIdx.x * RB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NW = (N + 31) / 32;  // mask words per row

  float* qrow = (float*)smem_raw;            // [RB][H]  qfix + q_first
  float* q = qrow + RB * H;                  // [RB][H]
  float* o = q + RB * H;                     // [RB][H]
  float* f = o + RB * H;                     // [RB][H]
  float* S = f + RB * H;                     // [RB][NH][N]
  float* u = S + (size_t)RB * NH * N;        // [RB][N]
  float2* xs = (float2*)(u + (size_t)RB * N);  // [N]
  int* tour = (int*)(xs + N);                // [RB][N]
  unsigned* vis = (unsigned*)(tour + RB * N);  // [RB][NW]
  int* cur = (int*)(vis + RB * NW);          // [RB]
  int* cnt = cur + RB;                       // [RB]
  int* nocand = cnt + RB;                    // [RB] a decode step found no finite logit: NaN reward

  const int src = clamp_index(src_a[b], N), tgt = clamp_index(tgt_a[b], N);
  const float* projb = proj + (size_t)b * N * PROJ;
  const float* embb = emb + (size_t)b * N * H;

  for (int i = tid; i < N; i += DF_THREADS) xs[i] = __ldg((const float2*)x + (size_t)b * N + i);
  for (int i = tid; i < RB * NW; i += DF_THREADS) vis[i] = 0u;
  __syncthreads();
  // step 0: visit(first[g]) with coupling, no decoder call (train_path_solver.py:256-257)
  if (tid < RB) {
    const int r = tid, g = g0 + r;
    int c = 0, n = 0;
    if (g < G) {
      int a = clamp_index(first[g], N);
      tour[r * N + n++] = a; vis[r * NW + (a >> 5)] |= 1u << (a & 31); c = a;
      if (a == src) { tour[r * N + n++] = tgt; vis[r * NW + (tgt >> 5)] |= 1u << (tgt & 31); c = tgt; }
      else if (a == tgt) { tour[r * N + n++] = src; vis[r * NW + (src >> 5)] |= 1u << (src & 31); c = src; }
    }
    cur[r] = c; cnt[r] = n; nocand[r] = 0;
  }
  for (int i = tid; i < RB * H; i += DF_THREADS) {
    const int r = i / H, c = i % H, g = min(g0 + r, G - 1);
    qrow[i] = __ldg(&qfix[(size_t)b * H + c]) + __ldg(&projb[(size_t)clamp_index(first[g], N) * PROJ + 3 * H + c]);
  }
  __syncthreads();

  const float inv_sqrt_h = sqrtf((float)H);
  for (int step = 0; step < N - 2; ++step) {
    // 1. glimpse query                                        (models.py:409-413)
    for (int i = tid; i < RB * H; i += DF_THREADS) {
      const int r = i / H, c = i % H;
      q[i] = qrow[i] + __ldg(&projb[(size_t)cur[r] * PROJ + 2 * H + c]);
    }
    __syncthreads();
    // 2. scores = q k^T / 4 + mask                            (utils.py:20-23)
    for (int idx = tid; idx < NH * N; idx += DF_THREADS) {
      const int h = idx / N, j = idx % N;
      float k[DK];
#pragma unroll
      for (int c = 0; c < DK; c += 4) {
        float4 t = __ldg((const float4*)(projb + (size_t)j * PROJ + h * DK + c));
        k[c] = t.x; k[c + 1] = t.y; k[c + 2] = t.z; k[c + 3] = t.w;
      }
#pragma unroll 4
      for (int r = 0; r < RB; ++r) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < DK; ++c) s = fmaf(q[r * H + h * DK + c], k[c], s);
        s = s / 4.0f;
        if ((vis[r * NW + (j >> 5)] >> (j & 31)) & 1u) s = -INFINITY;
        S[((size_t)r * NH + h) * N + j] = s;
      }
    }
    __syncthreads();
    // 3. softmax over keys, fp32                               (utils.py:26-27)
    for (int pr = warp; pr < RB * NH; pr += DF_THREADS / 32) {
      float* row = S + (size_t)pr * N;
      float m = -INFINITY;
      for (int j = lane; j < N; j += 32) m = fmaxf(m, row[j]);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
      float l = 0.f;
      for (int j = lane; j < N; j += 32) { float e = expf(row[j] - m); row[j] = e; l += e; }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
      for (int j = lane; j < N; j += 32) row[j] = row[j] / l;
    }
    __syncthreads();
    // 4. o = p v                                               (utils.py:26-29)
    {
      constexpr int HALF = (RB + 1) / 2;
      const int c = tid & (H - 1), r0 = (tid >> 7) * HALF, h = c / DK;
      float acc[HALF];
#pragma unroll
      for (int r = 0; r < HALF; ++r) acc[r] = 0.f;
      for (int j = 0; j < N; ++j) {
        float v = __ldg(&projb[(size_t)j * PROJ + H + c]);
#pragma unroll
        for (int r = 0; r < HALF; ++r)
          if (r0 + r < RB) acc[r] = fmaf(S[((size_t)(r0 + r) * NH + h) * N + j], v, acc[r]);
      }
#pragma unroll
      for (int r = 0; r < HALF; ++r)
        if (r0 + r < RB) o[(r0 + r) * H + c] = acc[r];
    }
    __syncthreads();
    // 5. final_q = multi_head_combine(o)                       (models.py:437)
    {
      constexpr int HALF = (RB + 1) / 2;
      const int c = tid & (H - 1), r0 = (tid >> 7) * HALF;
      float acc[HALF];
#pragma unroll
      for (int r = 0; r < HALF; ++r) acc[r] = 0.f;
      for (int i = 0; i < H; ++i) {
        float w = __ldg(&WcT[(size_t)i * H + c]);
#pragma unroll
        for (int r = 0; r < HALF; ++r)
          if (r0 + r < RB) acc[r] = fmaf(o[(r0 + r) * H + i], w, acc[r]);
      }
      const float bias = __ldg(&bc[c]);
#pragma unroll
      for (int r = 0; r < HALF; ++r)
        if (r0 + r < RB) f[(r0 + r) * H + c] = acc[r] + bias;
    }
    __syncthreads();
    // 6. pointer scores = final_q . emb_j / sqrt(H)             (models.py:438)
    for (int j = tid; j < N; j += DF_THREADS) {
      float acc[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) acc[r] = 0.f;
      const float4* e4 = (const float4*)(embb + (size_t)j * H);
      for (int c4 = 0; c4 < H / 4; ++c4) {
        float4 e = __ldg(&e4[c4]);
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const float4 ff = *(const float4*)&f[r * H + c4 * 4];
          acc[r] = fmaf(ff.x, e.x, acc[r]); acc[r] = fmaf(ff.y, e.y, acc[r]);
          acc[r] = fmaf(ff.z, e.z, acc[r]); acc[r] = fmaf(ff.w, e.w, acc[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        float z = 10.0f * tanhf(acc[r] / inv_sqrt_h);          // (models.py:439)
        if ((vis[r * NW + (j >> 5)] >> (j & 31)) & 1u) z = -INFINITY;   // (models.py:440)
        u[(size_t)r * N + j] = z;
      }
    }
    __syncthreads();
    // 7. greedy action (first index wins ties) + state update  (train_path_solver.py:263-264,45-72)
    for (int r = warp; r < RB; r += DF_THREADS / 32) {
      const int g = g0 + r;
      if (g >= G) continue;
      float best = -INFINITY; int bi = 0x7fffffff;
      for (int j = lane; j < N; j += 32) {
        float z = u[(size_t)r * N + j];
        if (logits_out != nullptr)
          logits_out[(((size_t)b * G + g) * (N - 2) + step) * N + j] = z;
        if (z > best) { best = z; bi = j; }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, best, off);
        int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) {
        int n = cnt[r];
        int a = (forced_pi != nullptr) ? clamp_index(forced_pi[((size_t)b * G + g) * N + min(n, N - 1)], N) : bi;
        if ((unsigned)a >= (unsigned)N) {          // no finite logit (non-finite embeddings): stay in range, flag the row
          a = cur[r];
          nocand[r] = 1;
        }
        int c = a;
        // (n < N always holds for a valid tour; a flagged forced tour that repeats nodes must not write past the row)
        if (n < N) tour[r * N + n++] = a;
        vis[r * NW + (a >> 5)] |= 1u << (a & 31);
        if (a == src) { if (n < N) tour[r * N + n++] = tgt; vis[r * NW + (tgt >> 5)] |= 1u << (tgt & 31); c = tgt; }
        else if (a == tgt) { if (n < N) tour[r * N + n++] = src; vis[r * NW + (src >> 5)] |= 1u << (src & 31); c = src; }
        cur[r] = c; cnt[r] = n;
      }
    }
    __syncthreads();
  }
  // tour + reward = -(closed length - fixed edge)              (train_path_solver.py:109-149)
  for (int r = warp; r < RB; r += DF_THREADS / 32) {
    const int g = g0 + r;
    if (g >= G) continue;
    // (a flagged row, or a forced tour that repeats nodes, may have recorded fewer than N nodes: every entry must be a node)
    for (int i = cnt[r] + lane; i < N; i += 32) tour[r * N + i] = cur[r];
    __syncwarp();
    float len = 0.f;
    for (int i = lane; i < N; i += 32) {
      const int a = tour[r * N + i], c = tour[r * N + (i + 1 == N ? 0 : i + 1)];
      pi[((size_t)b * G + g) * N + i] = a;
      const float dx = xs[a].x - xs[c].x, dy = xs[a].y - xs[c].y;
      len += sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) len += __shfl_xor_sync(0xffffffffu, len, off);
    if (lane == 0) {
      const float dx = xs[src].x - xs[tgt].x, dy = xs[src].y - xs[tgt].y;
      const float fixed = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      reward[(size_t)b * G + g] = nocand[r] ? __int_as_float(0x7fc00000) : -(len - fixed);
    }
  }
}

template <int RB>
static size_t decode_f32_smem(int N) {
  const int NW = (N + 31) / 32;
  return (size_t)4 * RB * H * 4 + (size_t)RB * NH * N * 4 + (size_t)RB * N * 4 + (size_t)N * 8 +
         (size_t)RB * N * 4 + (size_t)RB * NW * 4 + 3 * RB * 4;
}

template <int RB>
static int launch_rb(const float* blob, int n_layers, const float* x, const float* emb,
                     const float* proj, const float* qfix, const int32_t* src, const int32_t* tgt,
                     const int32_t* first, int Bp, int N, int G, const int32_t* forced_pi,
                     int32_t* pi, float* reward, float* logits_out, cudaStream_t s) {
  DecOff d = dec_off(n_layers);
  const size_t smem = decode_f32_smem<RB>(N);
  HTSP_CHECK_CUDA(cudaFuncSetAttribute(decode_f32_kernel<RB>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((G + RB - 1) / RB, Bp);
  kernel_timer_begin(s);
  decode_f32_kernel<RB><<<grid, DF_THREADS, smem, s>>>(
      x, emb, proj, qfix, blob + d.wproj + (size_t)4 * H * H, blob + d.bc, src, tgt, first, N, G,
      forced_pi, pi, reward, logits_out);
  kernel_timer_end(s);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

int launch_decode_f32(const float* blob, int n_layers, const float* x, const float* emb,
                      const float* proj, const float* qfix, const int32_t* src,
                      const int32_t* tgt, const int32_t* first, int Bp, int N, int G,
                      const int32_t* forced_pi, int32_t* pi, float* reward, float* logits_out,
                      cudaStream_t s) {
  const size_t limit = 200 * 1024;
  if (decode_f32_smem<16>(N) <= limit && G > 8)
    return launch_rb<16>(blob, n_layers, x, emb, proj, qfix, src, tgt, first, Bp, N, G, forced_pi, pi, reward, logits_out, s);
  if (decode_f32_smem<8>(N) <= limit && G > 4)
    return launch_rb<8>(blob, n_layers, x, emb, proj, qfix, src, tgt, first, Bp, N, G, forced_pi, pi, reward, logits_out, s);
  if (decode_f32_smem<4>(N) <= limit)
    return launch_rb<4>(blob, n_layers, x, emb, proj, qfix, src, tgt, first, Bp, N, G, forced_pi, pi, reward, logits_out, s);
  set_error("decode_f32: N=%d too large for shared memory", N);
  return HTSP_ERR_UNSUPPORTED;
}

}  // namespace htsp
