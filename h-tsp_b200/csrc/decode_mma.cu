// K3 (tensor-core variant, HTSP_MODE_TC_X3 / HTSP_MODE_TC_FAST): persistent POMO rollout kernel.
//
// One CTA owns up to 16*NWARPS rows (POMO starts) of one instance for all N-2 decode steps
// (PathDecoder.forward, rl4cop/models/models.py:404-446; GroupState, train_path_solver.py:45-72).
// Each compute warp owns 16 rows; per step and head it runs a flash-style pass
//   S = Q_h K_h^T (m16n8k16) -> online softmax with the visited mask -> O_h += P V_h
// entirely in registers, then the pointer logits  u = O K2^T + c0  (multi_head_combine folded
// into the keys, K2 = emb Wc, c0 = emb.bc), 10*tanh, mask, arg-max, and the state update.
//
// Arithmetic: every GEMM operand is split into fp16 hi + fp16 lo (x = hi + lo to ~2^-22) and
// each product is 3 tensor-core MMAs (hi*hi + hi*lo + lo*hi) with fp32 accumulation, which
// measures 3e-5 max abs error on the clipped logits against the fp32 reference (single fp16 /
// bf16 operands give 2e-2 / 2e-1).  HTSP_MODE_TC_FAST drops the lo
// terms of the glimpse scores only (3.5e-4).
//
// Operands: K_h, V_h (per head) and K2 (per key chunk) of the instance are pre-split, pre-swizzled
// fp16 tiles in global memory (L2-resident, 1536*NP bytes per instance) and are STREAMED through
// a shared-memory ring by one producer warp with cp.async.bulk (TMA bulk copy, SASS UBLKCP) +
// mbarrier full/empty pairs; consumers read them with ldmatrix.  Nothing else is staged.
#include "common.cuh"
#include "mma_util.cuh"

namespace htsp {

#ifndef HTSP_MMA_NWARPS
#define HTSP_MMA_NWARPS 13
#endif
#ifndef HTSP_MMA_SPLIT_ACC
#define HTSP_MMA_SPLIT_ACC 0
#endif
#ifndef HTSP_MMA_STAGES
#define HTSP_MMA_STAGES 3
#endif
constexpr int MM_STAGES = HTSP_MMA_STAGES;
constexpr int MM_KT = 48;            // keys per K2 tile (multiple of 16)
constexpr float kInvSqrtH = 0.08838834764831845f;            // 1/sqrt(128)

// image layout per instance (bytes): 8 KV tiles of 128*NP, then K2 tiles of 512*rows_t
__host__ __device__ inline size_t img_bytes(int NP) { return (size_t)1536 * NP; }
__host__ __device__ inline int n_k2_tiles(int NP) { return (NP + MM_KT - 1) / MM_KT; }
// keys per K/V tile: the whole (padded) key range while it fits one stage, else 128-key slices
__host__ __device__ inline int kv_tile_keys(int NP) { return NP <= 208 ? NP : 128; }
__host__ __device__ inline int n_kv_tiles(int NP) { return (NP + kv_tile_keys(NP) - 1) / kv_tile_keys(NP); }
__host__ __device__ inline int stage_bytes(int NP) {
  int a = 128 * kv_tile_keys(NP), b = 512 * MM_KT;
  return a > b ? a : b;
}

// ------------------------------------------------------------------------------------------
// pack: proj fp32 [Bp*N, 640] -> per-instance fp16 hi/lo operand image (swizzled tile order)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  __half2 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    float2 f = __half22float2(h[i]);
    l[i] = __floats2half2_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
  }
  hi = *reinterpret_cast<uint4*>(h);
  lo = *reinterpret_cast<uint4*>(l);
}

__global__ void __launch_bounds__(256)
decode_pack_kernel(const float* __restrict__ proj, int Bp, int N, int NP,
                   unsigned char* __restrict__ img) {
  // one thread per (instance, key row r < NP, matrix m in {K,V,K2}, 8-column chunk c16 < 16)
  const size_t total = (size_t)Bp * NP * 48;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c16 = (int)(i % 16);
    const int m = (int)((i / 16) % 3);
    const int r = (int)((i / 48) % NP);
    const int b = (int)(i / ((size_t)48 * NP));
    float v[8];
    if (r < N) {
      const int col = (m == 0 ? 0 : (m == 1 ? H : 4 * H)) + c16 * 8;
      const float4* p = (const float4*)(proj + ((size_t)b * N + r) * PROJ + col);
      float4 a = __ldg(p), c = __ldg(p + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = 0.f;
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    unsigned char* base = img + (size_t)b * img_bytes(NP);
    if (m < 2) {
      const int h = c16 >> 1, chunk = c16 & 1;
      const int T = kv_tile_keys(NP), j = r / T, rl = r - j * T;
      const int rows_t = min(T, NP - j * T);
      // head h, key slice j: tile = [K hi | K lo | V hi | V lo], each rows_t*32 bytes
      unsigned char* tile = base + (size_t)h * 128 * NP + (size_t)j * 128 * T;
      const int off = rl * 32 + ((chunk ^ ((rl >> 2) & 1)) * 16);
      *(uint4*)(tile + (size_t)(2 * m) * rows_t * 32 + off) = hi;
      *(uint4*)(tile + (size_t)(2 * m + 1) * rows_t * 32 + off) = lo;
    } else {
      const int t = r / MM_KT, rl = r - t * MM_KT;
      const int rows_t = min(MM_KT, NP - t * MM_KT);
      unsigned char* tile = base + (size_t)8 * 128 * NP + (size_t)t * 512 * MM_KT;
      const int off = rl * 256 + ((c16 ^ (rl & 7)) * 16);
      *(uint4*)(tile + off) = hi;
      *(uint4*)(tile + (size_t)rows_t * 256 + off) = lo;
    }
  }
}

// ------------------------------------------------------------------------------------------
// the rollout kernel
// ------------------------------------------------------------------------------------------
struct MmaArgs {
  const float* x;        // [Bp,N,2]
  const float* proj;     // [Bp*N,640]  (QL at +256, QF at +384)
  const float* qfix;     // [Bp,128]
  const float* c0;       // [Bp,N]
  const unsigned char* img;
  const int32_t* src;
  const int32_t* tgt;
  const int32_t* first;
  const int32_t* forced_pi;
  int32_t* pi;
  float* reward;
  float* logits_out;
  int N, NP, G, rows_per_cta, fast;
};

template <int NWARPS, int NCHUNK_T>
__global__ void __launch_bounds__((NWARPS + 1) * 32, 1) decode_mma_kernel(const MmaArgs A) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  const int N = A.N, NP = A.NP, G = A.G;
  const int g0 = blockIdx.x * A.rows_per_cta;
  const int rows_here = min(A.rows_per_cta, G - g0);
  const int active_warps = (rows_here + 15) / 16;
  const int NWORDS = (NP + 31) / 32;
  const int VSTRIDE = NWORDS | 1;       // odd stride: the 8 rows of a warp quarter hit distinct banks
  const int STAGE = stage_bytes(NP);
  const int NT2 = n_k2_tiles(NP);
  // the NCHUNK_T > 0 instantiation is the single-slice case (NP = 16*NCHUNK_T <= 208)
  const int KVT = NCHUNK_T > 0 ? NCHUNK_T * 16 : kv_tile_keys(NP);
  const int NKV = NCHUNK_T > 0 ? 1 : n_kv_tiles(NP);
  const int tiles_per_step = NH * NKV + NT2;

  // shared memory carve-up
  unsigned char* ring = smem;                                           // MM_STAGES * STAGE
  uint4* obuf = (uint4*)(ring + (size_t)MM_STAGES * STAGE);             // [NWARPS][NH][32][2]
  float* qfs = (float*)(obuf + (size_t)NWARPS * NH * 64);               // [H]   qfix
  float* c0s = qfs + H;                                                 // [NP]
  float2* xs = (float2*)(c0s + NP);                                     // [N]
  unsigned* vis = (unsigned*)(xs + N);                                  // [16*NWARPS][VSTRIDE]
  uint64_t* bars = (uint64_t*)(((uintptr_t)(vis + 16 * NWARPS * VSTRIDE) + 15) & ~(uintptr_t)15);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + MM_STAGES);

  const int src = clamp_index(A.src[b], A.N), tgt = clamp_index(A.tgt[b], A.N);
  const float* projb = A.proj + (size_t)b * N * PROJ;
  const unsigned char* imgb = A.img + (size_t)b * img_bytes(NP);

  if (tid == 0) {
    for (int s = 0; s < MM_STAGES; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, active_warps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = tid; i < N; i += blockDim.x) xs[i] = __ldg((const float2*)A.x + (size_t)b * N + i);
  for (int i = tid; i < NP; i += blockDim.x) c0s[i] = i < N ? __ldg(&A.c0[(size_t)b * N + i]) : 0.f;
  for (int i = tid; i < H; i += blockDim.x) qfs[i] = __ldg(&A.qfix[(size_t)b * H + i]);
  // visited masks: padded keys (>= N) are permanently visited
  for (int i = tid; i < 16 * NWARPS * VSTRIDE; i += blockDim.x) {
    const int w = i % VSTRIDE;
    unsigned m = 0u;
    if (w < NWORDS) {
      const int lo = w * 32;
      if (lo + 32 > N) m = (N <= lo) ? 0xffffffffu : (0xffffffffu << (N - lo));
    }
    vis[i] = m;
  }
  __syncthreads();

  const int n_steps = N - 2;
  // ======================================================================== producer warp
  if (warp == NWARPS) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int step = 0; step < n_steps; ++step) {
        for (int t = 0; t < tiles_per_step; ++t, ++it) {
          const uint32_t s = it % MM_STAGES, use = it / MM_STAGES;
          mbar_wait(empty0 + 8 * s, (use & 1) ^ 1);
          const unsigned char* srcp;
          uint32_t bytes;
          if (t < NH * NKV) {
            const int hh = t / NKV, j = t - hh * NKV;
            srcp = imgb + (size_t)hh * 128 * NP + (size_t)j * 128 * KVT;
            bytes = 128 * min(KVT, NP - j * KVT);
          } else {
            const int k2 = t - NH * NKV;
            srcp = imgb + (size_t)8 * 128 * NP + (size_t)k2 * 512 * MM_KT;
            bytes = 512 * min(MM_KT, NP - k2 * MM_KT);
          }
          mbar_arrive_expect_tx(full0 + 8 * s, bytes);
          bulk_g2s(smem_u32(ring + (size_t)s * STAGE), srcp, bytes, full0 + 8 * s);
        }
      }
    }
    return;
  }
  if (warp >= active_warps) return;

  // ======================================================================== compute warps
  const int q4 = lane & 3, rr = lane >> 2;
  const int rA = warp * 16 + rr, rB = rA + 8;          // rows inside the CTA
  const int gA = g0 + rA, gB = g0 + rB;
  const bool okA = rA < rows_here, okB = rB < rows_here;
  unsigned* visA = vis + rA * VSTRIDE;
  unsigned* visB = vis + rB * VSTRIDE;
  int curA = 0, curB = 0, cntA = 0, cntB = 0;
  bool badA = false, badB = false;
  int32_t* piA = A.pi + ((size_t)b * G + min(gA, G - 1)) * N;
  int32_t* piB = A.pi + ((size_t)b * G + min(gB, G - 1)) * N;

  // visit(a) with source<->target coupling (train_path_solver.py:45-66); all four lanes of a quad
  // execute it redundantly with identical values (same-value smem writes are benign)
  auto visit = [&](unsigned* v, int32_t* p, int& cur, int& cnt, int a) {
    if (q4 == 0 && cnt < A.N) p[cnt] = a;          // (a row without finite logits repeats a node: never write past the row)
    v[a >> 5] |= 1u << (a & 31);
    cnt++;
    cur = a;
    const int partner = (a == src) ? tgt : ((a == tgt) ? src : -1);
    if (partner >= 0) {
      if (q4 == 0 && cnt < A.N) p[cnt] = partner;
      v[partner >> 5] |= 1u << (partner & 31);
      cnt++;
      cur = partner;
    }
  };
  // step 0: no decoder call (train_path_solver.py:256-257)
  const int fA = clamp_index(A.first[min(gA, G - 1)], A.N), fB = clamp_index(A.first[min(gB, G - 1)], A.N);
  if (okA) visit(visA, piA, curA, cntA, fA);
  if (okB) visit(visB, piB, curB, cntB, fB);
  __syncwarp();

  uint32_t it = 0;
  uint4* my_o = obuf + (size_t)warp * NH * 64 + lane;         // hi at [h*64], lo at [h*64+32]: 16B lane stride
  const float* qfA = projb + (size_t)fA * PROJ + 3 * H + 2 * q4;   // Wq_first emb[first]  (models.py:393-397)
  const float* qfB = projb + (size_t)fB * PROJ + 3 * H + 2 * q4;
  const int mi = lane >> 3, l7 = lane & 7;                    // ldmatrix lane addressing

  for (int step = 0; step < n_steps; ++step) {
    const float* qlA = projb + (size_t)curA * PROJ + 2 * H + 2 * q4;   // Wq_last emb[cur] (models.py:409-411)
    const float* qlB = projb + (size_t)curB * PROJ + 2 * H + 2 * q4;
    // software prefetch of the per-head query pieces (L2-resident tables), one head ahead
    float2 pf[8];
    auto fetch_q = [&](int h) {
      const int c = h * DK;
      pf[0] = __ldg((const float2*)(qfA + c)); pf[1] = __ldg((const float2*)(qfA + c + 8));
      pf[2] = __ldg((const float2*)(qfB + c)); pf[3] = __ldg((const float2*)(qfB + c + 8));
      pf[4] = __ldg((const float2*)(qlA + c)); pf[5] = __ldg((const float2*)(qlA + c + 8));
      pf[6] = __ldg((const float2*)(qlB + c)); pf[7] = __ldg((const float2*)(qlB + c + 8));
    };
    fetch_q(0);
    // ------------------------------------------------------------ glimpse, one head at a time
#pragma unroll 1
    for (int h = 0; h < NH; ++h) {
      // q = ((q_graph+q_source+q_target) + q_first) + q_last  (models.py:413); 1/sqrt(dk)*log2(e) is
      // folded in so the MMA result is the base-2 exponent directly.  A fragments hi/lo.
      uint32_t qh[4], ql[4];
      {
        const float2 x0 = *(const float2*)(qfs + h * DK + 2 * q4);
        const float2 x1 = *(const float2*)(qfs + h * DK + 2 * q4 + 8);
        split2(((x0.x + pf[0].x) + pf[4].x) * kScoreScale, ((x0.y + pf[0].y) + pf[4].y) * kScoreScale, qh[0], ql[0]);
        split2(((x0.x + pf[2].x) + pf[6].x) * kScoreScale, ((x0.y + pf[2].y) + pf[6].y) * kScoreScale, qh[1], ql[1]);
        split2(((x1.x + pf[1].x) + pf[5].x) * kScoreScale, ((x1.y + pf[1].y) + pf[5].y) * kScoreScale, qh[2], ql[2]);
        split2(((x1.x + pf[3].x) + pf[7].x) * kScoreScale, ((x1.y + pf[3].y) + pf[7].y) * kScoreScale, qh[3], ql[3]);
      }
      if (h + 1 < NH) fetch_q(h + 1);
      // hi*hi and (hi*lo + lo*hi) are accumulated separately: independent MMA chains (ILP) and the
      // small cross terms do not get absorbed into the large sum
      float o0h[4] = {0.f, 0.f, 0.f, 0.f}, o1h[4] = {0.f, 0.f, 0.f, 0.f};
      float o0l[4] = {0.f, 0.f, 0.f, 0.f}, o1l[4] = {0.f, 0.f, 0.f, 0.f};
      // shift starts at a finite floor so (-inf) - m never produces NaN; any real score lifts it
      float mA = -1e30f, mB = -1e30f, lA = 0.f, lB = 0.f;
#pragma unroll 1
      for (int jt = 0; jt < NKV; ++jt, ++it) {   // key slices of this head (one slice when NP <= 208)
      const uint32_t s = it % MM_STAGES, use = it / MM_STAGES;
      mbar_wait(full0 + 8 * s, use & 1);
      const int rows_t = NCHUNK_T > 0 ? NCHUNK_T * 16 : min(KVT, NP - jt * KVT), kbase = jt * KVT;
      const int nchunk = NCHUNK_T > 0 ? NCHUNK_T : rows_t / 16;
      const uint32_t tile = smem_u32(ring + (size_t)s * STAGE);
      const uint32_t khi = tile, klo = tile + rows_t * 32, vhi = tile + 2 * rows_t * 32, vlo = tile + 3 * rows_t * 32;
      // S = Q K^T for the 16 keys of chunk kc (two n8 tiles); software-pipelined: the MMAs of chunk
      // kc+1 are issued before the softmax of chunk kc so the tensor pipe overlaps the ALU/MUFU work
      auto qk_chunk = [&](int kc, float (&s0)[4], float (&s1)[4]) {
        const int keyK = kc * 16 + (mi >> 1) * 8 + l7;
        const uint32_t offK = keyK * 32 + (((mi & 1) ^ ((keyK >> 2) & 1)) * 16);
        uint32_t bh0, bh1, bh2, bh3;
        ldsm_x4(khi + offK, bh0, bh1, bh2, bh3);
        s0[0] = s0[1] = s0[2] = s0[3] = 0.f;
        s1[0] = s1[1] = s1[2] = s1[3] = 0.f;
        mma16816(s0, qh, bh0, bh1);
        mma16816(s1, qh, bh2, bh3);
        if (!A.fast) {
          uint32_t bl0, bl1, bl2, bl3;
          ldsm_x4(klo + offK, bl0, bl1, bl2, bl3);
          mma16816(s0, qh, bl0, bl1);
          mma16816(s1, qh, bl2, bl3);
          mma16816(s0, ql, bh0, bh1);
          mma16816(s1, ql, bh2, bh3);
        }
      };
      float sn0[4], sn1[4];
      qk_chunk(0, sn0, sn1);
#pragma unroll
      for (int kc = 0; kc < nchunk; ++kc) {
        float s0[4] = {sn0[0], sn0[1], sn0[2], sn0[3]}, s1[4] = {sn1[0], sn1[1], sn1[2], sn1[3]};
        if (kc + 1 < nchunk) qk_chunk(kc + 1, sn0, sn1);
        // ---- mask + online softmax (utils.py:20-27), exponent base 2
        const int key0 = kbase + kc * 16 + 2 * q4;         // keys key0,key0+1 (tile 0), +8,+9 (tile 1)
        const unsigned wA = visA[key0 >> 5] >> (key0 & 31);
        const unsigned wB = visB[key0 >> 5] >> (key0 & 31);
        float t0 = (wA & 1u) ? -INFINITY : s0[0];
        float t1 = (wA & 2u) ? -INFINITY : s0[1];
        float t2 = (wA & 0x100u) ? -INFINITY : s1[0];
        float t3 = (wA & 0x200u) ? -INFINITY : s1[1];
        float u0 = (wB & 1u) ? -INFINITY : s0[2];
        float u1 = (wB & 2u) ? -INFINITY : s0[3];
        float u2 = (wB & 0x100u) ? -INFINITY : s1[2];
        float u3 = (wB & 0x200u) ? -INFINITY : s1[3];
        // lazy online softmax: the running shift m only moves when some row of the warp sees a
        // score more than kLazy above it (p stays <= 2^kLazy, well inside fp16), so most chunks
        // skip the max reduction and the O/l rescale entirely
        const float cA = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3));
        const float cB = fmaxf(fmaxf(u0, u1), fmaxf(u2, u3));
        if (__any_sync(0xffffffffu, (cA > mA + kLazy) || (cB > mB + kLazy))) {
          const float nmA = fmaxf(mA, quad_max(cA));
          const float nmB = fmaxf(mB, quad_max(cB));
          const float aA = ex2_approx(mA - nmA), aB = ex2_approx(mB - nmB);
          mA = nmA; mB = nmB;
          lA *= aA; lB *= aB;
          o0h[0] *= aA; o0h[1] *= aA; o1h[0] *= aA; o1h[1] *= aA;
          o0h[2] *= aB; o0h[3] *= aB; o1h[2] *= aB; o1h[3] *= aB;
#if HTSP_MMA_SPLIT_ACC
          o0l[0] *= aA; o0l[1] *= aA; o1l[0] *= aA; o1l[1] *= aA;
          o0l[2] *= aB; o0l[3] *= aB; o1l[2] *= aB; o1l[3] *= aB;
#endif
        }
        t0 = ex2_approx(t0 - mA); t1 = ex2_approx(t1 - mA); t2 = ex2_approx(t2 - mA); t3 = ex2_approx(t3 - mA);
        u0 = ex2_approx(u0 - mB); u1 = ex2_approx(u1 - mB); u2 = ex2_approx(u2 - mB); u3 = ex2_approx(u3 - mB);
        lA += (t0 + t1) + (t2 + t3);
        lB += (u0 + u1) + (u2 + u3);
        // ---- O += P V for these 16 keys
        uint32_t ph[4], pl[4];
        split2(t0, t1, ph[0], pl[0]);
        split2(u0, u1, ph[1], pl[1]);
        split2(t2, t3, ph[2], pl[2]);
        split2(u2, u3, ph[3], pl[3]);
        const int keyV = kc * 16 + (mi & 1) * 8 + l7;
        const uint32_t offV = keyV * 32 + (((mi >> 1) ^ ((keyV >> 2) & 1)) * 16);
        uint32_t vh0, vh1, vh2, vh3, vl0, vl1, vl2, vl3;
        ldsm_x4_t(vhi + offV, vh0, vh1, vh2, vh3);
        ldsm_x4_t(vlo + offV, vl0, vl1, vl2, vl3);
#if HTSP_MMA_SPLIT_ACC
        mma16816(o0h, ph, vh0, vh1);
        mma16816(o1h, ph, vh2, vh3);
        mma16816(o0l, ph, vl0, vl1);
        mma16816(o1l, ph, vl2, vl3);
        mma16816(o0l, pl, vh0, vh1);
        mma16816(o1l, pl, vh2, vh3);
#else
        mma16816(o0h, ph, vh0, vh1);
        mma16816(o1h, ph, vh2, vh3);
        mma16816(o0h, ph, vl0, vl1);
        mma16816(o1h, ph, vl2, vl3);
        mma16816(o0h, pl, vh0, vh1);
        mma16816(o1h, pl, vh2, vh3);
#endif
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * s);
      }   // key slices
      const float iA = 1.f / quad_sum(lA), iB = 1.f / quad_sum(lB);
      uint4 fh, fl;
      split2((o0h[0] + o0l[0]) * iA, (o0h[1] + o0l[1]) * iA, fh.x, fl.x);
      split2((o0h[2] + o0l[2]) * iB, (o0h[3] + o0l[3]) * iB, fh.y, fl.y);
      split2((o1h[0] + o1l[0]) * iA, (o1h[1] + o1l[1]) * iA, fh.z, fl.z);
      split2((o1h[2] + o1l[2]) * iB, (o1h[3] + o1l[3]) * iB, fh.w, fl.w);
      my_o[h * 64] = fh;
      my_o[h * 64 + 32] = fl;
    }
    uint32_t ohi[NH][4], olo[NH][4];
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const uint4 fh = my_o[h * 64], fl = my_o[h * 64 + 32];
      ohi[h][0] = fh.x; ohi[h][1] = fh.y; ohi[h][2] = fh.z; ohi[h][3] = fh.w;
      olo[h][0] = fl.x; olo[h][1] = fl.y; olo[h][2] = fl.z; olo[h][3] = fl.w;
    }
    // ------------------------------------------------------------ pointer logits + arg-max
    float bestA = -INFINITY, bestB = -INFINITY;
    int idxA = 0x7fffffff, idxB = 0x7fffffff;
    float* loA = A.logits_out ? A.logits_out + (((size_t)b * G + min(gA, G - 1)) * n_steps + step) * N : nullptr;
    float* loB = A.logits_out ? A.logits_out + (((size_t)b * G + min(gB, G - 1)) * n_steps + step) * N : nullptr;
#pragma unroll 1
    for (int t = 0; t < NT2; ++t, ++it) {
      const uint32_t s = it % MM_STAGES, use = it / MM_STAGES;
      mbar_wait(full0 + 8 * s, use & 1);
      const int rows_t = min(MM_KT, NP - t * MM_KT);
      const uint32_t k2hi = smem_u32(ring + (size_t)s * STAGE), k2lo = k2hi + rows_t * 256;
      // two n8 tiles (16 keys) per iteration, hi*hi and cross terms in separate accumulators:
      // four independent MMA chains
#pragma unroll 1
      for (int np = 0; np < rows_t / 16; ++np) {
        float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
        float x0[4] = {0.f, 0.f, 0.f, 0.f}, x1[4] = {0.f, 0.f, 0.f, 0.f};
        const int r0 = np * 16 + l7, r1 = r0 + 8;
#pragma unroll
        for (int kp = 0; kp < NH / 2; ++kp) {       // two k16 steps (= two heads) per ldmatrix.x4
          const uint32_t off0 = r0 * 256 + (((4 * kp + mi) ^ (r0 & 7)) * 16);
          const uint32_t off1 = r1 * 256 + (((4 * kp + mi) ^ (r1 & 7)) * 16);
          uint32_t bh0, bh1, bh2, bh3, bl0, bl1, bl2, bl3;
          uint32_t dh0, dh1, dh2, dh3, dl0, dl1, dl2, dl3;
          ldsm_x4(k2hi + off0, bh0, bh1, bh2, bh3);
          ldsm_x4(k2hi + off1, dh0, dh1, dh2, dh3);
          ldsm_x4(k2lo + off0, bl0, bl1, bl2, bl3);
          ldsm_x4(k2lo + off1, dl0, dl1, dl2, dl3);
#if HTSP_MMA_SPLIT_ACC
#define XACC0 x0
#define XACC1 x1
#else
#define XACC0 a0
#define XACC1 a1
#endif
          mma16816(a0, ohi[2 * kp], bh0, bh1);
          mma16816(a1, ohi[2 * kp], dh0, dh1);
          mma16816(XACC0, ohi[2 * kp], bl0, bl1);
          mma16816(XACC1, ohi[2 * kp], dl0, dl1);
          mma16816(a0, ohi[2 * kp + 1], bh2, bh3);
          mma16816(a1, ohi[2 * kp + 1], dh2, dh3);
          mma16816(XACC0, olo[2 * kp], bh0, bh1);
          mma16816(XACC1, olo[2 * kp], dh0, dh1);
          mma16816(XACC0, ohi[2 * kp + 1], bl2, bl3);
          mma16816(XACC1, ohi[2 * kp + 1], dl2, dl3);
          mma16816(XACC0, olo[2 * kp + 1], bh2, bh3);
          mma16816(XACC1, olo[2 * kp + 1], dh2, dh3);
        }
        // score/sqrt(H) -> 10*tanh -> mask -> running arg-max, first index wins (models.py:438-440)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const float* acc = half ? a1 : a0;
          const float* crs = half ? x1 : x0;
          const int key0 = t * MM_KT + np * 16 + half * 8 + 2 * q4;
          const float2 cc = *(const float2*)(c0s + key0);
          const unsigned wA = visA[key0 >> 5] >> (key0 & 31);
          const unsigned wB = visB[key0 >> 5] >> (key0 & 31);
          float zA0 = clip_tanh10(((acc[0] + crs[0]) + cc.x) * kInvSqrtH);
          float zA1 = clip_tanh10(((acc[1] + crs[1]) + cc.y) * kInvSqrtH);
          float zB0 = clip_tanh10(((acc[2] + crs[2]) + cc.x) * kInvSqrtH);
          float zB1 = clip_tanh10(((acc[3] + crs[3]) + cc.y) * kInvSqrtH);
          if (wA & 1u) zA0 = -INFINITY;
          if (wA & 2u) zA1 = -INFINITY;
          if (wB & 1u) zB0 = -INFINITY;
          if (wB & 2u) zB1 = -INFINITY;
          if (loA != nullptr) {
            if (okA && key0 < N) { loA[key0] = zA0; if (key0 + 1 < N) loA[key0 + 1] = zA1; }
            if (okB && key0 < N) { loB[key0] = zB0; if (key0 + 1 < N) loB[key0 + 1] = zB1; }
          }
          if (zA0 > bestA) { bestA = zA0; idxA = key0; }
          if (zA1 > bestA) { bestA = zA1; idxA = key0 + 1; }
          if (zB0 > bestB) { bestB = zB0; idxB = key0; }
          if (zB1 > bestB) { bestB = zB1; idxB = key0 + 1; }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * s);
    }
#pragma unroll
    for (int off = 1; off <= 2; off <<= 1) {
      const float obA = __shfl_xor_sync(0xffffffffu, bestA, off);
      const int oiA = __shfl_xor_sync(0xffffffffu, idxA, off);
      const float obB = __shfl_xor_sync(0xffffffffu, bestB, off);
      const int oiB = __shfl_xor_sync(0xffffffffu, idxB, off);
      if (obA > bestA || (obA == bestA && oiA < idxA)) { bestA = obA; idxA = oiA; }
      if (obB > bestB || (obB == bestB && oiB < idxB)) { bestB = obB; idxB = oiB; }
    }
    // ---- state update (train_path_solver.py:45-72); all quad lanes hold identical values
    __syncwarp();
    if ((unsigned)idxA >= (unsigned)N) { idxA = curA; badA = true; }     // a row without finite logits has no candidate:
    if ((unsigned)idxB >= (unsigned)N) { idxB = curB; badB = true; }     // stay in range, flag the row (NaN reward)
    if (okA) visit(visA, piA, curA, cntA, A.forced_pi ? clamp_index(A.forced_pi[((size_t)b * G + gA) * N + min(cntA, N - 1)], N) : idxA);
    if (okB) visit(visB, piB, curB, cntB, A.forced_pi ? clamp_index(A.forced_pi[((size_t)b * G + gB) * N + min(cntB, N - 1)], N) : idxB);
    __syncwarp();
  }
  // (a flagged row, or a forced tour that repeats nodes, may have recorded fewer than N nodes: every entry of its tour
  // must still be a node, the tour is read as indices below and by the callers)
  if (q4 == 0) {
    if (okA) for (; cntA < N; ++cntA) piA[cntA] = curA;
    if (okB) for (; cntB < N; ++cntB) piB[cntB] = curB;
  }
  // ---- reward = -(closed tour length - fixed edge)   (train_path_solver.py:109-149)
  __syncwarp();
  {
    float lenA = 0.f, lenB = 0.f;
    for (int i = q4; i < N; i += 4) {
      const int j = (i + 1 == N) ? 0 : i + 1;
      if (okA) {
        const float2 p = xs[piA[i]], c = xs[piA[j]];
        const float dx = p.x - c.x, dy = p.y - c.y;
        lenA += sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      }
      if (okB) {
        const float2 p = xs[piB[i]], c = xs[piB[j]];
        const float dx = p.x - c.x, dy = p.y - c.y;
        lenB += sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      }
    }
    lenA = quad_sum(lenA);
    lenB = quad_sum(lenB);
    const float dx = xs[src].x - xs[tgt].x, dy = xs[src].y - xs[tgt].y;
    const float fixed = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    if (q4 == 0) {
      if (okA) A.reward[(size_t)b * G + gA] = badA ? __int_as_float(0x7fc00000) : -(lenA - fixed);
      if (okB) A.reward[(size_t)b * G + gB] = badB ? __int_as_float(0x7fc00000) : -(lenB - fixed);
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int np_of(int N) { return (N + 15) / 16 * 16; }

size_t decode_mma_pack_bytes(int Bp, int N) { return align_up((size_t)Bp * img_bytes(np_of(N)), 256); }

template <int NWARPS>
static size_t mma_smem_bytes(int N, int NP) {
  const int NWORDS = (NP + 31) / 32;
  size_t b = (size_t)MM_STAGES * stage_bytes(NP);
  b += (size_t)H * 4 + (size_t)NP * 4 + (size_t)N * 8;
  b += (size_t)16 * NWARPS * (NWORDS | 1) * 4;
  b += 16 + 2 * MM_STAGES * 8;
  b += (size_t)NWARPS * NH * 32 * 32;
  return b;
}

template <int NWARPS, int NCHUNK_T>
static int launch_mma_t(const MmaArgs& a, int Bp, cudaStream_t s) {
  const size_t smem = mma_smem_bytes<NWARPS>(a.N, a.NP);
  if (smem > 227 * 1024) {
    set_error("decode_mma: N=%d needs %zu bytes of shared memory", a.N, smem);
    return HTSP_ERR_UNSUPPORTED;
  }
  HTSP_CHECK_CUDA(cudaFuncSetAttribute(decode_mma_kernel<NWARPS, NCHUNK_T>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((a.G + a.rows_per_cta - 1) / a.rows_per_cta, Bp);
  kernel_timer_begin(s);
  decode_mma_kernel<NWARPS, NCHUNK_T><<<grid, (NWARPS + 1) * 32, smem, s>>>(a);
  kernel_timer_end(s);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

int launch_decode_mma(const float* x, const float* proj, const float* qfix, const float* c0,
                      const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp, int N,
                      int G, int mode, const int32_t* forced_pi, int32_t* pi, float* reward,
                      float* logits_out, void* pack_ws, cudaStream_t s) {
  const int NP = np_of(N);
  {
    const size_t total = (size_t)Bp * NP * 48;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    decode_pack_kernel<<<blocks, 256, 0, s>>>(proj, Bp, N, NP, (unsigned char*)pack_ws);
    HTSP_CHECK_LAUNCH();
  }
  constexpr int NW = HTSP_MMA_NWARPS;
  MmaArgs a;
  a.x = x; a.proj = proj; a.qfix = qfix; a.c0 = c0; a.img = (const unsigned char*)pack_ws;
  a.src = src; a.tgt = tgt; a.first = first; a.forced_pi = forced_pi;
  a.pi = pi; a.reward = reward; a.logits_out = logits_out;
  a.N = N; a.NP = NP; a.G = G; a.fast = (mode == HTSP_MODE_TC_FAST) ? 1 : 0;
  const int n_cta = (G + 16 * NW - 1) / (16 * NW);
  a.rows_per_cta = (G + n_cta - 1) / n_cta;
  if (NP == 208) return launch_mma_t<NW, 13>(a, Bp, s);
  return launch_mma_t<NW, 0>(a, Bp, s);
}

}  // namespace htsp
