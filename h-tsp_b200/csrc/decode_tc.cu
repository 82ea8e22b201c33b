// K3, Blackwell-native variant (HTSP_MODE_TC_X3 / HTSP_MODE_TC_FAST, N <= 224): persistent POMO
// rollout kernel on tcgen05 tensor cores with TMEM accumulators.
//
// One CTA owns one instance (all G <= 256 POMO starts, two M = 128 row tiles) for all N-2 decode
// steps (PathDecoder.forward, rl4cop/models/models.py:404-446; GroupState / Env,
// rl4cop/train_path_solver.py:45-149, 256-288).  Per step and head h, for each row tile t:
//   S_h   = Q_h K_h^T            tcgen05.mma  A = query rows in shared memory, D = TMEM columns
//   P_h   = 2^(S_h - m) masked   softmax warps: tcgen05.ld -> mask/ex2/row sum -> fp16 hi/lo ->
//                                tcgen05.st IN PLACE over S_h (one thread owns one row = one lane)
//   O_h   = P_h V_h              tcgen05.mma  A = P_h from TMEM, D = 16 TMEM columns
//   O_h / l -> fp16 hi/lo into the shared-memory A tile (over the consumed Q_h slot)
// and once per step
//   U     = O K2^T               tcgen05.mma (multi_head_combine folded into the keys, K2 = emb Wc)
//   argmax_j 10 tanh((U + c0)/sqrt(128)) over unvisited j, state update, next query gather.
// Arithmetic is split-fp16 (x = hi + lo, three MMAs per product, fp32 accumulation; DESIGN.md
// section 4); HTSP_MODE_TC_FAST drops the lo terms of S only.
//
// Warp roles (320 threads): warp 0 = operand producer (cp.async.bulk of pre-swizzled K/V/K2 tiles
// from the L2-resident per-instance image into a 3-stage ring, shared by both row tiles), warp 1 =
// MMA issuer (one thread), warps 2-5 / 6-9 = softmax warps of tile 0 / 1 (warp%4 selects the TMEM
// lane quarter).  All synchronisation is mbarrier based; tcgen05.commit publishes accumulators and
// releases ring stages.
#include "common.cuh"
#include "mma_util.cuh"
#include "tc_util.cuh"

namespace htsp {

constexpr int TCD_THREADS = 320;
constexpr int TCD_STAGES = 3;
constexpr int TCD_TILE_STAGING = 65536;   // per row tile: [hi k0-63 | hi k64-127 | lo k0-63 | lo k64-127], 16 KB each
constexpr int TCD_OCOL = 448;             // O accumulators: column 448 + 32*tile + 16*(head&1)
constexpr float kTcdInvSqrtH = 0.08838834764831845f;   // 1/sqrt(128)
constexpr float kTcdFloor = -1e30f;       // initial softmax shift (integer valued, any real score lifts it)
constexpr float kTcdThresh = 32768.f;     // chunk sum above this => some p may leave fp16 range: move the shift

// ---- per-instance operand image (bytes) ----------------------------------------------------------
//   for p in 0..3:  KA_p [NP keys x 64 halfs]  = K hi/lo of heads 2p, 2p+1   (cols: hi_2p|lo_2p|hi_2p+1|lo_2p+1)
//                   VT_p [64 rows x keys]      = V^T hi/lo of heads 2p, 2p+1 (rows: hi_2p|lo_2p|hi_2p+1|lo_2p+1),
//                                                blocks of 64 keys (8 KB each)
//   K2 hi cols 0-63, K2 hi cols 64-127, K2 lo cols 0-63, K2 lo cols 64-127: [NP keys x 64 halfs] each
// every block is K-major SWIZZLE_128B and is copied verbatim into a ring stage.
__host__ __device__ inline int tcd_kb(int NP) { return (NP + 63) / 64; }
__host__ __device__ inline size_t tcd_szk(int NP) { return (size_t)NP * 128; }
__host__ __device__ inline size_t tcd_szv(int NP) { return (size_t)tcd_kb(NP) * 8192; }
__host__ __device__ inline size_t tcd_img_bytes(int NP) { return 8 * tcd_szk(NP) + 4 * tcd_szv(NP); }
__host__ __device__ inline int tcd_stage_bytes(int NP) {
  const size_t a = tcd_szk(NP), b = tcd_szv(NP);
  return (int)(a > b ? a : b);
}
// i-th tile of a step's stream: KA_0, VT_0, KA_1, VT_1, KA_2, VT_2, KA_3, VT_3, K2 x 4
__device__ __forceinline__ void tcd_stream_tile(int i, int NP, size_t& off, uint32_t& bytes) {
  const size_t szk = tcd_szk(NP), szv = tcd_szv(NP);
  if (i < 8) {
    off = (size_t)(i >> 1) * (szk + szv) + ((i & 1) ? szk : 0);
    bytes = (uint32_t)((i & 1) ? szv : szk);
  } else {
    off = 4 * (szk + szv) + (size_t)(i - 8) * szk;
    bytes = (uint32_t)szk;
  }
}

__device__ __forceinline__ void tcd_split8(const float* v, uint4& hi, uint4& lo) {
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

// proj fp32 [Bp*N, 640] (K | V | QL | QF | K2) -> operand image
__global__ void __launch_bounds__(256)
decode_tc_pack_kernel(const float* __restrict__ proj, int Bp, int N, int NP, unsigned char* __restrict__ img) {
  const size_t szk = tcd_szk(NP), szv = tcd_szv(NP), per = tcd_img_bytes(NP);
  // K and K2: one thread per (instance, key r, matrix m in {K, K2}, 8-column chunk c16)
  const size_t total_a = (size_t)Bp * NP * 32;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_a;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c16 = (int)(i % 16);
    const int m = (int)((i / 16) % 2);
    const int r = (int)((i / 32) % NP);
    const int b = (int)(i / ((size_t)32 * NP));
    float v[8];
    if (r < N) {
      const float4* p = (const float4*)(proj + ((size_t)b * N + r) * PROJ + (m == 0 ? 0 : 4 * H) + c16 * 8);
      const float4 a = __ldg(p), c = __ldg(p + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = 0.f;
    }
    uint4 hi, lo;
    tcd_split8(v, hi, lo);
    unsigned char* base = img + (size_t)b * per;
    if (m == 0) {
      const int h = c16 >> 1, p = h >> 1, e = h & 1;
      unsigned char* blk = base + (size_t)p * (szk + szv);
      *(uint4*)(blk + tc::sw128_off(r, e * 4 + (c16 & 1))) = hi;
      *(uint4*)(blk + tc::sw128_off(r, e * 4 + 2 + (c16 & 1))) = lo;
    } else {
      unsigned char* blk = base + 4 * (szk + szv) + (size_t)(c16 >> 3) * szk;
      *(uint4*)(blk + tc::sw128_off(r, c16 & 7)) = hi;
      *(uint4*)(blk + 2 * szk + tc::sw128_off(r, c16 & 7)) = lo;
    }
  }
  // V^T: one thread per (instance, group of 8 keys r8, value column c)
  const int NG = tcd_kb(NP) * 8;
  const size_t total_v = (size_t)Bp * NG * 128;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_v;
       i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % 128);
    const int r8 = (int)((i / 128) % NG);
    const int b = (int)(i / ((size_t)128 * NG));
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r = r8 * 8 + k;
      v[k] = (r < N) ? __ldg(proj + ((size_t)b * N + r) * PROJ + H + c) : 0.f;
    }
    uint4 hi, lo;
    tcd_split8(v, hi, lo);
    const int h = c >> 4, d = c & 15, p = h >> 1, e = h & 1;
    unsigned char* blk = img + (size_t)b * per + (size_t)p * (szk + szv) + szk + (size_t)(r8 >> 3) * 8192;
    *(uint4*)(blk + tc::sw128_off(e * 32 + d, r8 & 7)) = hi;
    *(uint4*)(blk + tc::sw128_off(e * 32 + 16 + d, r8 & 7)) = lo;
  }
}

struct TcdArgs {
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
  float* dbg;            // DEBUG builds: [S raw 2*8*128*NP | O 256*128 | U raw 256*NP] of instance 0 at dbg_step
  int N, NP, G, fast, dbg_step;
};

// barrier slots
enum { TB_FULL = 0, TB_EMPTY = 3, TB_QFULL = 6, TB_SFULL = 8, TB_PFULL = 10, TB_OFULL = 12, TB_OREADY = 16,
       TB_UFULL = 18, TB_COUNT = 20 };

__device__ __forceinline__ unsigned tcd_word(const unsigned (&vis)[7], int w) {
  unsigned r = vis[0];
#pragma unroll
  for (int i = 1; i < 7; ++i) r = (w == i) ? vis[i] : r;
  return r;
}

__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

template <bool LOGITS, bool DEBUG>
__global__ void __launch_bounds__(TCD_THREADS, 1) decode_tc_kernel(const TcdArgs A) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x;
  const int N = A.N, NP = A.NP, G = A.G;
  const int nch = NP >> 4;                       // 16-key chunks
  const int n_steps = N - 2;
  const int STAGE = tcd_stage_bytes(NP);
  const int n_tiles = G > 128 ? 2 : 1;

  unsigned char* staging = smem;                                         // 2 x 64 KB
  unsigned char* ring = smem + 2 * TCD_TILE_STAGING;                     // TCD_STAGES x STAGE
  float* c0s = (float*)(ring + (size_t)TCD_STAGES * STAGE);              // [NP]
  uint64_t* bars = (uint64_t*)(c0s + NP);
  uint32_t* tmem_slot = (uint32_t*)(bars + TB_COUNT);
  const uint32_t bar0 = tc::smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  const unsigned char* imgb = A.img + (size_t)b * tcd_img_bytes(NP);
  const float* projb = A.proj + (size_t)b * N * PROJ;

  if ((tc::smem_u32(smem) & 1023u) != 0u) __trap();   // SWIZZLE_128B atoms need 1024-byte alignment
  if (tid == 0) {
    for (int s = 0; s < TCD_STAGES; ++s) { tc::mbar_init(BAR(TB_FULL + s), 1); tc::mbar_init(BAR(TB_EMPTY + s), 1); }
    for (int t = 0; t < 2; ++t) {
      const int rows_t = min(128, G - 128 * t);
      const int warps_t = rows_t > 0 ? (rows_t + 31) / 32 : 1;
      tc::mbar_init(BAR(TB_QFULL + t), warps_t);
      tc::mbar_init(BAR(TB_SFULL + t), 1);
      tc::mbar_init(BAR(TB_PFULL + t), warps_t);
      tc::mbar_init(BAR(TB_OFULL + 2 * t), 1);
      tc::mbar_init(BAR(TB_OFULL + 2 * t + 1), 1);
      tc::mbar_init(BAR(TB_OREADY + t), warps_t);
      tc::mbar_init(BAR(TB_UFULL + t), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero the A tiles (rows that never receive a query must hold finite values)
  for (int i = tid; i < 2 * TCD_TILE_STAGING / 16; i += TCD_THREADS) ((uint4*)staging)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < NP; i += TCD_THREADS) c0s[i] = i < N ? __ldg(&A.c0[(size_t)b * N + i]) : 0.f;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc::fence_async_smem();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================================================== operand producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int step = 0; step < n_steps; ++step) {
        for (int i = 0; i < 12; ++i, ++it) {
          const uint32_t s = it % TCD_STAGES, use = it / TCD_STAGES;
          tc::mbar_wait(BAR(TB_EMPTY + s), (use & 1) ^ 1);
          size_t off;
          uint32_t bytes;
          tcd_stream_tile(i, NP, off, bytes);
          tc::mbar_expect_tx(BAR(TB_FULL + s), bytes);
          tc::bulk_g2s(tc::smem_u32(ring + (size_t)s * STAGE), imgb + off, bytes, BAR(TB_FULL + s));
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idS = tc::idesc_f16(NP), idO = tc::idesc_f16(16);
      const uint32_t stg0 = tc::smem_u32(staging), ring0 = tc::smem_u32(ring);
      auto stage_addr = [&](uint32_t k) { return ring0 + (k % TCD_STAGES) * (uint32_t)STAGE; };
      auto stage_wait = [&](uint32_t k) {
        tc::mbar_wait(BAR(TB_FULL + (k % TCD_STAGES)), (k / TCD_STAGES) & 1);
        tc::fence_after();
      };
      auto stage_free = [&](uint32_t k) { tc::commit(BAR(TB_EMPTY + (k % TCD_STAGES))); };
      // S_h(t) = Q_h K_h^T into TMEM columns [t*NP, t*NP+NP)
      auto issue_S = [&](int t, int h, uint32_t ka) {
        const uint32_t a_hi = stg0 + t * TCD_TILE_STAGING + (h >> 2) * 16384 + (h & 3) * 32, a_lo = a_hi + 32768;
        const uint32_t b_hi = ka + (h & 1) * 64, b_lo = b_hi + 32;
        const uint32_t d = tmem_base + (uint32_t)(t * NP);
        tc::mma_ss(d, tc::desc_sw128(a_hi), tc::desc_sw128(b_hi), idS, 0);
        if (!A.fast) {
          tc::mma_ss(d, tc::desc_sw128(a_hi), tc::desc_sw128(b_lo), idS, 1);
          tc::mma_ss(d, tc::desc_sw128(a_lo), tc::desc_sw128(b_hi), idS, 1);
        }
        tc::commit(BAR(TB_SFULL + t));
      };
      // O_h(t) = P_h V_h: A = P hi/lo in TMEM (8 + 8 columns per 16 keys), B = V^T rows of head h
      auto issue_PV = [&](int t, int h, uint32_t va) {
        const uint32_t d = tmem_base + (uint32_t)(TCD_OCOL + 32 * t + 16 * (h & 1));
        const uint32_t pa = tmem_base + (uint32_t)(t * NP);
        const uint32_t vb = va + (h & 1) * 4096;
        for (int ks = 0; ks < nch; ++ks) {
          const uint32_t v_hi = vb + (ks >> 2) * 8192 + (ks & 3) * 32, v_lo = v_hi + 2048;
          const uint32_t p_hi = pa + 16 * ks, p_lo = p_hi + 8;
          tc::mma_ts(d, p_hi, tc::desc_sw128(v_hi), idO, ks > 0);
          tc::mma_ts(d, p_hi, tc::desc_sw128(v_lo), idO, 1);
          tc::mma_ts(d, p_lo, tc::desc_sw128(v_hi), idO, 1);
        }
        tc::commit(BAR(TB_OFULL + 2 * t + (h & 1)));
      };
      for (int step = 0; step < n_steps; ++step) {
        const uint32_t k0 = (uint32_t)step * 12;
        stage_wait(k0);
        for (int t = 0; t < n_tiles; ++t) {
          tc::mbar_wait(BAR(TB_QFULL + t), step & 1);
          tc::fence_after();
          issue_S(t, 0, stage_addr(k0));
        }
        for (int h = 0; h < NH; ++h) {
          const int p = h >> 1;
          const uint32_t kv = k0 + 2 * p + 1;
          if ((h & 1) == 0) stage_wait(kv);
          for (int t = 0; t < n_tiles; ++t) {
            tc::mbar_wait(BAR(TB_PFULL + t), (step * NH + h) & 1);
            tc::fence_after();
            issue_PV(t, h, stage_addr(kv));
            if (h + 1 < NH) {
              const uint32_t kk = k0 + 2 * ((h + 1) >> 1);
              if (t == 0 && ((h + 1) & 1) == 0) stage_wait(kk);
              issue_S(t, h + 1, stage_addr(kk));
            }
          }
          if ((h & 1) == 0) stage_free(k0 + 2 * p);   // KA_p: S of both heads of the pair issued
          else stage_free(kv);                         // VT_p: PV of both heads issued
        }
        // pointer logits U(t) = O K2^T into the (consumed) S columns
        for (int i = 0; i < 4; ++i) {
          const uint32_t k = k0 + 8 + i;
          stage_wait(k);
          const uint32_t kb = stage_addr(k);
          const int half = i & 1;                      // O columns 64*half .. +63
          for (int t = 0; t < n_tiles; ++t) {
            if (i == 0) {
              tc::mbar_wait(BAR(TB_OREADY + t), step & 1);
              tc::fence_after();
            }
            const uint32_t d = tmem_base + (uint32_t)(t * NP);
            const uint32_t o_hi = stg0 + t * TCD_TILE_STAGING + half * 16384, o_lo = o_hi + 32768;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t bd = tc::desc_sw128(kb + kk * 32);
              tc::mma_ss(d, tc::desc_sw128(o_hi + kk * 32), bd, idS, (i | kk) != 0);
              if (i < 2) tc::mma_ss(d, tc::desc_sw128(o_lo + kk * 32), bd, idS, 1);
            }
          }
          stage_free(k);
        }
        for (int t = 0; t < n_tiles; ++t) tc::commit(BAR(TB_UFULL + t));
      }
    }
  } else {
    // ================================================================== softmax / state warps
    const int t = (warp - 2) >> 2, q = warp & 3;
    const int rows_t = min(128, G - 128 * t);
    if (q * 32 < rows_t) {
      const int row = q * 32 + lane;                 // row inside the tile = TMEM lane
      const int g = t * 128 + row;
      const bool valid = g < G;
      const int gc = min(g, G - 1);
      const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
      const uint32_t tS = tmem_base + lane_addr + (uint32_t)(t * NP);
      const uint32_t tO = tmem_base + lane_addr + (uint32_t)(TCD_OCOL + 32 * t);
      unsigned char* stg = staging + t * TCD_TILE_STAGING;
      const int src = A.src[b], tgt = A.tgt[b];
      const int first = A.first[gc];
      int32_t* pirow = A.pi + ((size_t)b * G + gc) * N;
      const float4 qfix4 = __ldg((const float4*)(A.qfix + (size_t)b * H) + lane);

      unsigned vis[7];
#pragma unroll
      for (int w = 0; w < 7; ++w) {           // padded keys (>= N) are permanently visited
        const int lo = w * 32;
        vis[w] = (lo + 32 <= N) ? 0u : ((N <= lo) ? 0xffffffffu : (0xffffffffu << (N - lo)));
      }
      int cur = 0, cnt = 0;
      auto mark = [&](int a) {
#pragma unroll
        for (int w = 0; w < 7; ++w) vis[w] |= ((a >> 5) == w) ? (1u << (a & 31)) : 0u;
      };
      // visit(a) with source<->target coupling (train_path_solver.py:45-66)
      auto visit = [&](int a) {
        if (valid) pirow[cnt] = a;
        mark(a);
        cnt++;
        cur = a;
        const int partner = (a == src) ? tgt : ((a == tgt) ? src : -1);
        if (partner >= 0) {
          if (valid) pirow[cnt] = partner;
          mark(partner);
          cnt++;
          cur = partner;
        }
      };
      // query rows of this warp's 32 starts -> A tile (fp16 hi/lo, K-major SWIZZLE_128B):
      // q = ((q_graph+q_source+q_target) + q_first) + q_last (models.py:413), 1/sqrt(dk)*log2(e) folded in
      auto gather_q = [&]() {
        const uint32_t sub = (uint32_t)(lane >> 4) * 16384u + (uint32_t)(lane & 1) * 8u;
        const uint32_t chunk = (uint32_t)(lane & 15) >> 1;
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
          const int rr = q * 32 + r;
          if (t * 128 + rr >= G) break;
          const int cr = __shfl_sync(0xffffffffu, cur, r), fr = __shfl_sync(0xffffffffu, first, r);
          const float4 ql = __ldg((const float4*)(projb + (size_t)cr * PROJ + 2 * H) + lane);
          const float4 qf = __ldg((const float4*)(projb + (size_t)fr * PROJ + 3 * H) + lane);
          uint32_t h0, l0, h1, l1;
          split2(((qfix4.x + qf.x) + ql.x) * kScoreScale, ((qfix4.y + qf.y) + ql.y) * kScoreScale, h0, l0);
          split2(((qfix4.z + qf.z) + ql.z) * kScoreScale, ((qfix4.w + qf.w) + ql.w) * kScoreScale, h1, l1);
          unsigned char* p = stg + sub + tc::sw128_off((uint32_t)rr, chunk);
          *(uint2*)p = make_uint2(h0, h1);
          *(uint2*)(p + 32768) = make_uint2(l0, l1);
        }
        tc::fence_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(BAR(TB_QFULL + t));
      };

      // step 0: no decoder call (train_path_solver.py:256-257)
      visit(first);
      gather_q();

      float l_prev = 1.f;
      // DEBUG builds: cycles this warp spends per phase (0 wait S, 1 softmax, 2 wait O + drain,
      // 3 wait U, 4 arg-max, 5 state + query gather), written behind the dumps
      long long tph[6] = {0, 0, 0, 0, 0, 0};
      long long tlast = DEBUG ? clock64() : 0;
      auto tick = [&](int ph) {
        if (DEBUG) {
          const long long now = clock64();
          tph[ph] += now - tlast;
          tlast = now;
        }
      };
      // drain O_h: TMEM -> /l -> fp16 hi/lo -> A tile columns 16h..16h+15 (the consumed Q_h slot)
      auto drain_o = [&](int step, int h, float l) {
        tc::mbar_wait(BAR(TB_OFULL + 2 * t + (h & 1)), (step * 4 + (h >> 1)) & 1);
        tc::fence_after();
        uint32_t o[16];
        tc::ld16(tO + 16 * (h & 1), o);
        tc::ld16_wait(o);
        const float inv = 1.f / l;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(o[j]) * inv;
        if (DEBUG && b == 0 && step == A.dbg_step) {
          float* dO = A.dbg + (size_t)2 * 8 * 128 * NP + (size_t)(t * 128 + row) * 128 + 16 * h;
#pragma unroll
          for (int j = 0; j < 16; ++j) dO[j] = v[j];
        }
        uint4 h0, l0, h1, l1;
        tcd_split8(v, h0, l0);
        tcd_split8(v + 8, h1, l1);
        unsigned char* p = stg + (h >> 2) * 16384;
        const uint32_t c = (uint32_t)(h & 3) * 2;
        *(uint4*)(p + tc::sw128_off((uint32_t)row, c)) = h0;
        *(uint4*)(p + tc::sw128_off((uint32_t)row, c + 1)) = h1;
        *(uint4*)(p + 32768 + tc::sw128_off((uint32_t)row, c)) = l0;
        *(uint4*)(p + 32768 + tc::sw128_off((uint32_t)row, c + 1)) = l1;
      };

      for (int step = 0; step < n_steps; ++step) {
        // ------------------------------------------------------------ glimpse, one head at a time
#pragma unroll 1
        for (int h = 0; h < NH; ++h) {
          tc::mbar_wait(BAR(TB_SFULL + t), (step * NH + h) & 1);
          tc::fence_after();
          tick(0);
          float m = valid ? kTcdFloor : 0.f, l = 0.f;
          uint32_t sA[16], sB[16];
          tc::ld16(tS, sA);
          // one 16-key chunk: mask, 2^(s-m), row sum, fp16 hi/lo, store in place
          auto chunk = [&](int c, uint32_t (&s)[16], unsigned bits) {
            if (DEBUG && b == 0 && step == A.dbg_step) {
              float* dS = A.dbg + ((size_t)(t * 8 + h) * 128 + row) * NP + 16 * c;
#pragma unroll
              for (int j = 0; j < 16; ++j) dS[j] = __uint_as_float(s[j]);
            }
            float tv[16], p[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) tv[j] = ((bits >> j) & 1u) ? -INFINITY : __uint_as_float(s[j]);
            float csum;
            auto expo = [&]() {
              float2 acc = make_float2(0.f, 0.f);
              const float2 nm = make_float2(-m, -m);
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const float2 d = __fadd2_rn(make_float2(tv[j], tv[j + 1]), nm);
                p[j] = ex2_approx(d.x);
                p[j + 1] = ex2_approx(d.y);
                acc = __fadd2_rn(acc, make_float2(p[j], p[j + 1]));
              }
              csum = acc.x + acc.y;
            };
            expo();
            const bool need = valid && !(csum <= kTcdThresh);
            if (__any_sync(0xffffffffu, need)) {
              // move the shift (to an integer, so rescaling earlier chunks by 2^(m-m') is exact)
              float cm = tv[0];
#pragma unroll
              for (int j = 1; j < 16; ++j) cm = fmaxf(cm, tv[j]);
              const float m_new = (need && cm > m) ? ceilf(cm) : m;
              const float alpha = ex2_approx(m - m_new);
              if (__any_sync(0xffffffffu, (m_new != m) && (l > 0.f))) {
                tc::st_wait();
                for (int cc = 0; cc < c; ++cc) {
                  uint32_t r[16];
                  tc::ld16(tS + 16 * cc, r);
                  tc::ld16_wait(r);
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    float2 f = __half22float2(*reinterpret_cast<__half2*>(&r[j]));
                    __half2 hh = __floats2half2_rn(f.x * alpha, f.y * alpha);
                    r[j] = *reinterpret_cast<uint32_t*>(&hh);
                  }
                  st16(tS + 16 * cc, r);
                }
              }
              l *= alpha;
              m = m_new;
              expo();
            }
            l += csum;
            uint32_t out[16];
#pragma unroll
            for (int j = 0; j < 16; j += 2) split2(p[j], p[j + 1], out[j >> 1], out[8 + (j >> 1)]);
            st16(tS + 16 * c, out);
          };
#pragma unroll 1
          for (int w = 0; 2 * w < nch; ++w) {
            const unsigned word = tcd_word(vis, w);
            tc::ld16_wait(sA);
            if (2 * w + 1 < nch) tc::ld16(tS + 16 * (2 * w + 1), sB);
            chunk(2 * w, sA, word & 0xffffu);
            if (2 * w + 1 < nch) {
              tc::ld16_wait(sB);
              if (2 * w + 2 < nch) tc::ld16(tS + 16 * (2 * w + 2), sA);
              chunk(2 * w + 1, sB, word >> 16);
            }
          }
          tc::st_wait();
          tc::fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(BAR(TB_PFULL + t));
          tick(1);
          if (h > 0) drain_o(step, h - 1, l_prev);
          l_prev = l;
          tick(2);
        }
        drain_o(step, NH - 1, l_prev);
        tc::fence_async_smem();
        tc::fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(BAR(TB_OREADY + t));
        tick(2);

        // ------------------------------------------------------------ pointer logits + arg-max
        tc::mbar_wait(BAR(TB_UFULL + t), step & 1);
        tc::fence_after();
        tick(3);
        float best = -INFINITY;
        int idx = 0x7fffffff;
        float* lo_out = nullptr;
        if (LOGITS && A.logits_out != nullptr)
          lo_out = A.logits_out + (((size_t)b * G + gc) * n_steps + step) * N;
        // pass(EXACT): EXACT compares the clipped logits 10*tanh(.) themselves (models.py:438-440);
        // otherwise the monotone pre-activation is compared, which picks the same key except when two
        // clipped logits round to the same float (handled by re-running EXACT near the clip)
        auto pass = [&](auto exact_tag) {
          constexpr bool EXACT = decltype(exact_tag)::value;
          best = -INFINITY;
          idx = 0x7fffffff;
          uint32_t u[16];
#pragma unroll 1
          for (int c = 0; c < nch; ++c) {
            tc::ld16(tS + 16 * c, u);
            const unsigned bits = (tcd_word(vis, c >> 1) >> ((c & 1) * 16)) & 0xffffu;
            tc::ld16_wait(u);
            if (DEBUG && b == 0 && step == A.dbg_step) {
              float* dU = A.dbg + (size_t)2 * 8 * 128 * NP + (size_t)256 * 128 + (size_t)(t * 128 + row) * NP + 16 * c;
#pragma unroll
              for (int j = 0; j < 16; ++j) dU[j] = __uint_as_float(u[j]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float z = (__uint_as_float(u[j]) + c0s[16 * c + j]) * kTcdInvSqrtH;
              if (EXACT) z = clip_tanh10(z);
              if ((bits >> j) & 1u) z = -INFINITY;
              if (LOGITS && lo_out != nullptr && valid && 16 * c + j < N) lo_out[16 * c + j] = z;
              if (z > best) { best = z; idx = 16 * c + j; }
            }
          }
        };
        if (LOGITS) {
          pass(std::true_type{});
        } else {
          pass(std::false_type{});
          if (__any_sync(0xffffffffu, valid && best > 3.0f)) pass(std::true_type{});
        }
        tick(4);
        // ---- state update (train_path_solver.py:45-72)
        const int a = A.forced_pi ? A.forced_pi[((size_t)b * G + gc) * N + cnt] : idx;
        visit(a);
        tc::fence_before();
        if (step + 1 < n_steps) gather_q();
        tick(5);
      }
      if (DEBUG && b == 0 && lane == 0) {
        float* dT = A.dbg + (size_t)2 * 8 * 128 * NP + (size_t)256 * 128 + (size_t)256 * NP + (warp - 2) * 8;
        for (int i = 0; i < 6; ++i) dT[i] = (float)tph[i];
      }
      // ---- reward = -(closed tour length - fixed edge)   (train_path_solver.py:109-149)
      if (valid) {
        const float2* xb = (const float2*)A.x + (size_t)b * N;
        float len = 0.f;
        float2 p0 = __ldg(&xb[pirow[0]]);
        const float2 pfirst = p0;
        for (int i = 1; i <= N; ++i) {
          const float2 c = (i == N) ? pfirst : __ldg(&xb[pirow[i]]);
          const float dx = p0.x - c.x, dy = p0.y - c.y;
          len += sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
          p0 = c;
        }
        const float2 ps = __ldg(&xb[src]), pt = __ldg(&xb[tgt]);
        const float dx = ps.x - pt.x, dy = ps.y - pt.y;
        const float fixed = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        A.reward[(size_t)b * G + g] = -(len - fixed);
      }
    }
  }
  tc::fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int tcd_np_of(int N) { return (N + 15) / 16 * 16; }

bool decode_tc_supported(int N, int G) { return tcd_np_of(N) <= 224 && G <= 256 && N >= 3; }
size_t decode_tc_pack_bytes(int Bp, int N) { return align_up((size_t)Bp * tcd_img_bytes(tcd_np_of(N)), 256) + 1024; }
size_t decode_tc_dbg_floats(int N) {
  const int NP = tcd_np_of(N);
  return (size_t)2 * 8 * 128 * NP + (size_t)256 * 128 + (size_t)256 * NP + 64;
}

template <bool LOGITS, bool DEBUG>
static int launch_tcd(const TcdArgs& a, int Bp, cudaStream_t s) {
  const size_t smem = (size_t)2 * TCD_TILE_STAGING + (size_t)TCD_STAGES * tcd_stage_bytes(a.NP) +
                      (size_t)a.NP * 4 + TB_COUNT * 8 + 16;
  if (smem > 227 * 1024) {
    set_error("decode_tc: N=%d needs %zu bytes of shared memory", a.N, smem);
    return HTSP_ERR_UNSUPPORTED;
  }
  HTSP_CHECK_CUDA(cudaFuncSetAttribute(decode_tc_kernel<LOGITS, DEBUG>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kernel_timer_begin(s);
  decode_tc_kernel<LOGITS, DEBUG><<<Bp, TCD_THREADS, smem, s>>>(a);
  kernel_timer_end(s);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

int launch_decode_tc(const float* x, const float* proj, const float* qfix, const float* c0,
                     const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp, int N,
                     int G, int mode, const int32_t* forced_pi, int32_t* pi, float* reward,
                     float* logits_out, void* pack_ws, float* dbg, int dbg_step, cudaStream_t s) {
  HTSP_REQUIRE(decode_tc_supported(N, G), "decode_tc shape");
  const int NP = tcd_np_of(N);
  unsigned char* img = (unsigned char*)(((uintptr_t)pack_ws + 1023) & ~(uintptr_t)1023);
  {
    const size_t total = (size_t)Bp * NP * 32;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    decode_tc_pack_kernel<<<blocks, 256, 0, s>>>(proj, Bp, N, NP, img);
    HTSP_CHECK_LAUNCH();
  }
  TcdArgs a;
  a.x = x; a.proj = proj; a.qfix = qfix; a.c0 = c0; a.img = img;
  a.src = src; a.tgt = tgt; a.first = first; a.forced_pi = forced_pi;
  a.pi = pi; a.reward = reward; a.logits_out = logits_out; a.dbg = dbg;
  a.N = N; a.NP = NP; a.G = G; a.fast = (mode == HTSP_MODE_TC_FAST) ? 1 : 0; a.dbg_step = dbg_step;
  if (dbg != nullptr) return launch_tcd<true, true>(a, Bp, s);
  if (logits_out != nullptr) return launch_tcd<true, false>(a, Bp, s);
  return launch_tcd<false, false>(a, Bp, s);
}

}  // namespace htsp
