// K3, Blackwell-native variant (HTSP_MODE_TC_X3 / HTSP_MODE_TC_FAST, 17 <= N <= 208): persistent POMO
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
// section 4); HTSP_MODE_TC_FAST stores the softmax probabilities P as a single fp16 (S, V and the logits keep
// the split).
//
// Warp roles (640 threads = 20 warps; dispatch at the end of the kernel): warp 0 = operand producer
// (cp.async.bulk of pre-swizzled K/V/K2 blocks from the L2-resident per-instance image into one 3-stage
// ring per row tile); warp 3 = MMA issuer of tile 0, warp 19 (or 2) = issuer of tile 1 (one elected
// lane issues); warps 4-19 = softmax / state warps: tile t, key half X, lane quarter q = warp % 4 --
// two warps share every 32-row lane quarter and split the keys between them ("split-K").  All
// synchronisation is mbarrier or named-barrier based; tcgen05.commit publishes accumulators and
// releases ring stages.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "mma_util.cuh"
#include "tc_util.cuh"
#include "tc_rollout_util.cuh"

namespace htsp {

constexpr int TCD_THREADS = 640;          // 20 warps: producer, 2 MMA issuers, (idle), 2 tiles x 2 halves x 4 softmax warps
constexpr int TCD_STAGES = 3;
constexpr int TCD_STAGE = 16384;          // ring stage (one key group of a K block, or one head of V^T)
constexpr int TCD_TILE_STAGING = 65536;   // per row tile: [hi k0-63 | hi k64-127 | lo k0-63 | lo k64-127], 16 KB each
constexpr int TCD_OCOL = 448;             // O accumulators: column 448 + 32*tile + 16*half (split-K partials)
constexpr int TCD_XCOL = 416;             // exchange columns: 416 + 16*tile: [head&1][half]{m,l} (8), chosen node (1)
constexpr float kTcdInvSqrtH = 0.08838834764831845f;   // 1/sqrt(128)
constexpr float kTcdFloor = -1e30f;       // initial softmax shift (integer valued, any real score lifts it)
constexpr float kTcdThresh = 32768.f;     // chunk sum above this => some p may leave fp16 range: move the shift

// ---- per-instance operand image (bytes) ----------------------------------------------------------
//   for p in 0..3:  KA_p [NP keys x 64 halfs]  = K hi/lo of heads 2p, 2p+1   (cols: hi_2p|lo_2p|hi_2p+1|lo_2p+1)
//                   VT_2p, VT_2p+1: [32 rows x keys] = V^T of one head (rows: hi d0-15 | lo d0-15),
//                                                blocks of 64 keys (4 KB each)
//   K2 hi cols 0-63, K2 hi cols 64-127, K2 lo cols 0-63, K2 lo cols 64-127: [NP keys x 64 halfs] each
// every block is K-major SWIZZLE_128B (rows of 128 bytes).  Key-major blocks (KA, K2) are streamed in
// two key halves ([0,K0) and [K0,NP), K0 = 16*floor(NP/32)) so that a ring stage is at most 16 KB.
__host__ __device__ inline int tcd_kb(int NP) { return (NP + 63) / 64; }
__host__ __device__ inline size_t tcd_szk(int NP) { return (size_t)NP * 128; }
__host__ __device__ inline size_t tcd_szvh(int NP) { return (size_t)tcd_kb(NP) * 4096; }
__host__ __device__ inline size_t tcd_pair_bytes(int NP) { return tcd_szk(NP) + 2 * tcd_szvh(NP); }
__host__ __device__ inline size_t tcd_img_bytes(int NP) { return 4 * tcd_pair_bytes(NP) + 4 * tcd_szk(NP); }
__host__ __device__ inline int tcd_k0(int NP) { return 16 * ((NP >> 4) >> 1); }

// proj fp32 [Bp*N, 640] (K | V | QL | QF | K2) -> operand image.  grid = (blocks per instance, Bp): all index
// math is 32-bit shifts (the first version spent its time in 64-bit divisions by NP)
__global__ void __launch_bounds__(256)
decode_tc_pack_kernel(const float* __restrict__ proj, int Bp, int N, int NP, unsigned char* __restrict__ img) {
  const size_t szk = tcd_szk(NP), szvh = tcd_szvh(NP), pair = tcd_pair_bytes(NP), per = tcd_img_bytes(NP);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (int b = blockIdx.y; b < Bp; b += gridDim.y) {
  const float* pb = proj + (size_t)b * N * PROJ;
  unsigned char* base = img + (size_t)b * per;
  // K and K2: one thread per (key r, matrix m in {K, K2}, 8-column chunk c16)
  if (i < NP * 32) {
    const int c16 = i & 15, m = (i >> 4) & 1, r = i >> 5;
    float v[8];
    if (r < N) {
      const float4* p = (const float4*)(pb + (size_t)r * PROJ + (m == 0 ? 0 : 4 * H) + c16 * 8);
      const float4 a = __ldg(p), c = __ldg(p + 1);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = 0.f;
    }
    uint4 hi, lo;
    tcd_split8(v, hi, lo);
    if (m == 0) {
      const int h = c16 >> 1, p = h >> 1, e = h & 1;
      unsigned char* blk = base + (size_t)p * pair;
      *(uint4*)(blk + tc::sw128_off(r, e * 4 + (c16 & 1))) = hi;
      *(uint4*)(blk + tc::sw128_off(r, e * 4 + 2 + (c16 & 1))) = lo;
    } else {
      unsigned char* blk = base + 4 * pair + (size_t)(c16 >> 3) * szk;
      *(uint4*)(blk + tc::sw128_off(r, c16 & 7)) = hi;
      *(uint4*)(blk + 2 * szk + tc::sw128_off(r, c16 & 7)) = lo;
    }
  }
  // V^T: one thread per (group of 8 keys r8, value column c)
  const int NG = tcd_kb(NP) * 8;
  if (i < NG * 128) {
    const int c = i & 127, r8 = i >> 7;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int r = r8 * 8 + k;
      v[k] = (r < N) ? __ldg(pb + (size_t)r * PROJ + H + c) : 0.f;
    }
    uint4 hi, lo;
    tcd_split8(v, hi, lo);
    const int h = c >> 4, d = c & 15, p = h >> 1, e = h & 1;
    unsigned char* blk = base + (size_t)p * pair + szk + (size_t)e * szvh + (size_t)(r8 >> 3) * 4096;
    *(uint4*)(blk + tc::sw128_off(d, r8 & 7)) = hi;
    *(uint4*)(blk + tc::sw128_off(16 + d, r8 & 7)) = lo;
  }
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
  float* dbg;            // DEBUG builds: [S raw 2*8*128*NP | O 256*128 | U raw 256*NP | phase cycles] of instance 0
  int N, NP, G, fast, dbg_step;
};

// barrier slots of one row tile
enum { TB_FULL = 0, TB_EMPTY = 3, TB_QFULL = 6, TB_SFULL = 7, TB_PFULL = 9, TB_OFULL = 11, TB_OREADY = 13,
       TB_UFULL = 14, TB_PER_TILE = 15 };   // SFULL / PFULL / OFULL: one per key half

// SAMPLE (htsp_decode_sample): instead of the arg-max every row draws its next node from softmax(logits)
// (train_path_solver.py:266-271, probs.multinomial(1)) by the Gumbel-max rule arg-max_j (z_j + g_j); the 64-bit
// seed travels in A.dbg (unused outside DEBUG builds: TcdArgs must not grow).
// PSINGLE (HTSP_MODE_TC_FAST): the softmax probabilities are stored as one fp16 (no lo part), the row
// sum is taken from the ROUNDED values so that the normalisation stays consistent; S, V and the pointer
// logits keep the full split.  Logit error ~2.5e-4 (bound 1e-3) instead of ~3e-5.
template <bool LOGITS, bool DEBUG, bool PSINGLE, bool SPLIT, bool SAMPLE, bool P15 = false>
__global__ void __launch_bounds__(TCD_THREADS, 1) decode_tc_kernel(const TcdArgs A) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler too
  // Low-occupancy launches (fewer instances than SMs, e.g. one environment step of a few large-TSP
  // instances) split the POMO rows of an instance over several CTAs: rows are independent, the operand
  // image is shared, and a CTA with fewer softmax warps finishes its decode steps sooner.
  // TcdArgs must not grow (8 more bytes of kernel parameters cost 7 % of the full-occupancy throughput) and the
  // unsplit kernel must not change: the split kernels are separate instantiations that find their rows per CTA
  // in A.dbg_step (unused outside DEBUG builds).
  const int n_parts = SPLIT ? (A.G + A.dbg_step - 1) / A.dbg_step : 1;
  const int b = SPLIT ? blockIdx.x / n_parts : blockIdx.x;
  const int r0 = SPLIT ? (blockIdx.x - b * n_parts) * A.dbg_step : 0;
  const int N = A.N, NP = A.NP, GT = A.G;
  const int G = SPLIT ? min(A.dbg_step, GT - r0) : A.G;   // rows of this CTA, from row r0 of the instance
  const unsigned long long sample_seed = SAMPLE ? (unsigned long long)(uintptr_t)A.dbg : 0ull;
  const int n_steps = N - 2;
  const int n_tiles = G > 128 ? 2 : 1;
  const int K0 = tcd_k0(NP), K1 = NP - K0;     // key halves A = [0,K0), B = [K0,NP): one softmax warp each
  const int nks = NP >> 4;                                               // k16 steps of P V

  unsigned char* staging = smem;                                         // 2 x 64 KB
  unsigned char* rings = smem + 2 * TCD_TILE_STAGING;                    // 2 tiles x TCD_STAGES x 16 KB
  float* c0s = (float*)(rings + (size_t)2 * TCD_STAGES * TCD_STAGE);     // [NP]
  uint64_t* bars = (uint64_t*)(c0s + NP);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * TB_PER_TILE);
  const uint32_t bar0 = tc::smem_u32(bars);

  const unsigned char* imgb = A.img + (size_t)b * tcd_img_bytes(NP);
  const float* projb = A.proj + (size_t)b * N * PROJ;

  if ((tc::smem_u32(smem) & 1023u) != 0u) __trap();   // SWIZZLE_128B atoms need 1024-byte alignment
  if (tid == 0) {
    for (int t = 0; t < 2; ++t) {
      const uint32_t bt = bar0 + 8u * (uint32_t)(t * TB_PER_TILE);
      const int rows_t = min(128, G - 128 * t);
      const int warps_t = rows_t > 0 ? (rows_t + 31) / 32 : 1;
      for (int s = 0; s < TCD_STAGES; ++s) { tc::mbar_init(bt + 8 * (TB_FULL + s), 1); tc::mbar_init(bt + 8 * (TB_EMPTY + s), 1); }
      tc::mbar_init(bt + 8 * TB_QFULL, 2 * warps_t);
      tc::mbar_init(bt + 8 * TB_SFULL, 1);
      tc::mbar_init(bt + 8 * (TB_SFULL + 1), 1);
      tc::mbar_init(bt + 8 * TB_PFULL, warps_t);            // half A: its softmax warps
      tc::mbar_init(bt + 8 * (TB_PFULL + 1), 2 * warps_t);  // half B: its softmax warps + the primaries' "O drained"
      tc::mbar_init(bt + 8 * TB_OFULL, 1);
      tc::mbar_init(bt + 8 * (TB_OFULL + 1), 1);
      tc::mbar_init(bt + 8 * TB_OREADY, warps_t);
      tc::mbar_init(bt + 8 * TB_UFULL, 1);
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

  // The issuers sit on scheduler 3 (warp % 4 == 3), the least loaded one: row tile 1 of a 200-row
  // instance has no rows in lane quarter 3, so its softmax warp 19 is free to be tile 1's issuer
  // (when tile 1 does use that quarter, warp 2 issues instead).
  const int iss1_warp = (G - 128 > 96) ? 2 : 19;
  // The operand producer polls two rings; on scheduler 0 its polling competes with four softmax warps.  P15 (chosen
  // by the host for two row tiles of long rows, NP >= 144, with tile 1 leaving lane quarter 3 free): warp 15 (tile 1,
  // primary, quarter 3) has no rows and takes the role on scheduler 3 instead (+0.9 % at N = G = 200, +2.7 % at 150).
  // Short problems (one stage lasts < 1 us) and single-tile launches lose 5 % that way (the issuer shares that
  // scheduler); it is a template parameter so that their kernels stay exactly as they were (the same code with a
  // run-time choice cost them 1.9 %).
  constexpr int prod_warp = P15 ? 15 : 0;
  auto run_issuer = [&](const int t) {
    // The issuers sit on scheduler 3 (warp % 4 == 3), the least loaded one: row tile 1 of a 200-row
    // instance has no rows in lane quarter 3, so its softmax warp 19 is free to be tile 1's issuer
    // (when tile 1 does use that quarter, warp 2 issues instead).
    const uint32_t bt = bar0 + 8u * (uint32_t)(t * TB_PER_TILE);
    auto BAR = [&](int i) { return bt + 8u * (uint32_t)i; };
    const uint32_t ring0 = tc::smem_u32(rings + (size_t)t * TCD_STAGES * TCD_STAGE);
    if (t < n_tiles) {
      {
        // ================================================================== MMA issuer of tile t
        // (the whole warp runs the control flow so that addresses stay in uniform registers; one
        // elected lane issues the tcgen05 instructions).  The two key halves X = A, B of a tile are
        // independent sub-pipelines: as soon as half X has published P_h^X the issuer runs its turn
        //   O_X = P_h^X V_h[keys of X]   and   S_{h+1}^X = Q_{h+1} K_{h+1}[keys of X]^T
        // without waiting for the other half.
        const uint32_t idK[2] = {tc::idesc_f16(K0), tc::idesc_f16(K1)};
        const uint32_t idO = tc::idesc_f16(16);
        const uint32_t stg = tc::smem_u32(staging) + t * TCD_TILE_STAGING;
        const uint32_t dS = tmem_base + (uint32_t)(t * NP);
        const uint32_t dO = tmem_base + (uint32_t)(TCD_OCOL + 32 * t);
        const int n0 = nks >> 1;
        uint32_t it = 0;
        // DEBUG builds: cycles of this issuer per phase (0 ring-stage wait, 1 idle polling for P,
        // 2 P V / S turns, 3 wait Q, 4 wait O staged, 5 logits issue)
        long long tph[6] = {0, 0, 0, 0, 0, 0};
        long long tlast = DEBUG ? clock64() : 0;
        auto tick = [&](int ph) {
          if (DEBUG) {
            const long long now = clock64();
            tph[ph] += now - tlast;
            tlast = now;
          }
        };
        auto take = [&]() {                      // next ring stage, in stream order
          const uint32_t s = it % TCD_STAGES, use = it / TCD_STAGES;
          tick(2);
          tc::mbar_wait(BAR(TB_FULL + s), use & 1);
          tc::fence_after();
          tick(0);
          ++it;
          return s;
        };
        // every tcgen05.mma / commit sits inside an elect.sync block: one lane issues, and inside
        // the block ptxas keeps the operands in uniform registers (no per-lane predicates)
        auto release = [&](uint32_t s) {
          if (tc::elect_one()) tc::commit(BAR(TB_EMPTY + s));
          __syncwarp();
        };
        // S_h^X into the TMEM columns of half X; ka = ring slot holding K hi/lo of heads 2p,2p+1, keys of X
        auto issue_S = [&](int X, int h, uint32_t ka) {
          const uint32_t a_hi = (stg >> 4) + (h >> 2) * 1024 + (h & 3) * 2, a_lo = a_hi + 2048;
          const uint32_t b_hi = ((ring0 + ka * TCD_STAGE) >> 4) + (h & 1) * 4, b_lo = b_hi + 2;
          const uint32_t d = dS + (X ? K0 : 0);
          if (tc::elect_one()) {
            tc::mma_ss_lo(d, a_hi, b_hi, idK[X], 0);
            tc::mma_ss_lo(d, a_hi, b_lo, idK[X], 1);
            tc::mma_ss_lo(d, a_lo, b_hi, idK[X], 1);
            tc::commit(BAR(TB_SFULL + X));
          }
          __syncwarp();
        };
        // O_X = P_h^X V_h: A = P hi/lo in TMEM (8 + 8 columns per 16 keys), B = V^T of head h in slot sv
        auto issue_PV = [&](int X, uint32_t sv) {
          const uint32_t vb = (ring0 + sv * TCD_STAGE) >> 4;
          const int k_lo = X ? n0 : 0, k_hi = X ? nks : n0;
          if (tc::elect_one()) {
#pragma unroll 1
            for (int ks = k_lo; ks < k_hi; ++ks) {
              const uint32_t v_hi = vb + (ks >> 2) * 256 + (ks & 3) * 2, v_lo = v_hi + 128;
              const uint32_t p_hi = dS + 16 * ks, p_lo = p_hi + 8;
              tc::mma_ts_lo(dO + 16 * X, p_hi, v_hi, idO, ks != k_lo);
              tc::mma_ts_lo(dO + 16 * X, p_hi, v_lo, idO, 1);
              if (!PSINGLE) tc::mma_ts_lo(dO + 16 * X, p_lo, v_hi, idO, 1);
            }
            tc::commit(BAR(TB_OFULL + X));
          }
          __syncwarp();
        };
        for (int step = 0; step < n_steps; ++step) {
          uint32_t ka[2];                        // ring slots of the K blocks (key halves A, B) in use
          ka[0] = take();
          ka[1] = take();
          tc::mbar_wait(BAR(TB_QFULL), step & 1);
          tc::fence_after();
          tick(3);
          issue_S(0, 0, ka[0]);
          issue_S(1, 0, ka[1]);
          for (int h = 0; h < NH; ++h) {
            const uint32_t sv = take();          // V^T of head h
            uint32_t kn[2] = {ka[0], ka[1]};
            if ((h & 1) && h + 1 < NH) {         // K blocks of the next head pair follow in the stream
              kn[0] = take();
              kn[1] = take();
            }
            const uint32_t par = (uint32_t)(step * NH + h) & 1u;
            bool done[2] = {false, false};
            uint32_t idle = 0;
            while (!(done[0] && done[1])) {
              bool any = false;
#pragma unroll
              for (int X = 0; X < 2; ++X) {
                if (!done[X] && tc::mbar_test(BAR(TB_PFULL + X), par)) {
                  tc::fence_after();
                  tick(1);
                  issue_PV(X, sv);
                  if (h + 1 < NH) {
                    issue_S(X, h + 1, kn[X]);
                    if (((h + 1) & 1)) release(kn[X]);   // second head of the pair: K block of X is done
                  }
                  done[X] = true;
                  any = true;
                  tick(2);
                }
              }
              if (!any && ++idle > (1u << 26)) __trap();
            }
            release(sv);
            ka[0] = kn[0];
            ka[1] = kn[1];
          }
          // pointer logits U = O K2^T into the (consumed) S columns
          for (int i = 0; i < 4; ++i) {
            const int half = i & 1;                    // O / K2 columns 64*half .. +63
            const uint32_t o_hi = (stg >> 4) + half * 1024, o_lo = o_hi + 2048;
            for (int gi = 0; gi < 2; ++gi) {
              const uint32_t sk = take();
              if (i == 0 && gi == 0) {
                tc::mbar_wait(BAR(TB_OREADY), step & 1);
                tc::fence_after();
                tick(4);
              }
              const uint32_t kb = (ring0 + sk * TCD_STAGE) >> 4;
              const uint32_t d = dS + (gi ? K0 : 0);
              if (tc::elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  tc::mma_ss_lo(d, o_hi + kk * 2, kb + kk * 2, idK[gi], (i | kk) != 0);
                  if (i < 2) tc::mma_ss_lo(d, o_lo + kk * 2, kb + kk * 2, idK[gi], 1);
                }
                tc::commit(BAR(TB_EMPTY + sk));
              }
              __syncwarp();
            }
          }
          if (tc::elect_one()) tc::commit(BAR(TB_UFULL));
          __syncwarp();
          tick(5);
        }
        if (DEBUG && b == 0 && lane == 0) {
          float* dT = A.dbg + (size_t)2 * 8 * 128 * NP + (size_t)256 * 128 + (size_t)256 * NP + (16 + t) * 8;
          for (int i = 0; i < 6; ++i) dT[i] = (float)tph[i];
        }
      }
    }
  };

  auto run_producer = [&]() {
    // ================================================================== operand producer (both row tiles)
    if (lane == 0) {
      const size_t szk = tcd_szk(NP), szvh = tcd_szvh(NP), pair = tcd_pair_bytes(NP);
      const int per_step = 24;                             // ring stages one tile consumes per decode step
      // back-off of the polling lane when no stage is free: a long sleep leaves scheduler 0's issue slots to its
      // softmax warps (+0.9 % at N = 200), but it must stay well below the time one stage lasts (1.7 us at N = 200,
      // 0.9 us at N = 100: 512 ns cost 4 % there, 1024 ns cost 4 % at N = 200)
      const unsigned poll_ns = (prod_warp == 0 && NP >= 144) ? 512u : 64u;     // on scheduler 3 a short back-off is free
      const uint32_t total = (uint32_t)n_steps * (uint32_t)per_step;
      uint32_t it[2] = {0, 0};
      int rem = n_tiles;
      uint32_t idle = 0;
      while (rem > 0) {
        bool pushed = false;
        for (int t = 0; t < n_tiles; ++t) {
          if (it[t] >= total) continue;
          const uint32_t bt = bar0 + 8u * (uint32_t)(t * TB_PER_TILE);
          const uint32_t s = it[t] % TCD_STAGES, use = it[t] / TCD_STAGES;
          if (!tc::mbar_test(bt + 8 * (TB_EMPTY + s), (use & 1) ^ 1)) continue;
          // i-th stage of the step: per head pair p [K_p half A, K_p half B, V 2p, V 2p+1], then the four K2
          // blocks by key half
          const int i = (int)(it[t] % (uint32_t)per_step);
          size_t off;
          uint32_t bytes;
          if (i < 16) {
            const int p = i >> 2, j = i & 3;
            if (j < 2) { off = (size_t)p * pair + (j ? (size_t)K0 * 128 : 0); bytes = (uint32_t)(j ? K1 : K0) * 128u; }
            else { off = (size_t)p * pair + szk + (size_t)(j - 2) * szvh; bytes = (uint32_t)szvh; }
          } else {
            const int i2 = i - 16, blk = i2 >> 1, gi = i2 & 1;
            off = 4 * pair + (size_t)blk * szk + (gi ? (size_t)K0 * 128 : 0);
            bytes = (uint32_t)(gi ? K1 : K0) * 128u;
          }
          tc::mbar_expect_tx(bt + 8 * (TB_FULL + s), bytes);
          tc::bulk_g2s(tc::smem_u32(rings + ((size_t)t * TCD_STAGES + s) * TCD_STAGE), imgb + off, bytes,
                       bt + 8 * (TB_FULL + s));
          if (++it[t] == total) --rem;
          pushed = true;
        }
        if (pushed) idle = 0;
        else if (++idle > (1u << 28)) __trap();
        else __nanosleep(poll_ns);
      }
    }
  };
  auto run_softmax = [&]() {
    // ================================================================== softmax / state warps
    // Two warps share every 32-row lane quarter of a tile ("split-K"): the primary (half 0) owns key
    // chunks [0,n0), the secondary (half 1) chunks [n0,nch).  Each runs its own online softmax
    // (own shift m, own row sum l) and its P V partial lands in its own accumulator O_half; the
    // primary merges the two partials when it drains them (O = sum_i 2^(m_i-M) O_i / sum_i 2^(m_i-M) l_i).
    // Only the primary carries the rollout state machine; it publishes the chosen node through a
    // TMEM exchange column so that the secondary can keep its copy of the visited mask.
    const int sw = warp - 4;
    const int t = sw >> 3, half = (sw >> 2) & 1, q = warp & 3;
    const int rows_t = min(128, G - 128 * t);
    if (q * 32 < rows_t) {
      const uint32_t bt = bar0 + 8u * (uint32_t)(t * TB_PER_TILE);
      auto BAR = [&](int i) { return bt + 8u * (uint32_t)i; };
      const int row = q * 32 + lane;                 // row inside the tile = TMEM lane
      const int g = t * 128 + row;
      const bool valid = g < G;
      const int gc = min(g, G - 1);
      const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
      const uint32_t tS = tmem_base + lane_addr + (uint32_t)(t * NP);
      const uint32_t tO = tmem_base + lane_addr + (uint32_t)(TCD_OCOL + 32 * t);
      const uint32_t tX = tmem_base + lane_addr + (uint32_t)(TCD_XCOL + 16 * t);
      unsigned char* stg = staging + t * TCD_TILE_STAGING;
      const int src = clamp_index(A.src[b], N), tgt = clamp_index(A.tgt[b], N);
      const int first = clamp_index(A.first[r0 + gc], N);
      int32_t* pirow = A.pi + ((size_t)b * GT + r0 + gc) * N;
      const int nch = NP >> 4, n0 = nch >> 1;
      const int c_lo = half ? n0 : 0, c_hi = half ? nch : n0;

      unsigned vis[7];
#pragma unroll
      for (int w = 0; w < 7; ++w) {           // padded keys (>= N) are permanently visited
        const int lo = w * 32;
        vis[w] = (lo + 32 <= N) ? 0u : ((N <= lo) ? 0xffffffffu : (0xffffffffu << (N - lo)));
      }
      int cur = 0, cnt = 0;
      bool nocand = false;                           // a decode step found no finite logit: the reward is NaN
      auto mark = [&](int a) {
#pragma unroll
        for (int w = 0; w < 7; ++w) vis[w] |= ((a >> 5) == w) ? (1u << (a & 31)) : 0u;
      };
      // visit(a) with source<->target coupling (train_path_solver.py:45-66); only the primary records pi
      auto visit = [&](int a) {
        if (valid && half == 0 && cnt < N) pirow[cnt] = a;     // (a forced tour, or a row without finite logits, may repeat nodes)
        mark(a);
        cnt++;
        cur = a;
        const int partner = (a == src) ? tgt : ((a == tgt) ? src : -1);
        if (partner >= 0) {
          if (valid && half == 0 && cnt < N) pirow[cnt] = partner;
          mark(partner);
          cnt++;
          cur = partner;
        }
      };

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

      // ---- online softmax of this warp's key chunks of head h; publishes (m, l) and arrives on PFULL
      auto softmax_pass = [&](int step, int h) {
        float m = valid ? kTcdFloor : 0.f, l = 0.f;
        uint32_t sA[32];             // only the first 16 entries are used (16-key chunks)
        // one 16-key chunk at column c0: mask, 2^(s-m), row sum, fp16 hi/lo split, store in place;
        // per 16 keys: 8 columns of hi pairs, then 8 columns of lo pairs (one MMA k-step each)
        auto chunk = [&](int c0, uint32_t (&s)[32], unsigned bits) {
          float p[16];
          if (DEBUG && b == 0 && step == A.dbg_step) {
            float* dS = A.dbg + ((size_t)(t * 8 + h) * 128 + row) * NP + c0;
#pragma unroll
            for (int j = 0; j < 16; ++j) dS[j] = __uint_as_float(s[j]);
          }
          float tv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) tv[j] = ((bits >> j) & 1u) ? -INFINITY : __uint_as_float(s[j]);
          float csum;
          uint32_t out[32];
          auto expo = [&]() {
            const float2 nm = make_float2(-m, -m);
            if (!PSINGLE) {
              float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float2 d0 = __fadd2_rn(make_float2(tv[j], tv[j + 1]), nm);
                const float2 d1 = __fadd2_rn(make_float2(tv[j + 2], tv[j + 3]), nm);
                p[j] = ex2_approx(d0.x);
                p[j + 1] = ex2_approx(d0.y);
                p[j + 2] = ex2_approx(d1.x);
                p[j + 3] = ex2_approx(d1.y);
                acc0 = __fadd2_rn(acc0, make_float2(p[j], p[j + 1]));
                acc1 = __fadd2_rn(acc1, make_float2(p[j + 2], p[j + 3]));
              }
              acc0 = __fadd2_rn(acc0, acc1);
              csum = acc0.x + acc0.y;
            } else {
              // single fp16 P: pack first, then sum the rounded halves (one FHADD each, four chains)
              float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float2 d0 = __fadd2_rn(make_float2(tv[j], tv[j + 1]), nm);
                const float2 d1 = __fadd2_rn(make_float2(tv[j + 2], tv[j + 3]), nm);
                const __half2 h0 = __floats2half2_rn(ex2_approx(d0.x), ex2_approx(d0.y));
                const __half2 h1 = __floats2half2_rn(ex2_approx(d1.x), ex2_approx(d1.y));
                const uint32_t u0 = *reinterpret_cast<const uint32_t*>(&h0), u1 = *reinterpret_cast<const uint32_t*>(&h1);
                out[j >> 1] = u0;
                out[(j >> 1) + 1] = u1;
                a0 = tcd_add_h_f(u0, 0, a0);
                a1 = tcd_add_h_f(u0, 1, a1);
                a2 = tcd_add_h_f(u1, 0, a2);
                a3 = tcd_add_h_f(u1, 1, a3);
              }
              csum = (a0 + a1) + (a2 + a3);
            }
          };
          // rows that have not seen a finite score yet (always the case in the first chunk) take their
          // shift from this chunk's maximum up front instead of tripping the overflow path below
          if (__any_sync(0xffffffffu, valid && m == kTcdFloor)) {
            float cm = tv[0];
#pragma unroll
            for (int j = 1; j < 16; ++j) cm = fmaxf(cm, tv[j]);
            if (valid && m == kTcdFloor && cm > kTcdFloor) m = ceilf(cm);
          }
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
              for (int cc = 16 * c_lo; cc < c0; cc += 16) {
                uint32_t r[16];
                tc::ld16(tS + cc, r);      // PSINGLE: columns 8-15 are stale scores, rescaled but never read
                tc::ld16_wait(r);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  float2 f = __half22float2(*reinterpret_cast<__half2*>(&r[j]));
                  __half2 hh = __floats2half2_rn(f.x * alpha, f.y * alpha);
                  r[j] = *reinterpret_cast<uint32_t*>(&hh);
                }
                tc::st16(tS + cc, r);
              }
            }
            l *= alpha;
            m = m_new;
            expo();
          }
          l += csum;
          if (!PSINGLE) {
#pragma unroll
            for (int j = 0; j < 16; j += 2)       // x = hi + lo; the residual hi - x is one exact FHADD per element
              tcd_split_pair(p[j], p[j + 1], out[j >> 1], out[8 + (j >> 1)]);
            tst<16>(tS + c0, out);
          } else {
            tst<8>(tS + c0, out);
          }
        };
        // 16-key chunks, two per 32-bit word of the visited mask; four warps per scheduler hide the
        // TMEM load latency
#pragma unroll 1
        for (int w = c_lo >> 1; 2 * w < c_hi; ++w) {
          const unsigned word = tcd_word(vis, w);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 2 * w + e;
            if (c >= c_lo && c < c_hi) {
              tld<16>(tS + 16 * c, sA);
              tld_wait16(sA);
              chunk(16 * c, sA, e ? (word >> 16) : (word & 0xffffu));
            }
          }
        }
        return make_float2(m, l);
      };
      // publish (m, l) of this half for the merge at drain time, then P itself
      auto publish = [&](int h, float2 ml) {
        const float m = ml.x, l = ml.y;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(tX + (h & 1) * 4 + half * 2),
                     "r"(__float_as_uint(m)), "r"(__float_as_uint(l))
                     : "memory");
        tc::st_wait();
        tc::fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(BAR(TB_PFULL + half));
      };

      const float4 qfix4 = __ldg((const float4*)(A.qfix + (size_t)b * H) + lane);
      const uint32_t pair_bar = 1u + (uint32_t)(t * 4 + q);     // named barrier of this warp pair (64 threads)
      auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory"); };
      // query rows -> A tile (fp16 hi/lo, K-major SWIZZLE_128B):
      // q = ((q_graph+q_source+q_target) + q_first) + q_last (models.py:413), 1/sqrt(dk)*log2(e) folded in.
      // The primary gathers rows 0-15 of the lane quarter, the secondary rows 16-31; lane l carries
      // columns 4l..4l+3 of every row; 8 rows (16 L2 loads) are in flight at a time.
      auto gather_q = [&]() {
        const uint32_t sub = (uint32_t)(lane >> 4) * 16384u + (uint32_t)(lane & 1) * 8u;
        const uint32_t chunk = (uint32_t)(lane & 15) >> 1;
        const int nrows = min(32, rows_t - q * 32);
        const int r_end = half ? nrows : min(16, nrows);
#pragma unroll 1
        for (int r0 = 16 * half; r0 < r_end; r0 += 8) {
          float4 ql[8], qf[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int cr = __shfl_sync(0xffffffffu, cur, r0 + i), fr = __shfl_sync(0xffffffffu, first, r0 + i);
            ql[i] = __ldg((const float4*)(projb + (size_t)cr * PROJ + 2 * H) + lane);
            qf[i] = __ldg((const float4*)(projb + (size_t)fr * PROJ + 3 * H) + lane);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (r0 + i < r_end) {
              uint32_t h0, l0, h1, l1;
              split2(((qfix4.x + qf[i].x) + ql[i].x) * kScoreScale, ((qfix4.y + qf[i].y) + ql[i].y) * kScoreScale, h0, l0);
              split2(((qfix4.z + qf[i].z) + ql[i].z) * kScoreScale, ((qfix4.w + qf[i].w) + ql[i].w) * kScoreScale, h1, l1);
              unsigned char* p = stg + sub + tc::sw128_off((uint32_t)(q * 32 + r0 + i), chunk);
              *(uint2*)p = make_uint2(h0, h1);
              *(uint2*)(p + 32768) = make_uint2(l0, l1);
            }
          }
        }
        tc::fence_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(BAR(TB_QFULL));
      };

      // step 0: no decoder call (train_path_solver.py:256-257)
      visit(first);
      gather_q();

      // drain O_h (primary): merge the two split-K partials, /l -> fp16 hi/lo -> A tile columns
      // 16h..16h+15 (the consumed Q_h slot)
      auto drain_o = [&](int step, int h) {
        tc::mbar_wait(BAR(TB_OFULL), (step * NH + h) & 1);
        tc::mbar_wait(BAR(TB_OFULL + 1), (step * NH + h) & 1);
        tc::fence_after();
        uint32_t oa[16], ob[16], x0, x1, x2, x3;
        tc::ld16(tO, oa);
        tc::ld16(tO + 16, ob);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                     : "r"(tX + (h & 1) * 4)
                     : "memory");
        tc::ld16_wait(oa);
        tc::ld16_wait(ob);
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3)::"memory");
        const float ma = __uint_as_float(x0), la = __uint_as_float(x1);
        const float mb = __uint_as_float(x2), lb = __uint_as_float(x3);
        const float M = fmaxf(ma, mb);
        // a half without a finite score has l = 0 and contributes nothing
        const float sa = (la > 0.f) ? ex2_approx(ma - M) : 0.f;
        const float sb = (lb > 0.f) ? ex2_approx(mb - M) : 0.f;
        float inv;         // MUFU reciprocal (1 ulp): the normalisation error is far below the fp16x2 split
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(la * sa + lb * sb));
        const float wa = sa * inv, wb = sb * inv;
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = (sa > 0.f) ? __uint_as_float(oa[j]) * wa : 0.f;
          v[j] = (sb > 0.f) ? fmaf(__uint_as_float(ob[j]), wb, a) : a;
        }
        if (DEBUG && b == 0 && step == A.dbg_step) {
          float* dO = A.dbg + (size_t)2 * 8 * 128 * NP + (size_t)(t * 128 + row) * 128 + 16 * h;
#pragma unroll
          for (int j = 0; j < 16; ++j) dO[j] = v[j];
        }
        uint4 h0, l0, h1, l1;
        tcd_split8_packed(v, h0, l0);
        tcd_split8_packed(v + 8, h1, l1);
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
          tc::mbar_wait(BAR(TB_SFULL + half), (step * NH + h) & 1);
          tc::fence_after();
          tick(0);
          const float2 ml = softmax_pass(step, h);
          tick(1);
          if (half == 0) {
            // the previous head's partials must be out of the O accumulators before this head's P V
            // overwrites them; by now both P V of head h-1 are long done, so the drain does not wait.
            // Then tell the issuer that O_B is free (second arrival on half B's barrier).
            if (h > 0) drain_o(step, h - 1);
            tc::fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(BAR(TB_PFULL + 1));
            tick(2);
          }
          publish(h, ml);
        }
        if (half == 0) {
          drain_o(step, NH - 1);
          tc::fence_async_smem();
          tc::fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(BAR(TB_OREADY));
          tick(2);
        }

        // ------------------------------------------------------------ pointer logits + arg-max
        // Each warp of the pair scans its own key chunks and reduces them to one candidate
        // (clipped logit, key); the primary merges the two (ties -> lower key = its own).
        tc::mbar_wait(BAR(TB_UFULL), step & 1);
        tc::fence_after();
        tick(3);
        float best = -INFINITY;
        int idx = 0x7fffffff;
        float* lo_out = nullptr;
        if (LOGITS && A.logits_out != nullptr)
          lo_out = A.logits_out + (((size_t)b * GT + r0 + gc) * n_steps + step) * N;
        // pass(EXACT): EXACT compares the clipped logits 10*tanh(.) themselves (models.py:438-440);
        // otherwise the monotone pre-activation is compared, which picks the same key except when two
        // clipped logits round to the same float (handled by re-running EXACT near the clip).
        // 16 independent running maxima (one per column slot of a 16-key chunk), merged at the end
        // with the lowest key winning ties (torch.argmax returns the first maximum).
        auto pass = [&](auto exact_tag) {
          constexpr bool EXACT = decltype(exact_tag)::value;
          float bj[16];
          int cj[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { bj[j] = -INFINITY; cj[j] = c_lo; }
          uint32_t u[32];
#pragma unroll 1
          for (int c = c_lo; c < c_hi; ++c) {
            tld<16>(tS + 16 * c, u);
            const unsigned bits = (tcd_word(vis, c >> 1) >> ((c & 1) * 16)) & 0xffffu;
            tld_wait16(u);
            if (DEBUG && b == 0 && step == A.dbg_step) {
              float* dU = A.dbg + (size_t)2 * 8 * 128 * NP + (size_t)256 * 128 + (size_t)(t * 128 + row) * NP + 16 * c;
#pragma unroll
              for (int j = 0; j < 16; ++j) dU[j] = __uint_as_float(u[j]);
            }
            const float4* cc4 = (const float4*)(c0s + 16 * c);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 cc = cc4[j4];
              const float ca[4] = {cc.x, cc.y, cc.z, cc.w};
              uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
              if (SAMPLE)
                rnd = tcd_philox(make_uint4((uint32_t)(4 * c + j4), (uint32_t)step, (uint32_t)(r0 + gc), (uint32_t)b),
                                 make_uint2((uint32_t)sample_seed, (uint32_t)(sample_seed >> 32)));
              const uint32_t ra[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int j = 4 * j4 + jj;
                float z = __uint_as_float(u[j]) + ca[jj];          // EXACT: 10 tanh(z / sqrt(128)); else the
                if (EXACT) z = clip_tanh10(z * kTcdInvSqrtH);       // monotone pre-activation itself
                if (SAMPLE) z += tcd_gumbel(ra[jj]);
                if ((bits >> j) & 1u) z = -INFINITY;
                if (LOGITS && lo_out != nullptr && valid && 16 * c + j < N) lo_out[16 * c + j] = z;
                if (z > bj[j]) { bj[j] = z; cj[j] = c; }
              }
            }
          }
          best = bj[0];
          idx = 16 * cj[0];
#pragma unroll
          for (int j = 1; j < 16; ++j) {
            const int key = 16 * cj[j] + j;
            if (bj[j] > best || (bj[j] == best && key < idx)) { best = bj[j]; idx = key; }
          }
        };
        if (LOGITS || SAMPLE) {
          pass(std::true_type{});
        } else {
          pass(std::false_type{});
          if (__any_sync(0xffffffffu, valid && best * kTcdInvSqrtH > 3.0f)) {
            pass(std::true_type{});
          } else {
            best = (best == -INFINITY) ? best : clip_tanh10(best * kTcdInvSqrtH);   // candidates are exchanged as clipped logits
          }
        }
        tick(4);
        int a;
        if (half == 1) {
          asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(tX + 9), "r"(__float_as_uint(best)),
                       "r"((uint32_t)idx)
                       : "memory");
          tc::st_wait();
          tc::fence_before();
          pair_sync();                // candidate published
          pair_sync();                // the primary has published the chosen node
          tc::fence_after();
          uint32_t av;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(av) : "r"(tX + 8) : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(av)::"memory");
          a = (int)av;
        } else {
          pair_sync();
          tc::fence_after();
          uint32_t bz, bi;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(bz), "=r"(bi) : "r"(tX + 9) : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(bz), "+r"(bi)::"memory");
          if (__uint_as_float(bz) > best) idx = (int)bi;     // ties keep the primary's (lower) key
          // ---- state update (train_path_solver.py:45-72); the chosen node goes to the secondary warp
          a = ((LOGITS || DEBUG) && A.forced_pi) ? clamp_index(A.forced_pi[((size_t)b * GT + r0 + gc) * N + min(cnt, N - 1)], N) : idx;   // teacher forcing: LOGITS / DEBUG kernels only
          if ((unsigned)a >= (unsigned)N) {            // a row without finite logits has no candidate: stay in range, flag the row
            a = cur;
            nocand = true;
          }
          asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tX + 8), "r"((uint32_t)a) : "memory");
          tc::st_wait();
          tc::fence_before();
          pair_sync();
        }
        visit(a);
        tc::fence_before();
        if (step + 1 < n_steps) gather_q();
        tick(5);
      }
      // (a flagged row, or a forced tour that repeats nodes, may have recorded fewer than N nodes: every entry of its
      // tour must still be a node, the tour is read as indices below and by the callers)
      if (valid && half == 0)
        for (; cnt < N; ++cnt) pirow[cnt] = cur;
      // ---- reward = -(closed tour length - fixed edge)   (train_path_solver.py:109-149)
      if (valid && half == 0) {
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
        A.reward[(size_t)b * GT + r0 + g] = nocand ? __int_as_float(0x7fc00000) : -(len - fixed);
      }
      if (DEBUG && b == 0 && lane == 0) {
        float* dT = A.dbg + (size_t)2 * 8 * 128 * NP + (size_t)256 * 128 + (size_t)256 * NP + sw * 8;
        for (int i = 0; i < 6; ++i) dT[i] = (float)tph[i];
      }
    }
  };

  // register budget: the role warpgroup (warps 0-3) hands registers to the softmax warpgroups; the
  // role dispatch is repeated inside each branch so that ptxas sees one register budget per path
  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;" ::: "memory");
    if (!P15 && warp == 0) run_producer();
    else if (warp == 3) run_issuer(0);
    else if (warp == 2 && iss1_warp == 2) run_issuer(1);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;" ::: "memory");
    if (warp == 19 && iss1_warp == 19) run_issuer(1);
    else if (P15 && warp == 15) run_producer();
    else run_softmax();
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

bool decode_tc_supported(int N, int G) { return tcd_np_of(N) <= 208 && tcd_np_of(N) >= 32 && G <= 256; }
size_t decode_tc_pack_bytes(int Bp, int N) { return align_up((size_t)Bp * tcd_img_bytes(tcd_np_of(N)), 256) + 1024; }
size_t decode_tc_dbg_floats(int N) {
  const int NP = tcd_np_of(N);
  return (size_t)2 * 8 * 128 * NP + (size_t)256 * 128 + (size_t)256 * NP + 256;
}

// CTAs per instance: 1 unless the launch would leave most SMs idle; then the rows are cut into 128-, 64- or
// 32-row parts as long as all CTAs still fit in one wave (HTSP_DECODE_SPLIT = 1 | 2 | 4 | 8 overrides: the
// requested number of parts, rounded to whole 32-row warps).
static void tcd_choose_split(int Bp, int G, bool debug, int& split, int& rows_per_cta) {
  split = 1;
  rows_per_cta = G;
  if (debug || G <= 32) return;
  int dev = 0, n_sm = 0;              // queried per call: the library serves several devices of one process
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0)
    n_sm = 148;
  const char* env = getenv("HTSP_DECODE_SPLIT");
  const int forced = env ? atoi(env) : 0;
  for (int want = 8; want >= 2; want >>= 1) {
    if (forced > 0 && want != forced) continue;
    const int rpc = ((G + want - 1) / want + 31) / 32 * 32;
    const int parts = (G + rpc - 1) / rpc;
    if (parts < 2) continue;
    if (forced > 0 || (long long)Bp * parts <= n_sm) {
      split = parts;
      rows_per_cta = rpc;
      return;
    }
  }
}

template <bool LOGITS, bool DEBUG, bool PSINGLE, bool SPLIT, bool SAMPLE = false, bool P15 = false>
static int launch_tcd(const TcdArgs& a, int Bp, cudaStream_t s) {
  const size_t smem = (size_t)2 * TCD_TILE_STAGING + (size_t)2 * TCD_STAGES * TCD_STAGE +
                      (size_t)a.NP * 4 + 2 * TB_PER_TILE * 8 + 16;   // 224 KB + c0 + barriers
  if (smem > 227 * 1024) {
    set_error("decode_tc: N=%d needs %zu bytes of shared memory", a.N, smem);
    return HTSP_ERR_UNSUPPORTED;
  }
  HTSP_CHECK_CUDA(cudaFuncSetAttribute(decode_tc_kernel<LOGITS, DEBUG, PSINGLE, SPLIT, SAMPLE, P15>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kernel_timer_begin(s);
  decode_tc_kernel<LOGITS, DEBUG, PSINGLE, SPLIT, SAMPLE, P15><<<SPLIT ? Bp * ((a.G + a.dbg_step - 1) / a.dbg_step) : Bp, TCD_THREADS, smem, s>>>(a);
  kernel_timer_end(s);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

int launch_decode_tc(const float* x, const float* proj, const float* qfix, const float* c0,
                     const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp, int N,
                     int G, int mode, const int32_t* forced_pi, int32_t* pi, float* reward,
                     float* logits_out, void* pack_ws, float* dbg, int dbg_step, const uint64_t* sample_seed,
                     cudaStream_t s) {
  HTSP_REQUIRE(decode_tc_supported(N, G), "decode_tc shape");
  HTSP_REQUIRE(sample_seed == nullptr || (dbg == nullptr && logits_out == nullptr && forced_pi == nullptr),
               "sampling rollouts take no teacher forcing, logit capture or debug dump");
  const int NP = tcd_np_of(N);
  unsigned char* img = (unsigned char*)(((uintptr_t)pack_ws + 1023) & ~(uintptr_t)1023);
  {
    const int items = max(NP * 32, tcd_kb(NP) * 8 * 128);
    decode_tc_pack_kernel<<<dim3((items + 255) / 256, min(Bp, 65535)), 256, 0, s>>>(proj, Bp, N, NP, img);
    HTSP_CHECK_LAUNCH();
  }
  TcdArgs a;
  a.x = x; a.proj = proj; a.qfix = qfix; a.c0 = c0; a.img = img;
  a.src = src; a.tgt = tgt; a.first = first; a.forced_pi = forced_pi;
  a.pi = pi; a.reward = reward; a.logits_out = logits_out; a.dbg = dbg;
  a.N = N; a.NP = NP; a.G = G; a.fast = (mode == HTSP_MODE_TC_FAST) ? 1 : 0; a.dbg_step = dbg_step;
  int n_split = 1, rows_per_cta = G;
  tcd_choose_split(Bp, G, dbg != nullptr, n_split, rows_per_cta);
  const bool split = n_split > 1;
  if (split) a.dbg_step = rows_per_cta;
  if (sample_seed != nullptr) {
    a.dbg = (float*)(uintptr_t)(*sample_seed);      // see SAMPLE above
    if (a.fast) return split ? launch_tcd<false, false, true, true, true>(a, Bp, s) : launch_tcd<false, false, true, false, true>(a, Bp, s);
    return split ? launch_tcd<false, false, false, true, true>(a, Bp, s) : launch_tcd<false, false, false, false, true>(a, Bp, s);
  }
  if (dbg != nullptr) return a.fast ? launch_tcd<false, true, true, false>(a, Bp, s) : launch_tcd<false, true, false, false>(a, Bp, s);
  if (logits_out != nullptr || forced_pi != nullptr) {     // teacher forcing is served by the LOGITS kernels
    if (a.fast) return split ? launch_tcd<true, false, true, true>(a, Bp, s) : launch_tcd<true, false, true, false>(a, Bp, s);
    return split ? launch_tcd<true, false, false, true>(a, Bp, s) : launch_tcd<true, false, false, false>(a, Bp, s);
  }
  // producer on warp 15 (see prod_warp in the kernel): unsplit greedy launches with two row tiles, tile 1 without
  // rows in lane quarter 3, long rows
  const bool p15 = !split && G > 128 && G - 128 <= 96 && NP >= 144;
  if (p15) return a.fast ? launch_tcd<false, false, true, false, false, true>(a, Bp, s) : launch_tcd<false, false, false, false, false, true>(a, Bp, s);
  if (a.fast) return split ? launch_tcd<false, false, true, true>(a, Bp, s) : launch_tcd<false, false, true, false>(a, Bp, s);
  return split ? launch_tcd<false, false, false, true>(a, Bp, s) : launch_tcd<false, false, false, false>(a, Bp, s);
}

}  // namespace htsp
