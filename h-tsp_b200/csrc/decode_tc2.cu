// K3, second tcgen05 rollout kernel (HTSP_MODE_TC_FAST, 33 <= N <= 208, and 208 < N <= 576 in key blocks -- see the
// template parameter NKB of the kernel): the glimpse probabilities stay in TMEM as the
// fp32 words they were computed in and feed a kind::tf32 MMA in place, and every row quarter is served by THREE
// softmax warps.
//
// Why (measured, tools/lab_tc.cu, tools/ab_rollout.py --g 32 ... 200): the first kernel
// (decode_tc.cu) is bound by the serial chain of one row quarter -- a launch with 32 POMO rows per instance takes
// 76 % of the time of one with 200 -- not by any pipe: S MMA -> 6-7 chunks of softmax in one warp -> drain -> P V
// -> S ... for eight heads, then logits -> arg-max -> gather.  Splitting the keys of a row over more warps
// shortens that chain, but split-fp16 probabilities need a per-warp shift (p must fit fp16) and therefore one
// accumulator per key part, and TMEM is full (2 x 208 score columns).  With P = 2^s kept as fp32 there is no
// shift (the exponent range is that of fp32; keys are centred per head so that scores sit around zero), all key
// parts of a tile accumulate into ONE accumulator, and the softmax inner loop is mask + ex2 + store.
//
// Per step and head h, row tile t (128 rows), key part j (3 parts of 48-80 keys):
//   S_h^j  = Q_h K_h[j]^T          tcgen05.mma kind::f16, split-fp16 (3 MMAs), K block K-major SWIZZLE_64B
//   P_h^j  = 2^(S_h^j) masked      softmax warp (t, j, lane quarter): tcgen05.ld -> ex2 -> tcgen05.st in place;
//                                  row sum of the values rounded to tf32 (exactly what the tensor core reads) in registers
//   O_h   += P_h^j [Vhi | Vlo]     tcgen05.mma kind::tf32, A = P in TMEM, B = V^T tf32 hi/lo stacked (N = 32)
//   O_h / sum_j l_j -> fp16 hi/lo into the shared-memory A tile (part-0 warps, before they publish P_{h+1})
// and once per step U = O K2^T (split-fp16), arg-max of 10 tanh((U + c0)/sqrt(128)) over the unvisited keys by the
// three warps of a row quarter, state update, query gather (PathDecoder.forward, rl4cop/models/models.py:404-446;
// GroupState / Env, rl4cop/train_path_solver.py:45-149, 256-288).
//
// Warps (896 threads): 0, 1 = MMA issuers of row tiles 0, 1 (parts served in the fixed order 0, 1, 2); 2 = operand
// producer (ONE ring of 8 x 12 KB for both tiles: every K / V / K2 block is fetched once per CTA and step, released
// by both issuers); 3 = TMEM allocation; 4-27 = softmax warps (tile, part, lane quarter = warp % 4).  Warps 3, 19, 23,
// 27 (the last three are the softmax warps of lane quarter 3 of tile 1, which has no rows: G <= 224) are issue
// helpers: the logit MMAs U and the first scores S_0 of key parts 1 and 2 write disjoint accumulators, so they are
// issued by their own threads in parallel (a tcgen05.mma costs its issuing thread ~60 cycles whatever its shape).
// Precision: logits within 1e-3 of the fp32 reference (P carries 10 mantissa bits, as a single fp16 would; V, Q, K,
// K2 and O keep their hi/lo splits); a row whose softmax sum leaves the fp32 range returns a NaN reward.
//
// Single-tile variant (template parameter ONE; launches with more CTAs than SMs): one row tile per CTA -- 512
// threads, 256 TMEM columns, a four-slot ring, 113 KB of shared memory -- and TWO CTAs per SM, so that an SM runs two
// decode chains whatever the number of POMO rows (instances with G <= 128 leave half of a two-tile CTA idle).
// Same arithmetic in the same order: tours, rewards and logits are bit-identical to the two-tile CTA's.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "mma_util.cuh"
#include "tc_util.cuh"
#include "tc_rollout_util.cuh"

namespace htsp {

constexpr int T2_THREADS = 896;
constexpr int T2_PARTS = 3;
constexpr int T2_SLOTS = 8;
// single-tile CTAs (template parameter ONE): one 128-row tile per CTA, TWO CTAs per SM -- half of everything
constexpr int T2_THREADS_ONE = 512;
constexpr int T2_SLOTS_ONE = 4;
#ifndef T2_SLOTS_KB
#define T2_SLOTS_KB 4                     // key-block kernels: ring slots of 8 KB (five would fit with two CTAs per SM:
                                          // N = 500 +2 %, N = 300 -1 %, both inside the run-to-run spread)
#endif
constexpr int T2_PER_STEP = 3 + 7 * 6 + 3 + 12;   // ring stages of one decode step
constexpr int T2_SLOT = 12288;            // ring slot: one K_h^j (<= 5 KB), V_h^j (<= 12 KB) or K2_i^j (<= 10 KB) block
constexpr int T2_SLOT_KB = 8192;          // key-block kernels: blocks of <= 192 keys have parts of <= 64 keys (K 4 KB, V 8 KB, K2 8 KB)
constexpr int T2_KB_MAX_NPB = 192;
constexpr int T2_TILE_STAGING = 65536;    // per row tile: [hi k0-63 | hi k64-127 | lo k0-63 | lo k64-127], 16 KB each
constexpr int T2_OCOL = 416;              // O accumulator of tile t: columns 416 + 32 t .. + 31 ([P Vhi | P Vlo])
constexpr int T2_XCOL = 480;              // exchange columns of tile t: 480 + 16 t: l[head & 1][part] (6), candidates (4), chosen (1)
#ifndef T2_GATHER_ROWS
#define T2_GATHER_ROWS 4                  // query rows in flight per gather round (6 spills: 23.0 vs 22.65 ms per 592 instances)
#endif
#ifndef T2_POLY
#define T2_POLY 0                         // of every 16 exponentials, this many go through the FMA-pipe polynomial
#endif
#ifndef T2_KB_CTAS_PER_SM
#define T2_KB_CTAS_PER_SM 2               // key-block kernels (N > 208): CTAs per SM the register budget is set for
#endif
constexpr float kT2InvSqrtH = 0.08838834764831845f;   // 1/sqrt(128)

// key parts in 16-key chunks, as even as possible (part 0, whose warps also drain O, is never the longest)
struct T2Parts {
  int cb[T2_PARTS + 1];
};
__host__ __device__ inline T2Parts t2_parts(int NP) {
  const int nch = NP >> 4;
  T2Parts p;
  int c1 = nch / 3;              // N = 200: 4 + 4 + 5 chunks (measured: 23.75 ms per 592 instances; 3 + 5 + 5: 23.9,
  if (c1 < 1) c1 = 1;            // 3 + 4 + 6, 2 + 5 + 6, 4 + 3 + 6: 24.2 - 24.3)
  p.cb[0] = 0;
  p.cb[1] = c1;
  p.cb[2] = c1 + (nch - c1) / 2;
  p.cb[3] = nch;
  return p;
}
// ---- per-instance operand image (bytes) ----------------------------------------------------------
//   K  : 8 heads x [NP keys x 64 B]   K-major SWIZZLE_64B, row = key: hi d0-15 | lo d0-15 (fp16), keys centred
//   V  : 8 heads x 3 parts x blocks of 32 keys [32 rows (hi d0-15 | lo d0-15) x 128 B] tf32, K-major SWIZZLE_128B
//   K2 : 4 blocks (hi cols 0-63, hi cols 64-127, lo cols 0-63, lo cols 64-127) x [NP keys x 128 B] SWIZZLE_128B
__host__ __device__ inline int t2_vblocks(int keys) { return (keys + 31) >> 5; }
__host__ __device__ inline size_t t2_szk(int NP) { return (size_t)NP * 64; }
__host__ __device__ inline size_t t2_szv(int NP) {
  const T2Parts p = t2_parts(NP);
  size_t b = 0;
  for (int j = 0; j < T2_PARTS; ++j) b += (size_t)t2_vblocks(16 * (p.cb[j + 1] - p.cb[j])) * 4096;
  return b;
}
__host__ __device__ inline size_t t2_voff(int NP, int j) {
  const T2Parts p = t2_parts(NP);
  size_t b = 0;
  for (int i = 0; i < j; ++i) b += (size_t)t2_vblocks(16 * (p.cb[i + 1] - p.cb[i])) * 4096;
  return b;
}
__host__ __device__ inline size_t t2_szk2(int NP) { return (size_t)NP * 128; }
// Key blocks (N > 208): the keys are processed in NKB blocks of NPB <= 208 keys (NPT = NKB * NPB padded keys in all);
// K and K2 keep one [NPT keys x row] block per head / column block (a key block is a row range of it), V has its
// part blocks per (head, key block).
__host__ __device__ inline size_t t2_img_bytes(int NPB, int NKB = 1) {
  return 8 * t2_szk(NPB * NKB) + 8 * (size_t)NKB * t2_szv(NPB) + 4 * t2_szk2(NPB * NKB);
}
__host__ __device__ constexpr int t2_per_step(int NKB) { return 3 + (8 * NKB - 1) * 6 + 3 + 12 * NKB; }

// stage i of a decode step's operand stream -- (offset into the instance's image, bytes), looked up by the producer.
// With it = h * NKB + kb running over (head, key block): K_0^{0,1,2}; per it: V_it^j, K_{it+1}^j (it + 1 < 8 NKB) for
// j = 0,1,2; then per key block kb: K2_i^j for i = 0..3, j = 0,1,2.
__host__ __device__ inline uint2 t2_stage_entry(int NPB, int i, int NKB = 1) {
  const T2Parts parts = t2_parts(NPB);
  const int NPT = NPB * NKB, NIT = 8 * NKB;
  const size_t szk = t2_szk(NPT), szvb = t2_szv(NPB), szk2 = t2_szk2(NPT);
  const int c1 = parts.cb[1], c2 = parts.cb[2], c3 = parts.cb[3];
  auto keys_of = [&](int j) { return (uint32_t)(16 * (j == 0 ? c1 : (j == 1 ? c2 - c1 : c3 - c2))); };
  auto krow_of = [&](int j) { return (uint32_t)(16 * (j == 0 ? 0 : (j == 1 ? c1 : c2))); };
  auto vbytes_of = [&](int j) { return (uint32_t)t2_vblocks((int)keys_of(j)) * 4096u; };
  auto voff_of = [&](int j) { return (size_t)((j > 0 ? vbytes_of(0) : 0u) + (j > 1 ? vbytes_of(1) : 0u)); };
  auto k_of = [&](int it, int j) { return (size_t)(it / NKB) * szk + ((size_t)(it % NKB) * NPB + krow_of(j)) * 64; };
  auto v_of = [&](int it, int j) { return 8 * szk + ((size_t)(it / NKB) * NKB + (size_t)(it % NKB)) * szvb + voff_of(j); };
  size_t off;
  uint32_t bytes;
  if (i < 3) {                                   // K_0^j
    off = k_of(0, i);
    bytes = keys_of(i) * 64u;
  } else if (i < 3 + 6 * (NIT - 1)) {            // it = 0 .. NIT-2: V_it^j, K_{it+1}^j
    const int r = i - 3, it = r / 6, e = r - 6 * it, j = e >> 1;
    if ((e & 1) == 0) { off = v_of(it, j); bytes = vbytes_of(j); }
    else { off = k_of(it + 1, j); bytes = keys_of(j) * 64u; }
  } else if (i < 6 + 6 * (NIT - 1)) {            // last (head, key block): V^j
    const int j = i - (3 + 6 * (NIT - 1));
    off = v_of(NIT - 1, j);
    bytes = vbytes_of(j);
  } else {                                       // K2_i^j of key block kb
    const int r = i - (6 + 6 * (NIT - 1)), q = r / 3, j = r - 3 * q, kb = q >> 2, blk = q & 3;
    off = 8 * szk + 8 * (size_t)NKB * szvb + (size_t)blk * szk2 + ((size_t)kb * NPB + krow_of(j)) * 128;
    bytes = keys_of(j) * 128u;
  }
  return make_uint2((uint32_t)off, bytes);
}

// mean key of every head (the centring vector): kbar[b][c] = mean_n K[b, n, c]; block 0 also writes the stage table
// that the producers of the single-tile CTAs read (they have no shared memory to spare for it)
__global__ void __launch_bounds__(128) decode_tc2_kbar_kernel(const float* __restrict__ proj, int N, int NPB, int NKB,
                                                              float* __restrict__ kbar, uint2* __restrict__ stage_tab) {
  const int b = blockIdx.x, c = threadIdx.x;
  if (b == 0)
    for (int i = c; i < t2_per_step(NKB); i += 128) stage_tab[i] = t2_stage_entry(NPB, i, NKB);
  const float* pb = proj + (size_t)b * N * PROJ + c;
  float s = 0.f;
  for (int n = 0; n < N; ++n) s += __ldg(pb + (size_t)n * PROJ);
  kbar[(size_t)b * H + c] = s / (float)N;
}

// proj fp32 [Bp*N, 640] (K | V | QL | QF | K2) -> operand image.  grid = (blocks per instance, Bp)
__global__ void __launch_bounds__(256)
decode_tc2_pack_kernel(const float* __restrict__ proj, const float* __restrict__ kbar, int Bp, int N, int NPB, int NKB,
                       unsigned char* __restrict__ img) {
  const int NP = NPB * NKB;                 // all padded keys
  const size_t szk = t2_szk(NP), szvb = t2_szv(NPB), szv = (size_t)NKB * szvb, szk2 = t2_szk2(NP), per = t2_img_bytes(NPB, NKB);
  const T2Parts parts = t2_parts(NPB);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  for (int b = blockIdx.y; b < Bp; b += gridDim.y) {
    const float* pb = proj + (size_t)b * N * PROJ;
    unsigned char* base = img + (size_t)b * per;
    // K (centred) and K2: one thread per (key r, matrix m in {K, K2}, 8-column chunk c16)
    if (i < NP * 32) {
      const int c16 = i & 15, m = (i >> 4) & 1, r = i >> 5;
      float v[8];
      if (r < N) {
        const float4* p = (const float4*)(pb + (size_t)r * PROJ + (m == 0 ? 0 : 4 * H) + c16 * 8);
        const float4 a = __ldg(p), c = __ldg(p + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
        if (m == 0) {
          const float4* kb = (const float4*)(kbar + (size_t)b * H + c16 * 8);
          const float4 ka = __ldg(kb), kc = __ldg(kb + 1);
          v[0] -= ka.x; v[1] -= ka.y; v[2] -= ka.z; v[3] -= ka.w; v[4] -= kc.x; v[5] -= kc.y; v[6] -= kc.z; v[7] -= kc.w;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.f;
      }
      uint4 hi, lo;
      tcd_split8(v, hi, lo);
      if (m == 0) {
        const int h = c16 >> 1;
        unsigned char* blk = base + (size_t)h * szk;
        *(uint4*)(blk + tc::sw64_off(r, (c16 & 1))) = hi;
        *(uint4*)(blk + tc::sw64_off(r, 2 + (c16 & 1))) = lo;
      } else {
        unsigned char* blk = base + 8 * szk + 8 * szv + (size_t)(c16 >> 3) * szk2;
        *(uint4*)(blk + tc::sw128_off(r, c16 & 7)) = hi;
        *(uint4*)(blk + 2 * szk2 + tc::sw128_off(r, c16 & 7)) = lo;
      }
    }
    // V^T: one thread per (group of 4 keys r4, value column c): tf32 hi = top 19 bits, lo = the rest
    if (i < (NP >> 2) * 128) {
      const int c = i & 127, r4 = i >> 7;
      float hi[4], lo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = r4 * 4 + k;
        const float v = (r < N) ? __ldg(pb + (size_t)r * PROJ + H + c) : 0.f;
        hi[k] = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        lo[k] = v - hi[k];
      }
      const int kb = (r4 * 4) / NPB;                            // key block of these keys
      const int ch = (r4 * 4 - kb * NPB) >> 4;                  // 16-key chunk inside the block
      const int j = ch >= parts.cb[2] ? 2 : (ch >= parts.cb[1] ? 1 : 0);
      const int kl = r4 * 4 - kb * NPB - 16 * parts.cb[j];      // key index inside the part
      const int h = c >> 4, d = c & 15;
      unsigned char* blk = base + 8 * szk + (size_t)h * szv + (size_t)kb * szvb + t2_voff(NPB, j) + (size_t)(kl >> 5) * 4096;
      *(float4*)(blk + tc::sw128_off(d, (kl & 31) >> 2)) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *(float4*)(blk + tc::sw128_off(16 + d, (kl & 31) >> 2)) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

struct T2Args {
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
  float* logp;           // SAMPLE: [Bp,G] sum over the decode steps of log p(sampled node), or NULL
  unsigned long long seed;
  int N, NP, G, rows_per_cta, dbg_step, tile1_delay;
};

// barrier slots: per row tile, then the shared ring (T2B_FULL, T2B_EMPTY, T2B_TOTAL depend on the variant)
enum { T2B_SFULL = 0, T2B_PFULL = 3, T2B_OFULL = 6, T2B_OREADY = 7, T2B_QFULL = 8, T2B_UFULL = 9, T2B_UDONE = 10, T2B_PER_TILE = 11 };
constexpr int T2_FULL_BARS = 8;     // "landed" barriers: stage `it` signals barrier it % 8 (see land() in run_helper)
__host__ __device__ constexpr int t2_bar_total(bool one, bool key_blocks = false) {
  return (one ? 1 : 2) * T2B_PER_TILE + T2_FULL_BARS + (one ? (key_blocks ? T2_SLOTS_KB : T2_SLOTS_ONE) : T2_SLOTS);
}

// ONE: the CTA owns ONE row tile (at most 128 rows, SPLIT launches only) with 512 threads, 256 TMEM columns, a
// four-slot ring and half the shared memory, so that TWO CTAs share an SM.  Why: a decode step is a latency chain
// (8 x [scores -> softmax -> P V] -> logits -> arg-max -> gather) and the two tiles of a two-tile CTA walk it in
// lock-step -- they share the operand ring -- so their softmax phases collide on the MUFU pipe and their serial tails
// leave it idle together.  Two independent CTAs drift apart and fill each other's gaps; the price is that every
// operand block is fetched from L2 once per tile instead of once per instance.
//
// NKB > 1 (N > 208, single-tile CTAs only, two per SM -- one when sampling, whose key-block loops do not fit 64
// registers): the keys of an instance are processed in NKB blocks of A.NP <= 192 keys.  P = 2^S carries no shift, so the blocks of a head simply accumulate into the same O and the same
// row sums: per head the chain is NKB x [S block -> softmax -> P V], the accumulator is drained after the last block;
// the pointer logits are computed and scanned block by block (the score columns hold one block at a time), with the
// running arg-max (and, when sampling, the sum of exponentials of the log-probability) carried across the blocks.
template <bool LOGITS, bool DEBUG, bool SPLIT, bool SAMPLE, bool ONE, int NKB = 1>
__global__ void __launch_bounds__(ONE ? T2_THREADS_ONE : T2_THREADS, ONE ? (NKB == 1 ? 2 : (SAMPLE ? 1 : T2_KB_CTAS_PER_SM)) : 1) decode_tc2_kernel(const T2Args A) {
  static_assert(!ONE || (SPLIT && !DEBUG), "single-tile CTAs: SPLIT launches, no debug dump");
  static_assert(NKB == 1 || ONE, "key blocks: single-tile CTAs");
  constexpr int NIT = NH * NKB;             // (head, key block) iterations of the glimpse per decode step
  constexpr int T2_THREADS = ONE ? htsp::T2_THREADS_ONE : htsp::T2_THREADS;     // (shadow the two-tile constants)
  constexpr int T2_SLOTS = ONE ? (NKB > 1 ? T2_SLOTS_KB : htsp::T2_SLOTS_ONE) : htsp::T2_SLOTS;
  constexpr int T2_SLOT = NKB > 1 ? htsp::T2_SLOT_KB : htsp::T2_SLOT;     // (the smaller slots leave room for c0 of all keys with two CTAs per SM)
  constexpr int TILES = ONE ? 1 : 2;
  constexpr int T2B_FULL = TILES * T2B_PER_TILE, T2B_EMPTY = T2B_FULL + T2_FULL_BARS, T2B_TOTAL = T2B_EMPTY + T2_SLOTS;
  constexpr int T2_OCOL = ONE ? 208 : htsp::T2_OCOL, T2_XCOL = ONE ? 240 : htsp::T2_XCOL, TCOLS = ONE ? 256 : 512;
  static_assert(T2B_TOTAL == t2_bar_total(ONE, NKB > 1), "barrier count");
  static_assert(T2_SLOTS <= T2_FULL_BARS, "a landed-barrier is reused only after its stage's slot has been released");
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler too
  const int n_cta_parts = SPLIT ? (A.G + A.rows_per_cta - 1) / A.rows_per_cta : 1;
  const int b = SPLIT ? blockIdx.x / n_cta_parts : blockIdx.x;
  const int r0 = SPLIT ? (blockIdx.x - b * n_cta_parts) * A.rows_per_cta : 0;
  const int N = A.N, NP = A.NP, GT = A.G;
  const int G = SPLIT ? min(A.rows_per_cta, GT - r0) : A.G;   // rows of this CTA, from row r0 of the instance
  const int n_steps = N - 2;
  const int n_tiles = (!ONE && G > 128) ? 2 : 1;
  const T2Parts parts = t2_parts(NP);      // NP = keys per block (= all padded keys when NKB == 1)
  const int NPT = NKB * NP;                // all padded keys
  constexpr int PER_STEP = t2_per_step(NKB);

  unsigned char* staging = smem;                                         // TILES x 64 KB
  unsigned char* ring = smem + TILES * T2_TILE_STAGING;                  // T2_SLOTS x 12 KB
  float* c0s = (float*)(ring + (size_t)T2_SLOTS * T2_SLOT);              // [NPT]
  uint64_t* bars = (uint64_t*)(c0s + NPT);
  uint32_t* tmem_slot = (uint32_t*)(bars + T2B_TOTAL);
  uint2* stage_tab = (uint2*)(tmem_slot + 2);                            // [PER_STEP] (image offset, bytes) of a step's stages (not ONE)
  const uint2* stage_tab_g = (const uint2*)(A.img - 2048);               // the same table in global memory (ONE)
  const uint32_t bar0 = tc::smem_u32(bars);
  auto TBAR = [&](int t, int i) { return bar0 + 8u * (uint32_t)(t * T2B_PER_TILE + i); };
  auto RBAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  const unsigned char* imgb = A.img + (size_t)b * t2_img_bytes(NP, NKB);
  const float* projb = A.proj + (size_t)b * N * PROJ;

  if ((tc::smem_u32(smem) & 1023u) != 0u) __trap();   // swizzle atoms need 1024-byte alignment
  if (G > (ONE ? 128 : 224)) __trap();                 // tile 1 has three lane quarters (the host splits larger instances)
  if (NKB > 1 && NP > T2_KB_MAX_NPB) __trap();         // key-block ring slots hold parts of at most 64 keys
  if (tid == 0) {
    for (int t = 0; t < TILES; ++t) {
      const int rows_t = min(128, G - 128 * t);
      const int warps_t = rows_t > 0 ? (rows_t + 31) / 32 : 1;
      for (int j = 0; j < T2_PARTS; ++j) {
        tc::mbar_init(TBAR(t, T2B_SFULL + j), 1);
        tc::mbar_init(TBAR(t, T2B_PFULL + j), warps_t);
      }
      tc::mbar_init(TBAR(t, T2B_OFULL), 1);
      tc::mbar_init(TBAR(t, T2B_OREADY), warps_t);
      tc::mbar_init(TBAR(t, T2B_QFULL), T2_PARTS * warps_t);
      tc::mbar_init(TBAR(t, T2B_UFULL), T2_PARTS);           // one commit per key part (issuer + two helpers)
      tc::mbar_init(TBAR(t, T2B_UDONE), T2_PARTS * warps_t); // (key blocks) the logits of a block have been scanned
    }
    for (int s = 0; s < T2_FULL_BARS; ++s) tc::mbar_init(RBAR(T2B_FULL + s), 1);
    for (int s = 0; s < T2_SLOTS; ++s) tc::mbar_init(RBAR(T2B_EMPTY + s), n_tiles);      // a slot is released by every tile's issuer
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero the A tiles (rows that never receive a query must hold finite values)
  for (int i = tid; i < TILES * T2_TILE_STAGING / 16; i += T2_THREADS) ((uint4*)staging)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < NPT; i += T2_THREADS) c0s[i] = i < N ? __ldg(&A.c0[(size_t)b * N + i]) : 0.f;
  if (!ONE && tid < PER_STEP) stage_tab[tid] = t2_stage_entry(NP, tid);
  if (warp == 3) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_slot)),
                 "r"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc::fence_async_smem();
  tc::fence_before();
  __syncthreads();
  tc::fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // stages of one decode step in stream order: K_0^{0,1,2}; per head h: V_h^j, K_{h+1}^j (h < 7) for j = 0,1,2;
  // K2_i^j for i = 0..3, j = 0,1,2

  // ================================================================== issue helper of (row tile t, key part j = 1, 2)
  // S_0^j at the start of a step and the logit MMAs U^j at its end: both wait on barriers that order them after
  // everything they depend on, and they write the score columns of part j only
  auto run_helper = [&](const int t, const int j) {
    if (t >= n_tiles) return;
    const uint32_t stg = tc::smem_u32(staging) + t * T2_TILE_STAGING;
    const uint32_t ring0 = tc::smem_u32(ring);
    const int cbj = j == 1 ? parts.cb[1] : parts.cb[2], cej = j == 1 ? parts.cb[2] : parts.cb[3];
    const uint32_t dSj = tmem_base + (uint32_t)(t * NP + 16 * cbj);
    const uint32_t idK = tc::idesc_f16(16 * (cej - cbj));
    // A parity wait tells the awaited use of a barrier from the previous one only: asked before the PREVIOUS use has
    // completed it answers "done".  The stages a consumer skips (they belong to the other issuers) must therefore keep
    // it away from that point.  Stage `it` lives in ring slot it % T2_SLOTS but signals "landed" barrier it % 8: with
    // the four slots of the single-tile CTAs a per-slot barrier's previous use is the stage four positions back,
    // issued just before the stage this consumer has just finished -- one L2 miss in that copy and the wait would pass
    // early (seen as a hung ring with two CTAs per SM).  Eight positions back is a stage whose slot has been released.
    auto land = [&](uint32_t it) {
      const uint32_t s = it % T2_SLOTS;
      tc::mbar_wait(RBAR(T2B_FULL + it % T2_FULL_BARS), (it / T2_FULL_BARS) & 1);
      tc::fence_after();
      return s;
    };
    for (int step = 0; step < n_steps; ++step) {
      const uint32_t base = (uint32_t)step * (uint32_t)PER_STEP;
      {
        const uint32_t sk = land(base + (uint32_t)j);              // K_0^j
        tc::mbar_wait(TBAR(t, T2B_QFULL), step & 1);
        tc::fence_after();
        const uint32_t a_hi = stg >> 4, a_lo = a_hi + 2048;
        const uint32_t b_hi = (ring0 + sk * T2_SLOT) >> 4, b_lo = b_hi + 2;
        if (tc::elect_one()) {
          tc::mma_ss_lo_b64(dSj, a_hi, b_hi, idK, 0);
          tc::mma_ss_lo_b64(dSj, a_hi, b_lo, idK, 1);
          tc::mma_ss_lo_b64(dSj, a_lo, b_hi, idK, 1);
          tc::commit(TBAR(t, T2B_SFULL + j));
          tc::commit(RBAR(T2B_EMPTY + sk));
        }
        __syncwarp();
      }
#pragma unroll 1
      for (int kb = 0; kb < NKB; ++kb) {
        if (NKB > 1 && kb > 0) {      // the softmax warps have scanned the previous block's logits
          tc::mbar_wait(TBAR(t, T2B_UDONE), (uint32_t)(step * NKB + kb - 1) & 1u);
          tc::fence_after();
        }
        for (int i = 0; i < 4; ++i) {
          const int half = i & 1;
          const uint32_t o_hi = (stg >> 4) + half * 1024, o_lo = o_hi + 2048;
          // (the parity of a ring barrier only tells the current use from the previous one: a stage may be waited for
          // once the stage eight positions earlier has been consumed by this tile -- true after OREADY, not before)
          if (i == 0 && kb == 0) {
            tc::mbar_wait(TBAR(t, T2B_OREADY), step & 1);
            tc::fence_after();
          }
          const uint32_t sk = land(base + (uint32_t)(6 + 6 * (NIT - 1) + 3 * (4 * kb + i) + j));
          const uint32_t kb_ = (ring0 + sk * T2_SLOT) >> 4;
          if (tc::elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              tc::mma_ss_lo(dSj, o_hi + kk * 2, kb_ + kk * 2, idK, (i | kk) != 0);
              if (i < 2) tc::mma_ss_lo(dSj, o_lo + kk * 2, kb_ + kk * 2, idK, 1);
            }
            tc::commit(RBAR(T2B_EMPTY + sk));
          }
          __syncwarp();
        }
        if (tc::elect_one()) tc::commit(TBAR(t, T2B_UFULL));
        __syncwarp();
      }
    }
  };

  // ================================================================== MMA issuer of row tile t
  auto run_issuer = [&](const int t) {
    if (t < n_tiles) {
      const uint32_t stg = tc::smem_u32(staging) + t * T2_TILE_STAGING;
      const uint32_t ring0 = tc::smem_u32(ring);
      const uint32_t dS = tmem_base + (uint32_t)(t * NP);
      const uint32_t dO = tmem_base + (uint32_t)(T2_OCOL + 32 * t);
      const uint32_t idO = tc::idesc_tf32(32);
      uint32_t idK[T2_PARTS];
      int kk8[T2_PARTS];                     // 8-key steps of P V per part
#pragma unroll
      for (int j = 0; j < T2_PARTS; ++j) {
        idK[j] = tc::idesc_f16(16 * (parts.cb[j + 1] - parts.cb[j]));
        kk8[j] = 2 * (parts.cb[j + 1] - parts.cb[j]);
      }
      long long tph[6] = {0, 0, 0, 0, 0, 0};   // DEBUG: 0 ring wait, 1 wait P, 2 issue, 3 wait Q, 4 wait O staged, 5 logits
      long long tlast = DEBUG ? clock64() : 0;
      auto tick = [&](int ph) {
        if (DEBUG) {
          const long long now = clock64();
          tph[ph] += now - tlast;
          tlast = now;
        }
      };
      auto land = [&](uint32_t it) {            // wait for stage `it` of the stream; returns its ring slot
        const uint32_t s = it % T2_SLOTS;
        tick(2);
        tc::mbar_wait(RBAR(T2B_FULL + it % T2_FULL_BARS), (it / T2_FULL_BARS) & 1);
        tc::fence_after();
        tick(0);
        return s;
      };
      auto release = [&](uint32_t s) {
        if (tc::elect_one()) tc::commit(RBAR(T2B_EMPTY + s));
        __syncwarp();
      };
      // S_h^j into the score columns of part j; sk = ring slot holding K_h^j ([keys x (hi | lo)], SWIZZLE_64B)
      auto issue_S = [&](int h, int j, uint32_t sk) {
        const uint32_t a_hi = (stg >> 4) + (h >> 2) * 1024 + (h & 3) * 2, a_lo = a_hi + 2048;
        const uint32_t b_hi = (ring0 + sk * T2_SLOT) >> 4, b_lo = b_hi + 2;
        const uint32_t d = dS + (uint32_t)(16 * parts.cb[j]);
        if (tc::elect_one()) {
          tc::mma_ss_lo_b64(d, a_hi, b_hi, idK[j], 0);
          tc::mma_ss_lo_b64(d, a_hi, b_lo, idK[j], 1);
          tc::mma_ss_lo_b64(d, a_lo, b_hi, idK[j], 1);
          tc::commit(TBAR(t, T2B_SFULL + j));
        }
        __syncwarp();
      };
      // O (+)= P_h^j [Vhi | Vlo]: A = fp32 words in the score columns of part j, B = V^T blocks of 32 keys in slot sv
      // (first: the first part of the first key block of a head starts the accumulator; last: the last part of the
      // last key block publishes it)
      auto issue_PV = [&](int j, uint32_t sv, bool first, bool last) {
        const uint32_t vb = (ring0 + sv * T2_SLOT) >> 4;
        const uint32_t pa = dS + (uint32_t)(16 * parts.cb[j]);
        const int n8 = kk8[j];
        if (tc::elect_one()) {
#pragma unroll 1
          for (int ks = 0; ks < n8; ++ks)
            tc::mma_ts_tf32_lo(dO, pa + 8 * ks, vb + (ks >> 2) * 256 + (ks & 3) * 2, idO, !(first && ks == 0));
          if (last) tc::commit(TBAR(t, T2B_OFULL));
        }
        __syncwarp();
      };
      for (int step = 0; step < n_steps; ++step) {
        const uint32_t base = (uint32_t)step * (uint32_t)PER_STEP;
        const uint32_t k0 = land(base);           // K_0^0 (the helpers issue S_0 of parts 1 and 2)
        tc::mbar_wait(TBAR(t, T2B_QFULL), step & 1);
        tc::fence_after();
        tick(3);
        issue_S(0, 0, k0);
        release(k0);
        for (int it = 0; it < NIT; ++it) {               // (head, key block)
          const uint32_t par = (uint32_t)(step * NIT + it) & 1u;
          const int kb = NKB == 1 ? 0 : it % NKB;
#pragma unroll
          for (int j = 0; j < T2_PARTS; ++j) {
            const uint32_t nv = base + (it + 1 < NIT ? (uint32_t)(3 + 6 * it + 2 * j) : (uint32_t)(3 + 6 * (NIT - 1) + j));
            const uint32_t sv = land(nv);                               // V of (head, block), part j
            const uint32_t sk = (it + 1 < NIT) ? land(nv + 1) : 0u;     // K of the next (head, block), part j
            tick(2);
            tc::mbar_wait(TBAR(t, T2B_PFULL + j), par);
            tc::fence_after();
            tick(1);
            issue_PV(j, sv, kb == 0 && j == 0, kb == NKB - 1 && j == T2_PARTS - 1);
            release(sv);
            if (it + 1 < NIT) {
              issue_S((it + 1) / NKB, j, sk);
              release(sk);
            }
          }
        }
        // pointer logits U = O K2^T into the (consumed) score columns of part 0 (parts 1, 2: helpers), key block by key block
#pragma unroll 1
        for (int kb = 0; kb < NKB; ++kb) {
          if (NKB > 1 && kb > 0) {
            tc::mbar_wait(TBAR(t, T2B_UDONE), (uint32_t)(step * NKB + kb - 1) & 1u);
            tc::fence_after();
          }
          for (int i = 0; i < 4; ++i) {
            const int half = i & 1;                    // O / K2 columns 64*half .. +63
            const uint32_t o_hi = (stg >> 4) + half * 1024, o_lo = o_hi + 2048;
            const uint32_t sk = land(base + (uint32_t)(6 + 6 * (NIT - 1) + 3 * (4 * kb + i)));
            if (i == 0 && kb == 0) {
              tc::mbar_wait(TBAR(t, T2B_OREADY), step & 1);
              tc::fence_after();
              tick(4);
            }
            const uint32_t kb_ = (ring0 + sk * T2_SLOT) >> 4;
            if (tc::elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                tc::mma_ss_lo(dS, o_hi + kk * 2, kb_ + kk * 2, idK[0], (i | kk) != 0);
                if (i < 2) tc::mma_ss_lo(dS, o_lo + kk * 2, kb_ + kk * 2, idK[0], 1);
              }
              tc::commit(RBAR(T2B_EMPTY + sk));
            }
            __syncwarp();
          }
          if (tc::elect_one()) tc::commit(TBAR(t, T2B_UFULL));
          __syncwarp();
        }
        tick(5);
      }
      if (DEBUG && b == 0 && lane == 0) {
        float* dT = A.dbg + (size_t)2 * 8 * 128 * NP + (size_t)256 * 128 + (size_t)256 * NP + (24 + t) * 8;
        for (int i = 0; i < 6; ++i) dT[i] = (float)tph[i];
      }
    }
  };
  // ================================================================== operand producer (one stream for all tiles)
  auto run_producer = [&]() {
    if (lane == 0) {
      const uint32_t total = (uint32_t)n_steps * (uint32_t)PER_STEP;
      uint32_t i = 0;
      uint2 nxt = ONE ? __ldg(stage_tab_g) : make_uint2(0u, 0u);
      for (uint32_t it = 0; it < total; ++it) {
        const uint32_t s = it % T2_SLOTS, use = it / T2_SLOTS;
        const uint2 st = ONE ? nxt : stage_tab[i];
        i = (i + 1 == (uint32_t)PER_STEP) ? 0u : i + 1;
        if (ONE) nxt = __ldg(stage_tab_g + i);      // (the next entry travels while this stage's slot is waited for)
        tc::mbar_wait(RBAR(T2B_EMPTY + s), (use & 1) ^ 1, it);
        const uint32_t fb = RBAR(T2B_FULL + it % T2_FULL_BARS);
        tc::mbar_expect_tx(fb, st.y);
        tc::bulk_g2s(tc::smem_u32(ring + (size_t)s * T2_SLOT), imgb + st.x, st.y, fb);
      }
    }
  };
  // ================================================================== softmax / state warps
  auto run_softmax = [&]() {
    const int sw = warp - 4;
    const int t = sw / 12, part = (sw % 12) >> 2, q = warp & 3;
    const int rows_t = min(128, G - 128 * t);
    if (q * 32 < rows_t) {
      const int row = q * 32 + lane;                 // row inside the tile = TMEM lane
      const int g = t * 128 + row;
      const bool valid = g < G;
      const int gc = min(g, G - 1);
      const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
      const uint32_t tS = tmem_base + lane_addr + (uint32_t)(t * NP);
      const uint32_t tO = tmem_base + lane_addr + (uint32_t)(T2_OCOL + 32 * t);
      const uint32_t tX = tmem_base + lane_addr + (uint32_t)(T2_XCOL + 16 * t);
      unsigned char* stg = staging + t * T2_TILE_STAGING;
      const int src = clamp_index(A.src[b], N), tgt = clamp_index(A.tgt[b], N);
      const int first = clamp_index(A.first[r0 + gc], N);
      int32_t* pirow = A.pi + ((size_t)b * GT + r0 + gc) * N;
      const int c_lo = part == 0 ? 0 : (part == 1 ? parts.cb[1] : parts.cb[2]);
      const int c_hi = part == 0 ? parts.cb[1] : (part == 1 ? parts.cb[2] : parts.cb[3]);
      bool bad = false;                              // softmax sum left the fp32 range (part 0 only)
      float logp_acc = 0.f;                          // SAMPLE, part 0: sum of log p(sampled node) over the steps

      // visited flags of THIS warp's keys only (per key block kb: bit k = key kb NP + 16 c_lo + k; at most 96 keys per
      // part and block); padded keys (>= N) are permanently visited.  (A key of another part or block sets, at most,
      // a bit beyond this warp's chunks, which nothing reads.)
      const int nc = c_hi - c_lo;
      unsigned pv[NKB][3];
#pragma unroll
      for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
        for (int w = 0; w < 3; ++w) {
          const int lo = kb * NP + 16 * c_lo + 32 * w;
          pv[kb][w] = (lo + 32 <= N) ? 0u : ((N <= lo) ? 0xffffffffu : (0xffffffffu << (N - lo)));
        }
      }
      int cur = 0, cnt = 0;
      auto mark = [&](int a) {
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
          const unsigned rel = (unsigned)(a - kb * NP - 16 * c_lo);
          const unsigned bit = rel < 96u ? (1u << (rel & 31u)) : 0u;
#pragma unroll
          for (int w = 0; w < 3; ++w) pv[kb][w] |= ((rel >> 5) == (unsigned)w) ? bit : 0u;
        }
      };
      // visit(a) with source<->target coupling (train_path_solver.py:45-66); only part 0 records pi
      auto visit = [&](int a) {
        if (valid && part == 0 && cnt < N) pirow[cnt] = a;     // (a forced tour, or a row without finite logits, may repeat nodes)
        mark(a);
        cnt++;
        cur = a;
        const int partner = (a == src) ? tgt : ((a == tgt) ? src : -1);
        if (partner >= 0) {
          if (valid && part == 0 && cnt < N) pirow[cnt] = partner;
          mark(partner);
          cnt++;
          cur = partner;
        }
      };
      long long tph[6] = {0, 0, 0, 0, 0, 0};   // DEBUG: 0 wait S, 1 softmax, 2 wait O + drain, 3 wait U, 4 arg-max, 5 state + gather
      long long tlast = DEBUG ? clock64() : 0;
      auto tick = [&](int ph) {
        if (DEBUG) {
          const long long now = clock64();
          tph[ph] += now - tlast;
          tlast = now;
        }
      };

      // ---- P = 2^S of this warp's key chunks of head h, in place; returns the row sum of what the tensor core
      // will read (the top 19 bits of every word)
      auto softmax_pass = [&](int step, int h, const unsigned (&pvb)[3]) {
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          if (i < nc) {
            const int c = c_lo + i;
            uint32_t s[32];
            tld<16>(tS + 16 * c, s);
            const unsigned bits = (pvb[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
            tld_wait16(s);
            if (DEBUG && b == 0 && step == A.dbg_step) {
              float* dSd = A.dbg + ((size_t)(t * 8 + h) * 128 + row) * NP + 16 * c;
#pragma unroll
              for (int j = 0; j < 16; ++j) dSd[j] = __uint_as_float(s[j]);
            }
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float p0, p1;
              if (j < T2_POLY) {      // FMA-pipe polynomial for some of the exponentials (MUFU.EX2 is the busiest pipe)
                const float2 r = ex2_poly2(__uint_as_float(s[j]), __uint_as_float(s[j + 1]));
                p0 = ((bits >> j) & 1u) ? 0.f : r.x;
                p1 = ((bits >> (j + 1)) & 1u) ? 0.f : r.y;
              } else {
                p0 = ex2_approx(((bits >> j) & 1u) ? -INFINITY : __uint_as_float(s[j]));
                p1 = ex2_approx(((bits >> (j + 1)) & 1u) ? -INFINITY : __uint_as_float(s[j + 1]));
              }
              // round to the 19 bits the tensor core reads (it would truncate: a mantissa-dependent bias that the
              // normalisation does not cancel -- 1.1e-3 instead of 0.9e-3 on the 12-layer stress weights) and sum
              // exactly those values
              s[j] = (__float_as_uint(p0) + 0x1000u) & 0xffffe000u;     // (cvt.rna.tf32.f32 does the same in one
              s[j + 1] = (__float_as_uint(p1) + 0x1000u) & 0xffffe000u; // instruction, but on the MUFU pipe: 25.1 vs 23.8 ms)
              acc = __fadd2_rn(acc, make_float2(__uint_as_float(s[j]), __uint_as_float(s[j + 1])));
            }
            tst<16>(tS + 16 * c, s);
          }
        }
        return acc.x + acc.y;
      };
      // publish the row sum of this part, then P itself
      auto publish = [&](int h, float l, bool with_sum) {
        if (with_sum)
          asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tX + (h & 1) * 3 + part), "r"(__float_as_uint(l))
                       : "memory");
        tc::st_wait();
        tc::fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(TBAR(t, T2B_PFULL + part));
      };

      const float4 qfix4 = __ldg((const float4*)(A.qfix + (size_t)b * H) + lane);
      const uint32_t trio_bar = 1u + (uint32_t)(t * 4 + q);     // named barrier of the three warps of this row quarter
      auto trio_sync = [&]() {
        if constexpr (ONE) {
          // immediate barrier numbers: with a register operand ptxas books all 16 named barriers for the CTA, and the
          // driver then places one CTA per SM (measured: the occupancy query answers 1 whatever the shared memory)
          if (q == 0) asm volatile("bar.sync 1, 96;" ::: "memory");
          else if (q == 1) asm volatile("bar.sync 2, 96;" ::: "memory");
          else if (q == 2) asm volatile("bar.sync 3, 96;" ::: "memory");
          else asm volatile("bar.sync 4, 96;" ::: "memory");
        } else {
          asm volatile("bar.sync %0, 96;" ::"r"(trio_bar) : "memory");
        }
      };
      // query rows -> A tile (fp16 hi/lo, K-major SWIZZLE_128B):
      // q = ((q_graph+q_source+q_target) + q_first) + q_last (models.py:413), 1/sqrt(dk)*log2(e) folded in.
      // The three warps of a row quarter gather a third of its rows each; lane l carries columns 4l..4l+3 of every
      // row; T2_GATHER_ROWS rows (two L2 loads each) are in flight at a time.
      auto gather_q = [&]() {
        const uint32_t sub = (uint32_t)(lane >> 4) * 16384u + (uint32_t)(lane & 1) * 8u;
        const uint32_t chunk = (uint32_t)(lane & 15) >> 1;
        const int nrows = min(32, rows_t - q * 32);
        const int r_beg = min(nrows, (32 * part) / 3), r_end = min(nrows, (32 * (part + 1)) / 3);
#pragma unroll 1
        for (int rr = r_beg; rr < r_end; rr += T2_GATHER_ROWS) {
          float4 ql[T2_GATHER_ROWS], qf[T2_GATHER_ROWS];
#pragma unroll
          for (int i = 0; i < T2_GATHER_ROWS; ++i) {
            const int cr = __shfl_sync(0xffffffffu, cur, (rr + i) & 31), fr = __shfl_sync(0xffffffffu, first, (rr + i) & 31);
            ql[i] = __ldg((const float4*)(projb + (size_t)cr * PROJ + 2 * H) + lane);
            qf[i] = __ldg((const float4*)(projb + (size_t)fr * PROJ + 3 * H) + lane);
          }
#pragma unroll
          for (int i = 0; i < T2_GATHER_ROWS; ++i) {
            if (rr + i < r_end) {
              uint32_t h0, l0, h1, l1;
              split2(((qfix4.x + qf[i].x) + ql[i].x) * kScoreScale, ((qfix4.y + qf[i].y) + ql[i].y) * kScoreScale, h0, l0);
              split2(((qfix4.z + qf[i].z) + ql[i].z) * kScoreScale, ((qfix4.w + qf[i].w) + ql[i].w) * kScoreScale, h1, l1);
              unsigned char* p = stg + sub + tc::sw128_off((uint32_t)(q * 32 + rr + i), chunk);
              *(uint2*)p = make_uint2(h0, h1);
              *(uint2*)(p + 32768) = make_uint2(l0, l1);
            }
          }
        }
        tc::fence_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(TBAR(t, T2B_QFULL));
      };

      if (t == 1 && A.tile1_delay > 0) {      // row tile 1 starts late: its head loop interleaves with tile 0's
        const long long t0c = clock64();
        while (clock64() - t0c < (long long)A.tile1_delay) {}
      }
      // step 0: no decoder call (train_path_solver.py:256-257)
      visit(first);
      gather_q();

      // drain O_h (part 0): (P Vhi + P Vlo) / sum of the three row sums -> fp16 hi/lo -> A tile columns
      // 16h..16h+15 (the consumed Q_h slot)
      auto drain_o = [&](int step, int h) {
        tc::mbar_wait(TBAR(t, T2B_OFULL), (step * NH + h) & 1);
        tc::fence_after();
        uint32_t oa[16], ob[16], x0, x1, x2, x3;
        tc::ld16(tO, oa);
        tc::ld16(tO + 16, ob);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                     : "r"(tX + (h & 1) * 3)
                     : "memory");
        tc::ld16_wait(oa);
        tc::ld16_wait(ob);
        asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3)::"memory");
        const float l = (__uint_as_float(x0) + __uint_as_float(x1)) + __uint_as_float(x2);
        float inv;         // MUFU reciprocal (1 ulp): the normalisation error is far below the tf32 probabilities
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(l));
        if (!(l > 0.f && l < 3.0e38f)) {   // empty or overflowed softmax: flag the row, keep the state machine finite
          inv = 0.f;
          bad = bad || valid;
        }
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (__uint_as_float(oa[j]) + __uint_as_float(ob[j])) * inv;
        if (!(l > 0.f && l < 3.0e38f)) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
        if (DEBUG && b == 0 && step == A.dbg_step) {
          float* dOd = A.dbg + (size_t)2 * 8 * 128 * NP + (size_t)(t * 128 + row) * 128 + 16 * h;
#pragma unroll
          for (int j = 0; j < 16; ++j) dOd[j] = v[j];
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
          float l = 0.f;                 // row sum of this part over the key blocks of the head
#pragma unroll
          for (int kb = 0; kb < NKB; ++kb) {
            tc::mbar_wait(TBAR(t, T2B_SFULL + part), (uint32_t)((step * NH + h) * NKB + kb) & 1u);
            tc::fence_after();
            tick(0);
            l += softmax_pass(step, h, pv[kb]);
            tick(1);
            // the previous head's accumulator must be drained before this head's first P V (part 0 of key block 0,
            // which the issuer serves first) overwrites it
            if (part == 0 && h > 0 && kb == 0) {
              drain_o(step, h - 1);
              tick(2);
            }
            publish(h, l, kb == NKB - 1);
          }
        }
        if (part == 0) {
          drain_o(step, NH - 1);
          tc::fence_async_smem();
          tc::fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(TBAR(t, T2B_OREADY));
          tick(2);
        }

        // ------------------------------------------------------------ pointer logits + arg-max
        // Each warp of the trio scans its own key chunks and reduces them to one candidate (clipped logit, key);
        // part 0 merges the three (ties -> lower key = the lower part).
        float best = -INFINITY;
        int idx = 0x7fffffff;
        float zbest = 0.f, sumexp = 0.f;       // SAMPLE with log-probabilities (train_path_solver.py:394-414)
        float* lo_out = nullptr;
        if (LOGITS && A.logits_out != nullptr)
          lo_out = A.logits_out + (((size_t)b * GT + r0 + gc) * n_steps + step) * N;
        // candidates of one key block (koff = its first key, pvb = its visited words)
        float bk, zk, sek;
        int ik;
        // pass(EXACT): EXACT compares the clipped logits 10*tanh(.) themselves (models.py:438-440); otherwise the
        // monotone pre-activation is compared, which picks the same key except when two clipped logits round to
        // the same float (handled by re-running EXACT near the clip).  Four independent running maxima (column
        // slot mod 4), merged with the lowest key winning ties (torch.argmax returns the first maximum).
        auto pass = [&](auto exact_tag, const int koff, const unsigned (&pvb)[3]) {
          constexpr bool EXACT = decltype(exact_tag)::value;
          float bj[4];
          int kj[4];
          float zj[4];                 // SAMPLE: clipped logit (without the noise) of each slot's running best
#pragma unroll
          for (int j = 0; j < 4; ++j) { bj[j] = -INFINITY; kj[j] = 0x7fffffff; zj[j] = 0.f; }
          float se = 0.f;              // SAMPLE: sum over this warp's unvisited keys of exp(z - 10) (z <= 10: no overflow)
          uint32_t u[32];
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            if (i >= nc) continue;
            const int c = c_lo + i;
            tld<16>(tS + 16 * c, u);
            const unsigned bits = (pvb[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
            tld_wait16(u);
            if (DEBUG && b == 0 && step == A.dbg_step) {
              float* dU = A.dbg + (size_t)2 * 8 * 128 * NP + (size_t)256 * 128 + (size_t)(t * 128 + row) * NP + 16 * c;
#pragma unroll
              for (int j = 0; j < 16; ++j) dU[j] = __uint_as_float(u[j]);
            }
            const float4* cc4 = (const float4*)(c0s + koff + 16 * c);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 cc = cc4[j4];
              const float ca[4] = {cc.x, cc.y, cc.z, cc.w};
              uint4 rnd = make_uint4(0u, 0u, 0u, 0u);
              if (SAMPLE)
                rnd = tcd_philox(make_uint4((uint32_t)((koff >> 2) + 4 * c + j4), (uint32_t)step, (uint32_t)(r0 + gc), (uint32_t)b),
                                 make_uint2((uint32_t)A.seed, (uint32_t)(A.seed >> 32)));
              const uint32_t ra[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int j = 4 * j4 + jj;
                float z = __uint_as_float(u[j]) + ca[jj];          // EXACT: 10 tanh(z / sqrt(128)); else the
                if (EXACT) z = clip_tanh10(z * kT2InvSqrtH);        // monotone pre-activation itself
                const float zc = z;                                 // SAMPLE: the policy's logit of this key
                if (SAMPLE) {
                  se += ((bits >> j) & 1u) ? 0.f : ex2_approx((zc - 10.f) * 1.4426950408889634f);
                  z += tcd_gumbel(ra[jj]);
                }
                if ((bits >> j) & 1u) z = -INFINITY;
                if (LOGITS && lo_out != nullptr && valid && koff + 16 * c + j < N) lo_out[koff + 16 * c + j] = z;
                if (z > bj[jj]) { bj[jj] = z; kj[jj] = koff + 16 * c + j; if (SAMPLE) zj[jj] = zc; }
              }
            }
          }
          bk = bj[0];
          ik = kj[0];
          zk = zj[0];
          sek = se;
#pragma unroll
          for (int j = 1; j < 4; ++j)
            if (bj[j] > bk || (bj[j] == bk && kj[j] < ik)) { bk = bj[j]; ik = kj[j]; zk = zj[j]; }
        };
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
          // (key blocks: the score columns hold the logits of one block at a time)
          tc::mbar_wait(TBAR(t, T2B_UFULL), (uint32_t)(step * NKB + kb) & 1u);
          tc::fence_after();
          tick(3);
          if (LOGITS || SAMPLE) {
            pass(std::true_type{}, kb * NP, pv[kb]);
          } else {
            pass(std::false_type{}, kb * NP, pv[kb]);
            if (__any_sync(0xffffffffu, valid && bk * kT2InvSqrtH > 3.0f)) {
              pass(std::true_type{}, kb * NP, pv[kb]);
            } else {
              bk = (bk == -INFINITY) ? bk : clip_tanh10(bk * kT2InvSqrtH);   // candidates are exchanged as clipped logits
            }
          }
          if (NKB > 1) {               // this warp has read the block's logits: the issuers may overwrite them
            tc::fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(TBAR(t, T2B_UDONE));
          }
          if (kb == 0 || bk > best) { best = bk; idx = ik; zbest = zk; }     // ties keep the lower block's key
          sumexp += sek;
        }
        tick(4);
        int a;
        if (part != 0) {
          asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(tX + 4 + 2 * part), "r"(__float_as_uint(best)),
                       "r"((uint32_t)idx)
                       : "memory");
          if (SAMPLE)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(tX + 9 + 2 * part), "r"(__float_as_uint(zbest)),
                         "r"(__float_as_uint(sumexp))
                         : "memory");
          tc::st_wait();
          tc::fence_before();
          trio_sync();                // candidates published
          trio_sync();                // part 0 has published the chosen node
          tc::fence_after();
          uint32_t av;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(av) : "r"(tX + 10) : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(av)::"memory");
          a = (int)av;
        } else {
          trio_sync();
          tc::fence_after();
          uint32_t z1, i1, z2, i2;
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(z1), "=r"(i1), "=r"(z2), "=r"(i2)
                       : "r"(tX + 6)
                       : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(z1), "+r"(i1), "+r"(z2), "+r"(i2)::"memory");
          int win = 0;
          // ties keep the lower key (one key block: that of the lower part)
          if (__uint_as_float(z1) > best || (NKB > 1 && __uint_as_float(z1) == best && (int)i1 < idx)) { best = __uint_as_float(z1); idx = (int)i1; win = 1; }
          if (__uint_as_float(z2) > best || (NKB > 1 && __uint_as_float(z2) == best && (int)i2 < idx)) { best = __uint_as_float(z2); idx = (int)i2; win = 2; }
          if (SAMPLE) {
            // log p(sampled node) = z - logsumexp(z over the unvisited keys) (softmax of the masked clipped logits,
            // models.py:442; train_path_solver.py:394-414 sums the logarithms over the decode steps)
            uint32_t e1, s1, e2, s2;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(e1), "=r"(s1), "=r"(e2), "=r"(s2)
                         : "r"(tX + 11)
                         : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(e1), "+r"(s1), "+r"(e2), "+r"(s2)::"memory");
            const float zw = win == 0 ? zbest : (win == 1 ? __uint_as_float(e1) : __uint_as_float(e2));
            const float tot = (sumexp + __uint_as_float(s1)) + __uint_as_float(s2);
            logp_acc += (zw - 10.f) - __logf(tot);
          }
          // ---- state update (train_path_solver.py:45-72); the chosen node goes to the other two warps
          a = ((LOGITS || DEBUG) && A.forced_pi) ? clamp_index(A.forced_pi[((size_t)b * GT + r0 + gc) * N + min(cnt, N - 1)], N) : idx;   // teacher forcing: LOGITS / DEBUG kernels only
          if ((unsigned)a >= (unsigned)N) a = cur;     // a NaN row has no candidate: stay in range (the row is flagged)
          asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tX + 10), "r"((uint32_t)a) : "memory");
          tc::st_wait();
          tc::fence_before();
          trio_sync();
        }
        visit(a);
        tc::fence_before();
        if (step + 1 < n_steps) gather_q();
        tick(5);
      }
      // (a flagged row -- no finite logits, or a forced tour that repeats nodes -- may have recorded fewer than N nodes:
      // every entry of its tour must still be a node, the tour is read as indices below and by the callers)
      if (valid && part == 0)
        for (; cnt < N; ++cnt) pirow[cnt] = cur;
      // ---- reward = -(closed tour length - fixed edge)   (train_path_solver.py:109-149)
      if (valid && part == 0) {
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
        A.reward[(size_t)b * GT + r0 + g] = bad ? __int_as_float(0x7fc00000) : -(len - fixed);
        if (SAMPLE && A.logp != nullptr) A.logp[(size_t)b * GT + r0 + g] = bad ? __int_as_float(0x7fc00000) : logp_acc;
      }
      if (DEBUG && b == 0 && lane == 0) {
        float* dT = A.dbg + (size_t)2 * 8 * 128 * NP + (size_t)256 * 128 + (size_t)256 * NP + sw * 8;
        for (int i = 0; i < 6; ++i) dT[i] = (float)tph[i];
      }
    }
  };

  if constexpr (ONE) {
    // warps 0-3: issuer, issue helper of key part 2, producer, issue helper of key part 1; 4-15: softmax warps (part, lane
    // quarter).  2 CTAs x 512 threads leave 64 registers per thread, which the softmax path fits without spills in its
    // loops (a setmaxnreg hand-over from the role warps makes the driver count one CTA per SM: measured).
#ifdef T2_ONE_ISSUER_WARP3
    if (warp == 3) run_issuer(0);
    else if (warp == 1) run_helper(0, 2);
    else if (warp == 2) run_producer();
    else if (warp == 0) run_helper(0, 1);
    else run_softmax();
#else
    if (warp == 0) run_issuer(0);
    else if (warp == 1) run_helper(0, 2);
    else if (warp == 2) run_producer();
    else if (warp == 3) run_helper(0, 1);
    else run_softmax();
#endif
  } else {
    if (warp < 2) run_issuer(warp);
    else if (warp == 2) run_producer();
    else if (warp == 3) run_helper(0, 1);
    else if (warp == 19 || warp == 23 || warp == 27) run_helper(warp == 19 ? 0 : 1, warp == 19 ? 2 : (warp == 23 ? 1 : 2));
    else if (warp >= 4) run_softmax();
  }

  tc::fence_before();
  __syncthreads();
  if (warp == 3) {
    tc::fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static int t2_np_of(int N) { return (N + 15) / 16 * 16; }
// key blocks: one block of all (padded) keys up to 208 keys, else NKB equal blocks of NPB <= 208 keys
static void t2_blocks_of(int N, int& NPB, int& NKB) {
  const char* env = getenv("HTSP_T2_NKB");       // (A/B experiments: key blocks for a subproblem that fits one block)
  const int forced = env ? atoi(env) : 0;
  if (t2_np_of(N) <= 208 && !((forced == 2 || forced == 3) && N >= 48 * forced)) {
    NKB = 1;
    NPB = t2_np_of(N);
  } else if (t2_np_of(N) <= 208) {
    NKB = forced;
    NPB = t2_np_of((N + NKB - 1) / NKB);
  } else {
    NKB = (t2_np_of(N) + T2_KB_MAX_NPB - 1) / T2_KB_MAX_NPB;
    NPB = t2_np_of((N + NKB - 1) / NKB);
  }
}
constexpr int T2_MAX_NKB = 3;
constexpr size_t T2_TAB_BYTES = 2048;     // stage table in front of the operand images (<= 180 entries of 8 bytes)

bool decode_tc2_supported(int N, int G) {
  int NPB, NKB;
  t2_blocks_of(N, NPB, NKB);
  if (NKB == 1) return NPB >= 48 && G <= 256;
  return NKB <= T2_MAX_NKB && G >= 1;        // key blocks: single-tile CTAs of 128 rows, any number per instance
}
size_t decode_tc2_pack_bytes(int Bp, int N) {
  // [slack to a 1024-byte boundary | stage table, 2 KB | operand images | mean keys]
  int NPB, NKB;
  t2_blocks_of(N, NPB, NKB);
  return 1024 + T2_TAB_BYTES + align_up((size_t)Bp * t2_img_bytes(NPB, NKB), 256) + align_up((size_t)Bp * H * 4, 256);
}
size_t decode_tc2_dbg_floats(int N) {
  const int NP = t2_np_of(N);
  return (size_t)2 * 8 * 128 * NP + (size_t)256 * 128 + (size_t)256 * NP + 256;
}

// geometry of the operand stream for N nodes (host logic, checked on the CPU by tests/test_abi.py): keys per block,
// key blocks, ring stages per decode step, operand-image bytes per instance, ring-slot bytes
bool decode_tc2_geometry(int N, int32_t out[5]) {
  if (N < 3 || !decode_tc2_supported(N, 1)) return false;
  int NPB, NKB;
  t2_blocks_of(N, NPB, NKB);
  out[0] = NPB;
  out[1] = NKB;
  out[2] = t2_per_step(NKB);
  out[3] = (int32_t)t2_img_bytes(NPB, NKB);
  out[4] = NKB > 1 ? T2_SLOT_KB : T2_SLOT;
  return true;
}
bool decode_tc2_stage(int N, int i, uint32_t out[2]) {
  int32_t g[5];
  if (!decode_tc2_geometry(N, g) || i < 0 || i >= g[2]) return false;
  const uint2 e = t2_stage_entry(g[0], i, g[1]);
  out[0] = e.x;
  out[1] = e.y;
  return true;
}

// CTAs per instance: 1 unless the launch would leave most SMs idle (then the rows are cut into 128-, 64- or 32-row
// parts as long as all CTAs still fit in one wave; HTSP_DECODE_SPLIT = 1 | 2 | 4 | 8 overrides) or the instance has
// more than 224 rows (a two-tile CTA has three lane quarters in tile 1).
static void t2_choose_split(int Bp, int G, bool debug, int& split, int& rows_per_cta) {
  split = 1;
  rows_per_cta = G;
  if (G > 224) {
    split = 2;
    rows_per_cta = 128;
  }
  if (debug || G <= 32) return;
  int dev = 0, n_sm = 0;
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

// dynamic shared memory of a CTA: A tiles, ring, c0, barriers, TMEM slot (+ the stage table of the two-tile variant)
static size_t t2_smem_bytes(int NP, bool one, bool key_blocks = false) {
  return (size_t)(one ? 1 : 2) * T2_TILE_STAGING +
         (size_t)(one ? (key_blocks ? T2_SLOTS_KB : T2_SLOTS_ONE) : T2_SLOTS) * (key_blocks ? T2_SLOT_KB : T2_SLOT) + (size_t)NP * 4 +
         (size_t)t2_bar_total(one, key_blocks) * 8 + 8 + (one ? 0 : T2_PER_STEP * 8);   // (single-tile CTAs, 208 keys: 115712 bytes -- two per SM with not a byte to spare)
}

template <bool LOGITS, bool DEBUG, bool SPLIT, bool SAMPLE = false, bool ONE = false, int NKB = 1>
static int launch_t2(const T2Args& a, int Bp, cudaStream_t s) {
  const size_t smem = t2_smem_bytes(a.NP * NKB, ONE, NKB > 1);
  if (smem > 227 * 1024) {
    set_error("decode_tc2: N=%d needs %zu bytes of shared memory", a.N, smem);
    return HTSP_ERR_UNSUPPORTED;
  }
  auto kern = decode_tc2_kernel<LOGITS, DEBUG, SPLIT, SAMPLE, ONE, NKB>;
  HTSP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (ONE) HTSP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  kernel_timer_begin(s);
  kern<<<SPLIT ? Bp * ((a.G + a.rows_per_cta - 1) / a.rows_per_cta) : Bp, ONE ? T2_THREADS_ONE : T2_THREADS, smem, s>>>(a);
  kernel_timer_end(s);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

// Single-tile CTAs, two per SM (template parameter ONE), pay off when the launch has more CTAs than SMs; they need the
// SM to really hold two of them (checked once per device: shared memory 2 x (112.97 KB + 1 KB reserved) of 228 KB).
// HTSP_T2_ONE = 0 | 1 overrides the choice.
static bool t2_use_one(int Bp, int G, int NP, bool debug) {
  if (debug) return false;
  const char* env = getenv("HTSP_T2_ONE");
  if (env && atoi(env) == 0) return false;
  int dev = 0, n_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0)
    return false;
  static int fits[64];                 // per device: 0 unknown, 1 two CTAs per SM, -1 not
  const int slot = dev & 63;
  if (fits[slot] == 0) {
    // Resource arithmetic instead of cudaOccupancyMaxActiveBlocksPerMultiprocessor: the occupancy query answers 1 for
    // every kernel that contains tcgen05.alloc, whatever its shared memory and registers, while the hardware does
    // place two such CTAs on an SM when each allocates at most 256 columns (measured on B200: 296 CTAs of a
    // 256-column, 113 KB, 512-thread kernel start within 1 us of each other and finish in one wave).
    auto kern = decode_tc2_kernel<false, false, true, false, true>;
    const size_t smem = t2_smem_bytes(208, true);
    int smem_sm = 0, smem_res = 0, regs_sm = 0;
    cudaFuncAttributes fa;
    const bool ok = cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev) == cudaSuccess &&
                    cudaDeviceGetAttribute(&smem_res, cudaDevAttrReservedSharedMemoryPerBlock, dev) == cudaSuccess &&
                    cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev) == cudaSuccess &&
                    cudaFuncGetAttributes(&fa, kern) == cudaSuccess;
    if (!ok) (void)cudaGetLastError();
    const size_t per_cta = align_up(smem + (size_t)smem_res, 1024);
    fits[slot] = (ok && 2 * per_cta <= (size_t)smem_sm && 2 * T2_THREADS_ONE * ((fa.numRegs + 7) / 8 * 8) <= regs_sm) ? 1 : -1;
    if (getenv("HTSP_T2_VERBOSE"))
      fprintf(stderr, "[decode_tc2] single-tile CTAs: two per SM %s (dynamic smem %zu + reserved %d of %d per SM; regs %d x %d of %d)\n",
              fits[slot] > 0 ? "fit" : "do NOT fit", smem, smem_res, smem_sm, ok ? fa.numRegs : -1, T2_THREADS_ONE, regs_sm);
  }
  if (fits[slot] < 0) return false;
  if (env && atoi(env) != 0) return true;
  (void)NP;
  // Measured (B200, 1184 instances, tools/ab_rollout.py): G <= 128 -- the two-tile CTA holds ONE chain per SM -- gains
  // 22 ... 51 % (N = G = 100: 95.3 K -> 133.7 K instances/s; N = 200, G = 128: 37.0 K -> 45.1 K; N = G = 64: 181 K ->
  // 273 K); two-tile instances gain 0.7 - 1 % (N = G = 200: 24.82 K -> 24.99 K).  Below one CTA per SM the row-split
  // two-tile kernels are the faster choice (one chain per SM either way, deeper ring).
  return (long long)Bp * ((G + 127) / 128) > n_sm && (G <= 128 || Bp >= n_sm);
}

// CTAs of the rollout kernel that share an SM for such a call: 2 = single-tile CTAs, 1 = two-tile (or row-split) CTAs
int decode_tc2_ctas_per_sm(int Bp, int N, int G) {
  if (!decode_tc2_supported(N, G)) return 1;
  if (t2_np_of(N) > 208) return T2_KB_CTAS_PER_SM;
  return t2_use_one(Bp, G, t2_np_of(N), false) ? 2 : 1;
}

int launch_decode_tc2(const float* x, const float* proj, const float* qfix, const float* c0,
                      const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp, int N,
                      int G, const int32_t* forced_pi, int32_t* pi, float* reward,
                      float* logits_out, void* pack_ws, float* dbg, int dbg_step, const uint64_t* sample_seed,
                      float* logp, cudaStream_t s) {
  HTSP_REQUIRE(decode_tc2_supported(N, G), "decode_tc2 shape");
  HTSP_REQUIRE(logp == nullptr || sample_seed != nullptr, "log-probabilities are those of a sampling rollout");
  HTSP_REQUIRE(sample_seed == nullptr || (dbg == nullptr && logits_out == nullptr && forced_pi == nullptr),
               "sampling rollouts take no teacher forcing, logit capture or debug dump");
  HTSP_REQUIRE(dbg == nullptr || G <= 224, "debug dump: at most 224 rows");
  int NP, NKB;                            // keys per block (all padded keys when NKB == 1), key blocks
  t2_blocks_of(N, NP, NKB);
  HTSP_REQUIRE(NKB == 1 || dbg == nullptr, "decode_tc2 with key blocks (N > 208): no debug dump");
  unsigned char* img = (unsigned char*)(((uintptr_t)pack_ws + 1023) & ~(uintptr_t)1023) + T2_TAB_BYTES;   // (the stage table sits in front)
  float* kbar = (float*)(img + align_up((size_t)Bp * t2_img_bytes(NP, NKB), 256));
  decode_tc2_kbar_kernel<<<Bp, 128, 0, s>>>(proj, N, NP, NKB, kbar, (uint2*)(img - T2_TAB_BYTES));
  HTSP_CHECK_LAUNCH();
  {
    const int items = max(NP * NKB * 32, (NP * NKB >> 2) * 128);
    decode_tc2_pack_kernel<<<dim3((items + 255) / 256, min(Bp, 65535)), 256, 0, s>>>(proj, kbar, Bp, N, NP, NKB, img);
    HTSP_CHECK_LAUNCH();
  }
  T2Args a;
  a.x = x; a.proj = proj; a.qfix = qfix; a.c0 = c0; a.img = img;
  a.src = src; a.tgt = tgt; a.first = first; a.forced_pi = forced_pi;
  a.pi = pi; a.reward = reward; a.logits_out = logits_out; a.dbg = dbg; a.logp = logp;
  a.seed = sample_seed ? *sample_seed : 0ull;
  a.N = N; a.NP = NP; a.G = G; a.dbg_step = dbg_step;
  a.tile1_delay = getenv("HTSP_TILE1_DELAY") ? atoi(getenv("HTSP_TILE1_DELAY")) : 0;
  if (NKB > 1) {
    // key blocks: single-tile CTAs of up to 128 rows, as many per instance as its rows need (rows spread evenly in
    // multiples of 32)
    const int n_ctas = (G + 127) / 128;
    a.rows_per_cta = min(128, ((G + n_ctas - 1) / n_ctas + 31) / 32 * 32);
    const bool lg = logits_out != nullptr || forced_pi != nullptr;
    if (sample_seed != nullptr)
      return NKB == 2 ? launch_t2<false, false, true, true, true, 2>(a, Bp, s) : launch_t2<false, false, true, true, true, 3>(a, Bp, s);
    if (NKB == 2) return lg ? launch_t2<true, false, true, false, true, 2>(a, Bp, s) : launch_t2<false, false, true, false, true, 2>(a, Bp, s);
    return lg ? launch_t2<true, false, true, false, true, 3>(a, Bp, s) : launch_t2<false, false, true, false, true, 3>(a, Bp, s);
  }
  int n_split = 1, rows_per_cta = G;
  if (t2_use_one(Bp, G, NP, dbg != nullptr)) {
    // two-CTA instances: 128 rows + the rest (a multiple of 32 rows in the first CTA keeps the number of lane quarters
    // minimal; 96 / 104 / 112 / 120 rows at G = 200 measured within 0.4 % of it)
    a.rows_per_cta = min(G, 128);
    if (getenv("HTSP_T2_ROWS") && atoi(getenv("HTSP_T2_ROWS")) >= 32 && atoi(getenv("HTSP_T2_ROWS")) <= 128 && 2 * atoi(getenv("HTSP_T2_ROWS")) >= G)
      a.rows_per_cta = atoi(getenv("HTSP_T2_ROWS"));     // (A/B experiments)
    if (sample_seed != nullptr) return launch_t2<false, false, true, true, true>(a, Bp, s);
    if (logits_out != nullptr || forced_pi != nullptr) return launch_t2<true, false, true, false, true>(a, Bp, s);
    return launch_t2<false, false, true, false, true>(a, Bp, s);
  }
  t2_choose_split(Bp, G, dbg != nullptr, n_split, rows_per_cta);
  a.rows_per_cta = rows_per_cta;
  const bool split = n_split > 1;
  if (sample_seed != nullptr) return split ? launch_t2<false, false, true, true>(a, Bp, s) : launch_t2<false, false, false, true>(a, Bp, s);
  if (dbg != nullptr) return launch_t2<false, true, false>(a, Bp, s);
  if (logits_out != nullptr || forced_pi != nullptr)      // teacher forcing is served by the LOGITS kernels
    return split ? launch_t2<true, false, true>(a, Bp, s) : launch_t2<true, false, false>(a, Bp, s);
  return split ? launch_t2<false, false, true>(a, Bp, s) : launch_t2<false, false, false>(a, Bp, s);
}

}  // namespace htsp
