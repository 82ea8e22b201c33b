// Encoder self-attention on tensor cores (utils.multi_head_attention, rl4cop/utils/utils.py:12-29, as
// used by MHAEncoderLayer.forward, models.py:30-34): softmax(q k^T / 4) v per (instance, head), no
// mask.  One CTA per (instance, head, block of 208 query rows -- one block up to N = 208): K_h / V_h of ALL keys
// (fp16 hi/lo, ldmatrix-swizzled, 128 bytes per key: N <= 1024) are staged once per CTA in
// shared memory, every warp owns 16 query rows and runs the same flash-style register pipeline as
// the rollout kernel (mma.sync m16n8k16, split-fp16 x3, lazy online softmax).  Inputs and outputs
// are the fp16 hi/lo activations the tcgen05 GEMMs produce/consume: qkv [M,384], out [M,128].
#include "common.cuh"
#include "mma_util.cuh"

namespace htsp {

#ifndef ATT_MIN_CTAS
#define ATT_MIN_CTAS 2
#endif
template <int NCHUNK_T>
__global__ void __launch_bounds__(416, ATT_MIN_CTAS)
enc_attention_mma_kernel(const __half* __restrict__ qkv_hi, const __half* __restrict__ qkv_lo,
                         __half* __restrict__ out_hi, __half* __restrict__ out_lo, int N, int NP) {
  extern __shared__ __align__(128) unsigned char att_smem[];
  const int h = blockIdx.x, b = blockIdx.y, q0 = blockIdx.z * 208;     // first query row of this CTA
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NCHUNK = NCHUNK_T > 0 ? NCHUNK_T : NP / 16;
  const size_t row0 = (size_t)b * N;
  // tiles: K hi | K lo | V hi | V lo, each NP x 16 halves (32 B rows), 16B chunk c of row r stored at
  // c ^ ((r>>2)&1): conflict-free for ldmatrix.  Thread tid stages 16-byte chunk (tid & 1) of key tid >> 1
  // of all four tiles (2 NP <= 416 threads up to N = 208)
  for (int it = tid; it < 2 * NP; it += 416) {
    const int key = it >> 1, chunk = it & 1;
    uint4 v[4];
#pragma unroll
    for (int mat = 0; mat < 4; ++mat) {
      v[mat] = make_uint4(0u, 0u, 0u, 0u);
      if (key < N) {
        const __half* src = ((mat & 1) ? qkv_lo : qkv_hi) + (row0 + key) * (3 * H) + (mat < 2 ? H : 2 * H) +
                            h * DK + chunk * 8;
        v[mat] = __ldg((const uint4*)src);
      }
    }
#pragma unroll
    for (int mat = 0; mat < 4; ++mat)
      *(uint4*)(att_smem + (size_t)mat * NP * 32 + key * 32 + ((chunk ^ ((key >> 2) & 1)) * 16)) = v[mat];
  }
  const uint32_t khi = smem_u32(att_smem), klo = khi + NP * 32, vhi = khi + 2 * NP * 32, vlo = khi + 3 * NP * 32;
  const int q4 = lane & 3, rr = lane >> 2, mi = lane >> 3, l7 = lane & 7;
  const int rA = q0 + warp * 16 + rr, rB = rA + 8;
  const bool okA = rA < N, okB = rB < N;
  // query A fragments: q = hi + lo (exact in fp32), scaled by 1/sqrt(dk)*log2(e), re-split
  uint32_t qh[4], ql[4];
  if (q0 + warp * 16 < N) {
    auto load2 = [&](int row, int col) -> float2 {
      const size_t off = (row0 + min(row, N - 1)) * (3 * H) + h * DK + col;
      const float2 a = __half22float2(*(const __half2*)(qkv_hi + off));
      const float2 c = __half22float2(*(const __half2*)(qkv_lo + off));
      return make_float2((a.x + c.x) * kScoreScale, (a.y + c.y) * kScoreScale);
    };
    const float2 a0 = load2(rA, 2 * q4), a1 = load2(rA, 2 * q4 + 8);
    const float2 b0 = load2(rB, 2 * q4), b1 = load2(rB, 2 * q4 + 8);
    split2p(a0.x, a0.y, qh[0], ql[0]);
    split2p(b0.x, b0.y, qh[1], ql[1]);
    split2p(a1.x, a1.y, qh[2], ql[2]);
    split2p(b1.x, b1.y, qh[3], ql[3]);
  }
  __syncthreads();
  if (q0 + warp * 16 >= N) return;
  float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
  float mA = -1e30f, mB = -1e30f, lA = 0.f, lB = 0.f;
  // two 16-key chunks per iteration: one lazy-rescale vote per 32 keys, and the two chunks' MMA / MUFU
  // chains interleave
#pragma unroll
  for (int kc0 = 0; kc0 < NCHUNK; kc0 += 2) {
    float sv[2][8];            // [chunk][t0 t1 u0 u1 t2 t3 u2 u3] = accumulators s0[0..3], s1[0..3]
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int kc = kc0 + e;
      if (kc < NCHUNK) {
        const int keyK = kc * 16 + (mi >> 1) * 8 + l7;
        const uint32_t offK = keyK * 32 + (((mi & 1) ^ ((keyK >> 2) & 1)) * 16);
        uint32_t bh0, bh1, bh2, bh3, bl0, bl1, bl2, bl3;
        ldsm_x4(khi + offK, bh0, bh1, bh2, bh3);
        ldsm_x4(klo + offK, bl0, bl1, bl2, bl3);
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
        mma16816(s0, qh, bh0, bh1);
        mma16816(s1, qh, bh2, bh3);
        mma16816(s0, qh, bl0, bl1);
        mma16816(s1, qh, bl2, bl3);
        mma16816(s0, ql, bh0, bh1);
        mma16816(s1, ql, bh2, bh3);
#pragma unroll
        for (int i = 0; i < 4; ++i) { sv[e][i] = s0[i]; sv[e][4 + i] = s1[i]; }
        // padded keys (>= N) exist only in the last chunk (the fixed-size instance has N > 16 (NCHUNK-1))
        if (NCHUNK_T == 0 || kc == NCHUNK_T - 1) {
          const int key0 = kc * 16 + 2 * q4;      // this lane's keys: key0, key0+1, key0+8, key0+9
          if (key0 >= N) { sv[e][0] = -INFINITY; sv[e][2] = -INFINITY; }
          if (key0 + 1 >= N) { sv[e][1] = -INFINITY; sv[e][3] = -INFINITY; }
          if (key0 + 8 >= N) { sv[e][4] = -INFINITY; sv[e][6] = -INFINITY; }
          if (key0 + 9 >= N) { sv[e][5] = -INFINITY; sv[e][7] = -INFINITY; }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) sv[e][i] = -INFINITY;
      }
    }
    // rows A: entries 0,1,4,5; rows B: entries 2,3,6,7
    float cA = fmaxf(fmaxf(sv[0][0], sv[0][1]), fmaxf(sv[0][4], sv[0][5]));
    float cB = fmaxf(fmaxf(sv[0][2], sv[0][3]), fmaxf(sv[0][6], sv[0][7]));
    if (kc0 + 1 < NCHUNK) {
      cA = fmaxf(cA, fmaxf(fmaxf(sv[1][0], sv[1][1]), fmaxf(sv[1][4], sv[1][5])));
      cB = fmaxf(cB, fmaxf(fmaxf(sv[1][2], sv[1][3]), fmaxf(sv[1][6], sv[1][7])));
    }
    if (__any_sync(0xffffffffu, (cA > mA + kLazy) || (cB > mB + kLazy))) {
      const float nmA = fmaxf(mA, quad_max(cA)), nmB = fmaxf(mB, quad_max(cB));
      const float aA = ex2_approx(mA - nmA), aB = ex2_approx(mB - nmB);
      mA = nmA; mB = nmB;
      lA *= aA; lB *= aB;
      o0[0] *= aA; o0[1] *= aA; o1[0] *= aA; o1[1] *= aA;
      o0[2] *= aB; o0[3] *= aB; o1[2] *= aB; o1[3] *= aB;
    }
    const float2 nA = make_float2(-mA, -mA), nB = make_float2(-mB, -mB);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int kc = kc0 + e;
      if (kc < NCHUNK) {
        const float2 dt0 = __fadd2_rn(make_float2(sv[e][0], sv[e][1]), nA);
        const float2 du0 = __fadd2_rn(make_float2(sv[e][2], sv[e][3]), nB);
        const float2 dt1 = __fadd2_rn(make_float2(sv[e][4], sv[e][5]), nA);
        const float2 du1 = __fadd2_rn(make_float2(sv[e][6], sv[e][7]), nB);
        const float t0 = ex2_approx(dt0.x), t1 = ex2_approx(dt0.y), t2 = ex2_approx(dt1.x), t3 = ex2_approx(dt1.y);
        const float u0 = ex2_approx(du0.x), u1 = ex2_approx(du0.y), u2 = ex2_approx(du1.x), u3 = ex2_approx(du1.y);
        lA += (t0 + t1) + (t2 + t3);
        lB += (u0 + u1) + (u2 + u3);
        uint32_t ph[4], pl[4];
        split2p(t0, t1, ph[0], pl[0]);
        split2p(u0, u1, ph[1], pl[1]);
        split2p(t2, t3, ph[2], pl[2]);
        split2p(u2, u3, ph[3], pl[3]);
        const int keyV = kc * 16 + (mi & 1) * 8 + l7;
        const uint32_t offV = keyV * 32 + (((mi >> 1) ^ ((keyV >> 2) & 1)) * 16);
        uint32_t vh0, vh1, vh2, vh3, vl0, vl1, vl2, vl3;
        ldsm_x4_t(vhi + offV, vh0, vh1, vh2, vh3);
        ldsm_x4_t(vlo + offV, vl0, vl1, vl2, vl3);
        mma16816(o0, ph, vh0, vh1);
        mma16816(o1, ph, vh2, vh3);
        mma16816(o0, ph, vl0, vl1);
        mma16816(o1, ph, vl2, vl3);
        mma16816(o0, pl, vh0, vh1);
        mma16816(o1, pl, vh2, vh3);
      }
    }
  }
  const float iA = 1.f / quad_sum(lA), iB = 1.f / quad_sum(lB);
  uint32_t h0, l0, h1, l1;
  if (okA) {
    const size_t off = (row0 + rA) * H + h * DK + 2 * q4;
    split2p(o0[0] * iA, o0[1] * iA, h0, l0);
    split2p(o1[0] * iA, o1[1] * iA, h1, l1);
    *(uint32_t*)(out_hi + off) = h0; *(uint32_t*)(out_lo + off) = l0;
    *(uint32_t*)(out_hi + off + 8) = h1; *(uint32_t*)(out_lo + off + 8) = l1;
  }
  if (okB) {
    const size_t off = (row0 + rB) * H + h * DK + 2 * q4;
    split2p(o0[2] * iB, o0[3] * iB, h0, l0);
    split2p(o1[2] * iB, o1[3] * iB, h1, l1);
    *(uint32_t*)(out_hi + off) = h0; *(uint32_t*)(out_lo + off) = l0;
    *(uint32_t*)(out_hi + off + 8) = h1; *(uint32_t*)(out_lo + off + 8) = l1;
  }
}

int launch_enc_attention_mma(const __half* qkv_hi, const __half* qkv_lo, __half* out_hi, __half* out_lo,
                             int Bp, int N, cudaStream_t s) {
  const int NP = (N + 15) / 16 * 16;
  HTSP_REQUIRE(N >= 1 && NP <= 1024, "enc_attention_mma supports N <= 1024");
  const size_t smem = (size_t)4 * NP * 32;
  dim3 grid(NH, Bp, (NP + 207) / 208);
  if (NP == 208) {
    enc_attention_mma_kernel<13><<<grid, 416, smem, s>>>(qkv_hi, qkv_lo, out_hi, out_lo, N, NP);
  } else {
    if (smem > 48 * 1024)
      HTSP_CHECK_CUDA(cudaFuncSetAttribute(enc_attention_mma_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    enc_attention_mma_kernel<0><<<grid, 416, smem, s>>>(qkv_hi, qkv_lo, out_hi, out_lo, N, NP);
  }
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

}  // namespace htsp
