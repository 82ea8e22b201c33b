// Blackwell-native dense layers for the encoder and the decoder projections (K2, prep):
//   Y[M,Nout] = epi( X[M,K] * W[Nout,K]^T )      (torch.nn.Linear, rl4cop/models/models.py:22-38,47,
//                                                  375-402), split-fp16 arithmetic:
//   X = Xhi + Xlo, W = Whi + Wlo (fp16 each);  acc = Xhi*Whi + Xhi*Wlo + Xlo*Whi  in fp32.
//
// One persistent, warp-specialised CTA per SM:
//   warp 0   : TMA producer  -- cp.async.bulk.tensor.2d (SWIZZLE_128B boxes of 128 rows x 64 halves)
//              of the four operand tiles per k-block into a 3-stage shared-memory ring
//   warp 1   : MMA issuer    -- tcgen05.mma.cta_group::1.kind::f16, M=128, N=128, K=16, three MMAs
//              per k16 step (hi*hi, hi*lo, lo*hi) into one of two TMEM accumulators (128 columns each)
//   warps 2-17: epilogue     -- four warps per TMEM lane quarter (32 of the tile's 128 columns each):
//              the residual rows are requested BEFORE the accumulator is waited for, then tcgen05.ld
//              (32 lanes x 32 columns), transpose through shared memory (two 16-column halves, XOR-swizzled
//              16-byte slots, st/ld.shared.v4) so that global loads / stores run along the rows,
//              bias / ReLU / residual / eval-BatchNorm affine in registers, stores fp32 and/or the fp16
//              hi/lo split that the next layer's TMA reads; launches that only emit the fp16 split (QKV, FF1)
//              write each 32 x 32 box of halves into a SWIZZLE_64B tile and hand it to a TMA store;
//              overlaps with the next tile's MMAs through the second accumulator.
// Synchronisation is mbarrier-only (full/empty per stage, tmem_full/tmem_empty per accumulator);
// tcgen05.commit releases shared-memory stages and publishes accumulators.
#include <cuda.h>
#include "common.cuh"
#include "mma_util.cuh"

namespace htsp {

constexpr int GT_BM = 128, GT_BN = 128, GT_BK = 64, GT_STAGES = 3;
constexpr int GT_TILE_BYTES = GT_BM * GT_BK * 2;            // 16 KB, one operand tile
constexpr int GT_STAGE_BYTES = 4 * GT_TILE_BYTES;           // A_hi, A_lo, B_hi, B_lo
constexpr int GT_EPI_WARPS = 16;
constexpr int GT_THREADS = 64 + 32 * GT_EPI_WARPS;          // TMA producer, MMA issuer, 16 epilogue warps
constexpr uint32_t GT_TMEM_COLS = 256;
constexpr int GT_EPI_BYTES = GT_EPI_WARPS * 32 * 16 * 4;    // one 32 x 16 fp32 transpose tile per epilogue warp
constexpr size_t GT_SMEM_BYTES = (size_t)GT_STAGES * GT_STAGE_BYTES + 1024 /*align*/ + 512 /*barriers*/ + GT_EPI_BYTES;

struct GemmEpi {
  const float* bias;       // [Nout] or null
  const float* residual;   // [M,Nout] fp32 or null
  const float* bn_alpha;   // [Nout] or null
  const float* bn_beta;
  float* y32;              // [M,Nout] or null
  __half* yhi;             // [M,Nout] or null (both or neither)
  __half* ylo;
  int M, Nout, K, relu;
};

__device__ __forceinline__ uint32_t gt_smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void gt_mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void gt_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void gt_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug traps (reported as a CUDA error) instead of hanging the GPU
__device__ __forceinline__ void gt_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void gt_tma_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start>>4 | LBO=0 | SBO = 8 rows * 128 B = 1024 B (>>4) | version 1 | layout SWIZZLE_128B (2)
__device__ __forceinline__ uint64_t gt_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void gt_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void gt_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void gt_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void gt_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void gt_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void gt_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 gt_lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// TMA store of one 32-row x 32-column fp16 box (64-byte rows, SWIZZLE_64B) from shared memory
__device__ __forceinline__ void gt_tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// every earlier bulk store of this thread has finished READING its shared-memory source
__device__ __forceinline__ void gt_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_x3_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                  const __grid_constant__ CUtensorMap tmY_hi, const __grid_constant__ CUtensorMap tmY_lo,
                  const GemmEpi E) {
  extern __shared__ unsigned char gt_smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  unsigned char* smem = (unsigned char*)(((uintptr_t)gt_smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + (size_t)GT_STAGES * GT_STAGE_BYTES);
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * GT_STAGES + 4);
  float* epi_smem = (float*)(smem + (size_t)GT_STAGES * GT_STAGE_BYTES + 512);   // 512-byte aligned: SWIZZLE_64B atoms
  const uint32_t full0 = gt_smem_u32(bars), empty0 = gt_smem_u32(bars + GT_STAGES);
  const uint32_t tfull0 = gt_smem_u32(bars + 2 * GT_STAGES), tempty0 = gt_smem_u32(bars + 2 * GT_STAGES + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int num_m = (E.M + GT_BM - 1) / GT_BM, num_n = E.Nout / GT_BN;
  const int num_tiles = num_m * num_n, num_kb = E.K / GT_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < GT_STAGES; ++s) { gt_mbar_init(full0 + 8 * s, 1); gt_mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { gt_mbar_init(tfull0 + 8 * a, 1); gt_mbar_init(tempty0 + 8 * a, GT_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {   // TMEM allocation is warp-collective; the same warp frees it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gt_smem_u32(tmem_slot)),
                 "r"(GT_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  gt_fence_before();
  __syncthreads();
  gt_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % GT_STAGES, use = it / GT_STAGES;
          gt_mbar_wait(empty0 + 8 * s, (use & 1) ^ 1);
          const uint32_t dst = gt_smem_u32(smem + (size_t)s * GT_STAGE_BYTES);
          gt_mbar_expect_tx(full0 + 8 * s, GT_STAGE_BYTES);
          gt_tma_2d(dst + 0 * GT_TILE_BYTES, &tmA_hi, kb * GT_BK, m_blk * GT_BM, full0 + 8 * s);
          gt_tma_2d(dst + 1 * GT_TILE_BYTES, &tmA_lo, kb * GT_BK, m_blk * GT_BM, full0 + 8 * s);
          gt_tma_2d(dst + 2 * GT_TILE_BYTES, &tmB_hi, kb * GT_BK, n_blk * GT_BN, full0 + 8 * s);
          gt_tma_2d(dst + 3 * GT_TILE_BYTES, &tmB_lo, kb * GT_BK, n_blk * GT_BN, full0 + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D=F32 (bit 4), A=B=F16,
      // both K-major, N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | ((uint32_t)(GT_BN >> 3) << 17) | ((uint32_t)(GT_BM >> 4) << 24);
      uint32_t it = 0, local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const uint32_t acc = local & 1, acc_use = local >> 1;
        gt_mbar_wait(tempty0 + 8 * acc, (acc_use & 1) ^ 1);
        gt_fence_after();
        const uint32_t tmem_d = tmem_base + acc * GT_BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % GT_STAGES, use = it / GT_STAGES;
          gt_mbar_wait(full0 + 8 * s, use & 1);
          gt_fence_after();
          const uint32_t base = gt_smem_u32(smem + (size_t)s * GT_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < GT_BK / 16; ++k) {
            const uint64_t a_hi = gt_desc(base + 0 * GT_TILE_BYTES + k * 32);
            const uint64_t a_lo = gt_desc(base + 1 * GT_TILE_BYTES + k * 32);
            const uint64_t b_hi = gt_desc(base + 2 * GT_TILE_BYTES + k * 32);
            const uint64_t b_lo = gt_desc(base + 3 * GT_TILE_BYTES + k * 32);
            gt_mma(tmem_d, a_hi, b_hi, idesc, (kb | k) != 0);
            gt_mma(tmem_d, a_hi, b_lo, idesc, 1);
            gt_mma(tmem_d, a_lo, b_hi, idesc, 1);
          }
          gt_commit(empty0 + 8 * s);          // frees the stage once these MMAs have read it
        }
        gt_commit(tfull0 + 8 * acc);          // accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;                   // TMEM lane quarter this warp may access
    const int c = (warp - 2) >> 2;            // its 32-column chunk of the tile
    // transpose tile of this warp: 32 rows x 16 columns fp32 (64-byte rows = four 16-byte slots); row r keeps
    // slot s at s ^ ((r >> 1) & 3): the quarter-warp phases of both the row-wise writes (lane = row) and the
    // reads (8 rows x 4 slots per pass) touch 8 distinct 16-byte bank groups
    const uint32_t tr = gt_smem_u32(epi_smem) + (uint32_t)(warp - 2) * (32 * 16 * 4);
    const uint32_t wr_row = tr + (uint32_t)lane * 64, wr_x = (uint32_t)(lane >> 1) & 3;
    const int rsub = lane >> 2, slot = lane & 3;             // read phase: 8 rows x 4 slots per pass
    const bool packed = E.y32 == nullptr && E.residual == nullptr && E.bn_alpha == nullptr && E.yhi != nullptr;
    const bool raw32 = E.y32 != nullptr && E.yhi == nullptr && E.residual == nullptr && E.bn_alpha == nullptr &&
                       E.bias == nullptr && !E.relu;
    uint32_t local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const uint32_t acc = local & 1, acc_use = local >> 1;
      const int col0 = n_blk * GT_BN + c * 32;
      const int rowb = m_blk * GT_BM + q * 32;
      // residual rows first: the loads fly while the MMAs of this tile are still running
      float4 res[2][4];
      if (E.residual != nullptr) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) {
            const int row = rowb + pp * 8 + rsub;
            res[hh][pp] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < E.M) res[hh][pp] = __ldg((const float4*)(E.residual + (size_t)row * E.Nout + col0 + hh * 16 + slot * 4));
          }
      }
      gt_mbar_wait(tfull0 + 8 * acc, acc_use & 1);
      gt_fence_after();
      uint32_t r[32];
      gt_tmem_ld32(tmem_base + acc * GT_BN + c * 32 + ((uint32_t)(q * 32) << 16), r);
      // the accumulator is in registers: hand the TMEM buffer back before the stores
      gt_fence_before();
      __syncwarp();
      if (lane == 0) gt_mbar_arrive(tempty0 + 8 * acc);
      if (packed) {
        // fp16 hi/lo outputs only (QKV, FF1): bias / ReLU / split in the row-owner lane (TMEM lane = output row)
        uint32_t hl[2][16];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
          if (E.bias != nullptr) bb = __ldg((const float4*)(E.bias + col0 + 4 * k));
          float v0 = __uint_as_float(r[4 * k]) + bb.x, v1 = __uint_as_float(r[4 * k + 1]) + bb.y;
          float v2 = __uint_as_float(r[4 * k + 2]) + bb.z, v3 = __uint_as_float(r[4 * k + 3]) + bb.w;
          if (E.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
          split2p(v0, v1, hl[0][2 * k], hl[1][2 * k]);
          split2p(v2, v3, hl[0][2 * k + 1], hl[1][2 * k + 1]);
        }
        // the row-owner lane writes its 64-byte row of halves into the SWIZZLE_64B box (16-byte chunk j at
        // j ^ ((row >> 1) & 3): conflict-free), one TMA store per part moves the 32 x 32 box out -- rows past M
        // are clipped by the tensor map
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          if (lane == 0) gt_bulk_wait_read();       // the previous store out of this buffer has read it
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            gt_sts128(wr_row + (((uint32_t)j ^ wr_x) << 4), hl[part][4 * j], hl[part][4 * j + 1], hl[part][4 * j + 2],
                      hl[part][4 * j + 3]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) gt_tma_store_2d(part ? &tmY_lo : &tmY_hi, tr, col0, rowb);
        }
        continue;
      }
      if (raw32) {
        // fp32 output without an epilogue (the decoder projection tables): the row-owner lane writes its 16
        // columns of a half chunk (64 bytes) into the SWIZZLE_64B box, one TMA store per half
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (lane == 0) gt_bulk_wait_read();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            gt_sts128(wr_row + (((uint32_t)j ^ wr_x) << 4), r[hh * 16 + 4 * j], r[hh * 16 + 4 * j + 1],
                      r[hh * 16 + 4 * j + 2], r[hh * 16 + 4 * j + 3]);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) gt_tma_store_2d(&tmY_hi, tr, col0 + hh * 16, rowb);
        }
        continue;
      }
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        // TMEM hands every lane one ROW; global memory wants lanes along the columns
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), al4 = make_float4(1.f, 1.f, 1.f, 1.f), be4 = b4;
        if (E.bias != nullptr) b4 = __ldg((const float4*)(E.bias + col0 + hh * 16 + slot * 4));
        if (E.bn_alpha != nullptr) {
          al4 = __ldg((const float4*)(E.bn_alpha + col0 + hh * 16 + slot * 4));
          be4 = __ldg((const float4*)(E.bn_beta + col0 + hh * 16 + slot * 4));
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
          gt_sts128(wr_row + (((uint32_t)j ^ wr_x) << 4), r[hh * 16 + 4 * j], r[hh * 16 + 4 * j + 1],
                    r[hh * 16 + 4 * j + 2], r[hh * 16 + 4 * j + 3]);
        __syncwarp();
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
          const int rl = pp * 8 + rsub;
          const int row = rowb + rl;
          const float4 t4 = gt_lds128(tr + (uint32_t)rl * 64 + ((((uint32_t)slot) ^ (((uint32_t)rl >> 1) & 3)) << 4));
          float v[4] = {t4.x, t4.y, t4.z, t4.w};
          if (row < E.M) {
            const size_t off = (size_t)row * E.Nout + col0 + hh * 16 + slot * 4;
            v[0] += b4.x; v[1] += b4.y; v[2] += b4.z; v[3] += b4.w;
            if (E.relu) {
#pragma unroll
              for (int t = 0; t < 4; ++t) v[t] = fmaxf(v[t], 0.f);
            }
            if (E.residual != nullptr) {
              const float4 x = res[hh][pp];
              v[0] = x.x + v[0]; v[1] = x.y + v[1]; v[2] = x.z + v[2]; v[3] = x.w + v[3];
            }
            if (E.bn_alpha != nullptr) {
              v[0] = fmaf(v[0], al4.x, be4.x); v[1] = fmaf(v[1], al4.y, be4.y);
              v[2] = fmaf(v[2], al4.z, be4.z); v[3] = fmaf(v[3], al4.w, be4.w);
            }
            if (E.y32 != nullptr) *(float4*)(E.y32 + off) = make_float4(v[0], v[1], v[2], v[3]);
            if (E.yhi != nullptr) {
              __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
              const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
              __half2 l0 = __floats2half2_rn(v[0] - f0.x, v[1] - f0.y);
              __half2 l1 = __floats2half2_rn(v[2] - f1.x, v[3] - f1.y);
              *(uint2*)(E.yhi + off) = make_uint2(*(uint32_t*)&h0, *(uint32_t*)&h1);
              *(uint2*)(E.ylo + off) = make_uint2(*(uint32_t*)&l0, *(uint32_t*)&l1);
            }
          }
        }
      }
    }
  }
  if (warp >= 2 && lane == 0) gt_bulk_wait_read();     // no bulk store may still read this CTA's shared memory
  gt_fence_before();
  __syncthreads();
  if (warp == 1) {
    gt_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(GT_TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Fused feed-forward block of an encoder layer (rl4cop/models/models.py:22-26, 36-37):
//     Y = BN2( X + relu(X W1^T + b1) W2^T + b2 ),   X [M,128], W1 [512,128], W2 [128,512]
// in ONE kernel: the [M,512] hidden activation (fp16 hi/lo: 2 KB per row written by FF1 and read back by FF2, 3.4 GB
// of HBM traffic per layer at config 2) never leaves the SM.  Per 128-row tile the hidden units are processed in
// eight chunks of 64:
//     acc1[c&1] (TMEM, 64 columns)  = X W1[64c..64c+63]^T                 24 MMAs (K = 128: 8 k16 steps x 3 split products)
//     H (shared, K-major SWIZZLE_128B, fp16 hi/lo) = split(relu(acc1 + b1))    epilogue-1 warps, row = TMEM lane
//     acc2 (TMEM, 128 columns)     += H W2[:, 64c..64c+63]^T               12 MMAs (K = 64)
// and after the last chunk the epilogue-2 warps apply b2 / residual / BatchNorm and write fp32 + fp16 hi/lo rows.
// The MMAs, their order and the epilogue arithmetic are those of the two unfused launches: the result is bit-identical.
// Warps: 0 TMA producer (X tile once per tile; W1 chunks through two stages, W2 chunks through one), 1 MMA issuer
// (GEMM1 of chunk c is issued before GEMM2 of chunk c-1, so the tensor pipe has work while epilogue 1 runs), 2-9
// epilogue 1 (lane quarter x 32-column half), 10-17 epilogue 2 (lane quarter x two 32-column chunks; overlaps the next
// tile through the second acc2 buffer).
constexpr int FF_CH = 64, FF_NCH = FF / FF_CH;
constexpr int FF_A_BYTES = 4 * GT_TILE_BYTES;        // X tile: hi k0-63 | hi k64-127 | lo k0-63 | lo k64-127
constexpr int FF_W1_TILE = FF_CH * GT_BK * 2;        // [64 hidden x 64 k] halves = 8 KB
constexpr int FF_W1_BYTES = 4 * FF_W1_TILE;          // hi k0 | hi k1 | lo k0 | lo k1
constexpr int FF_W2_BYTES = 2 * GT_TILE_BYTES;       // [128 out x 64 k] hi | lo
constexpr int FF_H_BYTES = 2 * GT_TILE_BYTES;        // [128 rows x 64 k] hi | lo
constexpr int FF_E1_WARPS = 8, FF_E2_WARPS = 8;
constexpr int FF_THREADS = 64 + 32 * (FF_E1_WARPS + FF_E2_WARPS);
constexpr int FF_EPI_BYTES = FF_E2_WARPS * 32 * 16 * 4;
constexpr int FF_NBARS = 20;
constexpr size_t FF_SMEM_BYTES = 1024 + (size_t)FF_A_BYTES + 2 * FF_W1_BYTES + FF_W2_BYTES + FF_H_BYTES + FF_EPI_BYTES + 256;
static_assert(FF_SMEM_BYTES <= 227 * 1024, "fused feed-forward: shared memory");

struct FfArgs {
  const float* b1;         // [512]
  const float* b2;         // [128]
  const float* residual;   // [M,128] fp32 (= X)
  const float* bn_alpha;   // [128]
  const float* bn_beta;
  float* y32;              // [M,128]
  __half* yhi;             // [M,128]
  __half* ylo;
  int M;
};

__global__ void __launch_bounds__(FF_THREADS, 1)
ff_fused_tc_kernel(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                   const __grid_constant__ CUtensorMap tmW1_hi, const __grid_constant__ CUtensorMap tmW1_lo,
                   const __grid_constant__ CUtensorMap tmW2_hi, const __grid_constant__ CUtensorMap tmW2_lo,
                   const FfArgs E) {
  extern __shared__ unsigned char gt_smem_raw[];
  unsigned char* smem = (unsigned char*)(((uintptr_t)gt_smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;
  unsigned char* sW1 = sA + FF_A_BYTES;               // two stages
  unsigned char* sW2 = sW1 + 2 * FF_W1_BYTES;
  unsigned char* sH = sW2 + FF_W2_BYTES;
  float* epi_smem = (float*)(sH + FF_H_BYTES);
  uint64_t* bars = (uint64_t*)((unsigned char*)epi_smem + FF_EPI_BYTES);
  uint32_t* tmem_slot = (uint32_t*)(bars + FF_NBARS);
  const uint32_t b0 = gt_smem_u32(bars);
  // barrier slots
  const uint32_t a_full = b0, a_empty = b0 + 8, w1_full = b0 + 16 /*[2]*/, w1_empty = b0 + 32 /*[2]*/, w2_full = b0 + 48,
                 w2_empty = b0 + 56, acc1_full = b0 + 64 /*[2]*/, acc1_empty = b0 + 80 /*[2]*/, h_full = b0 + 96,
                 h_empty = b0 + 104, acc2_full = b0 + 112 /*[2]*/, acc2_empty = b0 + 128 /*[2]*/;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (E.M + GT_BM - 1) / GT_BM;

  if (threadIdx.x == 0) {
    gt_mbar_init(a_full, 1); gt_mbar_init(a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      gt_mbar_init(w1_full + 8 * i, 1); gt_mbar_init(w1_empty + 8 * i, 1);
      gt_mbar_init(acc1_full + 8 * i, 1); gt_mbar_init(acc1_empty + 8 * i, FF_E1_WARPS);
      gt_mbar_init(acc2_full + 8 * i, 1); gt_mbar_init(acc2_empty + 8 * i, FF_E2_WARPS);
    }
    gt_mbar_init(w2_full, 1); gt_mbar_init(w2_empty, 1);
    gt_mbar_init(h_full, FF_E1_WARPS); gt_mbar_init(h_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gt_smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  gt_fence_before();
  __syncthreads();
  gt_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_acc1 = tmem_base, tmem_acc2 = tmem_base + 2 * FF_CH;     // acc1: 2 x 64 columns, acc2: 2 x 128

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t tl = 0, cw = 0;
      for (int m_blk = blockIdx.x; m_blk < num_m; m_blk += gridDim.x, ++tl) {
        gt_mbar_wait(a_empty, (tl & 1) ^ 1);
        gt_mbar_expect_tx(a_full, FF_A_BYTES);
        const uint32_t dA = gt_smem_u32(sA);
        gt_tma_2d(dA + 0 * GT_TILE_BYTES, &tmX_hi, 0, m_blk * GT_BM, a_full);
        gt_tma_2d(dA + 1 * GT_TILE_BYTES, &tmX_hi, GT_BK, m_blk * GT_BM, a_full);
        gt_tma_2d(dA + 2 * GT_TILE_BYTES, &tmX_lo, 0, m_blk * GT_BM, a_full);
        gt_tma_2d(dA + 3 * GT_TILE_BYTES, &tmX_lo, GT_BK, m_blk * GT_BM, a_full);
        // W1 chunks run one chunk ahead of the W2 chunks: W2 has ONE buffer, which is free only when GEMM 2 of the
        // previous chunk has completed, and nothing may queue behind that wait but the W2 load itself
        auto load_w1 = [&](uint32_t cwi, int c) {
          const uint32_t s = cwi & 1, use = cwi >> 1;
          gt_mbar_wait(w1_empty + 8 * s, (use & 1) ^ 1);
          gt_mbar_expect_tx(w1_full + 8 * s, FF_W1_BYTES);
          const uint32_t d1 = gt_smem_u32(sW1 + (size_t)s * FF_W1_BYTES);
          gt_tma_2d(d1 + 0 * FF_W1_TILE, &tmW1_hi, 0, c * FF_CH, w1_full + 8 * s);
          gt_tma_2d(d1 + 1 * FF_W1_TILE, &tmW1_hi, GT_BK, c * FF_CH, w1_full + 8 * s);
          gt_tma_2d(d1 + 2 * FF_W1_TILE, &tmW1_lo, 0, c * FF_CH, w1_full + 8 * s);
          gt_tma_2d(d1 + 3 * FF_W1_TILE, &tmW1_lo, GT_BK, c * FF_CH, w1_full + 8 * s);
        };
        for (int c = 0; c < FF_NCH; ++c, ++cw) {
          if (c == 0) load_w1(cw, 0);
          if (c + 1 < FF_NCH) load_w1(cw + 1, c + 1);
          gt_mbar_wait(w2_empty, (cw & 1) ^ 1);
          gt_mbar_expect_tx(w2_full, FF_W2_BYTES);
          const uint32_t d2 = gt_smem_u32(sW2);
          gt_tma_2d(d2, &tmW2_hi, c * FF_CH, 0, w2_full);
          gt_tma_2d(d2 + GT_TILE_BYTES, &tmW2_lo, c * FF_CH, 0, w2_full);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc1 = (1u << 4) | ((uint32_t)(FF_CH >> 3) << 17) | ((uint32_t)(GT_BM >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | ((uint32_t)(GT_BN >> 3) << 17) | ((uint32_t)(GT_BM >> 4) << 24);
      const uint32_t aA = gt_smem_u32(sA), aW2 = gt_smem_u32(sW2), aH = gt_smem_u32(sH);
      uint32_t tl = 0, cw = 0, cg = 0;
      // acc2 (+)= H W2_c^T for the chunk whose H tile and W2 tile are in shared memory
      auto gemm2 = [&](uint32_t acc2, bool first) {
        gt_mbar_wait(w2_full, cg & 1);
        gt_mbar_wait(h_full, cg & 1);
        gt_fence_after();
#pragma unroll
        for (int k = 0; k < GT_BK / 16; ++k) {
          const uint64_t a_hi = gt_desc(aH + k * 32), a_lo = gt_desc(aH + GT_TILE_BYTES + k * 32);
          const uint64_t b_hi = gt_desc(aW2 + k * 32), b_lo = gt_desc(aW2 + GT_TILE_BYTES + k * 32);
          gt_mma(acc2, a_hi, b_hi, idesc2, !(first && k == 0));
          gt_mma(acc2, a_hi, b_lo, idesc2, 1);
          gt_mma(acc2, a_lo, b_hi, idesc2, 1);
        }
        gt_commit(h_empty);
        gt_commit(w2_empty);
        ++cg;
      };
      for (int m_blk = blockIdx.x; m_blk < num_m; m_blk += gridDim.x, ++tl) {
        const uint32_t b2 = tl & 1;
        gt_mbar_wait(acc2_empty + 8 * b2, ((tl >> 1) & 1) ^ 1);
        gt_mbar_wait(a_full, tl & 1);
        gt_fence_after();
        const uint32_t acc2 = tmem_acc2 + b2 * GT_BN;
        for (int c = 0; c < FF_NCH; ++c, ++cw) {
          const uint32_t s = cw & 1, use = cw >> 1;
          gt_mbar_wait(w1_full + 8 * s, use & 1);
          gt_mbar_wait(acc1_empty + 8 * s, (use & 1) ^ 1);
          gt_fence_after();
          const uint32_t acc1 = tmem_acc1 + s * FF_CH;
          const uint32_t aW1 = gt_smem_u32(sW1 + (size_t)s * FF_W1_BYTES);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int k = 0; k < GT_BK / 16; ++k) {
              const uint64_t a_hi = gt_desc(aA + (0 + kb) * GT_TILE_BYTES + k * 32);
              const uint64_t a_lo = gt_desc(aA + (2 + kb) * GT_TILE_BYTES + k * 32);
              const uint64_t b_hi = gt_desc(aW1 + (0 + kb) * FF_W1_TILE + k * 32);
              const uint64_t b_lo = gt_desc(aW1 + (2 + kb) * FF_W1_TILE + k * 32);
              gt_mma(acc1, a_hi, b_hi, idesc1, (kb | k) != 0);
              gt_mma(acc1, a_hi, b_lo, idesc1, 1);
              gt_mma(acc1, a_lo, b_hi, idesc1, 1);
            }
          gt_commit(w1_empty + 8 * s);
          gt_commit(acc1_full + 8 * s);
          if (c == FF_NCH - 1) gt_commit(a_empty);       // the X tile may be replaced by the next one
          if (c >= 1) gemm2(acc2, c == 1);
        }
        gemm2(acc2, false);
        gt_commit(acc2_full + 8 * b2);
      }
    }
  } else if (warp < 2 + FF_E1_WARPS) {
    // ------------------------------------------------------------------ epilogue 1: acc1 -> relu -> H (fp16 hi/lo, A operand of GEMM 2)
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t row = (uint32_t)(q * 32 + lane);
    uint32_t ce = 0;
    for (int m_blk = blockIdx.x; m_blk < num_m; m_blk += gridDim.x) {
      for (int c = 0; c < FF_NCH; ++c, ++ce) {
        const uint32_t s = ce & 1, use = ce >> 1;
        gt_mbar_wait(acc1_full + 8 * s, use & 1);
        gt_fence_after();
        uint32_t r[32];
        gt_tmem_ld32(tmem_acc1 + s * FF_CH + half * 32 + ((uint32_t)(q * 32) << 16), r);
        gt_fence_before();
        __syncwarp();
        if (lane == 0) gt_mbar_arrive(acc1_empty + 8 * s);
        uint32_t hi[16], lo[16];
        const float* bp = E.b1 + c * FF_CH + half * 32;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 bb = __ldg((const float4*)(bp + 4 * k));
          const float v0 = fmaxf(__uint_as_float(r[4 * k]) + bb.x, 0.f), v1 = fmaxf(__uint_as_float(r[4 * k + 1]) + bb.y, 0.f);
          const float v2 = fmaxf(__uint_as_float(r[4 * k + 2]) + bb.z, 0.f), v3 = fmaxf(__uint_as_float(r[4 * k + 3]) + bb.w, 0.f);
          split2p(v0, v1, hi[2 * k], lo[2 * k]);
          split2p(v2, v3, hi[2 * k + 1], lo[2 * k + 1]);
        }
        gt_mbar_wait(h_empty, (ce & 1) ^ 1);             // GEMM 2 of the previous chunk has read H
        const uint32_t hb = gt_smem_u32(sH) + row * 128u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = (((uint32_t)(4 * half + j)) ^ (row & 7u)) << 4;     // K-major SWIZZLE_128B
          gt_sts128(hb + off, hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
          gt_sts128(hb + GT_TILE_BYTES + off, lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) gt_mbar_arrive(h_full);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue 2: acc2 -> + b2 + residual -> BatchNorm -> fp32 and fp16 hi/lo rows
    const int q = warp & 3, cgp = (warp - 2 - FF_E1_WARPS) >> 2;
    const uint32_t tr = gt_smem_u32(epi_smem) + (uint32_t)(warp - 2 - FF_E1_WARPS) * (32 * 16 * 4);
    const uint32_t wr_row = tr + (uint32_t)lane * 64, wr_x = (uint32_t)(lane >> 1) & 3;
    const int rsub = lane >> 2, slot = lane & 3;
    uint32_t tl = 0;
    for (int m_blk = blockIdx.x; m_blk < num_m; m_blk += gridDim.x, ++tl) {
      const uint32_t b2 = tl & 1;
      const int rowb = m_blk * GT_BM + q * 32;
      gt_mbar_wait(acc2_full + 8 * b2, (tl >> 1) & 1);
      gt_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int col0 = (2 * cgp + cc) * 32;
        uint32_t r[32];
        gt_tmem_ld32(tmem_acc2 + b2 * GT_BN + col0 + ((uint32_t)(q * 32) << 16), r);
        if (cc == 1) {       // both chunks of this warp are in registers: hand the accumulator back
          gt_fence_before();
          __syncwarp();
          if (lane == 0) gt_mbar_arrive(acc2_empty + 8 * b2);
        }
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const float4 b4 = __ldg((const float4*)(E.b2 + col0 + hh * 16 + slot * 4));
          const float4 al4 = __ldg((const float4*)(E.bn_alpha + col0 + hh * 16 + slot * 4));
          const float4 be4 = __ldg((const float4*)(E.bn_beta + col0 + hh * 16 + slot * 4));
          float4 res[4];
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) {
            const int rw = rowb + pp * 8 + rsub;
            res[pp] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rw < E.M) res[pp] = __ldg((const float4*)(E.residual + (size_t)rw * H + col0 + hh * 16 + slot * 4));
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            gt_sts128(wr_row + (((uint32_t)j ^ wr_x) << 4), r[hh * 16 + 4 * j], r[hh * 16 + 4 * j + 1],
                      r[hh * 16 + 4 * j + 2], r[hh * 16 + 4 * j + 3]);
          __syncwarp();
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) {
            const int rl = pp * 8 + rsub;
            const int rw = rowb + rl;
            const float4 t4 = gt_lds128(tr + (uint32_t)rl * 64 + ((((uint32_t)slot) ^ (((uint32_t)rl >> 1) & 3)) << 4));
            float v[4] = {t4.x, t4.y, t4.z, t4.w};
            if (rw < E.M) {
              const size_t off = (size_t)rw * H + col0 + hh * 16 + slot * 4;
              v[0] += b4.x; v[1] += b4.y; v[2] += b4.z; v[3] += b4.w;
              const float4 x = res[pp];
              v[0] = x.x + v[0]; v[1] = x.y + v[1]; v[2] = x.z + v[2]; v[3] = x.w + v[3];
              v[0] = fmaf(v[0], al4.x, be4.x); v[1] = fmaf(v[1], al4.y, be4.y);
              v[2] = fmaf(v[2], al4.z, be4.z); v[3] = fmaf(v[3], al4.w, be4.w);
              *(float4*)(E.y32 + off) = make_float4(v[0], v[1], v[2], v[3]);
              __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
              const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
              __half2 l0 = __floats2half2_rn(v[0] - f0.x, v[1] - f0.y);
              __half2 l1 = __floats2half2_rn(v[2] - f1.x, v[3] - f1.y);
              *(uint2*)(E.yhi + off) = make_uint2(*(uint32_t*)&h0, *(uint32_t*)&h1);
              *(uint2*)(E.ylo + off) = make_uint2(*(uint32_t*)&l0, *(uint32_t*)&l1);
            }
          }
        }
      }
    }
  }
  gt_fence_before();
  __syncthreads();
  if (warp == 1) {
    gt_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// fp32 -> fp16 hi/lo split (x = hi + lo), used where a producer kernel only writes fp32
__global__ void split_f16_kernel(const float* __restrict__ x, __half* __restrict__ hi,
                                 __half* __restrict__ lo, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg((const float4*)x + i);
    __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
    ((uint2*)hi)[i] = make_uint2(*(uint32_t*)&h0, *(uint32_t*)&h1);
    ((uint2*)lo)[i] = make_uint2(*(uint32_t*)&l0, *(uint32_t*)&l1);
  }
}

int launch_split_f16(const float* x, __half* hi, __half* lo, size_t n, cudaStream_t s) {
  HTSP_REQUIRE(n % 4 == 0, "split size");
  const size_t n4 = n / 4;
  int blocks = (int)((n4 + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  split_f16_kernel<<<blocks, 256, 0, s>>>(x, hi, lo, n4);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

// ---- host: tensor maps -------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// row-major fp16 matrix [rows, K] -> boxes of 128 rows x 64 halves (128 B), SWIZZLE_128B
static int make_map(CUtensorMap* m, const __half* ptr, int rows, int K) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return HTSP_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)GT_BK, (cuuint32_t)GT_BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%d K=%d", (int)r, rows, K);
    return HTSP_ERR_CUDA;
  }
  return HTSP_OK;
}

// row-major fp16 output [rows, cols] -> boxes of 32 rows x 32 halves (64 B), SWIZZLE_64B (epilogue TMA stores)
static int make_out_map(CUtensorMap* m, const __half* ptr, int rows, int cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return HTSP_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (output) failed (%d) rows=%d cols=%d", (int)r, rows, cols);
    return HTSP_ERR_CUDA;
  }
  return HTSP_OK;
}

// row-major fp32 output [rows, cols] -> boxes of 32 rows x 16 floats (64 B), SWIZZLE_64B
static int make_out_map_f32(CUtensorMap* m, const float* ptr, int rows, int cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return HTSP_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {16, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32 output) failed (%d) rows=%d cols=%d", (int)r, rows, cols);
    return HTSP_ERR_CUDA;
  }
  return HTSP_OK;
}

int launch_linear_tc(const __half* xhi, const __half* xlo, const __half* whi, const __half* wlo,
                     const float* bias, const float* residual, const float* bn_alpha,
                     const float* bn_beta, bool relu, float* y32, __half* yhi, __half* ylo, int M,
                     int Nout, int K, cudaStream_t s) {
  HTSP_REQUIRE(M > 0 && Nout % GT_BN == 0 && K % GT_BK == 0, "linear_tc shape");
  HTSP_REQUIRE((yhi == nullptr) == (ylo == nullptr), "hi/lo outputs come in pairs");
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  if ((rc = make_map(&ma_hi, xhi, M, K))) return rc;
  if ((rc = make_map(&ma_lo, xlo, M, K))) return rc;
  if ((rc = make_map(&mb_hi, whi, Nout, K))) return rc;
  if ((rc = make_map(&mb_lo, wlo, Nout, K))) return rc;
  // the packed (fp16 hi/lo only) epilogue stores through TMA; other launches never touch these maps
  CUtensorMap my_hi = ma_hi, my_lo = ma_lo;
  if (yhi != nullptr && y32 == nullptr && residual == nullptr && bn_alpha == nullptr) {
    if ((rc = make_out_map(&my_hi, yhi, M, Nout))) return rc;
    if ((rc = make_out_map(&my_lo, ylo, M, Nout))) return rc;
  } else if (y32 != nullptr && yhi == nullptr && residual == nullptr && bn_alpha == nullptr && bias == nullptr && !relu) {
    if ((rc = make_out_map_f32(&my_hi, y32, M, Nout))) return rc;     // epilogue-free fp32 output: TMA stores too
  }
  GemmEpi e;
  e.bias = bias; e.residual = residual; e.bn_alpha = bn_alpha; e.bn_beta = bn_beta;
  e.y32 = y32; e.yhi = yhi; e.ylo = ylo; e.M = M; e.Nout = Nout; e.K = K; e.relu = relu ? 1 : 0;
  HTSP_CHECK_CUDA(cudaFuncSetAttribute(gemm_x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)GT_SMEM_BYTES));
  const int tiles = ((M + GT_BM - 1) / GT_BM) * (Nout / GT_BN);
  const int grid = tiles < 148 ? tiles : 148;
  gemm_x3_tc_kernel<<<grid, GT_THREADS, GT_SMEM_BYTES, s>>>(ma_hi, ma_lo, mb_hi, mb_lo, my_hi, my_lo, e);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

// row-major fp16 matrix [rows, K] -> boxes of box_rows rows x 64 halves (128 B), SWIZZLE_128B
static int make_map_rows(CUtensorMap* m, const __half* ptr, int rows, int K, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return HTSP_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)GT_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%d K=%d box_rows=%d", (int)r, rows, K, box_rows);
    return HTSP_ERR_CUDA;
  }
  return HTSP_OK;
}

// Y = BN( X + relu(X W1^T + b1) W2^T + b2 ): X [M,128] as fp16 hi/lo + fp32 (the residual), W1 [512,128], W2 [128,512]
int launch_ff_fused_tc(const __half* xhi, const __half* xlo, const float* x32, const __half* w1hi, const __half* w1lo,
                       const float* b1, const __half* w2hi, const __half* w2lo, const float* b2, const float* bn_alpha,
                       const float* bn_beta, float* y32, __half* yhi, __half* ylo, int M, cudaStream_t s) {
  HTSP_REQUIRE(M > 0 && xhi && xlo && x32 && y32 && yhi && ylo, "ff_fused shape");
  CUtensorMap mx_hi, mx_lo, m1_hi, m1_lo, m2_hi, m2_lo;
  int rc;
  if ((rc = make_map(&mx_hi, xhi, M, H))) return rc;
  if ((rc = make_map(&mx_lo, xlo, M, H))) return rc;
  if ((rc = make_map_rows(&m1_hi, w1hi, FF, H, FF_CH))) return rc;
  if ((rc = make_map_rows(&m1_lo, w1lo, FF, H, FF_CH))) return rc;
  if ((rc = make_map(&m2_hi, w2hi, H, FF))) return rc;
  if ((rc = make_map(&m2_lo, w2lo, H, FF))) return rc;
  FfArgs e;
  e.b1 = b1; e.b2 = b2; e.residual = x32; e.bn_alpha = bn_alpha; e.bn_beta = bn_beta;
  e.y32 = y32; e.yhi = yhi; e.ylo = ylo; e.M = M;
  HTSP_CHECK_CUDA(cudaFuncSetAttribute(ff_fused_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FF_SMEM_BYTES));
  const int tiles = (M + GT_BM - 1) / GT_BM;
  ff_fused_tc_kernel<<<tiles < 148 ? tiles : 148, FF_THREADS, FF_SMEM_BYTES, s>>>(mx_hi, mx_lo, m1_hi, m1_lo, m2_hi, m2_lo, e);
  HTSP_CHECK_LAUNCH();
  return HTSP_OK;
}

}  // namespace htsp
