// FeedForward of the BasicTransformerBlock in ONE kernel (reference transformer.py:80-120, 14-77, 298-303):
//     X += W2 . snake(W1 . Xn + b1) + b2         snake(h) = h + sin^2(h e^alpha) / (e^beta + 1e-9)
// The 4C-wide hidden activation never leaves the SM: the unfused schedule wrote it to HBM as bf16 [rows, 4C] and read it
// back (38 % of the decode's DRAM traffic) and ran FF1 epilogue-bound.
//
// CTA pair (tcgen05 cta_group::2), 256 rows per pair (128 per CTA), persistent over row tiles.  Per CTA:
//   smem  Xn tile [128, C] bf16, resident for the whole tile (TMA, 128B swizzle, one 64-column block per mbarrier)
//         W1 ring: [32, 64] blocks (this CTA's half of a 64-row hidden chunk),  W2 ring: 2 x [C/4, 64] blocks (its half of W2's
//         rows for the two accumulator halves),  per-warp staging of the tile epilogue
//   TMEM  Y accumulator [128, C] fp32 at columns [0, C);  two hidden buffers H[0], H[1] of 64 columns at 384 / 448
// Hidden chunks of 64 columns are software-pipelined on the tensor pipe:   MMA1_0, MMA1_1, MMA2_0, MMA1_2, MMA2_1, ...
//   MMA1_j : H[j&1] = Xn . W1_j^T                   (M 256, N 64, K = C; A and B from shared memory)
//   epi_j  : 8 warps per CTA read H[j&1] (tcgen05.ld), add b1, SnakeBeta, and write the bf16 result P_j back INTO the same TMEM
//            columns (tcgen05.st; each warp only overwrites columns it has already read)
//   MMA2_j : Y += P_j . W2_j^T                      (M 256, N = 2 x C/2, K 64; A operand read from TMEM, B from shared memory)
// so the epilogue math of chunk j runs under MMA1_{j+1} / MMA2_{j-1} and P never touches shared memory (no generic->async proxy
// fence, no smem bandwidth for the A operand of the second GEMM).  MMA1 and MMA2 are issued by two different warps: an N = 64
// MMA lasts 32 cycles, less than one warp needs to issue it next to the second GEMM (measured: issue-bound at ~100 cycles per MMA
// with a single issuer).  At the end of a tile the warps add Y + b2 into the fp32 residual stream with the GEMM epilogues of
// gemm.cuh (TMA reduce-add, or coalesced load / add / store plus the masked bf16 copy the next convolution / skip connection
// consumes).  The first two MMA1 of the next tile already run during that.
#pragma once
#include <cuda.h>

#include "gemm.cuh"
#include "ptx.cuh"

namespace cfm {

struct FfParams {
  int M, C;             // rows, model width (hidden = 4 C)
  const float* b1;      // [4C]
  const float* ea;      // [4C] exp(alpha)
  const float* ib;      // [4C] 1 / (exp(beta) + 1e-9)
  const float* b2;      // [C]
  float* X;             // [M, ldx] fp32 residual stream, updated in place
  long long ldx;
  bf16* copy;           // optional: masked bf16 copy of the updated rows (nullptr = none)
  long long ld_copy;
  const int* row_info;  // validity flags for the copy
  unsigned long long* prof;  // debug: CTA 0 writes per-role cycle counters (see the kernel); nullptr = off
};

// Per-column constants of the SnakeBeta epilogue, passed BY VALUE as a kernel parameter: they then live in the constant bank and
// are read with warp-uniform LDC, which does not go through the L1 / shared-memory pipe (that pipe is saturated by the operand
// reads of the N = 64 MMAs: staging the constants in shared memory or reading them with __ldg cost ~1000 cycles per chunk).
struct FfConsts {
  float b1[4 * 384];
  float ea[4 * 384];
  float ib[4 * 384];
};

struct FfCfg {
  static constexpr int BM = 128, BK = 64, HC = 64;  // rows per CTA, K block, hidden columns per chunk
  static constexpr int MAX_C = 384;
  static constexpr int A_KB_BYTES = BM * BK * 2;            // 16 KB per 64-column block of Xn
  static constexpr int W1_KB_BYTES = (HC / 2) * BK * 2;     // 4 KB: this CTA's 32 rows of the chunk, one K block
  static constexpr int W1_STAGE_BYTES = 2 * W1_KB_BYTES;    // a stage = two K blocks (one barrier round trip per 8 MMAs: the N = 64
                                                            // MMAs last 32-52 cycles, per-stage overhead of the issuing warp is ~250)
  static constexpr int W1_STAGES = 5;
  static constexpr int W2_STAGES = 2;
  static constexpr int N_EPI_WARPS = 8;
  static constexpr int EPI_LD = 36;                         // staging row (floats) of the transposing tile epilogue (gemm.cuh)
  static constexpr int EPI_WARP_BYTES = 4608;               // 32 x 36 floats; the TMA-store path uses 4096-byte units at 4096 w
  static constexpr int THREADS = 128 + 32 * N_EPI_WARPS;
  static constexpr int TMEM_COLS = 512;
  static constexpr int H_COL0 = 384;  // H[b] at H_COL0 + 64 b
  __host__ __device__ static constexpr int w2_stage_bytes(int C) { return 2 * (C / 4) * BK * 2; }
  __host__ __device__ static constexpr int smem_bytes(int C) {
    return (C / BK) * A_KB_BYTES + W1_STAGES * W1_STAGE_BYTES + W2_STAGES * w2_stage_bytes(C) + N_EPI_WARPS * EPI_WARP_BYTES +
           1024 /*barriers*/ + 1024 /*alignment*/;
  }
};

__global__ void __launch_bounds__(FfCfg::THREADS, 1)
ff_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmX, const FfParams p,
                const __grid_constant__ FfConsts cst)
#ifdef CFM_FF_KERNEL_TU
{
  using Cfg = FfCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int C = p.C, NKB = C / Cfg::BK, NCH = 4 * C / Cfg::HC, NH = C / 2;
  const int W2_STAGE = Cfg::w2_stage_bytes(C), W2_HALF = W2_STAGE / 2;
  uint8_t* sA = smem;
  uint8_t* sW1 = sA + NKB * Cfg::A_KB_BYTES;
  uint8_t* sW2 = sW1 + Cfg::W1_STAGES * Cfg::W1_STAGE_BYTES;
  uint8_t* sEpi = sW2 + Cfg::W2_STAGES * W2_STAGE;  // per-warp staging of the tile epilogue
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + Cfg::N_EPI_WARPS * Cfg::EPI_WARP_BYTES);
  uint64_t* a_full = bars;                              // [6]  leader: both CTAs' Xn bytes
  uint64_t* a_empty = a_full + 6;                       // [6]  per CTA: last MMA1 of the tile has read block kb
  uint64_t* w1_full = a_empty + 6;                      // [W1_STAGES] leader
  uint64_t* w1_empty = w1_full + Cfg::W1_STAGES;        // [W1_STAGES] per CTA
  uint64_t* w2_full = w1_empty + Cfg::W1_STAGES;        // [W2_STAGES] leader
  uint64_t* w2_empty = w2_full + Cfg::W2_STAGES;        // [W2_STAGES] per CTA
  uint64_t* h_full = w2_empty + Cfg::W2_STAGES;         // [2] per CTA: MMA1_j complete
  uint64_t* h_free = h_full + 2;                        // [2] leader: MMA2_j has read P_j, the columns may be overwritten by MMA1_{j+2}
  uint64_t* p_full = h_free + 2;                        // [2] leader, 2 x N_EPI_WARPS arrivals: P_j written in both CTAs
  uint64_t* y_full = p_full + 2;                        // per CTA: last MMA2 of the tile complete
  uint64_t* y_free = y_full + 1;                        // leader, 2 x N_EPI_WARPS arrivals: Y drained in both CTAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();  // 0 = leader
  const int n_tiles = (p.M + 2 * Cfg::BM - 1) / (2 * Cfg::BM);
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  // Every pair walks the hidden chunks in its own rotation: the persistent pairs run in lock step, and 74 pairs asking L2 for the
  // same 48 KB of W1 / W2 at the same moment serialise on the few slices that own those lines (measured: the W1 ring, 1.8 chunks
  // deep, still starved the MMAs for ~1.3 k cycles per chunk).  The order of the chunk sum does not matter mathematically.
  const int rot = (pair_id * 7) % NCH;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW1);
    ptx::prefetch_tmap(&tmW2);
    ptx::prefetch_tmap(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 6; ++i) ptx::mbar_init(&a_full[i], 1), ptx::mbar_init(&a_empty[i], 1);
    for (int i = 0; i < Cfg::W1_STAGES; ++i) ptx::mbar_init(&w1_full[i], 1), ptx::mbar_init(&w1_empty[i], 1);
    for (int i = 0; i < Cfg::W2_STAGES; ++i) ptx::mbar_init(&w2_full[i], 1), ptx::mbar_init(&w2_empty[i], 1);
    for (int i = 0; i < 2; ++i)
      ptx::mbar_init(&h_full[i], 1), ptx::mbar_init(&h_free[i], 1), ptx::mbar_init(&p_full[i], 2 * Cfg::N_EPI_WARPS);
    ptx::mbar_init(y_full, 1);
    ptx::mbar_init(y_free, 2 * Cfg::N_EPI_WARPS);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();  // both CTAs' barriers and TMEM are set up before any cross-CTA traffic
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();
  const bool prof0 = p.prof != nullptr && blockIdx.x == 0;

  if (warp == 0) {
    // ===================== TMA producer 1 (both CTAs): the W1 ring, 6 small loads per chunk =====================
    int s1 = 0;
    uint32_t ph1 = 0;
    unsigned long long w_1 = 0;
    const long long t_start = clock64();
    for (int tile = pair_id; tile < n_tiles; tile += n_pairs) {
      for (int j = 0; j < NCH; ++j) {
        const int row = ((j + rot) % NCH) * Cfg::HC + rank * (Cfg::HC / 2);
        for (int kb = 0; kb < NKB; kb += 2) {
          const int nb = min(2, NKB - kb);
          mbar_wait_prof(&w1_empty[s1], ph1 ^ 1, prof0, w_1);
          if (rank == 0) ptx::mbar_expect_tx_elect(&w1_full[s1], 2 * nb * Cfg::W1_KB_BYTES);
          for (int i = 0; i < nb; ++i)
            ptx::tma_load_2d_pair_elect(sW1 + s1 * Cfg::W1_STAGE_BYTES + i * Cfg::W1_KB_BYTES, &tmW1, &w1_full[s1], (kb + i) * Cfg::BK, row);
          if (++s1 == Cfg::W1_STAGES) s1 = 0, ph1 ^= 1;
        }
      }
    }
    if (prof0 && lane == 0) p.prof[0] = (unsigned long long)(clock64() - t_start), p.prof[2] = w_1;
  } else if (warp == 2) {
    // ===================== TMA producer 2 (both CTAs): the Xn tile and the W2 ring =====================
    // Separate from producer 1 so that a wait for a free W2 slot or Xn block never delays the W1 stream the short MMAs live on.
    int s2 = 0, it = 0;
    uint32_t ph2 = 0;
    unsigned long long w_a = 0, w_2 = 0;
    for (int tile = pair_id; tile < n_tiles; tile += n_pairs, ++it) {
      const int m0 = (tile * 2 + rank) * Cfg::BM;
      for (int kb = 0; kb < NKB; ++kb) {
        mbar_wait_prof(&a_empty[kb], (it & 1) ^ 1, prof0, w_a);
        if (rank == 0) ptx::mbar_expect_tx_elect(&a_full[kb], 2 * Cfg::A_KB_BYTES);
        ptx::tma_load_2d_pair_elect(sA + kb * Cfg::A_KB_BYTES, &tmA, &a_full[kb], kb * Cfg::BK, m0);
      }
      for (int j = 0; j < NCH; ++j) {
        mbar_wait_prof(&w2_empty[s2], ph2 ^ 1, prof0, w_2);
        if (rank == 0) ptx::mbar_expect_tx_elect(&w2_full[s2], 2 * W2_STAGE);
        for (int hf = 0; hf < 2; ++hf)  // this CTA's half of the W2 rows of accumulator half hf
          ptx::tma_load_2d_pair_elect(sW2 + s2 * W2_STAGE + hf * W2_HALF, &tmW2, &w2_full[s2], ((j + rot) % NCH) * Cfg::HC,
                                      hf * NH + rank * (NH / 2));
        if (++s2 == Cfg::W2_STAGES) s2 = 0, ph2 ^= 1;
      }
    }
    if (prof0 && lane == 0) p.prof[1] = w_a, p.prof[3] = w_2;
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA; whole warp in uniform control flow, elect.sync inside the asm) =========
    // One thread issues both GEMMs in the order the tensor pipe should run them: MMA1_{j+1} (24 short MMAs), then MMA2_j (8 long
    // ones).  Two issuing warps were tried: their instructions interleave arbitrarily on the pipe, and alternating between the two
    // MMA shapes (SS N = 64 / TS N = 192) made every MMA1 take ~120 cycles instead of ~52.
    if (rank == 0) {
      const uint32_t idesc1 = ptx::umma_idesc_bf16(2 * Cfg::BM, Cfg::HC);
      const uint32_t idesc2 = ptx::umma_idesc_bf16(2 * Cfg::BM, NH);
      int s1 = 0, s2 = 0, it = 0;
      uint32_t ph1 = 0, ph2 = 0, g0 = 0;
      unsigned long long w_a = 0, w_1 = 0, w_p = 0, w_2 = 0, w_y = 0;
      const long long t_start = clock64();
      const uint64_t a_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(sA)), b_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(sW1));
      const uint64_t w_desc0 = ptx::umma_desc_sw128(ptx::smem_u32(sW2));
      for (int tile = pair_id; tile < n_tiles; tile += n_pairs, ++it, g0 += NCH) {
        for (int j = 0; j <= NCH; ++j) {
          if (j < NCH) {  // ---- MMA1_j: H[b] = Xn . W1_j^T  (H[b] was read by MMA2_{j-2}, issued earlier by this thread: in order)
            const uint32_t g = g0 + j, b = g & 1;
            const uint32_t tmem_h = tmem_base + Cfg::H_COL0 + b * Cfg::HC;
            if (prof0 && lane == 0 && g >= 8 && g < 16) p.prof[32 + (g - 8) * 8 + 0] = (unsigned long long)clock64();
            for (int kb = 0; kb < NKB; kb += 2) {
              const int nb = min(2, NKB - kb);
              if (j == 0) {
                mbar_wait_prof(&a_full[kb], it & 1, prof0, w_a);
                if (nb == 2) mbar_wait_prof(&a_full[kb + 1], it & 1, prof0, w_a);
              }
              mbar_wait_prof(&w1_full[s1], ph1, prof0, w_1);
              ptx::tc_fence_after();
              for (int i = 0; i < nb; ++i)
                ptx::umma_bf16_pair_k64_elect(tmem_h, a_desc0 + (uint64_t)((kb + i) * (Cfg::A_KB_BYTES >> 4)),
                                              b_desc0 + (uint64_t)(s1 * (Cfg::W1_STAGE_BYTES >> 4) + i * (Cfg::W1_KB_BYTES >> 4)), idesc1,
                                              (kb + i) > 0 ? 1u : 0u);
              ptx::umma_commit_pair_elect(&w1_empty[s1], 3);
              if (j == NCH - 1) {  // the tile's last reader of these Xn blocks
                ptx::umma_commit_pair_elect(&a_empty[kb], 3);
                if (nb == 2) ptx::umma_commit_pair_elect(&a_empty[kb + 1], 3);
              }
              if (++s1 == Cfg::W1_STAGES) s1 = 0, ph1 ^= 1;
            }
            ptx::umma_commit_pair_elect(&h_full[b], 3);
            if (prof0 && lane == 0 && g >= 8 && g < 16) p.prof[32 + (g - 8) * 8 + 1] = (unsigned long long)clock64();
          }
          if (j >= 1) {  // ---- MMA2_{j-1}: Y += P . W2^T   (A = P in tensor memory)
            const uint32_t g = g0 + j - 1, b = g & 1;
            if (j == 1) mbar_wait_prof(y_free, (it & 1) ^ 1, prof0, w_y);  // previous tile's Y drained by both CTAs
            mbar_wait_prof(&p_full[b], (g >> 1) & 1, prof0, w_p);
            if (prof0 && lane == 0 && g >= 8 && g < 16) p.prof[32 + (g - 8) * 8 + 6] = (unsigned long long)clock64();
            mbar_wait_prof(&w2_full[s2], ph2, prof0, w_2);
            ptx::tc_fence_after();
            if (prof0 && lane == 0 && g >= 8 && g < 16) p.prof[32 + (g - 8) * 8 + 4] = (unsigned long long)clock64();
            // hidden columns [16 k, 16 k + 16) of the chunk were packed by epilogue column half k / 2 at its own column offset
            const uint32_t tp = tmem_base + Cfg::H_COL0 + b * Cfg::HC;
            const uint64_t d0 = w_desc0 + (uint64_t)(s2 * (W2_STAGE >> 4));
            ptx::umma_bf16_pair_ts_k64x2_elect(tmem_base, tmem_base + NH, tp, tp + 8, tp + 32, tp + 40, d0, d0 + (uint64_t)(W2_HALF >> 4), idesc2,
                                               j > 1 ? 1u : 0u);
            ptx::umma_commit_pair_elect(&w2_empty[s2], 3);
            if (prof0 && lane == 0 && g >= 8 && g < 16) p.prof[32 + (g - 8) * 8 + 5] = (unsigned long long)clock64();
            if (++s2 == Cfg::W2_STAGES) s2 = 0, ph2 ^= 1;
          }
        }
        ptx::umma_commit_pair_elect(y_full, 3);
      }
      if (prof0 && lane == 0)
        p.prof[4] = (unsigned long long)(clock64() - t_start), p.prof[5] = w_a, p.prof[6] = w_1, p.prof[7] = w_p, p.prof[8] = w_2, p.prof[9] = w_y,
        p.prof[15] = (unsigned long long)it;
    }
    ptx::pdl_launch_dependents();
  } else if (warp >= 4) {
    // ===================== epilogue warps (both CTAs): SnakeBeta per chunk, residual update per tile =====================
    const int ew = warp - 4, q = ew & 3, hh = ew >> 2;  // TMEM lane quarter, column half
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    // staging of the tile epilogue: 1024-aligned 4 KB units for the TMA reduce-add, 32 x 36 floats for the transposing path
    const uint32_t stg = ptx::smem_u32(sEpi) + ew * (p.copy == nullptr ? 4096 : Cfg::EPI_WARP_BYTES);
    int it = 0;
    uint32_t g = 0;
    const bool prof = prof0 && warp == 4;
    unsigned long long w_h = 0, w_y = 0, t_tail = 0, t_ld = 0, t_math = 0, t_st = 0;
    const long long t_start = clock64();
    GemmParams gp = {};  // the tile epilogue reuses the GEMM epilogues of gemm.cuh (EPI_RESID); every other field zero
    gp.M = p.M, gp.N = C, gp.mode = EPI_RESID, gp.bias = p.b2, gp.row_mul = 1, gp.row_add = 0, gp.row_info = p.row_info;
    gp.resid = p.X, gp.ld_resid = p.ldx, gp.out_f32 = p.X, gp.ld_f32 = p.ldx, gp.out_act = p.copy, gp.ld_act = p.ld_copy;
    gp.fused_stats = 0, gp.stats = nullptr;
    auto release_y = [&] {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) ptx::mbar_arrive(y_free);
        else ptx::mbar_arrive_remote_relaxed(y_free, 0);
      }
    };
    for (int tile = pair_id; tile < n_tiles; tile += n_pairs, ++it) {
      for (int j = 0; j < NCH; ++j, ++g) {
        const uint32_t b = g & 1;
        const uint32_t th = tmem_base + Cfg::H_COL0 + b * Cfg::HC + hh * 32 + lane_off;
        const int n0 = ((j + rot) % NCH) * Cfg::HC + hh * 32;
        mbar_wait_prof(&h_full[b], (g >> 1) & 1, prof, w_h);
        ptx::tc_fence_after();
        uint32_t r0[16], r1[16];
        const long long c0 = prof ? clock64() : 0;
        if (prof && lane == 0 && g >= 8 && g < 16) p.prof[32 + (g - 8) * 8 + 2] = (unsigned long long)c0;
        ptx::tmem_ld16(th, r0);
        ptx::tmem_ld16(th + 16, r1);
        ptx::tmem_ld_wait();
        const long long c1 = prof ? clock64() : 0;
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {  // per-column constants: warp-uniform, L1-resident
          const float4 b1 = *reinterpret_cast<const float4*>(&cst.b1[n0 + i]);
          const float4 ea = *reinterpret_cast<const float4*>(&cst.ea[n0 + i]);
          const float4 ib = *reinterpret_cast<const float4*>(&cst.ib[n0 + i]);
          float h0 = __uint_as_float(i < 16 ? r0[i] : r1[i - 16]) + b1.x;
          float h1 = __uint_as_float(i < 16 ? r0[i + 1] : r1[i - 15]) + b1.y;
          float h2 = __uint_as_float(i < 16 ? r0[i + 2] : r1[i - 14]) + b1.z;
          float h3 = __uint_as_float(i < 16 ? r0[i + 3] : r1[i - 13]) + b1.w;
          const float s0 = __sinf(h0 * ea.x), s1 = __sinf(h1 * ea.y), s2 = __sinf(h2 * ea.z), s3 = __sinf(h3 * ea.w);
          h0 = fmaf(s0 * s0, ib.x, h0), h1 = fmaf(s1 * s1, ib.y, h1), h2 = fmaf(s2 * s2, ib.z, h2), h3 = fmaf(s3 * s3, ib.w, h3);
          __nv_bfloat162 p01 = __floats2bfloat162_rn(h0, h1), p23 = __floats2bfloat162_rn(h2, h3);
          pk[i / 2] = *reinterpret_cast<uint32_t*>(&p01);
          pk[i / 2 + 1] = *reinterpret_cast<uint32_t*>(&p23);
        }
        const long long c2 = prof ? clock64() : 0;
        ptx::tmem_st16(th, pk);  // bf16 pairs over the first 16 of the 32 columns this warp has just read
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) ptx::mbar_arrive(&p_full[b]);
          else ptx::mbar_arrive_remote_relaxed(&p_full[b], 0);
        }
        if (prof) {
          const long long c3 = clock64();
          if (lane == 0 && g >= 8 && g < 16) p.prof[32 + (g - 8) * 8 + 3] = (unsigned long long)c3;
          t_ld += (unsigned long long)(c1 - c0), t_math += (unsigned long long)(c2 - c1), t_st += (unsigned long long)(c3 - c2);
        }
      }
      // ---- tile epilogue: X += Y + b2 through the GEMM epilogues (coalesced): TMA reduce-add, or load / add / store + masked copy
      mbar_wait_prof(y_full, it & 1, prof, w_y);
      ptx::tc_fence_after();
      const long long t_tail0 = clock64();
      const int m0 = (tile * 2 + rank) * Cfg::BM + q * 32;
      const uint32_t taddr = tmem_base + lane_off;
      auto tail = [&](auto bn_tag) {
        constexpr int BN = decltype(bn_tag)::value;
        if (p.copy == nullptr) {
          epilogue_tile_tma<BN, EPI_RESID>(gp, &tmX, taddr, stg, m0, 0, hh, lane, release_y);
        } else {
          epilogue_tile<BN, EPI_RESID, Cfg::EPI_LD, Cfg::N_EPI_WARPS>(gp, taddr, stg, m0, 0, ew, lane, 0, 0u, 0u);
          release_y();
        }
      };
      switch (C) {
        case 64: tail(std::integral_constant<int, 64>{}); break;
        case 128: tail(std::integral_constant<int, 128>{}); break;
        case 192: tail(std::integral_constant<int, 192>{}); break;
        case 256: tail(std::integral_constant<int, 256>{}); break;
        case 320: tail(std::integral_constant<int, 320>{}); break;
        default: tail(std::integral_constant<int, 384>{}); break;
      }
      t_tail += (unsigned long long)(clock64() - t_tail0);
    }
    if (p.copy == nullptr) ptx::bulk_wait_all_elect();  // staging read / global updates landed before the CTA exits
    if (prof && lane == 0) p.prof[10] = (unsigned long long)(clock64() - t_start), p.prof[11] = w_h, p.prof[12] = w_y, p.prof[13] = t_tail,
                              p.prof[17] = t_ld, p.prof[18] = t_math, p.prof[19] = t_st;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();  // neither CTA leaves while its partner may still touch its smem / TMEM / barriers
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}
#else
;
#endif

KernelInfo kinfo_ff_fused();  // defined in ff_inst.cu; smem = FfCfg::smem_bytes(MAX_C)

}  // namespace cfm
