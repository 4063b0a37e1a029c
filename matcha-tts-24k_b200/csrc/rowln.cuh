// Linear + residual add + LayerNorm in ONE kernel (reference transformer.py:262-283 / 298-303 followed by the next norm,
// :179 / :215):      X += A . W^T + b          Xn = LayerNorm(X) * gamma + beta   (bf16, the next GEMM's operand)
// Used for the attention output projection (-> norm3) and for the second feed-forward GEMM when another transformer block
// follows (-> its norm1).  The unfused schedule ran the GEMM with N split over 2-3 CTAs (so no CTA ever saw a whole row) and
// a separate LayerNorm pass that re-read the fp32 stream: 180 launches and ~6 % of a decode.
//
// One CTA owns 128 rows x ALL N <= 384 columns (A-stationary: the activation tile is read once, not once per n-tile):
//   warp 0   TMA producer: stages of [128 x 64] A + [N x 64] W (two boxes of N/2 rows), 128B swizzle
//   warp 1   tcgen05.mma issuer: two M128 x (N/2) MMAs per 16-wide k step, fp32 accumulator [128, N] in tensor memory
//   warps 4.. epilogue (4 per TMEM lane quarter), one ROW per thread (tcgen05.ld shape), each warp N/4 columns of its 32 rows:
//     pass 1  x unit (32 x 32 fp32) arrives by TMA in a swizzled staging tile, y = acc + b + x goes back through the same tile
//             (TMA store), shifted sum / sum of squares, y written back over the accumulator (tcgen05.st)
//     --      the column shares of a row exchange (mean, M2) through shared memory (Chan's parallel-variance combine)
//     pass 2  y from tensor memory -> (y - mean) rstd gamma + beta -> bf16 units -> TMA store
// The accumulator is single-buffered (384 of the 512 columns), so MMAs and epilogue of a CTA alternate; the epilogue's staging
// is the tail of the operand ring (lent between a tile's last MMA and the end of its epilogue), the producer keeps the other
// stages filled with the next tile's operands.
//
// MEASURED (cfg2, B200): correct (same parity as the unfused schedule) but NOT faster - out-proj + norm3 37.0 us vs 23.0 + 14.3 us
// at full resolution, 22.2 vs 16.4 + 10.3 us at half resolution in the per-launch timeline, 27.9 vs 27.7 ms per decode in the
// graph; B = 1: 8.1 vs 6.6 ms.  The 144 small TMA operations per tile (32-row boxes with 128-byte rows: the widest a swizzled fp32
// box can be) cost ~150 cycles each on the SM's TMA unit and set the epilogue at ~16 k cycles per tile whatever the number of
// warps (8 -> 16 warps: 19 k -> 15.6 k); a row-per-thread epilogue with direct 256-bit global accesses is slower still (32 k:
// every warp instruction touches 32 lines).  The separate LayerNorm pass reads the rows coalesced at ~4.8 TB/s.  Kept as an
// option (cfm_set_option "rowln"), off by default.
#pragma once
#include <cuda.h>

#include "gemm.cuh"
#include "ptx.cuh"

namespace cfm {

struct RowLnParams {
  int M, N, K;          // rows, output width (= LayerNorm width, N % 64 == 0, N <= 384), reduction length (K % 64 == 0)
  const float* bias;    // [N]
  const float* gamma;   // [N]
  const float* beta;    // [N]
  float* X;             // [M, ldx] fp32 residual stream, updated in place
  long long ldx;
  bf16* Xn;             // [M, ldn] normalised rows
  long long ldn;
  float eps;
  unsigned long long* prof;
};

template <int WPQ>  // epilogue warps per TMEM lane quarter: each takes N / WPQ columns of its 32 rows
struct RowLnCfgT {
  static constexpr int BM = 128, BK = 64, MAX_N = 384;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int N_EPI_WARPS = 4 * WPQ;
  static constexpr int THREADS = 128 + 32 * N_EPI_WARPS;
  static constexpr int CTRL_BYTES = 1024 + WPQ * BM * 8;  // barriers + TMEM slot, (mean, M2) exchange [WPQ][128] float2
  static constexpr int MAX_SMEM = 227 * 1024;
  static constexpr int EPI_WARP_BYTES = 8192;                      // per-warp staging of the epilogue
  static constexpr int LEND_BYTES = N_EPI_WARPS * EPI_WARP_BYTES;  // tail of the operand ring lent to the epilogue between tiles
  __host__ __device__ static constexpr int stage_bytes(int N) { return A_BYTES + N * BK * 2; }
  __host__ __device__ static constexpr int stages(int N) {
    const int s = (MAX_SMEM - 1024 - CTRL_BYTES) / stage_bytes(N);
    return s > 8 ? 8 : s;
  }
  __host__ __device__ static constexpr int smem_bytes(int N) { return stages(N) * stage_bytes(N) + CTRL_BYTES + 1024; }
};

using RowLnCfg = RowLnCfgT<4>;

template <int WPQ>
__global__ void __launch_bounds__(RowLnCfgT<WPQ>::THREADS, 1)
gemm_rowln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmN, const RowLnParams p)
#ifdef CFM_ROWLN_KERNEL_TU
{
  using Cfg = RowLnCfgT<WPQ>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int N = p.N, NH = N >> 1, NW = N / WPQ;  // MMA width (two per k step), columns per epilogue warp
  const int STAGES = Cfg::stages(N), STAGE_BYTES = Cfg::stage_bytes(N);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tfull_bar = empty_bar + 8;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint64_t* xbars = tempty_bar + 1;  // [8 epilogue warps][2]: x units of pass 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xbars + 2 * Cfg::N_EPI_WARPS);
  float2* xch = reinterpret_cast<float2*>(smem + STAGES * STAGE_BYTES + 1024);  // [2][128] (mean, M2) of a row's column half

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + Cfg::BM - 1) / Cfg::BM;
  const int k_iters = p.K / Cfg::BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmN);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    ptx::mbar_init(tfull_bar, 1);
    ptx::mbar_init(tempty_bar, Cfg::N_EPI_WARPS);
    for (int i = 0; i < 2 * Cfg::N_EPI_WARPS; ++i) ptx::mbar_init(&xbars[i], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();
  const bool prof0 = p.prof != nullptr && blockIdx.x == 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    unsigned long long w_empty = 0;
    const long long t_start = clock64();
    // stages that overlap the last LEND_BYTES of the ring are the epilogue's staging while a tile is being normalised
    const int first_lent = (STAGES * STAGE_BYTES - Cfg::LEND_BYTES) / STAGE_BYTES;
    int local = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++local) {
      const int m0 = tile * Cfg::BM;
      for (int kb = 0; kb < k_iters; ++kb) {
        if (stage >= first_lent && local > 0) mbar_wait_prof(tempty_bar, (local - 1) & 1, prof0, w_empty);  // previous tile's epilogue done
        mbar_wait_prof(&empty_bar[stage], phase ^ 1, prof0, w_empty);
        ptx::mbar_expect_tx_elect(&full_bar[stage], STAGE_BYTES);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        ptx::tma_load_2d_elect(sa, &tmA, &full_bar[stage], kb * Cfg::BK, m0);
        ptx::tma_load_2d_elect(sa + Cfg::A_BYTES, &tmW, &full_bar[stage], kb * Cfg::BK, 0);
        ptx::tma_load_2d_elect(sa + Cfg::A_BYTES + NH * 128, &tmW, &full_bar[stage], kb * Cfg::BK, NH);
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
    }
    if (prof0 && lane == 0) p.prof[0] = (unsigned long long)(clock64() - t_start), p.prof[1] = w_empty;
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = ptx::umma_idesc_bf16(Cfg::BM, NH);
    int stage = 0, local = 0;
    uint32_t phase = 0;
    unsigned long long w_full = 0, w_tempty = 0;
    const long long t_start = clock64();
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++local) {
      mbar_wait_prof(tempty_bar, (local & 1) ^ 1, prof0, w_tempty);  // previous tile's rows normalised and stored
      ptx::tc_fence_after();
      for (int it = 0; it < k_iters; ++it) {
        mbar_wait_prof(&full_bar[stage], phase, prof0, w_full);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
        for (int k = 0; k < Cfg::BK / 16; ++k) {
          const uint64_t da = ptx::umma_desc_sw128(a_addr + k * 32);
          ptx::umma_bf16_elect(tmem_base, da, ptx::umma_desc_sw128(b_addr + k * 32), idesc, (it > 0 || k > 0) ? 1u : 0u);
          ptx::umma_bf16_elect(tmem_base + NH, da, ptx::umma_desc_sw128(b_addr + NH * 128 + k * 32), idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit_elect(&empty_bar[stage]);
        if (++stage == STAGES) stage = 0, phase ^= 1;
      }
      ptx::umma_commit_elect(tfull_bar);
    }
    if (prof0 && lane == 0) p.prof[2] = (unsigned long long)(clock64() - t_start), p.prof[3] = w_full, p.prof[4] = w_tempty;
    ptx::pdl_launch_dependents();
  } else if (warp >= 4) {
    // ===================== epilogue: one row per thread, two warps (column halves) per TMEM lane quarter =====================
    const int ew = warp - 4, q = ew & 3, hh = ew >> 2;
    const int r = q * 32 + lane;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + hh * NW;
    const int n_chunks = NW / 32;
    const int c0 = hh * NW;  // first column of this thread's share of the row
    // per-warp staging inside the lent tail of the operand ring: pass 1 = two 4 KB fp32 units (x in, y out, in place),
    // pass 2 = four 2 KB bf16 units
    uint8_t* stg_ptr = smem + STAGES * STAGE_BYTES - Cfg::LEND_BYTES + ew * Cfg::EPI_WARP_BYTES;
    const uint32_t stg = ptx::smem_u32(stg_ptr);
    uint64_t* xbar = xbars + ew * 2;
    uint32_t xuse0 = 0, xuse1 = 0;  // completed uses of the two x buffers (mbarrier phase)
    const bool prof = prof0 && warp == 4;
    unsigned long long w_tfull = 0, t_p1 = 0, t_x = 0, t_p2 = 0, w_x = 0;
    const long long t_start = clock64();
    int local = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += gridDim.x, ++local) {
      const int mq = tile * Cfg::BM + q * 32;  // first row of this warp's 32 x 32 units
      mbar_wait_prof(tfull_bar, local & 1, prof, w_tfull);  // all MMAs of the tile done: accumulator complete, lent stages consumed
      ptx::tc_fence_after();
      const long long t0 = prof ? clock64() : 0;
      // ---- pass 1: y = acc + b + x -> X, statistics, y back into tensor memory
      ptx::mbar_expect_tx_elect(&xbar[0], 4096);
      ptx::tma_load_2d_elect(stg_ptr, &tmX, &xbar[0], c0, mq);
      float shift = 0.f, s = 0.f, ss = 0.f;
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        const int b = c & 1;
        const uint32_t buf = stg + b * 4096;
        if (c + 1 < n_chunks) {  // the other buffer's last store (chunk c - 1) must have read it before the next x unit lands there
          ptx::bulk_wait_read_elect<0>();
          __syncwarp();
          ptx::mbar_expect_tx_elect(&xbar[b ^ 1], 4096);
          ptx::tma_load_2d_elect(stg_ptr + (b ^ 1) * 4096, &tmX, &xbar[b ^ 1], c0 + (c + 1) * 32, mq);
        }
        uint32_t acc[32];
        ptx::tmem_ld16(taddr + c * 32, *reinterpret_cast<uint32_t(*)[16]>(&acc[0]));
        ptx::tmem_ld16(taddr + c * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&acc[16]));
        ptx::tmem_ld_wait();
        float y[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(acc[i]);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + c * 32 + i));
          y[i] += bv.x, y[i + 1] += bv.y, y[i + 2] += bv.z, y[i + 3] += bv.w;
        }
        {
          const uint32_t use = b ? xuse1 : xuse0;
          mbar_wait_prof(&xbar[b], use & 1, prof, w_x);
          if (b) ++xuse1; else ++xuse0;
        }
        const uint32_t row = buf + lane * 128;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 xv = ptx::lds128(row + ((k ^ (lane & 7)) << 4));
          y[4 * k] += xv.x, y[4 * k + 1] += xv.y, y[4 * k + 2] += xv.z, y[4 * k + 3] += xv.w;
        }
        if (c == 0) shift = y[0];
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float d0 = y[i] - shift, d1 = y[i + 1] - shift;
          s0 += d0, s1 += d1;
          q0 = fmaf(d0, d0, q0), q1 = fmaf(d1, d1, q1);
        }
        s += s0 + s1, ss += q0 + q1;
#pragma unroll
        for (int k = 0; k < 8; ++k) ptx::sts128(row + ((k ^ (lane & 7)) << 4), y[4 * k], y[4 * k + 1], y[4 * k + 2], y[4 * k + 3]);
        ptx::fence_proxy_async();
        __syncwarp();
        ptx::tma_store_2d_elect(&tmX, buf, c0 + c * 32, mq);
        ptx::bulk_commit_elect();
        uint32_t yy[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) yy[i] = __float_as_uint(y[i]);
        ptx::tmem_st16(taddr + c * 32, *reinterpret_cast<const uint32_t(*)[16]>(&yy[0]));
        ptx::tmem_st16(taddr + c * 32 + 16, *reinterpret_cast<const uint32_t(*)[16]>(&yy[16]));
      }
      ptx::tmem_st_wait();
      const long long t1 = prof ? clock64() : 0;
      // ---- combine the WPQ column shares of every row: (n, mean, M2) partials, Chan et al.
      const float inv_nw = 1.f / (float)NW;
      const float mean_h = shift + s * inv_nw;
      const float m2_h = fmaxf(ss - s * s * inv_nw, 0.f);
      xch[hh * Cfg::BM + r] = make_float2(mean_h, m2_h);
      asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "n"(32 * WPQ) : "memory");
      float mean = 0.f;
      float2 part[WPQ];
#pragma unroll
      for (int w2 = 0; w2 < WPQ; ++w2) part[w2] = xch[w2 * Cfg::BM + r], mean += part[w2].x;
      mean *= 1.f / (float)WPQ;
      float m2 = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < WPQ; ++w2) {
        const float dm = part[w2].x - mean;
        m2 += part[w2].y + dm * dm * (float)NW;
      }
      const float var = m2 / (float)N;
      const float rstd = rsqrtf(var + p.eps);
      const float nmr = -mean * rstd;
      ptx::bulk_wait_read_elect<0>();  // pass 2 reuses the staging of pass 1
      __syncwarp();
      const long long t2 = prof ? clock64() : 0;
      // ---- pass 2: normalise from tensor memory, bf16 units leave through TMA
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        const uint32_t buf = stg + (c & 3) * 2048;
        uint32_t acc[32];
        ptx::tmem_ld16(taddr + c * 32, *reinterpret_cast<uint32_t(*)[16]>(&acc[0]));
        ptx::tmem_ld16(taddr + c * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&acc[16]));
        ptx::tmem_ld_wait();
        float y[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) y[i] = __uint_as_float(acc[i]);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + c0 + c * 32 + i));
          const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + c0 + c * 32 + i));
          y[i] = fmaf(fmaf(y[i], rstd, nmr), g.x, be.x), y[i + 1] = fmaf(fmaf(y[i + 1], rstd, nmr), g.y, be.y);
          y[i + 2] = fmaf(fmaf(y[i + 2], rstd, nmr), g.z, be.z), y[i + 3] = fmaf(fmaf(y[i + 3], rstd, nmr), g.w, be.w);
        }
        if (c >= 4) {  // the unit staged four chunks ago has been read
          ptx::bulk_wait_read_elect<3>();
          __syncwarp();
        }
        const uint32_t row = buf + lane * 64;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint2 lo = pack4_bf16(y[8 * k], y[8 * k + 1], y[8 * k + 2], y[8 * k + 3]);
          const uint2 hi = pack4_bf16(y[8 * k + 4], y[8 * k + 5], y[8 * k + 6], y[8 * k + 7]);
          ptx::sts128_u32(row + ((k ^ ((lane >> 1) & 3)) << 4), lo.x, lo.y, hi.x, hi.y);
        }
        ptx::fence_proxy_async();
        __syncwarp();
        ptx::tma_store_2d_elect(&tmN, buf, c0 + c * 32, mq);
        ptx::bulk_commit_elect();
      }
      ptx::bulk_wait_read_elect<0>();  // the staging goes back to the operand ring
      __syncwarp();
      ptx::tc_fence_before();
      if (lane == 0) ptx::mbar_arrive(tempty_bar);
      if (prof) t_p1 += (unsigned long long)(t1 - t0), t_x += (unsigned long long)(t2 - t1), t_p2 += (unsigned long long)(clock64() - t2);
    }
    ptx::bulk_wait_all_elect();  // global writes have landed before the CTA exits
    if (prof && lane == 0)
      p.prof[5] = (unsigned long long)(clock64() - t_start), p.prof[6] = w_tfull, p.prof[7] = t_p1, p.prof[8] = t_x, p.prof[9] = t_p2,
      p.prof[10] = (unsigned long long)local, p.prof[11] = w_x;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}
#else
;
#endif

KernelInfo kinfo_rowln();  // defined in rowln_inst.cu; smem = RowLnCfg::smem_bytes(N) per launch, attribute set for MAX_N

}  // namespace cfm
