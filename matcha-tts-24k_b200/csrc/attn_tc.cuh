// Tensor-core flash attention for sm_100a (bf16, head_dim 64): softmax(Q K^T d^-1/2 + key_bias) V per (utterance, head).
//
// One CTA per (utterance, head, 128-query tile); 192 threads = TMA warp, MMA warp, 4 softmax warps (one query row per
// thread).  Keys are processed in 64-key tiles through a software pipeline:
//     S_j  = Q K_j^T    tcgen05.mma M128 N64 K64  (Q, K tiles K-major, 128B-swizzled by TMA)  -> TMEM S[j & 1]
//     P_j  = exp2(...)  tcgen05.ld S_j -> registers (once) -> softmax -> bf16 P_j in smem P[j & 1] (UMMA K-major layout)
//     O   += P_j V_j    tcgen05.mma M128 N64 K64  (V tile MN-major straight from TMA), accumulated IN TMEM
// S, P, K and V are double-buffered, so S_{j+1} is computed while the softmax of tile j runs and the softmax warps never
// wait on an MMA issued in the same iteration.  The running maximum is updated lazily (only when a tile's maximum exceeds
// the reference by more than 2^8, FlashAttention-4 style): P and the row sum stay relative to the same reference, so the
// result is unchanged, and O is touched by the softmax warps only on those rare rescales (tcgen05.ld / mul / tcgen05.st)
// and once at the end.  Two CTAs per SM (80 KB smem, 256 TMEM columns each).
// The per-key bias implements the reference's additive float mask in the packed formulation (attn.cuh header).
#pragma once
#include <cuda.h>
#include <cstring>
#include <string>

#include "attn.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace cfm {

// X3: fp32 in / out on the bf16 tensor pipe - every operand tile exists twice, hi = bf16(x) and lo = bf16(x - hi) (the caller
// hands a [hi | lo] copy of the QKV buffer, kernels.cuh split3_rows_kernel), S = Qh Kh^T + Ql Kh^T + Qh Kl^T and
// O += Ph Vh + Pl Vh + Ph Vl with fp32 accumulation (dropped terms ~2^-17 relative); softmax, row sums and output in fp32.
template <bool X3>
struct AttnTcCfgT {
  static constexpr int D = 64, QT = 128, KT = 64, NP = X3 ? 2 : 1;  // NP: tiles per operand (hi, lo)
  static constexpr int Q_BYTES = QT * D * 2, K_BYTES = KT * D * 2, V_BYTES = KT * D * 2, P_BYTES = QT * KT * 2;
  static constexpr int OFF_Q = 0, OFF_K = NP * Q_BYTES, OFF_V = OFF_K + 2 * NP * K_BYTES, OFF_P = OFF_V + 2 * NP * V_BYTES;
  static constexpr int OFF_BAR = OFF_P + 2 * NP * P_BYTES;  // 11 mbarriers + TMEM slot (128 B)
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
  static constexpr int N_SOFTMAX_WARPS = 4;
  static constexpr int THREADS = 64 + 32 * N_SOFTMAX_WARPS;
  static constexpr int TMEM_COLS = 256;  // S[2] at columns 0 / 64, PV at column 128
  static constexpr int MIN_CTAS = X3 ? 1 : 2;
};
using AttnTcCfg = AttnTcCfgT<false>;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Blackwell packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2) and the 3-input maximum (FMNMX3): halve the
// instruction count of the softmax inner loops, which are issue-bound.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

#ifdef CFM_ATTN_KERNEL_TU  // the kernel body is compiled only in attn_inst.cu; cfm.cu launches through kinfo_attn_tc()
template <bool X3>
__global__ void __launch_bounds__(AttnTcCfgT<X3>::THREADS, AttnTcCfgT<X3>::MIN_CTAS)
attn_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, int inner,
               const UttTable* __restrict__ utt, const int4* __restrict__ work, void* __restrict__ out_v, long long ldo,
               float scale_log2, unsigned long long* prof) {
  using Cfg = AttnTcCfgT<X3>;
  constexpr int NP = Cfg::NP;
  const int lo_col = 3 * inner;  // X3: first column of the lo parts in the [hi | lo] QKV copy
  const bool do_prof = prof != nullptr && blockIdx.x == 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  // Every barrier is per buffer (index j & 1): a waiter can then never be more than one phase behind its barrier.
  uint64_t *bar_q = bars, *bar_k = bars + 1, *bar_v = bars + 3, *bar_s = bars + 5, *bar_p = bars + 7, *bar_pv = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int4 w = work[blockIdx.x];  // plan-constant tables: safe to read before pdl_wait
  const UttTable u = utt[w.x];
  const int head = w.y, q0 = w.z;
  const int L = u.len, nk = u.len + 1;
  const int n_tiles = (nk + Cfg::KT - 1) / Cfg::KT;
  const int row0 = u.start;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_q);
    ptx::prefetch_tmap(&tm_kv);
    ptx::mbar_init(bar_q, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_k + i, 1);
      ptx::mbar_init(bar_v + i, 1);
      ptx::mbar_init(bar_s + i, 1);
      ptx::mbar_init(bar_p + i, 32 * Cfg::N_SOFTMAX_WARPS);
      ptx::mbar_init(bar_pv + i, 1);
    }
    ptx::fence_mbar_init();
    // Q and the first two K / V tiles are requested right away: their L2 latency overlaps the TMEM allocation and the
    // CTA-wide barrier below instead of following them (the barriers they complete on are initialised and fenced above).
    ptx::pdl_wait();
    ptx::mbar_expect_tx(bar_q, NP * Cfg::Q_BYTES);
    for (int pt = 0; pt < NP; ++pt) ptx::tma_load_2d(smem + Cfg::OFF_Q + pt * Cfg::Q_BYTES, &tm_q, bar_q, pt * lo_col + head * Cfg::D, row0 + q0);
    for (int j = 0; j < 2 && j < n_tiles; ++j) {
      ptx::mbar_expect_tx(bar_k + j, NP * Cfg::K_BYTES);
      ptx::mbar_expect_tx(bar_v + j, NP * Cfg::V_BYTES);
      for (int pt = 0; pt < NP; ++pt) {
        ptx::tma_load_2d(smem + Cfg::OFF_K + (j * NP + pt) * Cfg::K_BYTES, &tm_kv, bar_k + j, pt * lo_col + inner + head * Cfg::D, row0 + j * Cfg::KT);
        ptx::tma_load_2d(smem + Cfg::OFF_V + (j * NP + pt) * Cfg::V_BYTES, &tm_kv, bar_v + j, pt * lo_col + 2 * inner + head * Cfg::D, row0 + j * Cfg::KT);
      }
    }
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_pv = tmem_base + 128;
  ptx::pdl_wait();

  // Software pipeline over 64-key tiles j (buffer b = j & 1, use count u = j >> 1):
  //   tensor core : S_0, S_1, [P_0] PV_0, S_2, [P_1] PV_1, S_3, ...   S_{j+2} reuses S[b] as soon as softmax_j has read S_j
  //   softmax     : wait S_j -> max -> P_j = exp2(...) -> smem P[b] -> O += PV_{j-1} (finished long ago), O *= corr -> arrive P_j
  // so the softmax warps never wait on an MMA issued in the same iteration.
  if (warp == 0) {
    {  // ---------------- TMA producer, whole warp in uniform control flow (tiles 0 and 1 were requested in the prologue)
      for (int j = 2; j < n_tiles; ++j) {
        const int b = j & 1;
        const uint32_t prev = ((j >> 1) - 1) & 1;  // parity of the previous use of buffer b
        ptx::mbar_wait(bar_s + b, prev);  // S_{j-2} complete: K[b] free
        ptx::mbar_expect_tx_elect(bar_k + b, NP * Cfg::K_BYTES);
        for (int pt = 0; pt < NP; ++pt)
          ptx::tma_load_2d_elect(smem + Cfg::OFF_K + (b * NP + pt) * Cfg::K_BYTES, &tm_kv, bar_k + b, pt * lo_col + inner + head * Cfg::D, row0 + j * Cfg::KT);
        ptx::mbar_wait(bar_pv + b, prev);  // PV_{j-2} complete: V[b] free
        ptx::mbar_expect_tx_elect(bar_v + b, NP * Cfg::V_BYTES);
        for (int pt = 0; pt < NP; ++pt)
          ptx::tma_load_2d_elect(smem + Cfg::OFF_V + (b * NP + pt) * Cfg::V_BYTES, &tm_kv, bar_v + b, pt * lo_col + 2 * inner + head * Cfg::D, row0 + j * Cfg::KT);
      }
    }
  } else if (warp == 1) {
    {  // ---------------- MMA issuer: whole warp in uniform control flow, elect.sync picks the issuing lane (see ptx.cuh)
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(128, 64, 0);
      constexpr uint32_t idesc_pv = ptx::umma_idesc_bf16(128, 64, 1);  // B (= V tile) is MN-major
      const uint32_t q_addr = ptx::smem_u32(smem + Cfg::OFF_Q), k_addr = ptx::smem_u32(smem + Cfg::OFF_K);
      const uint32_t v_addr = ptx::smem_u32(smem + Cfg::OFF_V), p_addr = ptx::smem_u32(smem + Cfg::OFF_P);
      unsigned long long wq = 0, wk = 0, wp = 0, wv = 0;
      const long long t_start = clock64();
      const unsigned long long t_start_ns = prof != nullptr ? ptx::globaltimer_ns() : 0ull;
      auto issue_s = [&](int j) {
        const int b = j & 1;
        mbar_wait_prof(bar_k + b, (j >> 1) & 1, do_prof, wk);
        ptx::tc_fence_after();
#pragma unroll
        for (int pr = 0; pr < (X3 ? 3 : 1); ++pr) {  // (Q part, K part): hi hi, lo hi, hi lo
          const uint32_t qa = q_addr + (pr == 1 ? Cfg::Q_BYTES : 0), ka = k_addr + (b * NP + (pr == 2 ? 1 : 0)) * Cfg::K_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_elect(tmem_base + b * 64, ptx::umma_desc_sw128(qa + k * 32), ptx::umma_desc_sw128(ka + k * 32), idesc_s, (pr > 0 || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit_elect(bar_s + b);
      };
      mbar_wait_prof(bar_q, 0, do_prof, wq);
      issue_s(0);
      if (n_tiles > 1) issue_s(1);
      for (int j = 0; j < n_tiles; ++j) {
        const int b = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait_prof(bar_p + b, ph, do_prof, wp);  // P_j in smem; S_j and PV_{j-1} drained from TMEM
        mbar_wait_prof(bar_v + b, ph, do_prof, wv);
        ptx::tc_fence_after();
#pragma unroll
        for (int pr = 0; pr < (X3 ? 3 : 1); ++pr) {  // (P part, V part): hi hi, lo hi, hi lo
          const uint32_t pa = p_addr + (b * NP + (pr == 1 ? 1 : 0)) * Cfg::P_BYTES, va = v_addr + (b * NP + (pr == 2 ? 1 : 0)) * Cfg::V_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_elect(tmem_pv, ptx::umma_desc_sw128(pa + k * 32), ptx::umma_desc_sw128(va + k * 2048), idesc_pv, (j > 0 || pr > 0 || k > 0) ? 1u : 0u);
        }
        ptx::umma_commit_elect(bar_pv + b);
        if (j + 2 < n_tiles) issue_s(j + 2);
      }
      ptx::pdl_launch_dependents();  // last MMA issued: the next kernel's launch overlaps this CTA's final softmax pass and epilogue
      if (do_prof && lane == 0) {
        prof[0] = (unsigned long long)(clock64() - t_start), prof[1] = wq, prof[2] = wk, prof[3] = wp, prof[4] = wv;
        prof[5] = (unsigned long long)n_tiles;
      }
      if (prof != nullptr && lane == 0 && gridDim.x <= 4096) {  // every CTA: MMA-warp lifetime [cycles], start / end [ns], SM id
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long* q = prof + 32 + 4ull * blockIdx.x;
        q[0] = (unsigned long long)(clock64() - t_start), q[1] = t_start_ns, q[2] = ptx::globaltimer_ns(), q[3] = smid;
      }
    }
  } else {
    // ---------------- softmax / rare rescale / epilogue: one query row per thread
    const int quarter = warp & 3;  // TMEM lane quarter this warp may touch
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const float raw_pad_bias = u.pad_key_bias / (scale_log2 * 0.6931471805599453f);  // additive mask in raw-score units
    float mref = -INFINITY, lrun = 0.f;  // exponent reference (log2 domain) and row sum relative to it
    const bool sp = do_prof && warp == 2;
    unsigned long long ws = 0, wbar = 0, wpv = 0;
    const long long ts_start = clock64();
    uint32_t sv[64];  // raw scores of the current tile (this thread's row)
    auto load_s = [&](int j) {
      const uint32_t ts = tmem_base + (j & 1) * 64 + lane_off;
#pragma unroll
      for (int c = 0; c < 4; ++c) ptx::tmem_ld16(ts + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&sv[c * 16]));
    };
    mbar_wait_prof(bar_s, 0, sp, ws);
    ptx::tc_fence_after();
    load_s(0);
    for (int j = 0; j < n_tiles; ++j) {
      const int b = j & 1;
      const int k0 = j * Cfg::KT;
      const uint32_t p_row = ptx::smem_u32(smem + Cfg::OFF_P + b * NP * Cfg::P_BYTES + r * 128);
      ptx::tmem_ld_wait();
      if (k0 + Cfg::KT > L) {  // tile holds the pad token and / or rows past this utterance: fix the raw scores in place
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const int key = k0 + i;
          const float t = __uint_as_float(sv[i]);
          sv[i] = __float_as_uint(key < L ? t : (key == L ? t + raw_pad_bias : -INFINITY));
        }
      }
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; i += 4) {
        m0 = max3(m0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
        m1 = max3(m1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
      }
      const float mt = fmaxf(m0, m1) * scale_log2;
      // Lazy reference update: keep mref while the tile maximum stays within 2^8 of it (P <= 256, exact in the final O / l).
      const bool grow = mt > mref + 8.f;  // j == 0: mref = -inf and key 0 is always valid -> true
      const float mnew = grow ? mt : mref;
      if (j > 0 && __any_sync(0xffffffffu, grow)) {  // rare: rescale this warp's rows of O (in TMEM) and l
        const float corr = ex2_approx(mref - mnew);  // 1 for the rows that keep their reference
        mbar_wait_prof(bar_pv + ((j - 1) & 1), ((j - 1) >> 1) & 1, sp, wpv);  // every PV issued so far has landed in O
        ptx::tc_fence_after();
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
          uint32_t o[16];
          ptx::tmem_ld16(tmem_pv + lane_off + c, o);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
          ptx::tmem_st16(tmem_pv + lane_off + c, o);
        }
        ptx::tmem_st_wait();
        lrun *= corr;
      }
      mref = mnew;
      // P = exp2(s * scale - mref) -> bf16 -> smem (UMMA K-major 128B swizzle: 16-byte chunk c of row r at chunk c ^ (r & 7))
      {
        const f32x2 sc2 = pack2(scale_log2, scale_log2), nm2 = pack2(-mref, -mref);
        f32x2 acc01 = pack2(0.f, 0.f), acc23 = pack2(0.f, 0.f);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float pv[8];
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            const f32x2 t2 = fma2(pack2(__uint_as_float(sv[8 * g + i]), __uint_as_float(sv[8 * g + i + 1])), sc2, nm2);
            float t0, t1;
            unpack2(t2, t0, t1);
            pv[i] = ex2_approx(t0), pv[i + 1] = ex2_approx(t1);
            if ((i & 2) == 0) acc01 = add2(acc01, pack2(pv[i], pv[i + 1]));
            else acc23 = add2(acc23, pack2(pv[i], pv[i + 1]));
          }
          __nv_bfloat162 h0 = __floats2bfloat162_rn(pv[0], pv[1]), h1 = __floats2bfloat162_rn(pv[2], pv[3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(pv[4], pv[5]), h3 = __floats2bfloat162_rn(pv[6], pv[7]);
          ptx::sts128_u32(p_row + ((g ^ (r & 7)) << 4), *reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                          *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
          if constexpr (X3) {  // lo part of P: what bf16 rounding dropped
            const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1), f2 = __bfloat1622float2(h2), f3 = __bfloat1622float2(h3);
            __nv_bfloat162 l0 = __floats2bfloat162_rn(pv[0] - f0.x, pv[1] - f0.y), l1 = __floats2bfloat162_rn(pv[2] - f1.x, pv[3] - f1.y);
            __nv_bfloat162 l2 = __floats2bfloat162_rn(pv[4] - f2.x, pv[5] - f2.y), l3 = __floats2bfloat162_rn(pv[6] - f3.x, pv[7] - f3.y);
            ptx::sts128_u32(p_row + Cfg::P_BYTES + ((g ^ (r & 7)) << 4), *reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1),
                            *reinterpret_cast<uint32_t*>(&l2), *reinterpret_cast<uint32_t*>(&l3));
          }
        }
        float s0, s1;
        unpack2(add2(acc01, acc23), s0, s1);
        lrun += s0 + s1;
      }
      if (j + 1 < n_tiles) {  // S_{j+1} was issued before PV_{j-1}: normally complete by now; its load overlaps the hand-off
        mbar_wait_prof(bar_s + ((j + 1) & 1), ((j + 1) >> 1) & 1, sp, ws);
        ptx::tc_fence_after();
        load_s(j + 1);
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      ptx::mbar_arrive(bar_p + b);
    }
    mbar_wait_prof(bar_pv + ((n_tiles - 1) & 1), ((n_tiles - 1) >> 1) & 1, sp, wpv);  // O complete
    ptx::tc_fence_after();
    if (sp && lane == 0) prof[8] = (unsigned long long)(clock64() - ts_start), prof[9] = ws, prof[10] = wbar, prof[11] = wpv;
    const int qi = q0 + r;
    const float inv = 1.f / lrun;
    bf16* dst = static_cast<bf16*>(out_v) + (long long)(row0 + qi) * ldo + head * Cfg::D;
    float* dst_f = static_cast<float*>(out_v) + (long long)(row0 + qi) * ldo + head * Cfg::D;
#pragma unroll
    for (int c = 0; c < 64; c += 16) {
      uint32_t a[16];
      ptx::tmem_ld16(tmem_pv + lane_off + c, a);
      ptx::tmem_ld_wait();
      if constexpr (X3) {
        if (qi < nk) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<float4*>(dst_f + c + 4 * g) = make_float4(__uint_as_float(a[4 * g]) * inv, __uint_as_float(a[4 * g + 1]) * inv,
                                                                        __uint_as_float(a[4 * g + 2]) * inv, __uint_as_float(a[4 * g + 3]) * inv);
        }
        continue;
      }
      if (qi < nk) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(a[8 * g + 0]) * inv, __uint_as_float(a[8 * g + 1]) * inv);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(a[8 * g + 2]) * inv, __uint_as_float(a[8 * g + 3]) * inv);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(a[8 * g + 4]) * inv, __uint_as_float(a[8 * g + 5]) * inv);
          __nv_bfloat162 h3 = __floats2bfloat162_rn(__uint_as_float(a[8 * g + 6]) * inv, __uint_as_float(a[8 * g + 7]) * inv);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0), pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2), pk.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(dst + c + 8 * g) = pk;
        }
      }
    }
    ptx::tc_fence_before();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

KernelInfo kinfo_attn_tc(int x3) {
  if (x3) return KernelInfo{reinterpret_cast<const void*>(&attn_tc_kernel<true>), AttnTcCfgT<true>::THREADS, AttnTcCfgT<true>::SMEM_BYTES};
  return KernelInfo{reinterpret_cast<const void*>(&attn_tc_kernel<false>), AttnTcCfg::THREADS, AttnTcCfg::SMEM_BYTES};
}
#else

inline int attn_tc_set_attr(std::string* err) {
  for (int x3 = 0; x3 < 3; ++x3) {
    const KernelInfo k = x3 == 2 ? kinfo_attn_persist() : kinfo_attn_tc(x3);
    cudaError_t e = cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, k.smem);
    if (e != cudaSuccess) {
      *err = std::string("cudaFuncSetAttribute(attn_tc_kernel): ") + cudaGetErrorString(e);
      return -2;
    }
  }
  return 0;
}

template <typename Enc>
inline int launch_attn_tc(Enc encode, const void* qkv, long long ld, int inner, int M, const UttTable* utt, const int4* work,
                          int n_work, void* out, long long ldo, float scale, cudaStream_t s, std::string* err,
                          unsigned long long* prof = nullptr, bool pdl = false, bool x3 = false) {
  // x3: qkv is the bf16 [hi | lo] copy ([M, 2 ld], lo parts at column ld) of an fp32 QKV buffer, out is fp32
  CUtensorMap tm_q, tm_kv;
  cuuint64_t dims[2] = {(cuuint64_t)ld * (x3 ? 2 : 1), (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2 * (x3 ? 2 : 1)};
  cuuint32_t estr[2] = {1, 1};
  for (int i = 0; i < 2; ++i) {
    cuuint32_t box[2] = {64, (cuuint32_t)(i == 0 ? AttnTcCfg::QT : AttnTcCfg::KT)};
    CUresult r = encode(i == 0 ? &tm_q : &tm_kv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      *err = "cuTensorMapEncodeTiled failed for the attention QKV map";
      return -2;
    }
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  const KernelInfo ki = kinfo_attn_tc(x3 ? 1 : 0);
  cfg.gridDim = dim3(n_work), cfg.blockDim = dim3(ki.threads), cfg.dynamicSmemBytes = ki.smem, cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl ? 1 : 0;
  void* out_b = out;
  float scale_log2 = scale * 1.4426950408889634f;
  void* args[] = {&tm_q, &tm_kv, &inner, &utt, &work, &out_b, &ldo, &scale_log2, &prof};
  cudaError_t e = cudaLaunchKernelExC(&cfg, ki.fn, args);
  if (e != cudaSuccess) {
    *err = std::string("attn_tc_kernel launch: ") + cudaGetErrorString(e);
    return -2;
  }
  return 0;
}

#endif  // CFM_ATTN_KERNEL_TU

}  // namespace cfm
