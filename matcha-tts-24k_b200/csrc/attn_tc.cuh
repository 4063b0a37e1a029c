// Tensor-core flash attention for sm_100a (bf16, head_dim 64): softmax(Q K^T d^-1/2 + key_bias) V per (utterance, head).
//
// One CTA per (utterance, head, 128-query tile); 320 threads = TMA warp, MMA warp, 8 softmax warps (two threads per
// query row, each owning half of the key columns and half of the output columns).  Per 128-key tile:
//     S  = Q K^T      tcgen05.mma  M128 N128 K64   (Q, K tiles K-major, 128B-swizzled by TMA)    -> TMEM cols [0,128)
//     P  = exp2(...)  tcgen05.ld S -> registers -> online softmax -> bf16 P written to smem in the UMMA K-major layout
//     PV = P V        tcgen05.mma  M128 N64 K128   (V tile MN-major straight from TMA)            -> TMEM cols [128,192)
//     O  = O*corr + PV in registers (tcgen05.ld)
// Two CTAs are resident per SM (80 KB smem, 256 TMEM columns each) so one CTA's softmax overlaps the other's MMAs.
// The per-key bias implements the reference's additive float mask in the packed formulation (attn.cuh header).
#pragma once
#include <cuda.h>
#include <cstring>
#include <string>

#include "attn.cuh"
#include "gemm.cuh"
#include "ptx.cuh"

namespace cfm {

struct AttnTcCfg {
  static constexpr int D = 64, QT = 128, KT = 128;
  static constexpr int Q_BYTES = QT * D * 2, K_BYTES = KT * D * 2, V_BYTES = KT * D * 2, P_BYTES = QT * KT * 2;
  static constexpr int OFF_Q = 0, OFF_K = Q_BYTES, OFF_V = OFF_K + K_BYTES, OFF_P = OFF_V + V_BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;         // 6 mbarriers + TMEM slot (64 B)
  static constexpr int OFF_RED = OFF_BAR + 64;            // [2][128] floats: row max / row sum exchange between column halves
  static constexpr int SMEM_BYTES = OFF_RED + 2 * QT * 4 + 1024;
  static constexpr int N_SOFTMAX_WARPS = 8;
  static constexpr int THREADS = 64 + 32 * N_SOFTMAX_WARPS;
  static constexpr int TMEM_COLS = 256;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Row maximum of this thread's 64 raw scores (TMEM columns [taddr, taddr + 64)).  RAGGED: key index >= L is either the
// virtual pad token (key == L, raw bias added) or outside the utterance (-inf).
template <bool RAGGED>
__device__ __forceinline__ float attn_rowmax(uint32_t taddr, int k0, int L, float raw_pad_bias) {
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
  for (int c = 0; c < 64; c += 32) {
    uint32_t a[16], b[16];
    ptx::tmem_ld16(taddr + c, a);
    ptx::tmem_ld16(taddr + c + 16, b);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      float t[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        t[j] = __uint_as_float((i + j) < 16 ? a[i + j] : b[i + j - 16]);
        if (RAGGED) {
          const int key = k0 + c + i + j;
          t[j] = key < L ? t[j] : (key == L ? t[j] + raw_pad_bias : -INFINITY);
        }
      }
      m0 = fmaxf(m0, t[0]), m1 = fmaxf(m1, t[1]), m2 = fmaxf(m2, t[2]), m3 = fmaxf(m3, t[3]);
    }
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// P = exp2(s * scale_log2 - m) for this thread's 64 columns -> bf16 -> smem in the UMMA K-major 128B-swizzled layout
// (16-byte chunk c of row r lives at chunk c ^ (r & 7)); returns the row's partial sum.
template <bool RAGGED>
__device__ __forceinline__ float attn_exp_store(uint32_t taddr, uint8_t* p_row, int r, int k0, int L, float scale_log2,
                                                float mnew, float pad_bias_log2) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int c = 0; c < 64; c += 32) {
    uint32_t a[16], b[16];
    ptx::tmem_ld16(taddr + c, a);
    ptx::tmem_ld16(taddr + c + 16, b);
    ptx::tmem_ld_wait();
    float pv[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float t = fmaf(__uint_as_float(i < 16 ? a[i] : b[i - 16]), scale_log2, -mnew);
      if (RAGGED) {
        const int key = k0 + c + i;
        t = key < L ? t : (key == L ? t + pad_bias_log2 : -INFINITY);
      }
      pv[i] = ex2_approx(t);
    }
#pragma unroll
    for (int i = 0; i < 32; i += 4) s0 += pv[i], s1 += pv[i + 1], s2 += pv[i + 2], s3 += pv[i + 3];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int chunk = (c >> 3) + g;
      uint4 pk;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(pv[8 * g + 0], pv[8 * g + 1]);
      __nv_bfloat162 h1 = __floats2bfloat162_rn(pv[8 * g + 2], pv[8 * g + 3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(pv[8 * g + 4], pv[8 * g + 5]);
      __nv_bfloat162 h3 = __floats2bfloat162_rn(pv[8 * g + 6], pv[8 * g + 7]);
      pk.x = *reinterpret_cast<uint32_t*>(&h0);
      pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2);
      pk.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(p_row + ((chunk ^ (r & 7)) << 4)) = pk;
    }
  }
  return (s0 + s1) + (s2 + s3);
}

__global__ void __launch_bounds__(AttnTcCfg::THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, int inner, const UttTable* __restrict__ utt,
               const int4* __restrict__ work, bf16* __restrict__ out, long long ldo, float scale_log2,
               unsigned long long* prof) {
  using Cfg = AttnTcCfg;
  const bool do_prof = prof != nullptr && blockIdx.x == 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t *bar_q = bars + 0, *bar_k = bars + 1, *bar_v = bars + 2, *bar_s = bars + 3, *bar_p = bars + 4, *bar_pv = bars + 5;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
  float* red = reinterpret_cast<float*>(smem + Cfg::OFF_RED);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ptx::pdl_launch_dependents();
  const int4 w = work[blockIdx.x];  // plan-constant tables: safe to read before pdl_wait
  const UttTable u = utt[w.x];
  const int head = w.y, q0 = w.z;
  const int L = u.len, nk = u.len + 1;
  const int n_tiles = (nk + Cfg::KT - 1) / Cfg::KT;
  const int row0 = u.start;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_qkv);
    ptx::mbar_init(bar_q, 1);
    ptx::mbar_init(bar_k, 1);
    ptx::mbar_init(bar_v, 1);
    ptx::mbar_init(bar_s, 1);
    ptx::mbar_init(bar_p, 32 * Cfg::N_SOFTMAX_WARPS);
    ptx::mbar_init(bar_pv, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_pv = tmem_base + 128;
  ptx::pdl_wait();

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer
      ptx::mbar_expect_tx(bar_q, Cfg::Q_BYTES);
      ptx::tma_load_2d(smem + Cfg::OFF_Q, &tm_qkv, bar_q, head * Cfg::D, row0 + q0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j > 0) ptx::mbar_wait(bar_s, (j - 1) & 1);  // S_{j-1} complete: K buffer free
        ptx::mbar_expect_tx(bar_k, Cfg::K_BYTES);
        ptx::tma_load_2d(smem + Cfg::OFF_K, &tm_qkv, bar_k, inner + head * Cfg::D, row0 + j * Cfg::KT);
        if (j > 0) ptx::mbar_wait(bar_pv, (j - 1) & 1);  // PV_{j-1} complete: V buffer free
        ptx::mbar_expect_tx(bar_v, Cfg::V_BYTES);
        ptx::tma_load_2d(smem + Cfg::OFF_V, &tm_qkv, bar_v, 2 * inner + head * Cfg::D, row0 + j * Cfg::KT);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(128, 128, 0);
      constexpr uint32_t idesc_pv = ptx::umma_idesc_bf16(128, 64, 1);  // B (= V tile) is MN-major
      const uint32_t q_addr = ptx::smem_u32(smem + Cfg::OFF_Q), k_addr = ptx::smem_u32(smem + Cfg::OFF_K);
      const uint32_t v_addr = ptx::smem_u32(smem + Cfg::OFF_V), p_addr = ptx::smem_u32(smem + Cfg::OFF_P);
      unsigned long long wq = 0, wk = 0, wp = 0, wv = 0;
      const long long t_start = clock64();
      mbar_wait_prof(bar_q, 0, do_prof, wq);
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t ph = j & 1;
        mbar_wait_prof(bar_k, ph, do_prof, wk);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::umma_bf16(tmem_s, ptx::umma_desc_sw128(q_addr + k * 32), ptx::umma_desc_sw128(k_addr + k * 32), idesc_s, k > 0);
        ptx::umma_commit(bar_s);
        mbar_wait_prof(bar_p, ph, do_prof, wp);  // P_j in smem, S_j and PV_{j-1} drained from TMEM
        mbar_wait_prof(bar_v, ph, do_prof, wv);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          ptx::umma_bf16(tmem_pv, ptx::umma_desc_sw128(p_addr + (k >> 2) * (Cfg::QT * 128) + (k & 3) * 32),
                         ptx::umma_desc_sw128(v_addr + k * 2048), idesc_pv, k > 0);
        ptx::umma_commit(bar_pv);
      }
      if (do_prof) {
        prof[0] = (unsigned long long)(clock64() - t_start), prof[1] = wq, prof[2] = wk, prof[3] = wp, prof[4] = wv;
        prof[5] = (unsigned long long)n_tiles;
      }
    }
  } else {
    // ---------------- softmax / correction / epilogue: two threads per query row.
    // Thread (row r, half hc) owns key columns [64 hc, 64 hc + 64) of S and output columns [32 hc, 32 hc + 32) of O.
    const int quarter = warp & 3;        // TMEM lane quarter this warp may touch
    const int hc = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    const float pad_bias = u.pad_key_bias * 1.4426950408889634f;
    float o[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) o[d] = 0.f;
    float mrun = -INFINITY, lrun = 0.f;
    const bool sp = do_prof && warp == 2;
    unsigned long long ws = 0, wbar = 0, wpv = 0;
    const long long ts_start = clock64();
    uint8_t* p_row = smem + Cfg::OFF_P + hc * (Cfg::QT * 128) + r * 128;  // K-block hc of the P tile, row r
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t ph = j & 1;
      const int k0 = j * Cfg::KT + hc * 64;
      const bool ragged = (k0 + 64 > L);  // this half holds the pad token and/or rows past this utterance
      mbar_wait_prof(bar_s, ph, sp, ws);
      ptx::tc_fence_after();
      // The masking of ragged tiles (pad token, keys past the utterance) is hoisted out of the element loops: full tiles
      // run a branch-free path.  Max and sum use 4 independent accumulators to break the dependency chains.
      float mt;
      if (!ragged) mt = attn_rowmax<false>(tmem_s + lane_off + hc * 64, k0, L, 0.f);
      else mt = attn_rowmax<true>(tmem_s + lane_off + hc * 64, k0, L, pad_bias / scale_log2);
      red[hc * Cfg::QT + r] = mt;
      {
        const long long tb = sp ? clock64() : 0;
        named_bar_sync(1, 32 * Cfg::N_SOFTMAX_WARPS);
        if (sp) wbar += (unsigned long long)(clock64() - tb);
      }
      mt = fmaxf(mt, red[(hc ^ 1) * Cfg::QT + r]);
      const float mnew = fmaxf(mrun, mt * scale_log2);  // finite: key 0 of the first tile is always valid
      const float corr = ex2_approx(mrun - mnew);
      float psum;
      if (!ragged) psum = attn_exp_store<false>(tmem_s + lane_off + hc * 64, p_row, r, k0, L, scale_log2, mnew, 0.f);
      else psum = attn_exp_store<true>(tmem_s + lane_off + hc * 64, p_row, r, k0, L, scale_log2, mnew, pad_bias);
      lrun = lrun * corr + psum;
      mrun = mnew;
#pragma unroll
      for (int d = 0; d < 32; ++d) o[d] *= corr;
      ptx::tc_fence_before();
      ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      ptx::mbar_arrive(bar_p);
      mbar_wait_prof(bar_pv, ph, sp, wpv);
      ptx::tc_fence_after();
      {
        uint32_t a[16], b[16];
        ptx::tmem_ld16(tmem_pv + lane_off + hc * 32, a);
        ptx::tmem_ld16(tmem_pv + lane_off + hc * 32 + 16, b);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          o[i] += __uint_as_float(a[i]);
          o[16 + i] += __uint_as_float(b[i]);
        }
      }
    }
    if (sp && lane == 0) prof[8] = (unsigned long long)(clock64() - ts_start), prof[9] = ws, prof[10] = wbar, prof[11] = wpv;
    ptx::tc_fence_before();
    // total row sum = both halves
    red[hc * Cfg::QT + r] = lrun;
    named_bar_sync(1, 32 * Cfg::N_SOFTMAX_WARPS);
    lrun += red[(hc ^ 1) * Cfg::QT + r];
    const int qi = q0 + r;
    if (qi < nk) {
      const float inv = 1.f / lrun;
      bf16* dst = out + (long long)(row0 + qi) * ldo + head * Cfg::D + hc * 32;
#pragma unroll
      for (int d = 0; d < 32; d += 8) {
        uint4 pk;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(o[d + 0] * inv, o[d + 1] * inv);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(o[d + 2] * inv, o[d + 3] * inv);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(o[d + 4] * inv, o[d + 5] * inv);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(o[d + 6] * inv, o[d + 7] * inv);
        pk.x = *reinterpret_cast<uint32_t*>(&h0);
        pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2);
        pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(dst + d) = pk;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

inline int attn_tc_set_attr(std::string* err) {
  cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnTcCfg::SMEM_BYTES);
  if (e != cudaSuccess) {
    *err = std::string("cudaFuncSetAttribute(attn_tc_kernel): ") + cudaGetErrorString(e);
    return -2;
  }
  return 0;
}

template <typename Enc>
inline int launch_attn_tc(Enc encode, const void* qkv, long long ld, int inner, int M, const UttTable* utt, const int4* work,
                          int n_work, void* out, long long ldo, float scale, cudaStream_t s, std::string* err,
                          unsigned long long* prof = nullptr, bool pdl = false) {
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    *err = "cuTensorMapEncodeTiled failed for the attention QKV map";
    return -2;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(n_work), cfg.blockDim = dim3(AttnTcCfg::THREADS), cfg.dynamicSmemBytes = AttnTcCfg::SMEM_BYTES, cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, attn_tc_kernel, tm, inner, utt, work, static_cast<bf16*>(out), ldo,
                                     scale * 1.4426950408889634f, prof);
  if (e != cudaSuccess) {
    *err = std::string("attn_tc_kernel launch: ") + cudaGetErrorString(e);
    return -2;
  }
  return 0;
}

}  // namespace cfm
