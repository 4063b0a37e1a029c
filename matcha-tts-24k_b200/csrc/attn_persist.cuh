// Persistent form of the tensor-core flash attention (attn_tc.cuh, bf16, head_dim 64): one CTA per SM slot (2 per SM) walks the
// (utterance, head, query tile) items  w = blockIdx.x, blockIdx.x + gridDim.x, ...  and treats the item boundary as just another
// key-tile boundary of ONE continuous software pipeline:
//   * the per-tile barriers / buffers (K, V, S, P: index g & 1 of a tile counter g that runs across items) never drain;
//   * Q is double-buffered by item parity, so S_0 / S_1 of the next item are issued while the last tiles of the current item are
//     still in the softmax warps;
//   * O is double-buffered in tensor memory (columns 128 / 192), and the softmax warps write an item's output only after they have
//     handed P of the NEXT item's first tile to the tensor core (deferred epilogue), so the PV of the last tile is never waited for
//     with an idle tensor pipe.
// What this removes (gpu_diag attnprof on the non-persistent kernel, cfg2): a CTA lives 15.2 us and its successor's first MMA
// starts 2.0 us after it ended (exit, launch, barrier init, TMEM allocation, first Q / K / V round trip, pipeline fill) - 12 % of
// the kernel at 15 key tiles per item, 20 % at 8 (half resolution).  The work list is sorted longest-first (cfm_plan), so the static
// round-robin assignment is an LPT schedule for ragged batches.
#pragma once
#include <cuda.h>
#include <cstring>
#include <string>

#include "attn_tc.cuh"

namespace cfm {

struct AttnPersistCfg {
  static constexpr int D = 64, QT = 128, KT = 64;
  static constexpr int Q_BYTES = QT * D * 2, K_BYTES = KT * D * 2, V_BYTES = KT * D * 2, P_BYTES = QT * KT * 2;
  static constexpr int OFF_Q = 0, OFF_K = 2 * Q_BYTES, OFF_V = OFF_K + 2 * K_BYTES, OFF_P = OFF_V + 2 * V_BYTES;
  static constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;  // 14 mbarriers + TMEM slot (128 B)
  static constexpr int SMEM_BYTES = OFF_BAR + 128 + 1024;
  static constexpr int N_SOFTMAX_WARPS = 4;
  static constexpr int THREADS = 64 + 32 * N_SOFTMAX_WARPS;
  static constexpr int TMEM_COLS = 256;  // S[2] at columns 0 / 64, O[2] at 128 / 192
};

struct AttnItem {
  int L, n_tiles, row0, head, q0;
  float pad_key_bias;
};

#ifdef CFM_ATTN_KERNEL_TU
__device__ __forceinline__ AttnItem attn_load_item(const UttTable* __restrict__ utt, const int4* __restrict__ work, int w) {
  const int4 wi = __ldg(work + w);
  AttnItem it;
  it.L = __ldg(&utt[wi.x].len);
  it.row0 = __ldg(&utt[wi.x].start);
  it.pad_key_bias = __ldg(&utt[wi.x].pad_key_bias);
  it.n_tiles = (it.L + 1 + AttnPersistCfg::KT - 1) / AttnPersistCfg::KT;
  it.head = wi.y, it.q0 = wi.z;
  return it;
}

__global__ void __launch_bounds__(AttnPersistCfg::THREADS, 2)
attn_tc_persist_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, int inner,
                       const UttTable* __restrict__ utt, const int4* __restrict__ work, int n_work, bf16* __restrict__ out, long long ldo,
                       float scale_log2) {
  using Cfg = AttnPersistCfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t *bar_q = bars, *bar_k = bars + 2, *bar_v = bars + 4, *bar_s = bars + 6, *bar_p = bars + 8, *bar_pv = bars + 10, *bar_of = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = ((int)blockIdx.x < n_work) ? (n_work - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_q);
    ptx::prefetch_tmap(&tm_kv);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(bar_q + i, 1);
      ptx::mbar_init(bar_k + i, 1);
      ptx::mbar_init(bar_v + i, 1);
      ptx::mbar_init(bar_s + i, 1);
      ptx::mbar_init(bar_p + i, 32 * Cfg::N_SOFTMAX_WARPS);
      ptx::mbar_init(bar_pv + i, 1);
      ptx::mbar_init(bar_of + i, 32 * Cfg::N_SOFTMAX_WARPS);
    }
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
  ptx::pdl_wait();

  if (warp == 0) {
    // ---------------- TMA producer: tiles in stream order; buffer b = g & 1 is free once tile g - 2 has been consumed
    int g = 0;
    for (int it = 0; it < n_items; ++it) {
      const AttnItem I = attn_load_item(utt, work, blockIdx.x + it * gridDim.x);
      for (int j = 0; j < I.n_tiles; ++j, ++g) {
        const int b = g & 1;
        const uint32_t prev = ((g >> 1) - 1) & 1;  // parity of the previous use of buffer b
        if (g >= 2) ptx::mbar_wait(bar_s + b, prev);  // S_{g-2} complete: K[b] free - and, at j == 0, every S of item it - 2: Q[it & 1] free
        if (j == 0) {
          ptx::mbar_expect_tx_elect(bar_q + (it & 1), Cfg::Q_BYTES);
          ptx::tma_load_2d_elect(smem + Cfg::OFF_Q + (it & 1) * Cfg::Q_BYTES, &tm_q, bar_q + (it & 1), I.head * Cfg::D, I.row0 + I.q0);
        }
        ptx::mbar_expect_tx_elect(bar_k + b, Cfg::K_BYTES);
        ptx::tma_load_2d_elect(smem + Cfg::OFF_K + b * Cfg::K_BYTES, &tm_kv, bar_k + b, inner + I.head * Cfg::D, I.row0 + j * Cfg::KT);
        if (g >= 2) ptx::mbar_wait(bar_pv + b, prev);  // PV_{g-2} complete: V[b] free
        ptx::mbar_expect_tx_elect(bar_v + b, Cfg::V_BYTES);
        ptx::tma_load_2d_elect(smem + Cfg::OFF_V + b * Cfg::V_BYTES, &tm_kv, bar_v + b, 2 * inner + I.head * Cfg::D, I.row0 + j * Cfg::KT);
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: the S cursor runs two tiles ahead of the PV cursor, across item boundaries
    constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(128, 64, 0);
    constexpr uint32_t idesc_pv = ptx::umma_idesc_bf16(128, 64, 1);  // B (= V tile) is MN-major
    const uint32_t q_addr = ptx::smem_u32(smem + Cfg::OFF_Q), k_addr = ptx::smem_u32(smem + Cfg::OFF_K);
    const uint32_t v_addr = ptx::smem_u32(smem + Cfg::OFF_V), p_addr = ptx::smem_u32(smem + Cfg::OFF_P);
    // (the table entries of the NEXT item are requested one item ahead by both cursors: no global-memory latency at a boundary)
    int s_it = 0, s_j = 0, s_g = 0, s_nt = n_items > 0 ? attn_load_item(utt, work, blockIdx.x).n_tiles : 0;
    int s_nt_next = n_items > 1 ? attn_load_item(utt, work, blockIdx.x + gridDim.x).n_tiles : 0;
    auto issue_s = [&]() {
      if (s_it >= n_items) return;
      const int b = s_g & 1;
      if (s_j == 0) ptx::mbar_wait(bar_q + (s_it & 1), (s_it >> 1) & 1);
      ptx::mbar_wait(bar_k + b, (s_g >> 1) & 1);
      ptx::tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        ptx::umma_bf16_elect(tmem_base + b * 64, ptx::umma_desc_sw128(q_addr + (s_it & 1) * Cfg::Q_BYTES + k * 32),
                             ptx::umma_desc_sw128(k_addr + b * Cfg::K_BYTES + k * 32), idesc_s, k > 0);
      ptx::umma_commit_elect(bar_s + b);
      ++s_g;
      if (++s_j == s_nt) {
        s_j = 0, s_nt = s_nt_next;
        if (++s_it + 1 < n_items) s_nt_next = attn_load_item(utt, work, blockIdx.x + (s_it + 1) * gridDim.x).n_tiles;
      }
    };
    issue_s();
    issue_s();
    int g = 0;
    int nt_next = n_items > 0 ? attn_load_item(utt, work, blockIdx.x).n_tiles : 0;
    for (int it = 0; it < n_items; ++it) {
      const int nt = nt_next;
      if (it + 1 < n_items) nt_next = attn_load_item(utt, work, blockIdx.x + (it + 1) * gridDim.x).n_tiles;
      const uint32_t tmem_o = tmem_base + 128 + (it & 1) * 64;
      if (it >= 2) {  // O[it & 1] was drained by the epilogue of item it - 2
        ptx::mbar_wait(bar_of + (it & 1), ((it >> 1) - 1) & 1);
        ptx::tc_fence_after();
      }
      for (int j = 0; j < nt; ++j, ++g) {
        const int b = g & 1;
        const uint32_t ph = (g >> 1) & 1;
        ptx::mbar_wait(bar_p + b, ph);  // P_g in smem; S_g read
        ptx::mbar_wait(bar_v + b, ph);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::umma_bf16_elect(tmem_o, ptx::umma_desc_sw128(p_addr + b * Cfg::P_BYTES + k * 32),
                               ptx::umma_desc_sw128(v_addr + b * Cfg::V_BYTES + k * 2048), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
        ptx::umma_commit_elect(bar_pv + b);
        issue_s();  // S_{g+2} (possibly of the next item): its buffer was read by the softmax warps before they handed over P_g
      }
    }
    ptx::pdl_launch_dependents();
  } else {
    // ---------------- softmax warps: one query row per thread; deferred epilogue of the previous item
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quarter * 32) << 16;
    uint32_t sv[64];
    auto load_s = [&](int g) {
      const uint32_t ts = tmem_base + (g & 1) * 64 + lane_off;
#pragma unroll
      for (int c = 0; c < 4; ++c) ptx::tmem_ld16(ts + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&sv[c * 16]));
    };
    // writes the rows of item `E` (O buffer ob, row sum l) after PV of its last tile gl has landed
    auto epilogue = [&](const AttnItem& E, int ob, int gl, float l) {
      ptx::mbar_wait(bar_pv + (gl & 1), (gl >> 1) & 1);
      ptx::tc_fence_after();
      const int qi = E.q0 + r;
      const float inv = 1.f / l;
      bf16* dst = out + (long long)(E.row0 + qi) * ldo + E.head * Cfg::D;
      const uint32_t tmem_o = tmem_base + 128 + ob * 64 + lane_off;
#pragma unroll
      for (int c = 0; c < 64; c += 16) {
        uint32_t a[16];
        ptx::tmem_ld16(tmem_o + c, a);
        ptx::tmem_ld_wait();
        if (qi < E.L + 1) {
#pragma unroll
          for (int g2 = 0; g2 < 2; ++g2) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(__uint_as_float(a[8 * g2 + 0]) * inv, __uint_as_float(a[8 * g2 + 1]) * inv);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(__uint_as_float(a[8 * g2 + 2]) * inv, __uint_as_float(a[8 * g2 + 3]) * inv);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(a[8 * g2 + 4]) * inv, __uint_as_float(a[8 * g2 + 5]) * inv);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(__uint_as_float(a[8 * g2 + 6]) * inv, __uint_as_float(a[8 * g2 + 7]) * inv);
            uint4 pk;
            pk.x = *reinterpret_cast<uint32_t*>(&h0), pk.y = *reinterpret_cast<uint32_t*>(&h1);
            pk.z = *reinterpret_cast<uint32_t*>(&h2), pk.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(dst + c + 8 * g2) = pk;
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(bar_of + ob);
    };
    int g = 0;
    AttnItem P0;  // previous item (epilogue pending)
    float l_prev = 1.f;
    if (n_items > 0) {
      ptx::mbar_wait(bar_s, 0);
      ptx::tc_fence_after();
      load_s(0);
    }
    AttnItem Nx = n_items > 0 ? attn_load_item(utt, work, blockIdx.x) : AttnItem{};
    for (int it = 0; it < n_items; ++it) {
      const AttnItem I = Nx;
      if (it + 1 < n_items) Nx = attn_load_item(utt, work, blockIdx.x + (it + 1) * gridDim.x);  // next item's table entries travel now
      const int L = I.L;
      const float raw_pad_bias = I.pad_key_bias / (scale_log2 * 0.6931471805599453f);
      const uint32_t tmem_o = tmem_base + 128 + (it & 1) * 64;
      float mref = -INFINITY, lrun = 0.f;
      const bool more_items = it + 1 < n_items;
      for (int j = 0; j < I.n_tiles; ++j, ++g) {
        const int b = g & 1;
        const int k0 = j * Cfg::KT;
        const uint32_t p_row = ptx::smem_u32(smem + Cfg::OFF_P + b * Cfg::P_BYTES + r * 128);
        ptx::tmem_ld_wait();
        if (k0 + Cfg::KT > L) {
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
        const bool grow = mt > mref + 8.f;
        const float mnew = grow ? mt : mref;
        if (j > 0 && __any_sync(0xffffffffu, grow)) {  // rare: rescale this warp's rows of O (in TMEM) and l
          const float corr = ex2_approx(mref - mnew);
          ptx::mbar_wait(bar_pv + ((g - 1) & 1), ((g - 1) >> 1) & 1);  // every PV of this item issued so far has landed in O
          ptx::tc_fence_after();
#pragma unroll
          for (int c = 0; c < 64; c += 16) {
            uint32_t o[16];
            ptx::tmem_ld16(tmem_o + lane_off + c, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
            ptx::tmem_st16(tmem_o + lane_off + c, o);
          }
          ptx::tmem_st_wait();
          lrun *= corr;
        }
        mref = mnew;
        {
          const f32x2 sc2 = pack2(scale_log2, scale_log2), nm2 = pack2(-mref, -mref);
          f32x2 acc01 = pack2(0.f, 0.f), acc23 = pack2(0.f, 0.f);
#pragma unroll
          for (int gg = 0; gg < 8; ++gg) {
            float pv[8];
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
              const f32x2 t2 = fma2(pack2(__uint_as_float(sv[8 * gg + i]), __uint_as_float(sv[8 * gg + i + 1])), sc2, nm2);
              float t0, t1;
              unpack2(t2, t0, t1);
              pv[i] = ex2_approx(t0), pv[i + 1] = ex2_approx(t1);
              if ((i & 2) == 0) acc01 = add2(acc01, pack2(pv[i], pv[i + 1]));
              else acc23 = add2(acc23, pack2(pv[i], pv[i + 1]));
            }
            __nv_bfloat162 h0 = __floats2bfloat162_rn(pv[0], pv[1]), h1 = __floats2bfloat162_rn(pv[2], pv[3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(pv[4], pv[5]), h3 = __floats2bfloat162_rn(pv[6], pv[7]);
            ptx::sts128_u32(p_row + ((gg ^ (r & 7)) << 4), *reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                            *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
          }
          float s0, s1;
          unpack2(add2(acc01, acc23), s0, s1);
          lrun += s0 + s1;
        }
        const bool next_live = j + 1 < I.n_tiles || more_items;  // is there a next tile in this CTA's stream?
        if (j == 0 && it > 0) {
          // first tile of an item: hand P over first, then write the PREVIOUS item's rows (its last PV ran under this tile's
          // softmax), then fetch the next scores - the score registers are dead here, so the epilogue costs no registers
          ptx::tc_fence_before();
          ptx::fence_proxy_async();
          ptx::mbar_arrive(bar_p + b);
          epilogue(P0, (it - 1) & 1, g - 1, l_prev);
          if (next_live) {
            ptx::mbar_wait(bar_s + ((g + 1) & 1), ((g + 1) >> 1) & 1);
            ptx::tc_fence_after();
            load_s(g + 1);
          }
        } else {
          if (next_live) {
            ptx::mbar_wait(bar_s + ((g + 1) & 1), ((g + 1) >> 1) & 1);
            ptx::tc_fence_after();
            load_s(g + 1);
          }
          ptx::tc_fence_before();
          ptx::fence_proxy_async();
          ptx::mbar_arrive(bar_p + b);
        }
      }
      P0 = I, l_prev = lrun;
    }
    if (n_items > 0) epilogue(P0, (n_items - 1) & 1, g - 1, l_prev);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

KernelInfo kinfo_attn_persist() {
  return KernelInfo{reinterpret_cast<const void*>(&attn_tc_persist_kernel), AttnPersistCfg::THREADS, AttnPersistCfg::SMEM_BYTES};
}
#else

template <typename Enc>
inline int launch_attn_persist(Enc encode, const void* qkv, long long ld, int inner, int M, const UttTable* utt, const int4* work, int n_work,
                               void* out, long long ldo, float scale, int max_ctas, cudaStream_t s, std::string* err, bool pdl = false) {
  CUtensorMap tm_q, tm_kv;
  cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t estr[2] = {1, 1};
  for (int i = 0; i < 2; ++i) {
    cuuint32_t box[2] = {64, (cuuint32_t)(i == 0 ? AttnPersistCfg::QT : AttnPersistCfg::KT)};
    CUresult r = encode(i == 0 ? &tm_q : &tm_kv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      *err = "cuTensorMapEncodeTiled failed for the attention QKV map";
      return -2;
    }
  }
  const KernelInfo ki = kinfo_attn_persist();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(n_work < max_ctas ? n_work : max_ctas), cfg.blockDim = dim3(ki.threads), cfg.dynamicSmemBytes = ki.smem, cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl ? 1 : 0;
  bf16* out_b = static_cast<bf16*>(out);
  float scale_log2 = scale * 1.4426950408889634f;
  void* args[] = {&tm_q, &tm_kv, &inner, &utt, &work, &n_work, &out_b, &ldo, &scale_log2};
  cudaError_t e = cudaLaunchKernelExC(&cfg, ki.fn, args);
  if (e != cudaSuccess) {
    *err = std::string("attn_tc_persist_kernel launch: ") + cudaGetErrorString(e);
    return -2;
  }
  return 0;
}

#endif  // CFM_ATTN_KERNEL_TU

}  // namespace cfm
