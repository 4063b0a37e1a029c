// GEMM family for the estimator: every Conv1d (k=3 / k=3 stride 2 / k=1), the ConvTranspose1d
// phases and every Linear of reference decoder.py / transformer.py is one "tapped" GEMM
//     D[m, n] = sum_taps sum_k A_tap[m + row_shift, a_col + k] * W[w_row + n, k]
// over token-major rows, followed by a fused epilogue.  Two implementations share the same
// parameter block and epilogue code:
//   * gemm_tc_kernel<BN>  - sm_100a: TMA (128B swizzle) -> smem ring -> tcgen05.mma (bf16, fp32 accum in
//                           TMEM, double-buffered accumulators) -> tcgen05.ld epilogue; persistent,
//                           warp-specialised (1 TMA warp, 1 MMA warp, 1 TMEM-alloc warp, 8 epilogue warps).
//   * gemm_simt_kernel<T> - plain fp32-FMA tile kernel; the fp32 precision mode and the on-device
//                           cross-check for the tensor-core path.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ptx.cuh"

namespace cfm {

typedef __nv_bfloat16 bf16;

enum EpiMode : int {
  EPI_STORE = 0,  // out_act = acc + bias
  EPI_STATS = 1,  // out_f32 = acc + bias, GroupNorm partial sums per (utterance, group)
  EPI_RESID = 2,  // y = acc + bias + resid; out_f32 = y; out_act = valid ? y : 0 (optional)
  EPI_SNAKE = 3,  // h = acc + bias; out_act = h + sin^2(h * ea) * ib     (reference transformer.py:68-75)
  EPI_MASK = 4,   // out_act = valid ? acc + bias : 0
  EPI_ODE = 5,    // v = valid ? acc + bias : 0; y = base + c_v v + sum c_k[i] k_i; fixed-grid ODE stage update
};

constexpr int ROW_VALID = 1 << 30;   // row holds a valid mel frame (mask == 1)
constexpr int ROW_INSTAT = 1 << 29;  // row contributes to GroupNorm statistics (valid frames + halo row)
constexpr int ROW_UTT_MASK = (1 << 24) - 1;
constexpr int MAX_TAPS = 9;

struct GemmTap {
  int a_src;      // which A operand (0/1)
  int row_shift;  // A row = output row + row_shift (rows outside [0, a_rows) read as zero)
  int a_col;      // first A column of this tap's K range
  int w_row;      // first W row of this tap's N range
};

struct GemmParams {
  int M, N, K;  // output rows, real output columns, reduction length per tap
  int n_taps;
  GemmTap taps[MAX_TAPS];
  const void* A[2];
  long long lda[2];  // elements
  int a_rows[2];
  const void* W;  // [w_rows, ldw] K-major
  long long ldw;
  int w_rows;
  int mode;
  const float* bias;
  int row_mul, row_add;  // logical row (for row_info) = m * row_mul + row_add
  const int* row_info;   // nullptr: every row valid
  float* out_f32;
  long long ld_f32;
  void* out_act;
  long long ld_act;
  const float* resid;
  long long ld_resid;
  const float* ea;  // exp(alpha)            [N]
  const float* ib;  // 1 / (exp(beta)+1e-9)  [N]
  double* stats;    // [n_utt][8][2]
  int group_ch;     // channels per GroupNorm group
  int fused_stats;  // 1: accumulate stats in the epilogue (tensor-core path)
  float c_v;
  float c_k[3];
  const float* kin[3];
  float* kout;
  long long ld_k;
};

// ------------------------------------------------------------------------------------------------
template <typename T> struct ActIO;
template <> struct ActIO<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
  static __device__ __forceinline__ float fsin(float x) { return sinf(x); }
};
template <> struct ActIO<bf16> {
  static __device__ __forceinline__ float ld(const bf16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ void st(bf16* p, float v) { *p = __float2bfloat16_rn(v); }
  // Cody-Waite reduction to [-pi, pi] then MUFU.SIN; the result is rounded to bf16 anyway.
  static __device__ __forceinline__ float fsin(float x) {
    float k = rintf(x * 0.15915494309189535f);
    float r = fmaf(k, -6.2831854820251465f, x);
    r = fmaf(k, 1.7484555e-7f, r);
    return __sinf(r);
  }
};

template <typename T, int NV>
__device__ __forceinline__ void store_act(T* dst, const float (&v)[NV], bool vec_ok) {
  if constexpr (sizeof(T) == 2 && NV % 8 == 0) {
    if (vec_ok) {
#pragma unroll
      for (int i = 0; i < NV; i += 8) {
        uint4 u;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i + 0], v[i + 1]);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(v[i + 4], v[i + 5]);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(v[i + 6], v[i + 7]);
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        u.z = *reinterpret_cast<uint32_t*>(&h2);
        u.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(dst + i) = u;
      }
      return;
    }
  }
  if constexpr (sizeof(T) == 4 && NV % 4 == 0) {
    if (vec_ok) {
#pragma unroll
      for (int i = 0; i < NV; i += 4)
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) ActIO<T>::st(dst + i, v[i]);
}

template <int NV>
__device__ __forceinline__ void store_f32(float* dst, const float (&v)[NV], bool vec_ok) {
  if (vec_ok) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) dst[i] = v[i];
  }
}

template <int NV>
__device__ __forceinline__ void load_f32(const float* src, float (&v)[NV], bool vec_ok) {
  if (vec_ok) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      float4 t = *reinterpret_cast<const float4*>(src + i);
      v[i] = t.x, v[i + 1] = t.y, v[i + 2] = t.z, v[i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = src[i];
  }
}

__device__ __forceinline__ bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Elementwise part of every epilogue for NV consecutive columns [n0, n0+NV) of output row m.
// `acc` holds the raw accumulators on entry.  Columns >= N are dropped.  (NV % 4 == 0.)
template <typename T, int NV>
__device__ __forceinline__ void epi_apply(const GemmParams& p, int m, int n0, float (&acc)[NV], int info) {
  if (n0 >= p.N) return;
  const bool full = (n0 + NV <= p.N);
  const bool valid = (info & ROW_VALID) != 0;
  float x[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    int n = n0 + i;
    float b = (p.bias != nullptr && n < p.N) ? __ldg(p.bias + n) : 0.f;
    x[i] = acc[i] + b;
  }
  T* oact = reinterpret_cast<T*>(p.out_act);
  if (!full) {  // ragged tail (e.g. final_proj N=100): scalar, predicated
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      int n = n0 + i;
      if (n >= p.N) continue;
      float xi = x[i];
      switch (p.mode) {
        case EPI_STORE: ActIO<T>::st(oact + (long long)m * p.ld_act + n, xi); break;
        case EPI_STATS: p.out_f32[(long long)m * p.ld_f32 + n] = xi; break;
        case EPI_RESID: {
          float y = xi + p.resid[(long long)m * p.ld_resid + n];
          if (p.out_f32) p.out_f32[(long long)m * p.ld_f32 + n] = y;
          if (oact) ActIO<T>::st(oact + (long long)m * p.ld_act + n, valid ? y : 0.f);
        } break;
        case EPI_SNAKE: {
          float s = ActIO<T>::fsin(xi * __ldg(p.ea + n));
          ActIO<T>::st(oact + (long long)m * p.ld_act + n, fmaf(s * s, __ldg(p.ib + n), xi));
        } break;
        case EPI_MASK: ActIO<T>::st(oact + (long long)m * p.ld_act + n, valid ? xi : 0.f); break;
        case EPI_ODE: {
          float v = valid ? xi : 0.f;
          float y = (p.resid ? p.resid[(long long)m * p.ld_resid + n] : 0.f) + p.c_v * v;
#pragma unroll
          for (int j = 0; j < 3; ++j)
            if (p.kin[j]) y = fmaf(p.c_k[j], p.kin[j][(long long)m * p.ld_k + n], y);
          if (p.kout) p.kout[(long long)m * p.ld_k + n] = v;
          if (p.out_f32) p.out_f32[(long long)m * p.ld_f32 + n] = y;
          if (oact) ActIO<T>::st(oact + (long long)m * p.ld_act + n, valid ? y : 0.f);
        } break;
      }
    }
    return;
  }
  switch (p.mode) {
    case EPI_STORE: {
      T* d = oact + (long long)m * p.ld_act + n0;
      store_act<T, NV>(d, x, aligned16(d));
    } break;
    case EPI_STATS: {
      float* d = p.out_f32 + (long long)m * p.ld_f32 + n0;
      store_f32<NV>(d, x, aligned16(d));
    } break;
    case EPI_RESID: {
      const float* r = p.resid + (long long)m * p.ld_resid + n0;
      float rv[NV];
      load_f32<NV>(r, rv, aligned16(r));
#pragma unroll
      for (int i = 0; i < NV; ++i) x[i] += rv[i];
      if (p.out_f32) {
        float* d = p.out_f32 + (long long)m * p.ld_f32 + n0;
        store_f32<NV>(d, x, aligned16(d));
      }
      if (oact) {
        if (!valid) {
#pragma unroll
          for (int i = 0; i < NV; ++i) x[i] = 0.f;
        }
        T* d = oact + (long long)m * p.ld_act + n0;
        store_act<T, NV>(d, x, aligned16(d));
      }
    } break;
    case EPI_SNAKE: {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float s = ActIO<T>::fsin(x[i] * __ldg(p.ea + n0 + i));
        x[i] = fmaf(s * s, __ldg(p.ib + n0 + i), x[i]);
      }
      T* d = oact + (long long)m * p.ld_act + n0;
      store_act<T, NV>(d, x, aligned16(d));
    } break;
    case EPI_MASK: {
      if (!valid) {
#pragma unroll
        for (int i = 0; i < NV; ++i) x[i] = 0.f;
      }
      T* d = oact + (long long)m * p.ld_act + n0;
      store_act<T, NV>(d, x, aligned16(d));
    } break;
    case EPI_ODE: {
      float y[NV];
      if (!valid) {
#pragma unroll
        for (int i = 0; i < NV; ++i) x[i] = 0.f;
      }
      if (p.resid) {
        const float* r = p.resid + (long long)m * p.ld_resid + n0;
        load_f32<NV>(r, y, aligned16(r));
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) y[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) y[i] = fmaf(p.c_v, x[i], y[i]);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (p.kin[j]) {
          const float* kp = p.kin[j] + (long long)m * p.ld_k + n0;
          float kv[NV];
          load_f32<NV>(kp, kv, aligned16(kp));
#pragma unroll
          for (int i = 0; i < NV; ++i) y[i] = fmaf(p.c_k[j], kv[i], y[i]);
        }
      }
      if (p.kout) {
        float* d = p.kout + (long long)m * p.ld_k + n0;
        store_f32<NV>(d, x, aligned16(d));
      }
      if (p.out_f32) {
        float* d = p.out_f32 + (long long)m * p.ld_f32 + n0;
        store_f32<NV>(d, y, aligned16(d));
      }
      if (oact) {
        if (!valid) {
#pragma unroll
          for (int i = 0; i < NV; ++i) y[i] = 0.f;
        }
        T* d = oact + (long long)m * p.ld_act + n0;
        store_act<T, NV>(d, y, aligned16(d));
      }
    } break;
  }
}

__device__ __forceinline__ int load_row_info(const GemmParams& p, int m) {
  if (m >= p.M) return 0;
  if (p.row_info == nullptr) return ROW_VALID;
  return __ldg(p.row_info + (long long)m * p.row_mul + p.row_add);
}

// ------------------------------------------------------------------------------------------------
// SIMT fp32-FMA implementation: 64x64 tile, 16x16 threads, 4x4 outputs per thread.
template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmParams p) {
  __shared__ float As[16][64 + 1];
  __shared__ float Bs[16][64 + 1];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = tid % 16, ty = tid / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = tid / 4, lk = (tid % 4) * 4;  // loader mapping: 64 rows x 16 k, 4 k per thread
  for (int t = 0; t < p.n_taps; ++t) {
    const GemmTap tap = p.taps[t];
    const T* A = reinterpret_cast<const T*>(p.A[tap.a_src]);
    const T* W = reinterpret_cast<const T*>(p.W);
    const long long lda = p.lda[tap.a_src];
    const int arow = m0 + lr + tap.row_shift;
    const bool a_ok = (m0 + lr < p.M) && arow >= 0 && arow < p.a_rows[tap.a_src];
    const int wrow = n0 + lr;
    const bool w_ok = wrow < p.N;
    for (int k0 = 0; k0 < p.K; k0 += 16) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int k = k0 + lk + i;
        float a = 0.f, w = 0.f;
        if (a_ok && k < p.K) a = ActIO<T>::ld(A + (long long)arow * lda + tap.a_col + k);
        if (w_ok && k < p.K) w = ActIO<T>::ld(W + (long long)(tap.w_row + wrow) * p.ldw + k);
        As[lk + i][lr] = a;
        Bs[lk + i][lr] = w;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i], b[i] = Bs[kk][tx * 4 + i];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    int info = load_row_info(p, m);
    epi_apply<T, 4>(p, m, n0 + tx * 4, acc[i], info);
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 implementation.
template <int BN>
struct TcCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int MAX_SMEM = 227 * 1024;
  static constexpr int CTRL_BYTES = 256;
  static constexpr int STAGES_RAW = (MAX_SMEM - 1024 - CTRL_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + CTRL_BYTES;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int N_EPI_WARPS = 8;
  static constexpr int THREADS = 128 + 32 * N_EPI_WARPS;
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN must be a multiple of 32 in [32, 256]");
  static_assert(STAGES >= 3, "pipeline too shallow");
};

template <int BN>
__global__ void __launch_bounds__(TcCfg<BN>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmW, const GemmParams p) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + Cfg::BM - 1) / Cfg::BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int n_tiles_total = m_tiles * n_tiles;
  const int kb_per_tap = p.K / Cfg::BK;
  const int k_iters = p.n_taps * kb_per_tap;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA0);
    ptx::prefetch_tmap(&tmA1);
    ptx::prefetch_tmap(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], Cfg::N_EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (one lane) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * Cfg::BM, n0 = (tile % n_tiles) * BN;
        for (int t = 0; t < p.n_taps; ++t) {
          const GemmTap tap = p.taps[t];
          const CUtensorMap* tmA = tap.a_src ? &tmA1 : &tmA0;
          for (int kb = 0; kb < kb_per_tap; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            ptx::mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            ptx::tma_load_2d(sa, tmA, &full_bar[stage], tap.a_col + kb * Cfg::BK, m0 + tap.row_shift);
            ptx::tma_load_2d(sa + Cfg::A_BYTES, &tmW, &full_bar[stage], kb * Cfg::BK, tap.w_row + n0);
            if (++stage == STAGES) stage = 0, phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one lane) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(Cfg::BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        ptx::mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int it = 0; it < k_iters; ++it) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < Cfg::BK / 16; ++k) {
            ptx::umma_bf16(tmem_d, ptx::umma_desc_sw128(a_addr + k * 32), ptx::umma_desc_sw128(b_addr + k * 32), idesc,
                           (it > 0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
        ptx::umma_commit(&tfull_bar[acc]);
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps: TMEM -> registers -> global =====================
    const int q = warp & 3;              // TMEM lane quarter this warp may touch
    const int half = (warp - 4) >> 2;    // which half of the tile's columns
    constexpr int COLS = BN / 2;
    int local = 0;
    for (int tile = blockIdx.x; tile < n_tiles_total; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int m0 = (tile / n_tiles) * Cfg::BM, n0 = (tile % n_tiles) * BN + half * COLS;
      const int m = m0 + q * 32 + lane;
      const int info = load_row_info(p, m);
      const bool do_stats = (p.mode == EPI_STATS) && p.fused_stats;
      const int utt = info & ROW_UTT_MASK;
      const bool instat = (info & ROW_INSTAT) != 0;
      bool uniform = false;
      if (do_stats) uniform = __all_sync(0xffffffffu, utt == __shfl_sync(0xffffffffu, utt, 0));
      float gs = 0.f, gss = 0.f;
      int cur_g = -1;
      auto flush = [&]() {
        if (cur_g < 0) return;
        if (uniform) {
          float a = gs, b = gss;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
          }
          if (lane == 0 && (a != 0.f || b != 0.f)) {
            atomicAdd(p.stats + ((long long)utt * 8 + cur_g) * 2, (double)a);
            atomicAdd(p.stats + ((long long)utt * 8 + cur_g) * 2 + 1, (double)b);
          }
        } else if (instat) {
          atomicAdd(p.stats + ((long long)utt * 8 + cur_g) * 2, (double)gs);
          atomicAdd(p.stats + ((long long)utt * 8 + cur_g) * 2 + 1, (double)gss);
        }
        gs = gss = 0.f;
      };

      ptx::mbar_wait(&tfull_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + half * COLS + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < COLS; c += 16) {
        uint32_t r[16];
        ptx::tmem_ld16(taddr + c, r);
        ptx::tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        if (do_stats) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int nn = n0 + c + 8 * h;
            if (nn < p.N) {
              const int g = nn / p.group_ch;
              if (g != cur_g) {
                flush();
                cur_g = g;
              }
              if (instat) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float xv = v[8 * h + i] + __ldg(p.bias + nn + i);
                  gs += xv;
                  gss = fmaf(xv, xv, gss);
                }
              }
            }
          }
        }
        if (m < p.M) epi_apply<bf16, 16>(p, m, n0 + c, v, info);
      }
      if (do_stats) flush();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace cfm
