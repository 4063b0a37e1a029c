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

#include <type_traits>

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
  EPI_STATS16 = 6,  // out_act = bf16(acc + bias), GroupNorm partial sums from the fp32 values (tensor-core path only)
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
  const void* W3;  // fp32 mode only: bf16 [2 w_rows, ldw], rows [0, w_rows) = hi, [w_rows, 2 w_rows) = lo parts of W (bf16 x 3 split)
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
  int cluster;  // CTAs per cluster sharing one weight tile via TMA multicast (1, 2 or 4)
  int pair;     // 1: CTA-pair kernel (tcgen05 cta_group::2, M = 256 per pair)
  int direct_epi;  // 1: epilogue_tile_direct (bf16 modes: rows stored straight from registers with 256-bit stores, no smem)
  int act_f32;  // 1: out_act is fp32 (the fp32 precision mode on the tensor pipe, bf16x3 split operands): generic epilogue only, precise sine
  int tma_epi;  // 1: epilogue_tile_tma (row-per-thread math, swizzled smem staging, TMA store / reduce-add); see launch_gemm
  unsigned long long* prof;  // debug: CTA 0 writes per-role cycle counters (see gemm_tc_kernel); nullptr = off
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
  // sin.approx = (x / 2pi) -> MUFU.SIN, which reduces the range itself; absolute error ~ |x| * 4e-7, far below the
  // bf16 rounding of the result for the arguments SnakeBeta sees (|h * e^alpha| up to a few hundred).
  static __device__ __forceinline__ float fsin(float x) { return __sinf(x); }
};

// sine for the tensor-core fp32 mode: two-constant Cody-Waite reduction to [-pi, pi], then MUFU.SIN (absolute error ~5e-7 for the
// arguments SnakeBeta sees; sinf's slow path for large arguments cost 100 us per FF1 launch)
__device__ __forceinline__ float sin_reduced(float x) {
  const float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(k, -6.2831854820251465f, x);
  r = fmaf(k, 1.7484555e-07f, r);
  return __sinf(r);
}

template <typename T, int NV>
__device__ __forceinline__ void store_act(T* dst, const float (&v)[NV], bool /*unused*/ = true) {
  const uintptr_t addr = reinterpret_cast<uintptr_t>(dst);
  if constexpr (sizeof(T) == 2 && NV % 8 == 0) {
    if ((addr & 15) == 0) {
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
  if constexpr (sizeof(T) == 2 && NV % 4 == 0) {
    if ((addr & 7) == 0) {
#pragma unroll
      for (int i = 0; i < NV; i += 4) {
        uint2 u;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i + 0], v[i + 1]);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(v[i + 2], v[i + 3]);
        u.x = *reinterpret_cast<uint32_t*>(&h0);
        u.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(dst + i) = u;
      }
      return;
    }
  }
  if constexpr (sizeof(T) == 4 && NV % 4 == 0) {
    if ((addr & 15) == 0) {
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
// Lean epilogue of the tensor-core kernel for one 32x32 accumulator block of one warp.
// Phase 1 (caller): the thread that owns TMEM lane r wrote row r of the block into `stg` (row stride EPI_LD floats).
// Phase 2 (here): lane (rs = lane/8, cg = lane%8) owns columns n..n+3 of rows rs, rs+4, ..., rs+28, so that 8 lanes
// cover 32 consecutive columns of a row: global loads / stores are full 128-byte (fp32) or 64-byte (bf16) row segments.
// Everything that depends only on the column (bias, SnakeBeta constants) is loaded once per block.
__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(a, b), h1 = __floats2bfloat162_rn(c, d);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  return u;
}

template <int MODE, int EPI_LD>
__device__ __forceinline__ void epi_block(const GemmParams& p, uint32_t stg, int m0, int n, int lane, int info,
                                          uint32_t valid_mask, uint32_t instat_mask, bool stats_uniform, bool do_stats,
                                          uint32_t warp_slots, int g_first, int rb = 0) {
  // rb > 0: the warp's 32 rows hold exactly two utterances, rows [0, rb) the first and [rb, 32) the second ("split" warp)
  const int rs = lane >> 3, cg = lane & 7;
  const bool col_ok = n < p.N;  // N % 4 == 0: a lane's 4 columns are all inside or all outside
  if (!col_ok) n = 0;           // keep the lane in the shuffles below with harmless addresses, stores predicated off
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
  float4 ea4, ib4;
  if constexpr (MODE == EPI_SNAKE) {
    ea4 = __ldg(reinterpret_cast<const float4*>(p.ea + n));
    ib4 = __ldg(reinterpret_cast<const float4*>(p.ib + n));
  }
  float gs = 0.f, gss = 0.f, gsB = 0.f, gssB = 0.f;
  const bool split = rb > 0;
  int cur_utt = -1;
  bf16* oact = reinterpret_cast<bf16*>(p.out_act);
  auto store_act4 = [&](int m, float a0, float a1, float a2, float a3) {  // 4 columns of row m in the activation type
    if (p.act_f32) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out_act) + (long long)m * p.ld_act + n) = make_float4(a0, a1, a2, a3);
    else *reinterpret_cast<uint2*>(oact + (long long)m * p.ld_act + n) = pack4_bf16(a0, a1, a2, a3);
  };
  // All global loads of the block are issued before any store: the residual stream is updated in place, so the compiler
  // must otherwise order every load behind the previous row's store and exposes one memory round trip per row.
  float4 rv[8];
  if constexpr (MODE == EPI_RESID || MODE == EPI_ODE) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int m = m0 + it * 4 + rs;
      rv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < p.M && col_ok && p.resid) rv[it] = *reinterpret_cast<const float4*>(p.resid + (long long)m * p.ld_resid + n);
    }
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + rs;
    const int m = m0 + r;
    const float4 a = ptx::lds128(stg + (r * EPI_LD + cg * 4) * 4);
    int info_r = 0;
    if constexpr (MODE == EPI_STATS) {
      if (do_stats && !stats_uniform && !split) info_r = __shfl_sync(0xffffffffu, info, r);  // utterance id per row: multi-boundary warps only
    }
    if (m >= p.M || !col_ok) continue;
    float x0 = a.x + b4.x, x1 = a.y + b4.y, x2 = a.z + b4.z, x3 = a.w + b4.w;
    const bool valid = (valid_mask >> r) & 1u;
    if constexpr (MODE == EPI_STORE) {
      store_act4(m, x0, x1, x2, x3);
    } else if constexpr (MODE == EPI_STATS) {
      *reinterpret_cast<float4*>(p.out_f32 + (long long)m * p.ld_f32 + n) = make_float4(x0, x1, x2, x3);
      if (do_stats && (stats_uniform || split)) {
        // one utterance (or two, split at row rb): branch-free accumulation into the first / second set
        const bool in = ((instat_mask >> r) & 1u) != 0;
        const float s4 = in ? (x0 + x1) + (x2 + x3) : 0.f;
        const float q4 = in ? fmaf(x0, x0, x1 * x1) + fmaf(x2, x2, x3 * x3) : 0.f;
        if (!split || r < rb) gs += s4, gss += q4;
        else gsB += s4, gssB += q4;
      } else if (do_stats && ((instat_mask >> r) & 1u)) {
        // Three or more utterances inside 32 rows (very short utterances).  Rows are sorted by utterance, so a lane sees a
        // non-decreasing utterance id; the running sums are flushed whenever the id changes (and once at the end).
        const int ur = stats_uniform ? 0 : (info_r & ROW_UTT_MASK);
        if (ur != cur_utt) {
          if (cur_utt >= 0 && !stats_uniform) {
            const long long o = ((long long)cur_utt * 8 + n / p.group_ch) * 2;
            atomicAdd(p.stats + o, (double)gs);
            atomicAdd(p.stats + o + 1, (double)gss);
            gs = gss = 0.f;
          }
          cur_utt = ur;
        }
        gs += (x0 + x1) + (x2 + x3);
        gss = fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, fmaf(x3, x3, gss))));
      }
    } else if constexpr (MODE == EPI_RESID) {
      x0 += rv[it].x, x1 += rv[it].y, x2 += rv[it].z, x3 += rv[it].w;
      if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + (long long)m * p.ld_f32 + n) = make_float4(x0, x1, x2, x3);
      if (oact) {
        if (!valid) x0 = x1 = x2 = x3 = 0.f;
        store_act4(m, x0, x1, x2, x3);
      }
    } else if constexpr (MODE == EPI_SNAKE) {
      float s0, s1, s2, s3;
      if (p.act_f32) s0 = sin_reduced(x0 * ea4.x), s1 = sin_reduced(x1 * ea4.y), s2 = sin_reduced(x2 * ea4.z), s3 = sin_reduced(x3 * ea4.w);
      else s0 = ActIO<bf16>::fsin(x0 * ea4.x), s1 = ActIO<bf16>::fsin(x1 * ea4.y), s2 = ActIO<bf16>::fsin(x2 * ea4.z), s3 = ActIO<bf16>::fsin(x3 * ea4.w);
      x0 = fmaf(s0 * s0, ib4.x, x0), x1 = fmaf(s1 * s1, ib4.y, x1), x2 = fmaf(s2 * s2, ib4.z, x2), x3 = fmaf(s3 * s3, ib4.w, x3);
      store_act4(m, x0, x1, x2, x3);
    } else if constexpr (MODE == EPI_MASK) {
      if (!valid) x0 = x1 = x2 = x3 = 0.f;
      store_act4(m, x0, x1, x2, x3);
    } else {  // EPI_ODE
      if (!valid) x0 = x1 = x2 = x3 = 0.f;
      float4 y = rv[it];
      y.x = fmaf(p.c_v, x0, y.x), y.y = fmaf(p.c_v, x1, y.y), y.z = fmaf(p.c_v, x2, y.z), y.w = fmaf(p.c_v, x3, y.w);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (p.kin[j]) {
          const float4 kv = *reinterpret_cast<const float4*>(p.kin[j] + (long long)m * p.ld_k + n);
          y.x = fmaf(p.c_k[j], kv.x, y.x), y.y = fmaf(p.c_k[j], kv.y, y.y), y.z = fmaf(p.c_k[j], kv.z, y.z), y.w = fmaf(p.c_k[j], kv.w, y.w);
        }
      }
      if (p.kout) *reinterpret_cast<float4*>(p.kout + (long long)m * p.ld_k + n) = make_float4(x0, x1, x2, x3);
      if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + (long long)m * p.ld_f32 + n) = y;
      if (oact) {
        if (!valid) y = make_float4(0.f, 0.f, 0.f, 0.f);
        store_act4(m, y.x, y.y, y.z, y.w);
      }
    }
  }
  if constexpr (MODE == EPI_STATS) {
    if (do_stats && !stats_uniform && !split && cur_utt >= 0 && col_ok) {  // multi-boundary warp: each lane flushes its last segment
      const long long o = ((long long)cur_utt * 8 + n / p.group_ch) * 2;
      atomicAdd(p.stats + o, (double)gs);
      atomicAdd(p.stats + o + 1, (double)gss);
    }
    // Sums of one utterance -> global: combine the column quads of a GroupNorm group first (segmented scan over cg), then
    // only the segment tails issue fp64 atomics (same-address atomics serialise in L2, ~30 cycles each).
    auto flush_global = [&](float s_, float ss_, int utt_) {
      s_ += __shfl_xor_sync(0xffffffffu, s_, 8), ss_ += __shfl_xor_sync(0xffffffffu, ss_, 8);
      s_ += __shfl_xor_sync(0xffffffffu, s_, 16), ss_ += __shfl_xor_sync(0xffffffffu, ss_, 16);
      const int gid = col_ok ? n / p.group_ch : 64 + cg;
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const float us = __shfl_up_sync(0xffffffffu, s_, o, 8), uss = __shfl_up_sync(0xffffffffu, ss_, o, 8);
        const int ug = __shfl_up_sync(0xffffffffu, gid, o, 8);
        if (cg >= o && ug == gid) s_ += us, ss_ += uss;
      }
      const int ng = __shfl_down_sync(0xffffffffu, gid, 1, 8);
      if (rs == 0 && col_ok && (cg == 7 || ng != gid) && (s_ != 0.f || ss_ != 0.f)) {
        const long long o = ((long long)utt_ * 8 + gid) * 2;
        atomicAdd(p.stats + o, (double)s_);
        atomicAdd(p.stats + o + 1, (double)ss_);
      }
    };
    if (do_stats && split) {
      flush_global(gs, gss, __shfl_sync(0xffffffffu, info, 0) & ROW_UTT_MASK);
      flush_global(gsB, gssB, __shfl_sync(0xffffffffu, info, 31) & ROW_UTT_MASK);
    }
    if (do_stats && stats_uniform && !warp_slots) flush_global(gs, gss, __shfl_sync(0xffffffffu, info, 0) & ROW_UTT_MASK);
    if (do_stats && stats_uniform && warp_slots) {  // reduce over the 4 row phases, then this warp's smem slots
      gs += __shfl_xor_sync(0xffffffffu, gs, 8);
      gss += __shfl_xor_sync(0xffffffffu, gss, 8);
      gs += __shfl_xor_sync(0xffffffffu, gs, 16);
      gss += __shfl_xor_sync(0xffffffffu, gss, 16);
      {
        // The whole tile is one utterance: combine the 8 column groups of this block that fall into the same GroupNorm
        // group (segmented scan over cg), then the segment tails add into this warp's private smem slots.  No atomics,
        // fixed order -> bitwise reproducible; the last warp of the tile reduces the 8 warps' slots (gemm_tc_kernel).
        const int gid = col_ok ? n / p.group_ch - g_first : 64 + cg;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
          const float us = __shfl_up_sync(0xffffffffu, gs, o, 8), uss = __shfl_up_sync(0xffffffffu, gss, o, 8);
          const int ug = __shfl_up_sync(0xffffffffu, gid, o, 8);
          if (cg >= o && ug == gid) gs += us, gss += uss;
        }
        const int ng = __shfl_down_sync(0xffffffffu, gid, 1, 8);
        if (rs == 0 && col_ok && (cg == 7 || ng != gid)) {
          const uint32_t a = warp_slots + gid * 8;
          ptx::sts32(a, ptx::lds32(a) + gs);
          ptx::sts32(a + 4, ptx::lds32(a + 4) + gss);
        }
      }
    }
  }
}

// bf16-output modes (STORE / SNAKE / MASK): 8 columns per lane (4 lanes cover 32 columns of a row, 8 rows per instruction),
// one 16-byte store per lane and row.  Requires N % 8 == 0.
template <int MODE, int EPI_LD>
__device__ __forceinline__ void epi_block_wide(const GemmParams& p, uint32_t stg, int m0, int nblk, int lane, uint32_t valid_mask) {
  const int rs = lane >> 2, cg = lane & 3;
  const int n = nblk + cg * 8;
  if (n >= p.N) return;  // no shuffles below: a lane may leave
  float b[8], ea[8], ib[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) b[i] = 0.f;
  if (p.bias) {
    const float4 t0 = __ldg(reinterpret_cast<const float4*>(p.bias + n)), t1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
    b[0] = t0.x, b[1] = t0.y, b[2] = t0.z, b[3] = t0.w, b[4] = t1.x, b[5] = t1.y, b[6] = t1.z, b[7] = t1.w;
  }
  if constexpr (MODE == EPI_SNAKE) {
    const float4 e0 = __ldg(reinterpret_cast<const float4*>(p.ea + n)), e1 = __ldg(reinterpret_cast<const float4*>(p.ea + n + 4));
    const float4 i0 = __ldg(reinterpret_cast<const float4*>(p.ib + n)), i1 = __ldg(reinterpret_cast<const float4*>(p.ib + n + 4));
    ea[0] = e0.x, ea[1] = e0.y, ea[2] = e0.z, ea[3] = e0.w, ea[4] = e1.x, ea[5] = e1.y, ea[6] = e1.z, ea[7] = e1.w;
    ib[0] = i0.x, ib[1] = i0.y, ib[2] = i0.z, ib[3] = i0.w, ib[4] = i1.x, ib[5] = i1.y, ib[6] = i1.z, ib[7] = i1.w;
  }
  bf16* orow = reinterpret_cast<bf16*>(p.out_act) + (long long)(m0 + rs) * p.ld_act + n;
  const long long ostep = 8 * p.ld_act;
#pragma unroll
  for (int it = 0; it < 4; ++it, orow += ostep) {
    const int r = it * 8 + rs;
    if (m0 + r >= p.M) break;  // rows increase with it
    const float4 a0 = ptx::lds128(stg + (r * EPI_LD + cg * 8) * 4);
    const float4 a1 = ptx::lds128(stg + (r * EPI_LD + cg * 8) * 4 + 16);
    float x[8] = {a0.x + b[0], a0.y + b[1], a0.z + b[2], a0.w + b[3], a1.x + b[4], a1.y + b[5], a1.z + b[6], a1.w + b[7]};
    if constexpr (MODE == EPI_SNAKE) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float sn = ActIO<bf16>::fsin(x[i] * ea[i]);
        x[i] = fmaf(sn * sn, ib[i], x[i]);
      }
    }
    if constexpr (MODE == EPI_MASK) {
      if (!((valid_mask >> r) & 1u)) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = 0.f;
      }
    }
    const uint2 lo = pack4_bf16(x[0], x[1], x[2], x[3]), hi = pack4_bf16(x[4], x[5], x[6], x[7]);
    *reinterpret_cast<uint4*>(orow) = make_uint4(lo.x, lo.y, hi.x, hi.y);
  }
}

// Epilogue of one 128-row x BN accumulator tile for one warp (TMEM lane quarter q = warp & 3; the two warps of a
// quarter alternate over the 32-column blocks).  The next block's tcgen05.ld is issued before the current block's global
// I/O so TMEM latency overlaps it.  Shared by the 1-CTA and the CTA-pair kernels.
template <int BN, int MODE, int EPI_LD, int N_EPI_WARPS>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t taddr, uint32_t stg, int m0, int n0, int ewarp, int lane,
                                              int acc, uint32_t gn_slots, uint32_t gn_counters, int info_pre = -1) {
  const int half = ewarp >> 2;
  const int q = ewarp & 3;
  // STORE / SNAKE never look at the row flags; the other modes get them from the caller, which requests them before it waits
  // for the accumulator (info_pre >= 0), so that their L2 latency does not open the tile's epilogue
  const int info = (MODE == EPI_STORE || MODE == EPI_SNAKE) ? ROW_VALID : (info_pre >= 0 ? info_pre : load_row_info(p, m0 + lane));
  const uint32_t valid_mask = __ballot_sync(0xffffffffu, (info & ROW_VALID) != 0);
  const uint32_t instat_mask = __ballot_sync(0xffffffffu, (info & ROW_INSTAT) != 0);
  bool do_stats = false, uniform = false;
  uint32_t warp_slots = 0;
  int tile_utt = 0, g_first = 0, rb = 0;
  if constexpr (MODE == EPI_STATS) {
    do_stats = p.fused_stats != 0;
    const int utt = info & ROW_UTT_MASK;
    if (do_stats) {
      const int utt_first = __shfl_sync(0xffffffffu, utt, 0), utt_last = __shfl_sync(0xffffffffu, utt, 31);
      uniform = __all_sync(0xffffffffu, utt == utt_first);
      // exactly two utterances in the warp (rows are sorted by utterance): split row = number of rows of the first one.
      // Rows past M carry utterance 0 in `info`; such a warp takes the generic multi-boundary path.
      if (!uniform && __all_sync(0xffffffffu, utt == utt_first || utt == utt_last) && m0 + 31 < p.M)
        rb = __popc(__ballot_sync(0xffffffffu, utt == utt_first));
      // If the 128 rows of the tile are one utterance (the common case) the sums are gathered per CTA in smem.
      const int mt = m0 - q * 32;
      const int u_lo = __ldg(p.row_info + min(mt, p.M - 1)) & ROW_UTT_MASK;
      const int u_hi = __ldg(p.row_info + min(mt + 127, p.M - 1)) & ROW_UTT_MASK;
      if (u_lo == u_hi && mt < p.M && (BN % p.group_ch) == 0 && p.row_mul == 1) {
        warp_slots = gn_slots + ((acc * 8 + ewarp) * 16) * 4;
        tile_utt = u_lo, g_first = n0 / p.group_ch;
        uniform = true;  // rows past M carry no ROW_INSTAT flag and add nothing
        rb = 0;
      }
    }
  }
  constexpr bool WIDE = (MODE == EPI_STORE || MODE == EPI_SNAKE || MODE == EPI_MASK);
  const bool wide = WIDE && (p.N % 8 == 0) && !p.act_f32;
  uint32_t r0[16], r1[16];
  if (half < BN / 32) {
    ptx::tmem_ld16(taddr + half * 32, r0);
    ptx::tmem_ld16(taddr + half * 32 + 16, r1);
  }
#pragma unroll 1
  constexpr int BLK_STEP = N_EPI_WARPS / 4;  // warps per TMEM lane quarter alternate over the 32-column blocks
  for (int blk = half; blk < BN / 32; blk += BLK_STEP) {
    const int c = blk * 32;
    ptx::tmem_ld_wait();
    const uint32_t srow = stg + lane * EPI_LD * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ptx::sts128(srow + i * 16, __uint_as_float(r0[4 * i]), __uint_as_float(r0[4 * i + 1]), __uint_as_float(r0[4 * i + 2]),
                  __uint_as_float(r0[4 * i + 3]));
      ptx::sts128(srow + 64 + i * 16, __uint_as_float(r1[4 * i]), __uint_as_float(r1[4 * i + 1]), __uint_as_float(r1[4 * i + 2]),
                  __uint_as_float(r1[4 * i + 3]));
    }
    if (blk + BLK_STEP < BN / 32) {  // prefetch the next block's accumulators while this one goes out to global memory
      ptx::tmem_ld16(taddr + c + 32 * BLK_STEP, r0);
      ptx::tmem_ld16(taddr + c + 32 * BLK_STEP + 16, r1);
    }
    __syncwarp();
    if constexpr (WIDE) {
      if (wide) epi_block_wide<MODE, EPI_LD>(p, stg, m0, n0 + c, lane, valid_mask);
      else epi_block<MODE, EPI_LD>(p, stg, m0, n0 + c + (lane & 7) * 4, lane, info, valid_mask, instat_mask, uniform, do_stats, warp_slots, g_first, rb);
    } else {
      epi_block<MODE, EPI_LD>(p, stg, m0, n0 + c + (lane & 7) * 4, lane, info, valid_mask, instat_mask, uniform, do_stats, warp_slots, g_first, rb);
    }
    __syncwarp();
  }
  if constexpr (MODE == EPI_STATS) {
    if (warp_slots) {  // the last of the 8 epilogue warps to finish this tile reduces and flushes (16 atomics / tile)
      __syncwarp();
      uint32_t old = 0;
      if (lane == 0) {
        __threadfence_block();
        old = ptx::atoms_add_u32(gn_counters + acc * 4, 1);
      }
      old = __shfl_sync(0xffffffffu, old, 0);
      if (old == N_EPI_WARPS - 1) {
        __threadfence_block();
        if (lane < 16) {
          double v = 0.0;
#pragma unroll
          for (int w8 = 0; w8 < 8; ++w8) {
            const uint32_t a = gn_slots + ((acc * 8 + w8) * 16 + lane) * 4;
            v += (double)ptx::lds32(a);
            ptx::sts32(a, 0.f);
          }
          if (v != 0.0) atomicAdd(p.stats + ((long long)tile_utt * 8 + g_first + (lane >> 1)) * 2 + (lane & 1), v);
        }
        if (lane == 0) ptx::sts32(gn_counters + acc * 4, 0.f);
        __syncwarp();
        __threadfence_block();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// TMA-store epilogue (modes STORE / SNAKE / MASK -> bf16, RESID without activation copy -> fp32 add into the stream).
// tcgen05.ld gives each thread one accumulator row; instead of transposing through smem to get coalesced global accesses,
// the thread applies the epilogue math to its own row (column constants come from uniform, L1-resident __ldg), writes
// the converted row into a swizzled staging tile and one lane hands the 32 x 32 unit to the TMA engine:
//   bf16 : 32 rows x 64 B, SWIZZLE_64B, two units in flight per warp         (cp.async.bulk.tensor store)
//   fp32 : 32 rows x 128 B, SWIZZLE_128B, x += acc + bias done by L2         (cp.reduce.async.bulk.tensor .add)
// Per 128 x BN tile this moves 128*BN*2 (bf16) bytes into smem once and lets TMA read them once - no ld.shared, no
// st.global, no residual load - against write + read of an fp32 tile plus per-row global instructions in epilogue_tile;
// the GEMMs are shared-memory-bandwidth bound, so the staging bytes are what the epilogue costs.
// The accumulator is released (`release()`) as soon as the warp's last tcgen05.ld has completed, before the stores drain.
template <int BN, int MODE, typename Release>
__device__ __forceinline__ void epilogue_tile_tma(const GemmParams& p, const CUtensorMap* tm_out, uint32_t taddr, uint32_t stg,
                                                  int m0, int n0, int half, int lane, Release&& release) {
  constexpr int UNITS = BN / 32;
  constexpr bool F32 = (MODE == EPI_RESID);
  constexpr int UNIT_BYTES = F32 ? 4096 : 2048;
  constexpr int GROUP = 4096 / UNIT_BYTES;  // units staged per fence / TMA-issue round (per-warp staging: 4 KB of the 5 KB slot)
  bool valid = true;
  if constexpr (MODE == EPI_MASK) valid = (load_row_info(p, m0 + lane) & ROW_VALID) != 0;
  uint32_t ra[16], rb[16];
  bool released = false;
  if (half < UNITS) {
    ptx::tmem_ld16(taddr + half * 32, ra);
    ptx::tmem_ld16(taddr + half * 32 + 16, rb);
  }
  int staged = 0;        // units staged and not yet handed to the TMA engine
  int first_n = 0;       // first column of the oldest staged unit (staged units are 64 columns apart)
  auto flush = [&]() {   // generic-proxy smem writes -> async proxy, then one lane issues the stores of this round
    ptx::fence_proxy_async();
    __syncwarp();
    for (int i = 0; i < staged; ++i) {  // warp-uniform: the elected lane issues
      if constexpr (F32) ptx::tma_reduce_add_2d_elect(tm_out, stg + i * UNIT_BYTES, first_n + i * 64, m0);
      else ptx::tma_store_2d_elect(tm_out, stg + i * UNIT_BYTES, first_n + i * 64, m0);
    }
    ptx::bulk_commit_elect();
    staged = 0;
  };
#pragma unroll 1
  for (int u = half; u < UNITS; u += 2) {
    const int n = n0 + u * 32;
    ptx::tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(ra[i]), v[16 + i] = __uint_as_float(rb[i]);
    if (u + 2 < UNITS) {  // next unit's accumulators travel while this one is converted and staged
      ptx::tmem_ld16(taddr + (u + 2) * 32, ra);
      ptx::tmem_ld16(taddr + (u + 2) * 32 + 16, rb);
    } else {
      release();
      released = true;
    }
    if (n >= p.N || m0 >= p.M) continue;  // warp-uniform
    // ---- epilogue math on this thread's row; columns >= N are clipped by the TMA store
    if (p.bias) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        if (n + i < p.N) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n + i));
          v[i] += b.x, v[i + 1] += b.y, v[i + 2] += b.z, v[i + 3] += b.w;
        }
      }
    }
    if constexpr (MODE == EPI_SNAKE) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        if (n + i < p.N) {
          const float4 ea = __ldg(reinterpret_cast<const float4*>(p.ea + n + i));
          const float4 ib = __ldg(reinterpret_cast<const float4*>(p.ib + n + i));
          const float s0 = ActIO<bf16>::fsin(v[i] * ea.x), s1 = ActIO<bf16>::fsin(v[i + 1] * ea.y);
          const float s2 = ActIO<bf16>::fsin(v[i + 2] * ea.z), s3 = ActIO<bf16>::fsin(v[i + 3] * ea.w);
          v[i] = fmaf(s0 * s0, ib.x, v[i]), v[i + 1] = fmaf(s1 * s1, ib.y, v[i + 1]);
          v[i + 2] = fmaf(s2 * s2, ib.z, v[i + 2]), v[i + 3] = fmaf(s3 * s3, ib.w, v[i + 3]);
        }
      }
    }
    if constexpr (MODE == EPI_MASK) {
      if (!valid) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
      }
    }
    // ---- stage; before the first write of a round the previous round's stores must have read the staging area
    if (staged == 0) {
      ptx::bulk_wait_read_elect<0>();
      __syncwarp();
      first_n = n;
    }
    const uint32_t buf = stg + staged * UNIT_BYTES;
    if constexpr (F32) {
      const uint32_t row = buf + lane * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) ptx::sts128(row + ((c ^ (lane & 7)) << 4), v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    } else {
      const uint32_t row = buf + lane * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint2 lo = pack4_bf16(v[8 * c], v[8 * c + 1], v[8 * c + 2], v[8 * c + 3]);
        const uint2 hi = pack4_bf16(v[8 * c + 4], v[8 * c + 5], v[8 * c + 6], v[8 * c + 7]);
        ptx::sts128_u32(row + ((c ^ ((lane >> 1) & 3)) << 4), lo.x, lo.y, hi.x, hi.y);
      }
    }
    if (++staged == GROUP) flush();
  }
  if (staged > 0) flush();
  if (!released) release();
}

// ------------------------------------------------------------------------------------------------
// Direct epilogue (bf16-output modes): no shared-memory staging at all.  tcgen05.ld leaves each thread one accumulator row;
// 32 consecutive bf16 columns of it are exactly two 32-byte sectors, written with two 256-bit stores (st.global.v8.b32).  A warp
// instruction therefore touches 32 different rows, but every sector is written whole and nothing goes through the shared
// memory port that TMA writes, UMMA operand reads and the transposing epilogue's ld/st.shared compete for (ncu: the staging
// ld.shared take 2x their ideal wavefronts inside these kernels).  Requires N % 32 == 0, ld_act % 16 == 0, 32-byte aligned base.
__device__ __forceinline__ void stg256(void* dst, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
               "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
template <int BN, int MODE, typename Release>
__device__ __forceinline__ void epilogue_tile_direct(const GemmParams& p, uint32_t taddr, int m0, int n0, int half, int lane,
                                                     Release&& release, int info_pre = -1) {
  constexpr int UNITS = BN / 32;
  const int m = m0 + lane;
  bool valid = true;
  if constexpr (MODE == EPI_MASK) valid = ((info_pre >= 0 ? info_pre : load_row_info(p, m)) & ROW_VALID) != 0;
  bf16* orow = reinterpret_cast<bf16*>(p.out_act) + (long long)m * p.ld_act;
  uint32_t ra[16], rb[16];
  bool released = false;
  if (half < UNITS) {
    ptx::tmem_ld16(taddr + half * 32, ra);
    ptx::tmem_ld16(taddr + half * 32 + 16, rb);
  }
#pragma unroll 1
  for (int u = half; u < UNITS; u += 2) {
    const int n = n0 + u * 32;
    ptx::tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(ra[i]), v[16 + i] = __uint_as_float(rb[i]);
    if (u + 2 < UNITS) {
      ptx::tmem_ld16(taddr + (u + 2) * 32, ra);
      ptx::tmem_ld16(taddr + (u + 2) * 32 + 16, rb);
    } else {
      release();
      released = true;
    }
    if (n >= p.N) continue;
    if (p.bias) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n + i));
        v[i] += b.x, v[i + 1] += b.y, v[i + 2] += b.z, v[i + 3] += b.w;
      }
    }
    if constexpr (MODE == EPI_SNAKE) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 ea = __ldg(reinterpret_cast<const float4*>(p.ea + n + i));
        const float4 ib = __ldg(reinterpret_cast<const float4*>(p.ib + n + i));
        const float s0 = ActIO<bf16>::fsin(v[i] * ea.x), s1 = ActIO<bf16>::fsin(v[i + 1] * ea.y);
        const float s2 = ActIO<bf16>::fsin(v[i + 2] * ea.z), s3 = ActIO<bf16>::fsin(v[i + 3] * ea.w);
        v[i] = fmaf(s0 * s0, ib.x, v[i]), v[i + 1] = fmaf(s1 * s1, ib.y, v[i + 1]);
        v[i + 2] = fmaf(s2 * s2, ib.z, v[i + 2]), v[i + 3] = fmaf(s3 * s3, ib.w, v[i + 3]);
      }
    }
    if constexpr (MODE == EPI_MASK) {
      if (!valid) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
      }
    }
    if (m < p.M) {
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          __nv_bfloat162 hb = __floats2bfloat162_rn(v[16 * hlf + 2 * i], v[16 * hlf + 2 * i + 1]);
          w[i] = *reinterpret_cast<uint32_t*>(&hb);
        }
        stg256(orow + n + 16 * hlf, w);
      }
    }
  }
  if (!released) release();
}

// ------------------------------------------------------------------------------------------------
// Conv + GroupNorm statistics with a bf16 result (EPI_STATS16): row-per-thread like epilogue_tile_direct - the accumulator row
// leaves with 256-bit stores, no shared-memory transpose - and the statistics are taken from the fp32 values on the way: a
// thread sums its own row over the columns of the current GroupNorm group, the warp combines the 32 rows with shuffles when a
// group ends (a 32-column unit holds at most one group boundary) and one lane adds the pair to the fp64 sums.  Warps whose 32
// rows hold two utterances let every lane add for itself (rare: one warp per utterance boundary).
// The row flags are loaded by the caller BEFORE it waits for the accumulator (their L2 latency otherwise opens every tile's epilogue).
template <int BN, typename Release>
__device__ __forceinline__ void epilogue_tile_stats16(const GemmParams& p, uint32_t taddr, int m0, int n0, int half, int lane, int info,
                                                      Release&& release) {
  constexpr int UNITS = BN / 32;
  const int m = m0 + lane;
  const bool in = p.fused_stats != 0 && (info & ROW_INSTAT) != 0;
  const int utt = info & ROW_UTT_MASK;
  const uint32_t in_mask = __ballot_sync(0xffffffffu, in);
  const int wutt = __shfl_sync(0xffffffffu, utt, in_mask ? __ffs(in_mask) - 1 : 0);
  const bool uni = __all_sync(0xffffffffu, !in || utt == wutt);
  const uint32_t b_mask = __ballot_sync(0xffffffffu, in && utt != wutt);  // rows of a second utterance, if any
  const int utt_b = __shfl_sync(0xffffffffu, utt, b_mask ? __ffs(b_mask) - 1 : 0);
  const bool two = !uni && __all_sync(0xffffffffu, !in || utt == wutt || utt == utt_b);
  bf16* orow = reinterpret_cast<bf16*>(p.out_act) + (long long)m * p.ld_act;
  float cs = 0.f, css = 0.f;
  int cur_g = -1;
  auto flush = [&]() {  // the sums of group cur_g are complete for this warp's rows
    if (in_mask != 0u) {
      if (uni) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, o), css += __shfl_xor_sync(0xffffffffu, css, o);
        if (lane == 0) {
          atomicAdd(p.stats + ((long long)wutt * 8 + cur_g) * 2, (double)cs);
          atomicAdd(p.stats + ((long long)wutt * 8 + cur_g) * 2 + 1, (double)css);
        }
      } else if (two) {  // rows of two utterances (sorted): one reduction per utterance, two lanes add
        float sa = (in && utt == wutt) ? cs : 0.f, qa = (in && utt == wutt) ? css : 0.f;
        float sb = (in && utt == utt_b) ? cs : 0.f, qb = (in && utt == utt_b) ? css : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          sa += __shfl_xor_sync(0xffffffffu, sa, o), qa += __shfl_xor_sync(0xffffffffu, qa, o);
          sb += __shfl_xor_sync(0xffffffffu, sb, o), qb += __shfl_xor_sync(0xffffffffu, qb, o);
        }
        if (lane < 2) {
          const int uu = lane == 0 ? wutt : utt_b;
          atomicAdd(p.stats + ((long long)uu * 8 + cur_g) * 2, (double)(lane == 0 ? sa : sb));
          atomicAdd(p.stats + ((long long)uu * 8 + cur_g) * 2 + 1, (double)(lane == 0 ? qa : qb));
        }
      } else if (in) {  // three or more utterances inside 32 rows (very short utterances): every lane for itself
        atomicAdd(p.stats + ((long long)utt * 8 + cur_g) * 2, (double)cs);
        atomicAdd(p.stats + ((long long)utt * 8 + cur_g) * 2 + 1, (double)css);
      }
    }
    cs = css = 0.f;
  };
  uint32_t ra[16], rb[16];
  bool released = false;
  if (half < UNITS) {
    ptx::tmem_ld16(taddr + half * 32, ra);
    ptx::tmem_ld16(taddr + half * 32 + 16, rb);
  }
#pragma unroll 1
  for (int u = half; u < UNITS; u += 2) {
    const int n = n0 + u * 32;
    ptx::tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(ra[i]), v[16 + i] = __uint_as_float(rb[i]);
    if (u + 2 < UNITS) {
      ptx::tmem_ld16(taddr + (u + 2) * 32, ra);
      ptx::tmem_ld16(taddr + (u + 2) * 32 + 16, rb);
    } else {
      release();
      released = true;
    }
    if (n >= p.N) continue;
    if (p.bias) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n + i));
        v[i] += b.x, v[i + 1] += b.y, v[i + 2] += b.z, v[i + 3] += b.w;
      }
    }
    if (m < p.M) {
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          __nv_bfloat162 hb = __floats2bfloat162_rn(v[16 * hlf + 2 * i], v[16 * hlf + 2 * i + 1]);
          w[i] = *reinterpret_cast<uint32_t*>(&hb);
        }
        stg256(orow + n + 16 * hlf, w);
      }
    }
    if (p.fused_stats) {
      const int g_lo = n / p.group_ch;
      const int bnd = (g_lo + 1) * p.group_ch - n;  // columns [0, bnd) of the unit belong to group g_lo
      if (g_lo != cur_g) {
        if (cur_g >= 0) flush();
        cur_g = g_lo;
      }
      if (bnd >= 32) {  // the whole unit lies in one group (warp-uniform): four independent chains
        float sa = 0.f, sb = 0.f, sc = 0.f, sd = 0.f, qa = 0.f, qb = 0.f, qc = 0.f, qd = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          sa += v[i], sb += v[i + 1], sc += v[i + 2], sd += v[i + 3];
          qa = fmaf(v[i], v[i], qa), qb = fmaf(v[i + 1], v[i + 1], qb), qc = fmaf(v[i + 2], v[i + 2], qc), qd = fmaf(v[i + 3], v[i + 3], qd);
        }
        if (in) cs += (sa + sb) + (sc + sd), css += (qa + qb) + (qc + qd);
      } else {
        float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f, s2 = 0.f, q2 = 0.f, s3 = 0.f, q3 = 0.f;  // even / odd columns x (below, above) the boundary
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float x = v[i], y = v[i + 1];
          if (i < bnd) s0 += x, q0 = fmaf(x, x, q0); else s1 += x, q1 = fmaf(x, x, q1);
          if (i + 1 < bnd) s2 += y, q2 = fmaf(y, y, q2); else s3 += y, q3 = fmaf(y, y, q3);
        }
        if (in) cs += s0 + s2, css += q0 + q2;
        flush();
        cur_g = g_lo + 1;
        if (in) cs = s1 + s3, css = q1 + q3;
      }
    }
  }
  if (cur_g >= 0) flush();
  if (!released) release();
}

// ------------------------------------------------------------------------------------------------
// tcgen05 implementation.
template <int BN, int NEW = 8>  // NEW = epilogue warps (8, or 12 for the SnakeBeta GEMM whose epilogue is the bottleneck)
struct TcCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int MAX_SMEM = 227 * 1024;
  static constexpr int CTRL_BYTES = 2048;  // mbarriers + TMEM slot (256 B), GroupNorm partial sums [2 sets][8 warps][8 groups][2] + counters
  static constexpr int N_EPI_WARPS = NEW;
  static constexpr int EPI_LD = 36;  // padded row (floats): 16-byte aligned rows, conflict-free 128-bit accesses
  static constexpr int EPI_WARP_BYTES = 5120;  // per-warp staging: 32 x 36 floats (transposing epilogue) or 1024-aligned TMA-store units
  static constexpr int EPI_BYTES = N_EPI_WARPS * EPI_WARP_BYTES;
  static constexpr int STAGES_RAW = (MAX_SMEM - 1024 - CTRL_BYTES - EPI_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + CTRL_BYTES + EPI_BYTES;
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int THREADS = 128 + 32 * N_EPI_WARPS;
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN must be a multiple of 32 in [32, 256]");
  static_assert(STAGES >= 3, "pipeline too shallow");
};

// mbarrier wait that optionally accumulates the cycles spent waiting (debug profile of CTA 0).
__device__ __forceinline__ void mbar_wait_prof(uint64_t* bar, uint32_t parity, bool prof, unsigned long long& acc) {
  if (!prof) {
    ptx::mbar_wait(bar, parity);
    return;
  }
  const long long t0 = clock64();
  ptx::mbar_wait(bar, parity);
  acc += (unsigned long long)(clock64() - t0);
}

template <int BN, int NEW = 8>
__global__ void __launch_bounds__(TcCfg<BN, NEW>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut, const GemmParams p) {
  using Cfg = TcCfg<BN, NEW>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Cluster of CL CTAs = CL consecutive m-tiles of one n-tile ("super-tile"); every CTA of a cluster walks the same
  // super-tile list in lock step (a CTA whose m-tile lies past M still runs: its loads are zero-filled, stores predicated).
  const int CL = p.cluster;
  const int crank = CL > 1 ? (int)ptx::cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << CL) - 1);
  const int m_tiles = (p.M + Cfg::BM - 1) / Cfg::BM;
  const int m_super = (m_tiles + CL - 1) / CL;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int n_tiles_total = m_super * n_tiles;  // super-tiles
  const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
  const int kb_per_tap = p.K / Cfg::BK;
  const int k_iters = p.n_taps * kb_per_tap;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA0);
    ptx::prefetch_tmap(&tmA1);
    ptx::prefetch_tmap(&tmW);
    if (p.tma_epi) ptx::prefetch_tmap(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], CL);  // one tcgen05.commit arrival from every CTA of the cluster
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], Cfg::N_EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  // GroupNorm partial sums: [2 accumulator sets][8 epilogue warps][8 groups][2] floats, then 2 arrival counters
  const uint32_t gn_slots = ptx::smem_u32(smem + STAGES * Cfg::STAGE_BYTES + 256);
  const uint32_t gn_counters = gn_slots + 2 * 8 * 8 * 2 * 4;
  if (warp == 3) {
    for (int i = lane; i < 2 * 8 * 8 * 2 + 2; i += 32) ptx::sts32(gn_slots + i * 4, 0.f);
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();  // peers' barriers are initialised before any multicast can reach them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();  // everything above overlapped the previous kernel's tail; global memory is touched only below

  if (warp == 0) {
    // ===================== TMA producer (whole warp in uniform control flow; elect.sync picks the issuing lane) =========
    {
      int stage = 0;
      uint32_t phase = 0;
      const bool prof = p.prof != nullptr && blockIdx.x == 0;
      unsigned long long w_empty = 0;
      const long long t_start = clock64();
      const int b_rows = BN / CL;  // this CTA's slice of the weight tile
      for (int tile = cluster_id; tile < n_tiles_total; tile += n_clusters) {
        const int m0 = ((tile / n_tiles) * CL + crank) * Cfg::BM, n0 = (tile % n_tiles) * BN;
        for (int t = 0; t < p.n_taps; ++t) {
          const GemmTap tap = p.taps[t];
          const CUtensorMap* tmA = tap.a_src ? &tmA1 : &tmA0;
          for (int kb = 0; kb < kb_per_tap; ++kb) {
            mbar_wait_prof(&empty_bar[stage], phase ^ 1, prof, w_empty);  // every CTA of the cluster released this stage
            ptx::mbar_expect_tx_elect(&full_bar[stage], Cfg::STAGE_BYTES);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            ptx::tma_load_2d_elect(sa, tmA, &full_bar[stage], tap.a_col + kb * Cfg::BK, m0 + tap.row_shift);
            if (CL == 1)
              ptx::tma_load_2d_elect(sa + Cfg::A_BYTES, &tmW, &full_bar[stage], kb * Cfg::BK, tap.w_row + n0);
            else
              ptx::tma_load_2d_mcast_elect(sa + Cfg::A_BYTES + crank * b_rows * 128, &tmW, &full_bar[stage], kb * Cfg::BK,
                                           tap.w_row + n0 + crank * b_rows, cmask);
            if (++stage == STAGES) stage = 0, phase ^= 1;
          }
        }
      }
      if (prof && lane == 0) p.prof[0] = (unsigned long long)(clock64() - t_start), p.prof[1] = w_empty;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp in uniform control flow; elect.sync picks the issuing lane) ===========
    {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(Cfg::BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      const bool prof = p.prof != nullptr && blockIdx.x == 0;
      unsigned long long w_full = 0, w_tempty = 0;
      const long long t_start = clock64();
      for (int tile = cluster_id; tile < n_tiles_total; tile += n_clusters, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait_prof(&tempty_bar[acc], acc_phase ^ 1, prof, w_tempty);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait_prof(&full_bar[stage], phase, prof, w_full);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < Cfg::BK / 16; ++k) {
            ptx::umma_bf16_elect(tmem_d, ptx::umma_desc_sw128(a_addr + k * 32), ptx::umma_desc_sw128(b_addr + k * 32), idesc,
                                 (it > 0 || k > 0) ? 1u : 0u);
          }
          if (CL == 1) ptx::umma_commit_elect(&empty_bar[stage]);
          else ptx::umma_commit_mcast_elect(&empty_bar[stage], cmask);
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
        ptx::umma_commit_elect(&tfull_bar[acc]);
      }
      if (prof && lane == 0) p.prof[2] = (unsigned long long)(clock64() - t_start), p.prof[3] = w_full, p.prof[4] = w_tempty, p.prof[8] = (unsigned long long)local;
      // Programmatic dependent launch: the next kernel may be scheduled once every CTA has issued its last MMA, so its launch
      // latency and prologue overlap this kernel's last epilogue instead of its whole run (an early trigger lets the
      // dependent CTAs sit on SMs that this grid's later tiles still need).
      ptx::pdl_launch_dependents();
    }
  } else if (warp >= 4) {
    // ===================== epilogue warps: TMEM -> registers -> smem transpose -> coalesced global I/O ============
    // tcgen05.ld hands each thread one accumulator ROW; touching global memory in that shape makes every warp
    // instruction hit 32 different cache lines, so each 32x32 block goes through a per-warp smem tile (epi_block).
    // The mode switch is hoisted out of all loops: each mode runs its own specialised copy of the loop.
    const uint32_t stg = ptx::smem_u32(smem + STAGES * Cfg::STAGE_BYTES + Cfg::CTRL_BYTES) + (warp - 4) * Cfg::EPI_WARP_BYTES;
    auto run = [&](auto mode_tag) {
      constexpr int MODE = decltype(mode_tag)::value;
      const int q = warp & 3;            // TMEM lane quarter this warp may touch
      const int half = (warp - 4) >> 2;  // the two warps of a quarter alternate over the 32-column blocks
      const int cg = lane & 7;
      (void)cg;
      int local = 0;
      const bool prof = p.prof != nullptr && blockIdx.x == 0 && warp == 4;
      unsigned long long w_tfull = 0;
      const long long t_start = clock64();
      for (int tile = cluster_id; tile < n_tiles_total; tile += n_clusters, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        const int m0 = ((tile / n_tiles) * CL + crank) * Cfg::BM + q * 32, n0 = (tile % n_tiles) * BN;
        int info_pre = 0;  // row flags requested before the accumulator wait: their latency hides under the MMAs of this tile
        if constexpr (MODE != EPI_STORE && MODE != EPI_SNAKE) info_pre = load_row_info(p, m0 + lane);
        mbar_wait_prof(&tfull_bar[acc], acc_phase, prof, w_tfull);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
        if constexpr (MODE == EPI_STATS16) {
          epilogue_tile_stats16<BN>(p, taddr, m0, n0, half, lane, info_pre, [&] {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
          });
          continue;
        }
        if constexpr (MODE == EPI_STORE || MODE == EPI_SNAKE || MODE == EPI_MASK) {
          if (NEW == 8 && p.direct_epi) {
            epilogue_tile_direct<BN, MODE>(p, taddr, m0, n0, half, lane, [&] {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
            }, info_pre);
            continue;
          }
        }
        if constexpr (MODE == EPI_STORE || MODE == EPI_SNAKE || MODE == EPI_MASK || MODE == EPI_RESID) {
          if (NEW == 8 && p.tma_epi) {
            epilogue_tile_tma<BN, MODE>(p, &tmOut, taddr, stg, m0, n0, half, lane, [&] {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
            });
            continue;
          }
        }
        epilogue_tile<BN, MODE, Cfg::EPI_LD, Cfg::N_EPI_WARPS>(p, taddr, stg, m0, n0, warp - 4, lane, acc, gn_slots, gn_counters, info_pre);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc]);
      }
      if (p.tma_epi) ptx::bulk_wait_all_elect();  // smem staging is read / global writes land before the CTA exits
      if (prof && lane == 0) p.prof[5] = (unsigned long long)(clock64() - t_start), p.prof[6] = w_tfull;
    };
    switch (p.mode) {
      case EPI_STORE: run(std::integral_constant<int, EPI_STORE>{}); break;
      case EPI_STATS: run(std::integral_constant<int, EPI_STATS>{}); break;
      case EPI_RESID: run(std::integral_constant<int, EPI_RESID>{}); break;
      case EPI_SNAKE: run(std::integral_constant<int, EPI_SNAKE>{}); break;
      case EPI_MASK: run(std::integral_constant<int, EPI_MASK>{}); break;
      case EPI_STATS16: run(std::integral_constant<int, EPI_STATS16>{}); break;
      default: run(std::integral_constant<int, EPI_ODE>{}); break;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL > 1) ptx::cluster_sync_all();  // no CTA exits while a peer may still multicast into / arrive on its smem
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): two CTAs on neighbouring SMs compute one 256 x BN tile.  Each CTA loads the
// A rows of its own 128-row half and HALF of the weight tile (BN/2 rows); the leader CTA's single MMA thread issues
// tcgen05.mma.cta_group::2 (M = 256), which feeds both tensor cores from both shared memories, so every weight byte is
// read from shared memory once per pair instead of once per CTA.  In the 1-CTA kernel TMA writes plus UMMA operand
// reads ask for ~210 B/clk of shared-memory bandwidth at full MMA rate (the SM has 128): measured 55-60 % tensor
// utilisation.  Here the demand is ~125-145 B/clk.
//   full barrier   : leader's, count 1 (+ tx bytes of BOTH CTAs' loads, routed there by the .cta_group::2 TMA form)
//   empty barrier  : one per CTA, count 1, signalled by the leader's tcgen05.commit multicast to both CTAs
//   tmem-full      : one per CTA, count 1, same multicast commit
//   tmem-empty     : leader's, count 16 = 8 epilogue warps x 2 CTAs (the follower's warps arrive remotely)
template <int BN>
struct Tc2Cfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / 2) * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int MAX_SMEM = 227 * 1024;
  static constexpr int CTRL_BYTES = 2048;
  static constexpr int N_EPI_WARPS = 8;
  static constexpr int EPI_LD = 36;
  static constexpr int EPI_WARP_BYTES = 5120;
  static constexpr int EPI_BYTES = N_EPI_WARPS * EPI_WARP_BYTES;
  static constexpr int STAGES_RAW = (MAX_SMEM - 1024 - CTRL_BYTES - EPI_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + CTRL_BYTES + EPI_BYTES;
  static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int THREADS = 128 + 32 * N_EPI_WARPS;
  static_assert(BN % 32 == 0 && BN >= 64 && BN <= 256, "BN must be a multiple of 32 in [64, 256]");
};

template <int BN>
__global__ void __launch_bounds__(Tc2Cfg<BN>::THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut, const GemmParams p) {
  using Cfg = Tc2Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();  // 0 = leader
  const int m_pairs = (p.M + 2 * Cfg::BM - 1) / (2 * Cfg::BM);
  const int n_tiles = (p.N + BN - 1) / BN;
  const int n_tiles_total = m_pairs * n_tiles;
  const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int kb_per_tap = p.K / Cfg::BK;
  const int k_iters = p.n_taps * kb_per_tap;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA0);
    ptx::prefetch_tmap(&tmA1);
    ptx::prefetch_tmap(&tmW);
    if (p.tma_epi) ptx::prefetch_tmap(&tmOut);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);
      ptx::mbar_init(&tempty_bar[i], 2 * Cfg::N_EPI_WARPS);
    }
    ptx::fence_mbar_init();
  }
  const uint32_t gn_slots = ptx::smem_u32(smem + STAGES * Cfg::STAGE_BYTES + 256);
  const uint32_t gn_counters = gn_slots + 2 * 8 * 8 * 2 * 4;
  if (warp == 3) {
    for (int i = lane; i < 2 * 8 * 8 * 2 + 2; i += 32) ptx::sts32(gn_slots + i * 4, 0.f);
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

  if (warp == 0) {
    // ===================== TMA producer (warp 0 of each CTA, uniform control flow; elect.sync picks the lane) ===========
    {
      int stage = 0;
      uint32_t phase = 0;
      const bool prof = p.prof != nullptr && blockIdx.x == 0;
      unsigned long long w_empty = 0;
      const long long t_start = clock64();
      for (int tile = pair_id; tile < n_tiles_total; tile += n_pairs) {
        const int m0 = ((tile / n_tiles) * 2 + rank) * Cfg::BM, n0 = (tile % n_tiles) * BN + rank * (BN / 2);
        for (int t = 0; t < p.n_taps; ++t) {
          const GemmTap tap = p.taps[t];
          const CUtensorMap* tmA = tap.a_src ? &tmA1 : &tmA0;
          for (int kb = 0; kb < kb_per_tap; ++kb) {
            mbar_wait_prof(&empty_bar[stage], phase ^ 1, prof, w_empty);
            if (rank == 0) ptx::mbar_expect_tx_elect(&full_bar[stage], 2 * Cfg::STAGE_BYTES);  // both CTAs' bytes land here
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            ptx::tma_load_2d_pair_elect(sa, tmA, &full_bar[stage], tap.a_col + kb * Cfg::BK, m0 + tap.row_shift);
            ptx::tma_load_2d_pair_elect(sa + Cfg::A_BYTES, &tmW, &full_bar[stage], kb * Cfg::BK, tap.w_row + n0);
            if (++stage == STAGES) stage = 0, phase ^= 1;
          }
        }
      }
      if (prof && lane == 0) p.prof[0] = (unsigned long long)(clock64() - t_start), p.prof[1] = w_empty;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp 1 of the leader CTA, uniform control flow; elect.sync picks the lane) ======
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(2 * Cfg::BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      const bool prof = p.prof != nullptr && blockIdx.x == 0;
      unsigned long long w_full = 0, w_tempty = 0;
      const long long t_start = clock64();
      for (int tile = pair_id; tile < n_tiles_total; tile += n_pairs, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait_prof(&tempty_bar[acc], acc_phase ^ 1, prof, w_tempty);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait_prof(&full_bar[stage], phase, prof, w_full);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < Cfg::BK / 16; ++k) {
            ptx::umma_bf16_pair_elect(tmem_d, ptx::umma_desc_sw128(a_addr + k * 32), ptx::umma_desc_sw128(b_addr + k * 32), idesc,
                                      (it > 0 || k > 0) ? 1u : 0u);
          }
          ptx::umma_commit_pair_elect(&empty_bar[stage], 3);  // frees the stage in both CTAs
          if (++stage == STAGES) stage = 0, phase ^= 1;
        }
        ptx::umma_commit_pair_elect(&tfull_bar[acc], 3);  // accumulator ready in both CTAs
      }
      if (prof && lane == 0) p.prof[2] = (unsigned long long)(clock64() - t_start), p.prof[3] = w_full, p.prof[4] = w_tempty, p.prof[8] = (unsigned long long)local;
    }
    ptx::pdl_launch_dependents();  // leader: after its last MMA; follower: nothing left to issue
  } else if (warp >= 4) {
    // ===================== epilogue warps (both CTAs, each on its own 128 rows) =====================
    const uint32_t stg = ptx::smem_u32(smem + STAGES * Cfg::STAGE_BYTES + Cfg::CTRL_BYTES) + (warp - 4) * Cfg::EPI_WARP_BYTES;
    auto run = [&](auto mode_tag) {
      constexpr int MODE = decltype(mode_tag)::value;
      const int q = warp & 3;
      const int half = (warp - 4) >> 2;
      const int cg = lane & 7;
      (void)cg;
      int local = 0;
      const bool prof = p.prof != nullptr && blockIdx.x == 0 && warp == 4;
      unsigned long long w_tfull = 0;
      const long long t_start = clock64();
      for (int tile = pair_id; tile < n_tiles_total; tile += n_pairs, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        const int m0 = ((tile / n_tiles) * 2 + rank) * Cfg::BM + q * 32, n0 = (tile % n_tiles) * BN;
        int info_pre = 0;  // row flags requested before the accumulator wait: their latency hides under the MMAs of this tile
        if constexpr (MODE != EPI_STORE && MODE != EPI_SNAKE) info_pre = load_row_info(p, m0 + lane);
        mbar_wait_prof(&tfull_bar[acc], acc_phase, prof, w_tfull);
        ptx::tc_fence_after();
        const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
        if constexpr (MODE == EPI_STATS16) {
          epilogue_tile_stats16<BN>(p, taddr, m0, n0, half, lane, info_pre, [&] {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (rank == 0) ptx::mbar_arrive(&tempty_bar[acc]);
              else ptx::mbar_arrive_remote_relaxed(&tempty_bar[acc], 0);
            }
          });
          continue;
        }
        if constexpr (MODE == EPI_STORE || MODE == EPI_SNAKE || MODE == EPI_MASK) {
          if (p.direct_epi) {
            epilogue_tile_direct<BN, MODE>(p, taddr, m0, n0, half, lane, [&] {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (rank == 0) ptx::mbar_arrive(&tempty_bar[acc]);
                else ptx::mbar_arrive_remote_relaxed(&tempty_bar[acc], 0);
              }
            }, info_pre);
            continue;
          }
        }
        if constexpr (MODE == EPI_STORE || MODE == EPI_SNAKE || MODE == EPI_MASK || MODE == EPI_RESID) {
          if (p.tma_epi) {
            epilogue_tile_tma<BN, MODE>(p, &tmOut, taddr, stg, m0, n0, half, lane, [&] {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if (rank == 0) ptx::mbar_arrive(&tempty_bar[acc]);
                else ptx::mbar_arrive_remote_relaxed(&tempty_bar[acc], 0);
              }
            });
            continue;
          }
        }
        epilogue_tile<BN, MODE, Cfg::EPI_LD, Cfg::N_EPI_WARPS>(p, taddr, stg, m0, n0, warp - 4, lane, acc, gn_slots, gn_counters, info_pre);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {  // the accumulator of BOTH CTAs must be drained before the leader's MMA thread reuses it
          if (rank == 0) ptx::mbar_arrive(&tempty_bar[acc]);
          else ptx::mbar_arrive_remote_relaxed(&tempty_bar[acc], 0);
        }
      }
      if (p.tma_epi) ptx::bulk_wait_all_elect();
      if (prof && lane == 0) p.prof[5] = (unsigned long long)(clock64() - t_start), p.prof[6] = w_tfull;
    };
    switch (p.mode) {
      case EPI_STORE: run(std::integral_constant<int, EPI_STORE>{}); break;
      case EPI_STATS: run(std::integral_constant<int, EPI_STATS>{}); break;
      case EPI_RESID: run(std::integral_constant<int, EPI_RESID>{}); break;
      case EPI_SNAKE: run(std::integral_constant<int, EPI_SNAKE>{}); break;
      case EPI_MASK: run(std::integral_constant<int, EPI_MASK>{}); break;
      case EPI_STATS16: run(std::integral_constant<int, EPI_STATS16>{}); break;
      default: run(std::integral_constant<int, EPI_ODE>{}); break;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();  // neither CTA leaves while its partner may still touch its smem / TMEM / barriers
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

// The heavy kernels are instantiated in their own translation units (csrc/gemm_inst.cu / attn_inst.cu, one nvcc process per
// instantiation, see build.py).  cfm.cu never names the kernel templates: it launches through the function pointers these
// getters return (cudaLaunchKernelExC), so its own compilation does not instantiate them again.
struct KernelInfo {
  const void* fn;
  int threads;
  int smem;
};
#define CFM_FOR_EACH_TC(X) X(64, 8) X(128, 8) X(160, 8) X(192, 8) X(256, 8) X(256, 12)
#define CFM_FOR_EACH_TC2(X) X(128) X(160) X(192) X(256)
#define CFM_X_TC(BN, NEW) KernelInfo kinfo_tc_##BN##_##NEW();
#define CFM_X_TC2(BN) KernelInfo kinfo_tc2_##BN();
CFM_FOR_EACH_TC(CFM_X_TC)
CFM_FOR_EACH_TC2(CFM_X_TC2)
#undef CFM_X_TC
#undef CFM_X_TC2
KernelInfo kinfo_attn_tc(int x3);
KernelInfo kinfo_attn_persist();

}  // namespace cfm
