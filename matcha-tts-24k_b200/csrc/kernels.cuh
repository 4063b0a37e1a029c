// Bandwidth-side kernels of the estimator: pack / unpack between the caller's (B, F, T) tensors and the
// token-major packed layout, GroupNorm statistics + apply (+Mish, mask, time-embedding / residual add),
// LayerNorm, the time-embedding MLP, weight packing.  fp32 math everywhere; `T` is the activation storage
// type (bf16 in the tensor-core mode, float in the fp32 mode).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "gemm.cuh"

namespace cfm {

struct UttTable {        // one entry per utterance and resolution
  int start;             // first packed row
  int len;               // valid frames L
  int rows;              // rows owned in the packed buffer
  int bias_rows;         // max(P - 1, 0): padded frames whose conv output is exactly the bias
  int t_res;             // padded length T at this resolution (GroupNorm denominator)
  float pad_key_bias;    // log(P) - 1 for the virtual pad token's key, -inf when P == 0
  double inv_gn_count;   // 1 / (channels per GroupNorm group * t_res)
};

// (mean, rstd) of GroupNorm group g of utterance b from the fp64 sums: epilogue sums + bias_rows * (sum_c b_c, sum_c b_c^2)
// for the padded frames that are never materialised (DESIGN.md "pad-aware packing"), over group_ch * t_res elements
// (reference decoder.py:35-45 normalises over the PADDED length).  Evaluated by every consumer thread itself: 3 fp64
// operations, which is cheaper than a separate finalize launch between the conv and the apply pass.
__device__ __forceinline__ float2 gn_mean_rstd(const double* __restrict__ stats, const double* __restrict__ bias_gsum,
                                               const UttTable* __restrict__ utt, int b, int g) {
  const double2 st = *reinterpret_cast<const double2*>(stats + ((long long)b * 8 + g) * 2);
  const double2 bg = __ldg(reinterpret_cast<const double2*>(bias_gsum + g * 2));
  const double br = (double)__ldg(&utt[b].bias_rows);
  const double inv = __ldg(&utt[b].inv_gn_count);
  const double mean = fma(br, bg.x, st.x) * inv;
  const double var = fmax(fma(-mean, mean, fma(br, bg.y, st.y) * inv), 0.0);
  return make_float2((float)mean, rsqrtf((float)var + 1e-5f));
}

__device__ __forceinline__ float mish_f(float x) {
  // x * tanh(softplus(x)) = x * n / (n + 2),  n = e^x (e^x + 2)     (reference decoder.py:40 nn.Mish)
  if (x > 20.f) return x;
  float e = __expf(x);
  float n = e * (e + 2.f);
  return x * __fdividef(n, n + 2.f);
}
__device__ __forceinline__ float mish_precise(float x) {
  if (x > 20.f) return x;
  float e = expf(x);
  float n = e * (e + 2.f);
  return x * (n / (n + 2.f));
}

// 4 / 8 consecutive values of an fp32 or bf16 row (the conv output and the res_conv output are bf16 in the tensor-core mode)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void ld8(const float* p, float4& lo, float4& hi) { lo = ld4(p), hi = ld4(p + 4); }
__device__ __forceinline__ void ld8(const bf16* p, float4& lo, float4& hi) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.z)), d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.w));
  lo = make_float4(a.x, a.y, b.x, b.y), hi = make_float4(c.x, c.y, d.x, d.y);
}

// ---------------------------------------------------------------------------------- bf16 x 3 split (fp32 mode on the tensor pipe)
// a = hi + lo + r with hi = bf16(a), lo = bf16(a - hi), |r| <= 2^-18 |a|.  A product a * w is taken as hi*hi + lo*hi + hi*lo on
// the bf16 tensor pipe with fp32 accumulation (the dropped terms are ~2^-17 relative): the GEMM sees [hi | lo] as two K ranges of
// one wider operand and three (activation range, weight range) pairs per tap.
// src: [rows, ld] fp32 -> dst: [rows, 2 ld] bf16, hi in columns [0, ld), lo in [ld, 2 ld).  ld % 4 == 0.
static __global__ void split3_rows_kernel(const float* __restrict__ src, long long ld, long long rows, bf16* __restrict__ dst) {
  ptx::pdl_wait();
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one float4 per thread
  const long long per_row = ld >> 2;
  if (q >= rows * per_row) return;
  const long long r = q / per_row, c = (q - r * per_row) << 2;
  const float4 v = *reinterpret_cast<const float4*>(src + r * ld + c);
  ptx::pdl_launch_dependents();
  const float a[4] = {v.x, v.y, v.z, v.w};
  __nv_bfloat16 hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    hi[i] = __float2bfloat16_rn(a[i]);
    lo[i] = __float2bfloat16_rn(a[i] - __bfloat162float(hi[i]));
  }
  bf16* d = dst + r * 2 * ld + c;
  *reinterpret_cast<uint2*>(d) = *reinterpret_cast<const uint2*>(hi);
  *reinterpret_cast<uint2*>(d + ld) = *reinterpret_cast<const uint2*>(lo);
}
// Weights: [rows, ld] fp32 -> [2 rows, ld] bf16, hi in rows [0, rows), lo in rows [rows, 2 rows).
static __global__ void split3_weight_kernel(const float* __restrict__ src, long long n, bf16* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = src[i];
  const __nv_bfloat16 hi = __float2bfloat16_rn(a);
  dst[i] = hi;
  dst[n + i] = __float2bfloat16_rn(a - __bfloat162float(hi));
}

// ---------------------------------------------------------------------------------- pack / unpack
// (B, F, T) fp32 channels-first  ->  rows [start_b + t], columns [col0, col0 + F) of a token-major buffer.
// Rows t >= L of the utterance's segment are written as zero.  Optionally also writes the fp32 state.
template <typename T>
__global__ void pack_rows_kernel(const float* __restrict__ src, int F, int Tpad, const UttTable* __restrict__ utt,
                                 T* __restrict__ dst_act, long long ld_act, int col0, float* __restrict__ dst_f32,
                                 long long ld_f32, float scale) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const UttTable u = utt[b];
  const int t0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  if (t0 >= u.rows) return;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    int f = f0 + i, t = t0 + tx;
    float v = 0.f;
    if (f < F && t < u.len) v = src[((long long)b * F + f) * Tpad + t] * scale;
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int t = t0 + i, f = f0 + tx;
    if (t < u.rows && f < F) {
      float v = tile[tx][i];
      long long row = u.start + t;
      if (dst_act) ActIO<T>::st(dst_act + row * ld_act + col0 + f, v);
      if (dst_f32) dst_f32[row * ld_f32 + f] = v;
    }
  }
}

// token-major fp32 state -> (B, F, T); frames t >= L take `fill` (the injected noise z for a solve, because the
// masked velocity never moves padded frames; nullptr -> 0 for a bare estimator call).
static __global__ void unpack_rows_kernel(const float* __restrict__ state, long long ld, const UttTable* __restrict__ utt, int F,
                                   int Tpad, const float* __restrict__ fill, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const UttTable u = utt[b];
  const int t0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    int t = t0 + i, f = f0 + tx;
    float v = 0.f;
    if (t < u.len && f < F) v = state[(long long)(u.start + t) * ld + f];
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    int f = f0 + i, t = t0 + tx;
    if (f < F && t < Tpad) {
      long long o = ((long long)b * F + f) * Tpad + t;
      dst[o] = (t < u.len) ? tile[tx][i] : (fill ? fill[o] : 0.f);
    }
  }
}

// Upstream-style speaker conditioning: spks (B, S) broadcast over the valid frames of each utterance into columns
// [col0, col0 + S) of the token-major estimator input (rows t >= L stay zero, as x * mask does in the reference).
template <typename T>
__global__ void pack_speaker_kernel(const float* __restrict__ spks, int S, const UttTable* __restrict__ utt,
                                    T* __restrict__ dst, long long ld, int col0) {
  const int b = blockIdx.y;
  const UttTable u = utt[b];
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = idx / S, s_ = idx % S;
  if (t >= u.rows) return;
  ActIO<T>::st(dst + (long long)(u.start + t) * ld + col0 + s_, t < u.len ? spks[(long long)b * S + s_] : 0.f);
}

// ---------------------------------------------------------------------------------- GroupNorm
// Stand-alone statistics pass (fp32 mode and the debug path; the tensor-core GEMM fuses this into its epilogue).
template <typename HT>
__global__ void gn_stats_kernel(const HT* __restrict__ h, long long ld, int M, int C, int group_ch,
                                const int* __restrict__ row_info, double* __restrict__ stats) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (row, group)
  const int m = idx >> 3, g = idx & 7;
  if (m >= M) return;
  const int info = row_info[m];
  if (!(info & ROW_INSTAT)) return;
  const HT* p = h + (long long)m * ld + g * group_ch;
  float s = 0.f, ss = 0.f;
  for (int i = 0; i < group_ch; ++i) {
    float v = ActIO<HT>::ld(p + i);
    s += v;
    ss = fmaf(v, v, ss);
  }
  const int utt = info & ROW_UTT_MASK;
  atomicAdd(stats + ((long long)utt * 8 + g) * 2, (double)s);
  atomicAdd(stats + ((long long)utt * 8 + g) * 2 + 1, (double)ss);
}

// GroupNorm statistics -> (mean, rstd) per (utterance, group), in double from the fp64 sums.
// sums = epilogue sums + bias_rows * (sum_c b_c, sum_c b_c^2) for the padded frames that are never materialised
// (DESIGN.md "pad-aware packing"), over group_ch * t_res elements (reference decoder.py:35-45 normalises over the
// PADDED length).
static __global__ void gn_finalize_kernel(const double* __restrict__ stats, const double* __restrict__ bias_gsum,
                                   const UttTable* __restrict__ utt, int n_utt, int group_ch, float2* __restrict__ mr) {
  ptx::pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_utt * 8) return;
  const int b = i >> 3, g = i & 7;
  const UttTable u = utt[b];
  const double n = (double)group_ch * (double)u.t_res;
  const double s = stats[(long long)i * 2] + (double)u.bias_rows * bias_gsum[g * 2];
  const double ss = stats[(long long)i * 2 + 1] + (double)u.bias_rows * bias_gsum[g * 2 + 1];
  const double mean = s / n;
  const double var = fmax(ss / n - mean * mean, 0.0);
  mr[i] = make_float2((float)mean, (float)(1.0 / sqrt(var + 1e-5)));
}

// y = valid ? Mish(GN(h)) + addvec[c] : 0;  y += resid[m, c];  -> out_f32 and/or out_act.   8 channels per thread.
template <typename T, bool PRECISE, typename HT = float>  // HT: storage type of the conv output h and of resid
__global__ void gn_apply_kernel(const HT* __restrict__ h, long long ld_h, int M, int C, int group_ch,
                                const int* __restrict__ row_info, const double* __restrict__ stats,
                                const double* __restrict__ bias_gsum, const UttTable* __restrict__ utt,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ addvec, long long addvec_utt_stride, const HT* __restrict__ resid,
                                long long ld_resid, float* __restrict__ out_f32, long long ld_f32, T* __restrict__ out_act,
                                long long ld_act) {
  // addvec: time-embedding projection added after Mish (reference decoder.py:60); one vector for the batch (stride 0) or one
  // per utterance (stride > 0: per-sample t of the training forward, reference flow_matching.py:84-97)
  ptx::pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c8 = C >> 3;
  const int m = (int)(idx / c8);
  if (m >= M) return;
  const int c = (int)(idx % c8) * 8;
  // every load that does not depend on the row flags is issued up front: one memory round trip instead of three
  const int info = __ldg(row_info + m);
  float4 h0, h1;
  ld8(h + (long long)m * ld_h + c, h0, h1);
  float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
  if (resid) ld8(resid + (long long)m * ld_resid + c, r0, r1);
  const bool valid = (info & ROW_VALID) != 0;
  float y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) y[i] = 0.f;
  if (valid) {
    const int b = info & ROW_UTT_MASK;
    const float2 st = gn_mean_rstd(stats, bias_gsum, utt, b, c / group_ch);  // group_ch % 8 == 0: one group per thread
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
    const float x[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
    const float ga[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float be[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float z = fmaf((x[i] - st.x) * st.y, ga[i], be[i]);
      y[i] = PRECISE ? mish_precise(z) : mish_f(z);
    }
    if (addvec) {
      const float* av = addvec + (long long)b * addvec_utt_stride;
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(av + c)), a1 = __ldg(reinterpret_cast<const float4*>(av + c + 4));
      y[0] += a0.x, y[1] += a0.y, y[2] += a0.z, y[3] += a0.w, y[4] += a1.x, y[5] += a1.y, y[6] += a1.z, y[7] += a1.w;
    }
  }
  if (resid) {
    y[0] += r0.x, y[1] += r0.y, y[2] += r0.z, y[3] += r0.w, y[4] += r1.x, y[5] += r1.y, y[6] += r1.z, y[7] += r1.w;
  }
  ptx::pdl_launch_dependents();  // inputs consumed: the dependent kernel's launch overlaps the stores of the last wave
  if (out_f32) {
    *reinterpret_cast<float4*>(out_f32 + (long long)m * ld_f32 + c) = make_float4(y[0], y[1], y[2], y[3]);
    *reinterpret_cast<float4*>(out_f32 + (long long)m * ld_f32 + c + 4) = make_float4(y[4], y[5], y[6], y[7]);
  }
  if (out_act) {
    T* d = out_act + (long long)m * ld_act + c;
    if constexpr (sizeof(T) == 2) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(y[0], y[1]), p1 = __floats2bfloat162_rn(y[2], y[3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(y[4], y[5]), p3 = __floats2bfloat162_rn(y[6], y[7]);
      uint4 o;
      o.x = *reinterpret_cast<uint32_t*>(&p0), o.y = *reinterpret_cast<uint32_t*>(&p1);
      o.z = *reinterpret_cast<uint32_t*>(&p2), o.w = *reinterpret_cast<uint32_t*>(&p3);
      *reinterpret_cast<uint4*>(d) = o;
    } else {
      *reinterpret_cast<float4*>(d) = make_float4(y[0], y[1], y[2], y[3]);
      *reinterpret_cast<float4*>(d + 4) = make_float4(y[4], y[5], y[6], y[7]);
    }
  }
}

// Last GroupNorm of a ResnetBlock1D fused with the first LayerNorm of the transformer stack that follows it
// (reference decoder.py:62 -> transformer.py:249): one warp per row, NCH*128 channels.
//   x  = (valid ? Mish(GN(h)) : 0) + resid     -> fp32 residual stream
//   xn = LayerNorm(x) * ln_gamma + ln_beta      -> activation type (next GEMM's A operand)
// Saves one pass over the fp32 stream and one launch per stage.
template <typename T, bool PRECISE, int NCH, typename HT = float>
__global__ void gn_apply_ln_kernel(const HT* __restrict__ h, long long ld_h, int M, int group_ch,
                                   const int* __restrict__ row_info, const double* __restrict__ stats,
                                   const double* __restrict__ bias_gsum, const UttTable* __restrict__ utt,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const HT* __restrict__ resid, long long ld_resid, float* __restrict__ out_f32,
                                   long long ld_f32, const float* __restrict__ ln_gamma, const float* __restrict__ ln_beta,
                                   T* __restrict__ out_ln, long long ld_ln) {
  constexpr int C = NCH * 128;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  ptx::pdl_wait();
  if (row >= M) return;
  const int info = __ldg(row_info + row);
  const bool valid = (info & ROW_VALID) != 0;
  const int b = info & ROW_UTT_MASK;
  float y[NCH][4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 r = ld4(resid + (long long)row * ld_resid + c);
    y[i][0] = r.x, y[i][1] = r.y, y[i][2] = r.z, y[i][3] = r.w;
    const float4 hv = ld4(h + (long long)row * ld_h + c);  // unconditional: overlaps the flag load
    if (valid) {
      const float2 st = gn_mean_rstd(stats, bias_gsum, utt, b, c / group_ch);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c)), be = __ldg(reinterpret_cast<const float4*>(beta + c));
      const float x[4] = {hv.x, hv.y, hv.z, hv.w}, ga[4] = {g.x, g.y, g.z, g.w}, bb[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float z = fmaf((x[j] - st.x) * st.y, ga[j], bb[j]);
        y[i][j] += PRECISE ? mish_precise(z) : mish_f(z);
      }
    }
    *reinterpret_cast<float4*>(out_f32 + (long long)row * ld_f32 + c) = make_float4(y[i][0], y[i][1], y[i][2], y[i][3]);
    s += (y[i][0] + y[i][1]) + (y[i][2] + y[i][3]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float d = y[i][j] - mean;
      ss = fmaf(d, d, ss);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / (float)C + 1e-5f);
  ptx::pdl_launch_dependents();
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(ln_gamma + c)), be = __ldg(reinterpret_cast<const float4*>(ln_beta + c));
    const float o0 = fmaf((y[i][0] - mean) * rstd, g.x, be.x), o1 = fmaf((y[i][1] - mean) * rstd, g.y, be.y);
    const float o2 = fmaf((y[i][2] - mean) * rstd, g.z, be.z), o3 = fmaf((y[i][3] - mean) * rstd, g.w, be.w);
    T* d = out_ln + (long long)row * ld_ln + c;
    if constexpr (sizeof(T) == 2) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(o0, o1), p1 = __floats2bfloat162_rn(o2, o3);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&p0), u.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(d) = u;
    } else {
      *reinterpret_cast<float4*>(d) = make_float4(o0, o1, o2, o3);
    }
  }
}

// per-GroupNorm-site sums of the conv bias per group: [8][2] doubles (sum b, sum b^2)
static __global__ void bias_group_sums_kernel(const float* __restrict__ bias, int C, int group_ch, double* __restrict__ out) {
  const int g = threadIdx.x;
  if (g >= 8) return;
  double s = 0, ss = 0;
  for (int i = 0; i < group_ch; ++i) {
    double b = bias[g * group_ch + i];
    s += b;
    ss += b * b;
  }
  out[g * 2] = s;
  out[g * 2 + 1] = ss;
}

// ---------------------------------------------------------------------------------- LayerNorm
// One warp per row, C <= 32 * MAXV.  eps = 1e-5, biased variance (torch.nn.LayerNorm; reference transformer.py:179,215).
template <typename T, int MAXV>
__global__ void layernorm_kernel(const float* __restrict__ x, long long ldx, int M, int C, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, T* __restrict__ out, long long ldo) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  ptx::pdl_wait();
  if (warp >= M) return;
  const float* xr = x + (long long)warp * ldx;
  float v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = lane + 32 * i;
    v[i] = (c < C) ? xr[c] : 0.f;
    s += v[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = lane + 32 * i;
    float d = (c < C) ? v[i] - mean : 0.f;
    ss = fmaf(d, d, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / (float)C + 1e-5f);
  ptx::pdl_launch_dependents();
  T* orow = out + (long long)warp * ldo;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    int c = lane + 32 * i;
    if (c < C) ActIO<T>::st(orow + c, fmaf((v[i] - mean) * rstd, __ldg(gamma + c), __ldg(beta + c)));
  }
}

// Same for C = NCH * 128: a lane owns 4 consecutive channels of every 128-channel chunk, so the row is read with 16-byte
// loads and written with 8-byte (bf16) / 16-byte (fp32) stores - 3 + 3 memory instructions per lane for C = 384 instead
// of 12 + 12 scalar ones.
template <typename T, int NCH>
__global__ void layernorm_vec_kernel(const float* __restrict__ x, long long ldx, int M, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, T* __restrict__ out, long long ldo) {
  constexpr int C = NCH * 128;
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  ptx::pdl_wait();
  if (row >= M) return;
  float4 v[NCH];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    v[i] = *reinterpret_cast<const float4*>(x + (long long)row * ldx + i * 128 + lane * 4);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
    ss = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, ss))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / (float)C + 1e-5f);
  ptx::pdl_launch_dependents();
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = i * 128 + lane * 4;
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c)), be = __ldg(reinterpret_cast<const float4*>(beta + c));
    const float o0 = fmaf((v[i].x - mean) * rstd, g.x, be.x), o1 = fmaf((v[i].y - mean) * rstd, g.y, be.y);
    const float o2 = fmaf((v[i].z - mean) * rstd, g.z, be.z), o3 = fmaf((v[i].w - mean) * rstd, g.w, be.w);
    T* d = out + (long long)row * ldo + c;
    if constexpr (sizeof(T) == 2) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(o0, o1), p1 = __floats2bfloat162_rn(o2, o3);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&p0), u.y = *reinterpret_cast<uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(d) = u;
    } else {
      *reinterpret_cast<float4*>(d) = make_float4(o0, o1, o2, o3);
    }
  }
}

// ---------------------------------------------------------------------------------- time embedding
// Sinusoidal features (reference decoder.py:20-29) for NT time points: [NT][dim] = [sin | cos](1000 t f_k),
// f_k = exp(-k ln(1e4)/(half-1)).  The reference evaluates this in fp32; each fp32 rounding step is reproduced,
// with the transcendental functions taken in double and rounded once.
static __global__ void sinusoid_kernel(const float* __restrict__ t, int NT, int dim, float* __restrict__ out) {
  const int i = blockIdx.x, k = threadIdx.x;
  const int half = dim / 2;
  if (i >= NT || k >= half) return;
  const float step = (float)(-(log(10000.0) / (double)(half - 1)));
  const float prod = (float)k * step;
  const float freq = (float)exp((double)prod);
  const float a = 1000.f * t[i];
  const float arg = a * freq;
  out[(long long)i * dim + k] = (float)sin((double)arg);
  out[(long long)i * dim + half + k] = (float)cos((double)arg);
}

enum GemvAct : int { ACT_NONE = 0, ACT_SILU = 1, ACT_MISH = 2 };
// y[i][n] = act_out(bias[n] + sum_k W[n][k] * act_in(x[i][k]));  one warp per (i, n); fp32 weights.
static __global__ void gemv_rows_kernel(const float* __restrict__ x, long long ldx, int NT, const float* __restrict__ W, int N,
                                 int K, const float* __restrict__ bias, int act_in, int act_out, float* __restrict__ y,
                                 long long ldy) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= NT * N) return;
  const int i = gw / N, n = gw % N;
  const float* xr = x + (long long)i * ldx;
  const float* wr = W + (long long)n * K;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) {
    float xv = xr[k];
    if (act_in == ACT_MISH) xv = mish_precise(xv);
    s = fmaf(wr[k], xv, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    s += bias[n];
    if (act_out == ACT_SILU) s = s / (1.f + expf(-s));
    y[(long long)i * ldy + n] = s;
  }
}

// ---------------------------------------------------------------------------------- front / back of the decode
// Reference matcha/inference.py:146-172.  Front: integer phoneme durations -> alignment path (utils/model.py:24-40), encoder
// output expanded along the path (mu_x @ path, :155-162), averaged over frames 2t-1, 2t, 2t+1 (downsample, utils/model.py:57-68:
// avg_pool1d k3 s2 p1, zero padding, always / 3).  The path matrix is never materialised: frame t of the fine grid belongs to the
// phoneme i with cum[i-1] <= t < cum[i], found by binary search in the cumulative durations.
// Step 1: inclusive cumulative durations per utterance + fine length (sum, clamped to >= 1).  One warp per utterance.
static __global__ void front_cumsum_kernel(const float* __restrict__ dur, int B, int Tx, int* __restrict__ cum, int* __restrict__ fine_len) {
  const int b = blockIdx.x, lane = threadIdx.x;
  if (b >= B) return;
  int carry = 0;
  for (int i0 = 0; i0 < Tx; i0 += 32) {
    const int i = i0 + lane;
    int v = i < Tx ? (int)(long long)dur[(long long)b * Tx + i] : 0;  // .long() of the reference: truncation of an integral float
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (i < Tx) cum[(long long)b * Tx + i] = carry + v;
    carry += __shfl_sync(0xffffffffu, v, 31);
  }
  if (lane == 0) fine_len[b] = max(carry, 1);
}
__device__ __forceinline__ int front_phoneme_of(const int* __restrict__ cum, int Tx, int t) {  // first i with cum[i] > t, or -1
  int lo = 0, hi = Tx;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(cum + mid) > t) hi = mid; else lo = mid + 1;
  }
  return lo < Tx ? lo : -1;
}
// Step 2: mu_y[b, f, t] = (fine[2t-1] + fine[2t] + fine[2t+1]) / 3, fine[u] = mu_x[b, f, phoneme(u)] for u < fine length, else 0;
// also writes the float mask (t < ceil(fine_len / 2)) the decoder takes.  One thread per (b, t), loop over the mel bins.
static __global__ void front_expand_kernel(const float* __restrict__ mu_x, const int* __restrict__ cum, const int* __restrict__ fine_len,
                                           int B, int F, int Tx, int T, float* __restrict__ mu_y, float* __restrict__ y_mask) {
  const int b = blockIdx.y, t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int* c = cum + (long long)b * Tx;
  const int fl = min(fine_len[b], __ldg(c + Tx - 1) > 0 ? __ldg(c + Tx - 1) : 0);  // frames the path covers (0 for an all-zero row)
  int idx[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int u = 2 * t - 1 + k;
    idx[k] = (u >= 0 && u < fl) ? front_phoneme_of(c, Tx, u) : -1;
  }
  const float* src = mu_x + (long long)b * F * Tx;
  float* dst = mu_y + (long long)b * F * T + t;
  for (int f = 0; f < F; ++f) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (idx[k] >= 0) a += __ldg(src + (long long)f * Tx + idx[k]);
    dst[(long long)f * T] = a / 3.f;
  }
  if (y_mask) y_mask[(long long)b * T + t] = t < max((fine_len[b] + 1) / 2, 1) ? 1.f : 0.f;
}
// Back: mel = x * std + mean on the first t_out frames of every row (denormalize + the [:, :, :y_max_length] slice, :170-172).
static __global__ void back_denormalize_kernel(const float* __restrict__ x, long long rows, int T, int t_out, float mean, float stdv,
                                               float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * t_out) return;
  const long long r = i / t_out;
  const int t = (int)(i % t_out);
  out[i] = fmaf(x[r * T + t], stdv, mean);
}

// ---------------------------------------------------------------------------------- weight packing
// dst[(tap * n_stride + n) * ldd + k] = src[n * s_n + k * s_k + tap_k[tap] * s_t], zero for k in [K, ldd).
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ src, long long s_n, long long s_k, long long s_t, int n_taps,
                                   int4 tap_k, int N, int K, T* __restrict__ dst, long long ldd, int n_stride) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long per_tap = (long long)N * ldd;
  if (idx >= per_tap * n_taps) return;
  const int tap = (int)(idx / per_tap);
  const long long rem = idx % per_tap;
  const int n = (int)(rem / ldd), k = (int)(rem % ldd);
  const int tk = tap == 0 ? tap_k.x : tap == 1 ? tap_k.y : tap == 2 ? tap_k.z : tap_k.w;
  float v = (k < K) ? src[n * s_n + k * s_k + tk * s_t] : 0.f;
  ActIO<T>::st(dst + ((long long)tap * n_stride + n) * ldd + k, v);
}

static __global__ void snake_consts_kernel(const float* __restrict__ alpha, const float* __restrict__ beta, int N,
                                    float* __restrict__ ea, float* __restrict__ ib) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  ea[i] = expf(alpha[i]);
  ib[i] = 1.0f / (expf(beta[i]) + 1e-9f);  // reference transformer.py:59,75
}

}  // namespace cfm
