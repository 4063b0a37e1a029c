// Self-attention over the mel frames of one utterance (reference transformer.py:253-258 -> diffusers
// Attention / AttnProcessor2_0): softmax(Q K^T * d^-1/2 + key_bias) V, non-causal, per head.
// Packed formulation: keys are the utterance's L valid rows (bias 0) plus ONE virtual pad token at row L whose
// bias log(P) - 1 stands for the P identical padded frames the reference attends with additive mask 0 versus
// +1 for valid keys (DESIGN.md "pad-aware packing"); -inf (no pad key) when P == 0.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>

#include "gemm.cuh"
#include "kernels.cuh"

namespace cfm {

// fp32-FMA flash-style kernel: one thread per query, keys staged through shared memory 32 at a time.
// Used by the fp32 precision mode and as the on-device cross-check of the tensor-core kernel.
template <typename T, int D>
__global__ void __launch_bounds__(128) attn_simt_kernel(const T* __restrict__ qkv, long long ld, int inner,
                                                         const UttTable* __restrict__ utt, const int4* __restrict__ work,
                                                         T* __restrict__ out, long long ldo, float scale) {
  __shared__ float Ks[32][D];
  __shared__ float Vs[32][D];
  __shared__ float kbias[32];
  const int4 w = work[blockIdx.x];
  const UttTable u = utt[w.x];
  const int head = w.y, q0 = w.z;
  const int nk = u.len + 1;
  const int qi = q0 + threadIdx.x;
  const bool qok = qi < nk;
  const long long base = (long long)u.start;
  float q[D], o[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    q[d] = qok ? ActIO<T>::ld(qkv + (base + qi) * ld + head * D + d) * scale : 0.f;
    o[d] = 0.f;
  }
  float mrun = -INFINITY, lrun = 0.f;
  for (int k0 = 0; k0 < nk; k0 += 32) {
    for (int i = threadIdx.x; i < 32 * D; i += 128) {
      int j = i / D, d = i % D;
      int key = k0 + j;
      float kv = 0.f, vv = 0.f;
      if (key < nk) {
        kv = ActIO<T>::ld(qkv + (base + key) * ld + inner + head * D + d);
        vv = ActIO<T>::ld(qkv + (base + key) * ld + 2 * inner + head * D + d);
      }
      Ks[j][d] = kv;
      Vs[j][d] = vv;
    }
    if (threadIdx.x < 32) {
      int key = k0 + threadIdx.x;
      kbias[threadIdx.x] = key < u.len ? 0.f : (key == u.len ? u.pad_key_bias : -INFINITY);
    }
    __syncthreads();
    float s[32];
    float mt = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) a = fmaf(q[d], Ks[j][d], a);
      s[j] = a + kbias[j];
      mt = fmaxf(mt, s[j]);
    }
    const float mnew = fmaxf(mrun, mt);  // finite: key 0 of the first tile is always a valid frame
    const float corr = expf(mrun - mnew);
    lrun *= corr;
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] *= corr;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float pj = expf(s[j] - mnew);
      lrun += pj;
#pragma unroll
      for (int d = 0; d < D; ++d) o[d] = fmaf(pj, Vs[j][d], o[d]);
    }
    mrun = mnew;
    __syncthreads();
  }
  if (qok) {
    const float inv = 1.f / lrun;
#pragma unroll
    for (int d = 0; d < D; ++d) ActIO<T>::st(out + (base + qi) * ldo + head * D + d, o[d] * inv);
  }
}

// Instantiated in attn_simt_inst.cu (one object per variant: the unrolled head_dim loops are slow to compile); cfm.cu launches
// through these getters and never names the template.
KernelInfo kinfo_attn_simt_bf16_64();
KernelInfo kinfo_attn_simt_bf16_32();
KernelInfo kinfo_attn_simt_f32_64();
KernelInfo kinfo_attn_simt_f32_32();

}  // namespace cfm
