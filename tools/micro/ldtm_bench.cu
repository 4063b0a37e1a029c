// Microbenchmark: latency of tcgen05.ld issued by warps on each SM sub-partition while ONE warp keeps the tensor pipe's queue
// full of MMAs.  Question: do TMEM loads of a warp that shares its sub-partition with the MMA-issuing warp queue behind the MMAs?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I matcha-tts-24k_b200/csrc -o tools/micro/bin/ldtm_bench tools/micro/ldtm_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace cfm;

__global__ void __launch_bounds__(256, 1) bench(int N, int n_mma, int issuer_warp, int throttle, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 160 * 1024);  // [0] done, [1..2] throttle
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  volatile int* stop = reinterpret_cast<volatile int*>(slot + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::mbar_init(bar + 1, 1); ptx::mbar_init(bar + 2, 1); ptx::fence_mbar_init(); *stop = 0; }
  if (warp == 7) { ptx::tmem_alloc(slot, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == issuer_warp) {
    const uint32_t idesc = ptx::umma_idesc_bf16(128, N);
    const uint64_t a0 = ptx::umma_desc_sw128(ptx::smem_u32(smem)), b0 = ptx::umma_desc_sw128(ptx::smem_u32(smem + 64 * 1024));
    int kb = 0;
    for (int i = 0; i < n_mma; i += 4, ++kb) {
      if (throttle && kb >= throttle) ptx::mbar_wait(bar + 1 + (kb % 2 == 0 ? 0 : 1) * 0, ((kb - throttle) & 1));  // k-block kb - throttle complete
      asm volatile("{\n\t.reg .pred p, t, e;\n\t.reg .b64 a1, b1, a2, b2, a3, b3;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, %4, %4;\n\t"
                   "add.u64 a1, %1, 2;\n\tadd.u64 b1, %2, 2;\n\tadd.u64 a2, %1, 4;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 a3, %1, 6;\n\tadd.u64 b3, %2, 6;\n\t"
                   "elect.sync _|e, 0xffffffff;\n\t"
                   "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t@e tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
                   "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t@e tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n\t}\n" ::"r"(tmem),
                   "l"(a0), "l"(b0), "r"(idesc), "r"((uint32_t)(i > 0))
                   : "memory");
      if (throttle) ptx::umma_commit_elect(bar + 1);  // one barrier, phase flips once per k-block
    }
    ptx::umma_commit_elect(bar);
    ptx::mbar_wait(bar, 0);
    if (lane == 0) *stop = 1;
  } else if (warp < 4 || (warp >= 4 && warp < 7 && issuer_warp >= 4)) {
    // measuring warps: repeated tcgen05.ld (32 columns of this warp's lane quarter, columns 384.. = not touched by the MMA)
    const uint32_t t = tmem + 384 + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    unsigned long long sum = 0, mx = 0;
    int n = 0;
    while (!*stop && n < 100000) {
      uint32_t r[16];
      const long long t0 = clock64();
      ptx::tmem_ld16(t, r);
      ptx::tmem_ld_wait();
      float acc0 = __uint_as_float(r[0]), acc1 = __uint_as_float(r[1]), acc2 = __uint_as_float(r[2]), acc3 = __uint_as_float(r[3]);
#pragma unroll
      for (int k = 0; k < 64; ++k) {  // 256 independent-ish FMAs: ~64+ issue cycles for one warp alone
        acc0 = fmaf(acc0, 1.0001f, 0.5f), acc1 = fmaf(acc1, 1.0002f, 0.25f), acc2 = fmaf(acc2, 0.9999f, 0.125f), acc3 = fmaf(acc3, 0.9998f, 1.f);
      }
      r[0] = __float_as_uint(acc0 + acc1 + acc2 + acc3);
      const long long dt = clock64() - t0;
      if (r[0] == 0xdeadbeef) sum += 1;
      sum += (unsigned long long)dt;
      mx = dt > (long long)mx ? (unsigned long long)dt : mx;
      ++n;
      __nanosleep(200);
    }
    if (lane == 0 && blockIdx.x == 0) out[warp * 4] = sum, out[warp * 4 + 1] = (unsigned long long)n, out[warp * 4 + 2] = mx;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 7) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
  unsigned long long* out;
  cudaMallocManaged(&out, 64 * 8);
  const int smem = 162 * 1024 + 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int throttle : {0, 1, 2, 4})
    for (int N : {64, 192})
      for (int issuer : {0, 1, 5}) {
        for (int i = 0; i < 64; ++i) out[i] = 0;
        bench<<<148, 256, smem>>>(N, 40000, issuer, throttle, out);
        cudaError_t e = cudaDeviceSynchronize();
        printf("N=%3d issuer warp %d (SMSP %d) throttle %d:", N, issuer, issuer % 4, throttle);
        for (int w = 0; w < 7; ++w)
          if (out[w * 4 + 1]) printf("  w%d(SMSP%d) avg %5.0f max %6llu", w, w % 4, (double)out[w * 4] / out[w * 4 + 1], out[w * 4 + 2]);
        printf("  %s\n", cudaGetErrorString(e));
      }
  return 0;
}
