// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16 operands, K = 16) as a function of N, cta_group and A source.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I matcha-tts-24k_b200/csrc -o gpurun_out/mma_bench tools/micro/mma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace cfm;

__device__ __forceinline__ void mma1(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
               "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma1_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
               "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma2_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
               "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode: 0 = SS, 1 = TS.  PAIR: cta_group::2 (M = 256), else M = 128.  unroll: MMAs per asm-free loop iteration.
template <bool PAIR>
__global__ void __launch_bounds__(128, 1) bench(int N, int mode, int n_mma, int n_addr, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 160 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = PAIR ? (int)ptx::cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0 && lane == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  if (warp == 1) { if (PAIR) { ptx::tmem_alloc_pair(slot, 512); ptx::tmem_relinquish_pair(); } else { ptx::tmem_alloc(slot, 512); ptx::tmem_relinquish(); } }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR) ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc = ptx::umma_idesc_bf16(PAIR ? 256 : 128, N);
    const uint64_t a0 = ptx::umma_desc_sw128(ptx::smem_u32(smem)), b0 = ptx::umma_desc_sw128(ptx::smem_u32(smem + 64 * 1024));
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 4) {  // four K = 16 steps per asm statement (descriptor + 2 per step), like the real kernels
      if (mode == 0) {
        if (PAIR) ptx::umma_bf16_pair_k64_elect(tmem, a0, b0, idesc, i > 0);
        else {
          asm volatile("{\n\t.reg .pred p, t, e;\n\t.reg .b64 a1, b1, a2, b2, a3, b3;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.eq.b32 t, %4, %4;\n\t"
                       "add.u64 a1, %1, 2;\n\tadd.u64 b1, %2, 2;\n\tadd.u64 a2, %1, 4;\n\tadd.u64 b2, %2, 4;\n\tadd.u64 a3, %1, 6;\n\tadd.u64 b3, %2, 6;\n\t"
                       "elect.sync _|e, 0xffffffff;\n\t"
                       "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t@e tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, t;\n\t"
                       "@e tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, t;\n\t@e tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, t;\n\t}\n" ::"r"(tmem),
                       "l"(a0), "l"(b0), "r"(idesc), "r"((uint32_t)(i > 0))
                       : "memory");
        }
      } else {
        if (PAIR) ptx::umma_bf16_pair_ts_k64x2_elect(tmem, tmem + 0, tmem + 256, tmem + 264, tmem + 272, tmem + 280, b0, b0, idesc, i > 0), i += 4;
        else {
          mma1_ts(tmem, tmem + 256, b0, idesc, i > 0);
          mma1_ts(tmem, tmem + 264, b0 + 2, idesc, 1);
          mma1_ts(tmem, tmem + 272, b0 + 4, idesc, 1);
          mma1_ts(tmem, tmem + 280, b0 + 6, idesc, 1);
        }
      }
    }
    const long long t1 = clock64();
    if (PAIR) ptx::umma_commit_pair_elect(bar, 1); else ptx::umma_commit_elect(bar);
    ptx::mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0), out[1] = (unsigned long long)(t2 - t0);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (PAIR) ptx::cluster_sync_all();
  if (warp == 1) { ptx::tc_fence_after(); if (PAIR) ptx::tmem_dealloc_pair(tmem, 512); else ptx::tmem_dealloc(tmem, 512); }
}

int main() {
  unsigned long long* out;
  cudaMallocManaged(&out, 64);
  const int smem = 162 * 1024 + 1024;
  cudaFuncSetAttribute(bench<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n_mma = 2000;
  for (int pair = 0; pair < 2; ++pair)
    for (int mode = 0; mode < 2; ++mode)
      for (int N : {32, 64, 96, 128, 192, 256}) {
        if (pair && N % 32) continue;
        for (int grid_sms : {1, 148}) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(pair ? (grid_sms == 1 ? 2 : 148) : grid_sms), cfg.blockDim = dim3(128), cfg.dynamicSmemBytes = smem;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeClusterDimension;
          at[0].val.clusterDim.x = pair ? 2 : 1, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
          cfg.attrs = at, cfg.numAttrs = 1;
          out[0] = out[1] = 0;
          cudaError_t e = pair ? cudaLaunchKernelEx(&cfg, bench<true>, N, mode, n_mma, 4, out) : cudaLaunchKernelEx(&cfg, bench<false>, N, mode, n_mma, 4, out);
          cudaError_t e2 = cudaDeviceSynchronize();
          const double ideal = (pair ? 256.0 : 128.0) * N * 16 / (pair ? 8192.0 : 4096.0);
          printf("%s %s N=%3d grid=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (ideal %.0f)  %s %s\n", pair ? "pair M=256" : "1cta M=128", mode ? "TS" : "SS", N,
                 (int)cfg.gridDim.x, out[0] / (double)n_mma, out[1] / (double)n_mma, ideal, cudaGetErrorString(e), cudaGetErrorString(e2));
        }
      }
  return 0;
}
