// Microbenchmark: what slows the N = 64 SS pair MMA inside the fused feed-forward kernel (52 cycles alone, ~92-120 there)?
// One CTA pair per SM pair, warp 0 of the leader issues, per "chunk":  24 x MMA(SS, M 256, N 64)  [+ 8 x MMA(TS, N 192)]
// Variants (bit flags): 1 = walk 6 different 16 KB A blocks and 4 KB B stages (else reuse one block)
//                       2 = interleave the 8 TS N=192 MMAs after every 24 SS MMAs
//                       4 = 8 other warps per CTA keep doing tcgen05.ld / tcgen05.st on other TMEM columns
//                       8 = accumulators at columns 384 / 448 alternating (else column 0)
//                      16 = another warp streams global -> shared copies (cp.async.bulk 4 KB) into unused smem (TMA write traffic)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace cfm;

__global__ void __launch_bounds__(384, 1) bench(int flags, int n_chunks, const uint8_t* gsrc, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 200 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  volatile int* stop = reinterpret_cast<volatile int*>(slot + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)ptx::cluster_ctarank();
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::mbar_init(bar + 1, 1); ptx::mbar_init(bar + 2, 1 << 19); ptx::mbar_init(bar + 3, 1); ptx::fence_mbar_init(); *stop = 0; }
  if (warp == 2) { ptx::tmem_alloc_pair(slot, 512); ptx::tmem_relinquish_pair(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 0 && rank == 0) {
    const uint32_t idesc1 = ptx::umma_idesc_bf16(256, 64), idesc2 = ptx::umma_idesc_bf16(256, 192);
    const uint64_t a0 = ptx::umma_desc_sw128(ptx::smem_u32(smem)), b0 = ptx::umma_desc_sw128(ptx::smem_u32(smem + 96 * 1024));
    const uint64_t w0 = ptx::umma_desc_sw128(ptx::smem_u32(smem + 136 * 1024));
    const long long t0 = clock64();
    for (int c = 0; c < n_chunks; ++c) {
      const uint32_t d = (flags & 8) ? tmem + 384 + (c & 1) * 64 : tmem;
      for (int kb = 0; kb < 6; ++kb) {
        const uint64_t ao = (flags & 1) ? (uint64_t)(kb * 1024) : 0, bo = (flags & 1) ? (uint64_t)(((c * 6 + kb) % 10) * 256) : 0;
        if ((flags & 64) && (kb & 1) == 0) {  // a wait on an already completed phase + fence, as per pipeline stage
          ptx::mbar_wait(bar + 3, 1);
          ptx::tc_fence_after();
        }
        ptx::umma_bf16_pair_k64_elect(d, a0 + ao, b0 + bo, idesc1, kb > 0);
        if ((flags & 32) && (kb & 1)) ptx::umma_commit_pair_elect(bar + 2, 3);  // stage release: multicast commit per 8 MMAs
      }
      if (flags & 32) ptx::umma_commit_pair_elect(bar + 2, 3);
      if (flags & 2) {
        const uint32_t tp = (flags & 8) ? tmem + 384 + ((c + 1) & 1) * 64 : tmem + 448;
        ptx::umma_bf16_pair_ts_k64x2_elect((flags & 8) ? tmem : tmem + 64, (flags & 8) ? tmem + 192 : tmem + 256, tp, tp + 8, tp + 32, tp + 40, w0,
                                           w0 + 768, idesc2, c > 0);
      }
    }
    const long long t1 = clock64();
    ptx::umma_commit_pair_elect(bar, 1);
    ptx::mbar_wait(bar, 0);
    const long long t2 = clock64();
    if (lane == 0) *stop = 1;
    if (lane == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0), out[1] = (unsigned long long)(t2 - t0);
  } else if (warp >= 4 && (flags & 4)) {
    const uint32_t t = tmem + 320 + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    int n = 0;
    while (!*stop && n < 1000000) {
      uint32_t r[16];
      ptx::tmem_ld16(t, r);
      ptx::tmem_ld16(t + 16, r);
      ptx::tmem_ld_wait();
      ptx::tmem_st16(t + 32, r);
      ptx::tmem_st_wait();
      __nanosleep(500);
      ++n;
    }
  } else if (warp == 3 && (flags & 16)) {
    int n = 0;
    uint64_t* cbar = bar + 1;
    uint32_t ph = 0;
    while (!*stop && n < 1000000) {
      if (lane == 0) {
        ptx::mbar_expect_tx(cbar, 8192);
        for (int i = 0; i < 2; ++i)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ptx::smem_u32(smem + 184 * 1024 + i * 4096)),
                       "l"(gsrc + ((n * 2 + i) % 256) * 4096), "r"(4096), "r"(ptx::smem_u32(cbar))
                       : "memory");
      }
      ptx::mbar_wait(cbar, ph);
      ph ^= 1;
      ++n;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 2) { ptx::tc_fence_after(); ptx::tmem_dealloc_pair(tmem, 512); }
}

int main() {
  unsigned long long* out;
  cudaMallocManaged(&out, 64);
  uint8_t* gsrc;
  cudaMalloc(&gsrc, 1 << 20);
  cudaMemset(gsrc, 0, 1 << 20);
  const int smem = 202 * 1024 + 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n_chunks = 200;
  for (int flags : {3, 35, 67, 99, 127}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148), cfg.blockDim = dim3(384), cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    out[0] = out[1] = 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, bench, flags, n_chunks, (const uint8_t*)gsrc, out);
    cudaError_t e2 = cudaDeviceSynchronize();
    const double per_chunk = out[1] / (double)n_chunks;
    const double ideal = 24 * 32.0 + ((flags & 2) ? 8 * 96.0 : 0.0);
    printf("flags %2d [%s%s%s%s%s]: %.0f cycles per chunk (ideal %.0f)  => SS N=64 MMA ~ %.1f cycles  %s %s\n", flags, (flags & 1) ? "walk " : "", (flags & 2) ? "+TS192 " : "",
           (flags & 4) ? "ldtm/sttm " : "", (flags & 8) ? "cols384+ " : "", (flags & 16) ? "bulk-writes " : ((flags & 32) ? ((flags & 64) ? "commits waits " : "commits ") : ((flags & 64) ? "waits " : "")), per_chunk, ideal,
           (per_chunk - ((flags & 2) ? 768.0 : 0.0)) / 24.0, cudaGetErrorString(e), cudaGetErrorString(e2));
  }
  return 0;
}
