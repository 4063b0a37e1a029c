// The fused feed-forward kernel, compiled in its own nvcc process (see build.py); exports its launch information.
#define CFM_FF_KERNEL_TU 1
#include "ff_fused.cuh"

namespace cfm {
KernelInfo kinfo_ff_fused() {
  return KernelInfo{reinterpret_cast<const void*>(&ff_fused_kernel), FfCfg::THREADS, FfCfg::smem_bytes(FfCfg::MAX_C)};
}
}  // namespace cfm
