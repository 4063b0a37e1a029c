// The Linear + residual + LayerNorm kernel, compiled in its own nvcc process (see build.py); exports its launch information.
#define CFM_ROWLN_KERNEL_TU 1
#include "rowln.cuh"

namespace cfm {
KernelInfo kinfo_rowln() {
  return KernelInfo{reinterpret_cast<const void*>(&gemm_rowln_kernel<4>), RowLnCfg::THREADS, RowLnCfg::smem_bytes(RowLnCfg::MAX_N)};
}
}  // namespace cfm
