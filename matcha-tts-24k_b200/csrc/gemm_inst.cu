// One instantiation of a tensor-core GEMM kernel per compilation (build.py passes -DCFM_BN / -DCFM_NEW / -DCFM_PAIR): the
// kernels dominate the build time, so each gets its own nvcc process.  Exports the kernel's launch information.
#include "gemm.cuh"

namespace cfm {
#define CFM_CAT3(a, b, c) a##b##_##c
#define CFM_NAME_TC(bn, nw) CFM_CAT3(kinfo_tc_, bn, nw)
#define CFM_CAT2(a, b) a##b
#define CFM_NAME_TC2(bn) CFM_CAT2(kinfo_tc2_, bn)
#if CFM_PAIR
KernelInfo CFM_NAME_TC2(CFM_BN)() {
  return KernelInfo{reinterpret_cast<const void*>(&gemm_tc2_kernel<CFM_BN>), Tc2Cfg<CFM_BN>::THREADS, Tc2Cfg<CFM_BN>::SMEM_BYTES};
}
#else
KernelInfo CFM_NAME_TC(CFM_BN, CFM_NEW)() {
  return KernelInfo{reinterpret_cast<const void*>(&gemm_tc_kernel<CFM_BN, CFM_NEW>), TcCfg<CFM_BN, CFM_NEW>::THREADS,
                    TcCfg<CFM_BN, CFM_NEW>::SMEM_BYTES};
}
#endif
}  // namespace cfm
