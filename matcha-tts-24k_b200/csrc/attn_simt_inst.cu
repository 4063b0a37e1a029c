// One instantiation of the fp32-FMA attention kernel per compilation (build.py passes -DCFM_BF16=0/1 -DCFM_D=32/64).
#include "attn.cuh"

namespace cfm {
#define CFM_CAT(a, b) a##b
#define CFM_NAME(prefix, d) CFM_CAT(prefix, d)
#if CFM_BF16
KernelInfo CFM_NAME(kinfo_attn_simt_bf16_, CFM_D)() { return KernelInfo{reinterpret_cast<const void*>(&attn_simt_kernel<bf16, CFM_D>), 128, 0}; }
#else
KernelInfo CFM_NAME(kinfo_attn_simt_f32_, CFM_D)() { return KernelInfo{reinterpret_cast<const void*>(&attn_simt_kernel<float, CFM_D>), 128, 0}; }
#endif
}  // namespace cfm
