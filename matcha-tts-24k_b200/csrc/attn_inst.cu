// The tensor-core attention kernel, compiled in its own nvcc process (see build.py); exports its launch information.
#define CFM_ATTN_KERNEL_TU 1
#include "attn_tc.cuh"
#include "attn_persist.cuh"
