// tc_attn_tu.cu -- translation unit of the tcgen05 attention-channel kernel (tc_attn.cuh)
#define TC_ATTN_KERNEL_TU
#define CCSD_AUX_TU
#include "tc_attn.cuh"
