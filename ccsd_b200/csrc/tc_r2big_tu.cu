// tc_r2big_tu.cu -- translation unit of the large-E rank-2 tcgen05 GEMMs (tc_r2big.cuh)
#define TC_R2BIG_KERNEL_TU
#define CCSD_AUX_TU
#include "tc_r2big.cuh"
