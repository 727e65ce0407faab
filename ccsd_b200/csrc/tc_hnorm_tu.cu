// tc_hnorm_tu.cu -- translation unit of the Gram-quantity norm kernel (tc_hnorm.cuh)
#define TC_HNORM_KERNEL_TU
#define CCSD_AUX_TU
#include "tc_hnorm.cuh"
