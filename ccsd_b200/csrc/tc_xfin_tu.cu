// tc_xfin_tu.cu -- translation unit of the tcgen05 ScoreNetworkX final-MLP kernel (tc_xfin.cuh)
#define TC_XFIN_KERNEL_TU
#define CCSD_AUX_TU
#include "tc_xfin.cuh"
