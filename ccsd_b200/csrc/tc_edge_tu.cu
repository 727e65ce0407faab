// tc_edge_tu.cu -- translation unit of the tcgen05 per-edge MLP kernel (tc_edge.cuh)
#define TC_EDGE_KERNEL_TU
#define CCSD_AUX_TU
#include "tc_edge.cuh"
