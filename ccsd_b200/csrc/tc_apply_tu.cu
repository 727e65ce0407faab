// tc_apply_tu.cu -- the tcgen05 rank-2 apply kernel (tc_apply.cuh) for ONE ScoreNetworkF entry path:
// compiled once per -DTA_FMODE=k (k = 0..4), five pass-mode instantiations each.
#define TC_APPLY_KERNEL_TU
#define CCSD_AUX_TU
#include "tc_apply.cuh"

#ifndef TA_FMODE
#error "compile with -DTA_FMODE=0..4"
#endif
#define TA_CAT2(a, b) a##b
#define TA_CAT(a, b) TA_CAT2(a, b)

namespace ccsd {
int TA_CAT(tc_apply_launch_f, TA_FMODE)(const DevPlan *dP, int grid, const ApplyArgs &a, const TcApplyMaps &m, void *stream) {
  return tc_apply_launch_f<TA_FMODE>(dP, grid, a, m, stream);
}
}  // namespace ccsd
