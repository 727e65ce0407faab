// tc_gram.cuh -- tcgen05 / TMEM / TMA Gram kernel (placeholder until the tensor-core path lands).
#pragma once
#include "plan_dev.h"
namespace ccsd {
static inline int tc_gram_supported(int, int, int) { return 0; }
static inline int tc_gram_prepare() { return 0; }
static inline int tc_gram_launch(const DevPlan *, const DevPlan &, const float *, float *, float *, void *) { return -1; }
}  // namespace ccsd
