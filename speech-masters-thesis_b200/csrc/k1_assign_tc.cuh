// placeholder until the tcgen05 kernel lands
#pragma once
#include "vq_common.cuh"
#include "k1_prepare.cuh"
namespace vq {
inline const char* tc_unsupported_reason(const float*, int64_t, int, int64_t, int) { return "the tcgen05 kernel is not built yet"; }
inline int launch_assign_tc(const float*, int64_t, int, int64_t, const float*, int, int64_t*, float*, double*, const AssignWorkspace&, cudaStream_t) { return fail("tc path not built%s"); }
}
