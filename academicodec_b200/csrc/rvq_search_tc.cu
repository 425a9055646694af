// K1b placeholder: tensor-core (tcgen05) search.  Filled in by the tcgen05 implementation.
#include "acq_common.cuh"
namespace acq {
bool rvq_search_tc_supported(int S, int G, int K, int D, int flags, const char** why) {
    *why = "not built";
    return false;
}
int rvq_search_tc(const float*, const float* const*, const float*, int, int, int, int, int, int,
                  int, int64_t*, float*, float*, double*, cudaStream_t) {
    return fail(ACQ_ENOTIMPL, "tensor-core search not built");
}
}  // namespace acq
