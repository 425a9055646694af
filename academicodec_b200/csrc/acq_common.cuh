// Shared helpers for the acq_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/acq_b200.h"

namespace acq {

// Table of device pointers passed by value in the kernel parameter space
// (stage-major: entry s*G+g).  Avoids a device-side pointer array + its H2D copy.
struct PtrTable {
    const float* p[ACQ_MAX_TABLE];
};
struct MutPtrTable {
    float* p[ACQ_MAX_TABLE];
};

void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

constexpr int kNumSMs = 148;   // B200

// Process-wide choice of the tensor-core search variant.  Initialised from the environment
// (ACQ_TC_KERNEL, ACQ_TC_CLUSTER, ACQ_TC_SPLIT) on first use; acq_tc_configure overrides it at run
// time (tests sweep the variants inside one process).
struct TcConfig {
    int variant;   // 0 = by shape, 1 = single fp16 product + rigorous filter + exact re-score, 3 = three-product split
    int cluster;   // CTAs sharing one multicast codebook stream: 0 = automatic, 1, 2 or 4
    int split;     // small batches: one cluster per tile, codebook passes split across its CTAs
};
TcConfig& tc_config();

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

}  // namespace acq
