// EMA statistics exchange over NVLink peer memory (training, multi-GPU): the collective of the path.
//
// The reference has no collective here: each rank updates its codebooks from its local batch and DDP
// re-broadcasts rank 0's buffers before the next forward (core_vq.py:214-225 + main_launch.py:199-204).  This
// path sums the per-cluster statistics [S*K*D sums | S*K counts] over all ranks and applies the same update
// everywhere.  The first version called NCCL on a plain torch buffer (0.34 ms for 25 MB at 2 GPUs, a third
// of the recipe-batch step).  Here the buffer lives in symmetric memory (every rank maps every rank's copy,
// plus -- on NVSwitch systems -- one multicast address that fans out to all of them) and ONE kernel does the
// two-shot all-reduce in place:
//   rank r owns the elements [r*n/W, (r+1)*n/W) of the buffer
//   NVLS:  v = multimem.ld_reduce.add.v4.f32 [mc + i]   the switch adds the W copies on the way in
//          multimem.st.v4.f32 [mc + i], v               the switch writes the sum to all W copies
//   P2P:   v = sum over ranks in rank order of ld [peer[q] + i];  st [peer[q] + i], v for every q
// so every byte crosses the links once per direction (n/W in, n/W out per rank with NVLS) and every rank
// ends up with bit-identical sums (each element is reduced exactly once, by its owner).
// The caller brackets the launch with cross-rank barriers (all statistics written before / all slices
// stored after); torch's symmetric-memory handle provides both on the same stream.
#include "acq_common.cuh"

namespace acq {
namespace {

constexpr int MAXW = 16;
struct PeerPtrs { float* p[MAXW]; };

__device__ __forceinline__ float4 mm_ld_reduce(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];\n"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}
__device__ __forceinline__ void mm_st(float* mc, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(mc), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ float4 ld_sys(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];\n"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void st_sys(float* p, const float4& v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// n4 = number of float4 elements of the whole buffer; this rank reduces [lo4, hi4).
// Four independent 16-byte transactions per thread and iteration are in flight before the first is consumed:
// one NVLink round trip is ~2 us, and with one load per thread the kernel was latency-bound (90 us for 25 MB on
// two GPUs against 69 us for NCCL).
template <bool NVLS>
__global__ void __launch_bounds__(512) peer_allreduce_kernel(float* mc, const PeerPtrs peers, int world, size_t lo4,
                                                             size_t hi4) {
    constexpr int U = 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = lo4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi4; i0 += stride * U) {
        float4 v[U];
        if (NVLS) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t i = i0 + u * stride;
                if (i < hi4) v[u] = mm_ld_reduce(mc + 4 * i);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t i = i0 + u * stride;
                if (i < hi4) mm_st(mc + 4 * i, v[u]);
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const size_t i = i0 + u * stride;
                if (i < hi4) v[u] = ld_sys(peers.p[0] + 4 * i);
            }
            for (int q = 1; q < world; ++q) {
                float4 w[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const size_t i = i0 + u * stride;
                    if (i < hi4) w[u] = ld_sys(peers.p[q] + 4 * i);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const size_t i = i0 + u * stride;
                    if (i < hi4) {
                        v[u].x = __fadd_rn(v[u].x, w[u].x); v[u].y = __fadd_rn(v[u].y, w[u].y);
                        v[u].z = __fadd_rn(v[u].z, w[u].z); v[u].w = __fadd_rn(v[u].w, w[u].w);
                    }
                }
            }
            for (int q = 0; q < world; ++q) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const size_t i = i0 + u * stride;
                    if (i < hi4) st_sys(peers.p[q] + 4 * i, v[u]);
                }
            }
        }
    }
}

}  // namespace

int peer_allreduce(float* multicast, float* const* peers, int world, int rank, size_t n, cudaStream_t st) {
    if (world < 1 || world > MAXW || rank < 0 || rank >= world) return fail(ACQ_EINVAL, "peer_allreduce: world=%d rank=%d", world, rank);
    if (n % 4) return fail(ACQ_EINVAL, "peer_allreduce: element count %zu is not a multiple of 4", n);
    if (!multicast && !peers) return fail(ACQ_EINVAL, "peer_allreduce: neither a multicast nor peer pointers");
    if (world == 1 || n == 0) return 0;
    PeerPtrs pp;
    for (int q = 0; q < MAXW; ++q) pp.p[q] = nullptr;
    if (peers)
        for (int q = 0; q < world; ++q) {
            if (!peers[q] || ((uintptr_t)peers[q] & 15)) return fail(ACQ_EINVAL, "peer_allreduce: peer pointer %d null or unaligned", q);
            pp.p[q] = peers[q];
        }
    const size_t n4 = n / 4, per = (n4 + world - 1) / world;
    const size_t lo = per * rank < n4 ? per * rank : n4, hi = lo + per < n4 ? lo + per : n4;
    if (hi <= lo) return 0;
    // enough loads in flight to cover the NVLink round trip: 4 CTAs of 512 threads per SM
    size_t blocks = (hi - lo + 4 * 512 - 1) / (4 * 512);
    if (blocks > (size_t)kNumSMs * 4) blocks = (size_t)kNumSMs * 4;
    if (blocks < 1) blocks = 1;
    if (multicast) {
        if ((uintptr_t)multicast & 15) return fail(ACQ_EINVAL, "peer_allreduce: multicast pointer unaligned");
        peer_allreduce_kernel<true><<<(unsigned)blocks, 512, 0, st>>>(multicast, pp, world, lo, hi);
    } else {
        peer_allreduce_kernel<false><<<(unsigned)blocks, 512, 0, st>>>(nullptr, pp, world, lo, hi);
    }
    return check_cuda(cudaGetLastError(), "peer_allreduce launch");
}

}  // namespace acq
