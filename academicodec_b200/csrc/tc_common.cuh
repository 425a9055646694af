// Shared pieces of the tcgen05 search kernels: tile constants, operand-image layout, PTX wrappers.
#pragma once
#include "acq_common.cuh"
#include <cuda_fp16.h>

namespace acq {
namespace tc {

constexpr int BM = 128;            // frames per tile (UMMA M)
constexpr int BN = 256;            // codewords per pass (UMMA N)
// Operand images are K-major with one swizzle row of ROWB bytes per matrix row and ring stage:
//   ROWB = 128 -> SWIZZLE_128B, 64 channels per stage, 96 KiB stages, 2-deep ring
//   ROWB =  64 -> SWIZZLE_64B,  32 channels per stage, 48 KiB stages, 4-deep ring (finer
//                 prefetch: three stages of MMA work cover one TMA round trip)
constexpr int ROWB = 64;
constexpr int BK = ROWB / 2;       // channels (fp16 elements) per ring stage
constexpr int CPR = ROWB / 16;     // 16-byte chunks per row
constexpr int UK = 16;             // UMMA K for kind::f16
constexpr int A_BYTES = BM * ROWB; // one operand image (hi or lo) of a chunk
constexpr int B_BYTES = BN * ROWB;
constexpr int TMEM_COLS = 512;
constexpr int GMAX = 4;            // max channel groups
constexpr int KMAX = 1024;         // max codebook size (the scaled norms of one table live in shared memory)
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);


// ------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Blocking waits pass a suspend-time hint (ns): the thread sleeps in hardware until the phase completes or the
// time is up, instead of returning every ~80 cycles to a clock64 / compare / branch loop that competes for issue
// slots with the warps that have work (ACQ_TRYWAIT_HINT=0: no hint).
#ifndef ACQ_TRYWAIT_HINT
#define ACQ_TRYWAIT_HINT 0
#endif
__device__ __forceinline__ bool mbar_try_wait_blocking(uint32_t bar, uint32_t parity) {
#if ACQ_TRYWAIT_HINT > 0
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"((uint32_t)ACQ_TRYWAIT_HINT)
        : "memory");
    return ok != 0;
#else
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
#endif
}
// Non-suspending poll (mbarrier.test_wait): for the single-thread TMA / MMA roles, whose hand-offs are on the
// critical path of the operand ring.
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity, int* err, int code) {
    if (mbar_test_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_test_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) {
            if (err) atomicExch(err, code);
            __trap();
        }
    }
}
// Bounded wait: a protocol bug must not hang the GPU -- after ~4 s flag the error and trap.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err, int code) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_blocking(smem_u32(bar), parity)) {
        if (clock64() - t0 > 8000000000LL) {
            if (err) atomicExch(err, code);
            printf("[acq] mbarrier wait timed out: code %d, block %d, thread %d, parity %u\n", code, blockIdx.x,
                   threadIdx.x, parity);
            __trap();
        }
    }
}
// Same, accumulating the cycles spent waiting (stall attribution, ACQ_TC_DBG bit 512).
__device__ __forceinline__ void mbar_wait_t(uint64_t* bar, uint32_t parity, int* err, int code,
                                            unsigned long long& acc) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    mbar_wait(bar, parity, err, code);
    acc += (unsigned long long)(clock64() - t0);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// Multicast variant: the bytes land at the same CTA-relative offset of every CTA in `mask`, and each
// destination CTA's mbarrier (same offset) receives the complete_tx.
__device__ __forceinline__ void bulk_g2s_mc(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                            uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, "
        "[%3], %4;\n" ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
// ---- variants on 32-bit shared-memory addresses (single-thread producer loops: no generic-pointer arithmetic)
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_u32(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity, int* err, int code) {
    if (mbar_try_wait_u32(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_blocking(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) {
            if (err) atomicExch(err, code);
            __trap();
        }
    }
}
__device__ __forceinline__ void bulk_g2s_u32(uint32_t dst, const void* src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
        "l"(src_gmem), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc_u32(uint32_t dst, const void* src_gmem, uint32_t bytes, uint32_t bar,
                                                uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, "
        "[%3], %4;\n" ::"r"(dst),
        "l"(src_gmem), "r"(bytes), "r"(bar), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// ---- cross-CTA exchange inside a cluster (distributed shared memory) -------------------------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v2(uint32_t cluster_addr, float v, int i) {
    asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};\n" ::"r"(cluster_addr), "r"(__float_as_uint(v)), "r"(i)
                 : "memory");
}
__device__ __forceinline__ void fence_acq_rel_cluster() { asm volatile("fence.acq_rel.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_bar_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                     smem_u32(dst_smem)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(addr), "r"(cols)
                 : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                         uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// Arrive on the barrier at this offset in every CTA of `mask` once the MMAs issued so far retire.
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::
            "r"(smem_u32(bar)),
        "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// Asynchronous accumulator read: issue now, wait later (tmem_ld_wait ties the registers to the wait
// so that the compiler cannot consume them early).
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),
                   "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]),
                   "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]),
                   "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]),
                   "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// 16-column variants (rolled epilogue loops: the instruction footprint of the x32 version, fully unrolled,
// does not fit the L0 instruction cache -- ncu showed the sweeps stalled on instruction fetch)
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]),
                   "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]),
                   "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

// Walk one 256-column accumulator of this thread's row in ascending column order, calling
// f(column, score) with score = acc - xs * hn[column]; the TMEM read of the next 32 columns is in
// flight while the current 32 are processed, and hn comes from shared memory (warp-uniform
// 16-byte broadcasts).
template <typename F>
__device__ __forceinline__ void score_block32(const uint32_t (&r)[32], const float* hn, float nxs, int c,
                                              F& f) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 h = *reinterpret_cast<const float4*>(hn + j);
        f(c + j, fmaf(nxs, h.x, __uint_as_float(r[j])));
        f(c + j + 1, fmaf(nxs, h.y, __uint_as_float(r[j + 1])));
        f(c + j + 2, fmaf(nxs, h.z, __uint_as_float(r[j + 2])));
        f(c + j + 3, fmaf(nxs, h.w, __uint_as_float(r[j + 3])));
    }
}
template <typename F>
__device__ __forceinline__ void for_each_score(uint32_t taddr, const float* hn_pass, float nxs, F f) {
    uint32_t ra[32], rb[32];
    tmem_ld32_async(taddr, ra);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 64) {
        tmem_ld_wait(ra);
        tmem_ld32_async(taddr + c0 + 32, rb);
        score_block32(ra, hn_pass + c0, nxs, c0, f);
        tmem_ld_wait(rb);
        if (c0 + 64 < BN) tmem_ld32_async(taddr + c0 + 64, ra);
        score_block32(rb, hn_pass + c0 + 32, nxs, c0 + 32, f);
    }
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (sm_100 version 1):
//   [0,14) start>>4   [16,30) LBO>>4 (unused for swizzled K-major, 1)   [32,46) SBO>>4 = 1024 B
//   (8 rows x 128 B per swizzle atom)   [46,48) version=1   [61,64) layout=2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    // SBO = one swizzle atom = 8 rows x ROWB bytes; layout type 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)((8 * ROWB) >> 4) << 32) |
           (1ull << 46) | ((uint64_t)(ROWB == 128 ? 2 : 4) << 61);
}

// Power-of-two scale that brings a magnitude into [1024, 2048).
__device__ __host__ __forceinline__ float scale_for(float maxabs) {
    uint32_t bits;
#ifdef __CUDA_ARCH__
    bits = __float_as_uint(maxabs);
#else
    memcpy(&bits, &maxabs, 4);
#endif
    int e = (int)((bits >> 23) & 0xFF);
    if (e == 0 || e == 255) return 1.0f;   // zero / denormal / inf / nan rows: no scaling
    int se = 264 - e;                      // 2^(10 - (e - 127)) has exponent field 137 - (e - 127)
    se = se < 1 ? 1 : (se > 254 ? 254 : se);
    uint32_t sb = (uint32_t)se << 23;
#ifdef __CUDA_ARCH__
    return __uint_as_float(sb);
#else
    float f;
    memcpy(&f, &sb, 4);
    return f;
#endif
}

// byte offset of (row r, 16-byte chunk c) inside a K-major swizzled operand image:
// Swizzle<3,4,3> (128 B rows): chunk ^= r & 7;  Swizzle<2,4,3> (64 B rows): chunk ^= (r >> 1) & 3
__device__ __host__ __forceinline__ uint32_t sw_offset(int r, int c) {
    const int x = ROWB == 128 ? (r & 7) : ((r >> 1) & 3);
    return (uint32_t)((r >> 3) * (8 * ROWB) + (r & 7) * ROWB + ((c ^ x) << 4));
}

// One 32-byte (full L2 sector) store.  The loaders' image stores are scattered (one sector per
// frame row); 16-byte halves of a sector cost a partial-sector write each -- measured 0.29 ms of
// a 1.06 ms kernel -- whereas whole sectors are written without read-modify-write.
__device__ __forceinline__ void stg256(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(p), "r"(a.x), "r"(a.y),
                 "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}
// Store the two adjacent 16-byte chunks c (even) and c+1 of row r of a swizzled operand image: the
// XOR swizzle keeps them inside one 32-byte sector, possibly swapped.
__device__ __forceinline__ void store_chunk_pair(uint8_t* chunk_base, int r, int c, const uint4& v0,
                                                 const uint4& v1) {
    const int x = ROWB == 128 ? (r & 7) : ((r >> 1) & 3);
    uint8_t* dst = chunk_base + (r >> 3) * (8 * ROWB) + (r & 7) * ROWB + (((c ^ x) & ~1) << 4);
    if (x & 1) stg256(dst, v1, v0); else stg256(dst, v0, v1);
}

__device__ __forceinline__ uint32_t pack_half2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

// ------------------------------------------------------------------------------------ kernel
// Per-CTA scratch in global memory (L2 resident), double-buffered by tile parity:
//   Aimg[2][D/64 chunks][hi 16 KiB | lo 16 KiB]   fp16 operand images of the tile's residual
//   R   [2][128][D] fp32                          exact residual rows (only touched when S > 1)
__device__ __forceinline__ void fence_proxy_async_global() {
    asm volatile("fence.proxy.async.global;\n" ::: "memory");
}

// split 8 scaled fp32 values into fp16 hi / lo and pack each into one 16-byte chunk
__device__ __forceinline__ void split8(const float (&a)[8], float xs, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float v0 = a[2 * j] * xs, v1 = a[2 * j + 1] * xs;
        const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
        h[j] = pack_half2(h0, h1);
        l[j] = pack_half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1)));
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// One row of a tile's bias chunk: fp16 {w, w, w, 0 ...} in the first K16 slice of the row (32 bytes; the
// MMA that consumes the chunk reads nothing else).  w is a power of two; below 2^-24 it flushes to zero.
__device__ __forceinline__ void write_bias_row(uint8_t* bias_img, int row, float w) {
    const uint32_t h = (uint32_t)__half_as_ushort(__float2half_rn(w));
    *reinterpret_cast<uint4*>(bias_img + sw_offset(row, 0)) = make_uint4(h | (h << 16), h, 0u, 0u);
    *reinterpret_cast<uint4*>(bias_img + sw_offset(row, 1)) = make_uint4(0u, 0u, 0u, 0u);
}

// Between two residual stages: r <- r - e[i] in exact fp32 with the reference's operation order
// (core_vq.py:359, or the straight-through form :304/:339), then the next stage's per-row scale and
// fp16 operand image.  One warp per frame, lanes across channels; RB frames are in flight together
// so that the dependent codeword gathers (L2 latency) overlap -- processing the 32 frames one by one
// made this phase 4x longer than the tile's MMAs.
//   LO   also write the lo image (three-product kernel); the image then holds {hi | lo} per channel chunk
//        (ACH = 2 * A_BYTES bytes per chunk), otherwise hi only (ACH = A_BYTES)
//   SQ   also return, per row, ||hi||^2 and ||v*xs - hi||^2 of the SCALED new residual (the single-product
//        kernel's error bound): sq_g[2*row], sq_g[2*row+1]
// One batch of RB rows (row0 .. row0+RB-1); idxs = their winning codes.
// NDST > 1 (split mode of the 3-product kernel): the new image rows and row scales are written into the
// scratch / shared memory of all NDST CTAs of the cluster (img = rank 0's buffer, consecutive ranks
// cta_stride bytes apart; the scale goes out through distributed shared memory); the fp32 residual rows
// stay private to the CTA that owns these rows.
//   BIAS also write the row of the next stage's bias chunk (single-product kernel: the -xs*hn_k term of the
//        score is folded into the MMA as an extra K-slice): three fp16 copies of w = xs / bscale, where the
//        row scale is capped at xs_cap so that w stays representable
template <int RB, int JN, bool LO, bool SQ, int NDST = 1, bool BIAS = false>
__device__ __forceinline__ void residual_update_batch(int row0, int lane, int nf, const int (&idxs)[RB],
                                                      const float* __restrict__ cbp, int Dg, int D, int g,
                                                      float* R, uint8_t* img, float* sc_g, float* sq_g, bool ste,
                                                      size_t cta_stride = 0, uint8_t* bias_img = nullptr,
                                                      float xs_cap = 0.f, float inv_bscale = 0.f) {
    float4 e[RB][JN], r[RB][JN];
    bool live[RB];
#pragma unroll
    for (int u = 0; u < RB; ++u) {
        const int urow = row0 + u;
        const int idx = idxs[u];
        live[u] = urow < nf;                                   // tail rows stay zero
        const float* erow = cbp + (size_t)idx * Dg;
        const float* rrow = R + (size_t)urow * D + g * Dg;
#pragma unroll
        for (int j = 0; j < JN; ++j) {
            const int d = lane * 4 + 128 * j;
            if (live[u] && d < Dg) {
                e[u][j] = __ldg(reinterpret_cast<const float4*>(erow + d));
                r[u][j] = *reinterpret_cast<const float4*>(rrow + d);
            } else {
                e[u][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                r[u][j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    constexpr int ACH = LO ? 2 * A_BYTES : A_BYTES;     // image bytes per channel chunk
    float m[RB];
#pragma unroll
    for (int u = 0; u < RB; ++u) {
        m[u] = 0.f;
#pragma unroll
        for (int j = 0; j < JN; ++j) {
            float4 v = r[u][j];
            const float4 ee = e[u][j];
            if (ste) {
                v.x = __fsub_rn(v.x, __fadd_rn(v.x, __fsub_rn(ee.x, v.x)));
                v.y = __fsub_rn(v.y, __fadd_rn(v.y, __fsub_rn(ee.y, v.y)));
                v.z = __fsub_rn(v.z, __fadd_rn(v.z, __fsub_rn(ee.z, v.z)));
                v.w = __fsub_rn(v.w, __fadd_rn(v.w, __fsub_rn(ee.w, v.w)));
            } else {
                v.x = __fsub_rn(v.x, ee.x); v.y = __fsub_rn(v.y, ee.y);
                v.z = __fsub_rn(v.z, ee.z); v.w = __fsub_rn(v.w, ee.w);
            }
            r[u][j] = v;
            m[u] = fmaxf(m[u], fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
        for (int u = 0; u < RB; ++u) m[u] = fmaxf(m[u], __shfl_xor_sync(0xffffffffu, m[u], off));
    }
#pragma unroll
    for (int u = 0; u < RB; ++u) {
        if (!live[u]) continue;
        const int urow = row0 + u;
        float* rrow = R + (size_t)urow * D + g * Dg;
        float xs = scale_for(m[u]);
        if (BIAS) xs = fminf(xs, xs_cap);
        if (BIAS && lane == 0) write_bias_row(bias_img, urow, xs * inv_bscale);
        if (lane == 0) {
            if (NDST == 1) {
                sc_g[urow] = xs;
            } else {
                const uint32_t a = smem_u32(sc_g + urow);
#pragma unroll
                for (int c = 0; c < NDST; ++c)
                    asm volatile("st.shared::cluster.b32 [%0], %1;\n" ::"r"(mapa_u32(a, (uint32_t)c)),
                                 "r"(__float_as_uint(xs))
                                 : "memory");
            }
        }
        float qh = 0.f, qd = 0.f;
#pragma unroll
        for (int j = 0; j < JN; ++j) {
            const int d = lane * 4 + 128 * j;
            if (d < Dg) {
                const float4 v = r[u][j];
                *reinterpret_cast<float4*>(rrow + d) = v;
                const float v0 = v.x * xs, v1 = v.y * xs, v2 = v.z * xs, v3 = v.w * xs;
                const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1),
                             h2 = __float2half_rn(v2), h3 = __float2half_rn(v3);
                if (SQ) {
                    const float f0 = __half2float(h0), f1 = __half2float(h1), f2 = __half2float(h2),
                                f3 = __half2float(h3);
                    qh = fmaf(f0, f0, fmaf(f1, f1, fmaf(f2, f2, fmaf(f3, f3, qh))));
                    const float e0 = v0 - f0, e1 = v1 - f1, e2 = v2 - f2, e3 = v3 - f3;
                    qd = fmaf(e0, e0, fmaf(e1, e1, fmaf(e2, e2, fmaf(e3, e3, qd))));
                }
                const int dd = g * Dg + d;        // channel within the full latent
                uint8_t* dst = img + (size_t)(dd / BK) * ACH + sw_offset(urow, (dd % BK) >> 3) +
                               ((dd & 7) >> 2) * 8;
                const uint2 hv = make_uint2(pack_half2(h0, h1), pack_half2(h2, h3));
                uint2 lv = make_uint2(0u, 0u);
                if (LO)
                    lv = make_uint2(
                        pack_half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1))),
                        pack_half2(__float2half_rn(v2 - __half2float(h2)), __float2half_rn(v3 - __half2float(h3))));
#pragma unroll
                for (int c = 0; c < NDST; ++c) {
                    *reinterpret_cast<uint2*>(dst + c * cta_stride) = hv;
                    if (LO) *reinterpret_cast<uint2*>(dst + c * cta_stride + A_BYTES) = lv;
                }
            }
        }
        if (SQ) {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                qh += __shfl_xor_sync(0xffffffffu, qh, off);
                qd += __shfl_xor_sync(0xffffffffu, qd, off);
            }
            if (lane == 0) { sq_g[2 * urow] = qh; sq_g[2 * urow + 1] = qd; }
        }
    }
}
// Kernel parameters shared by the 3-product and the single-product search kernels.
struct TcParams {
    const float* x;
    PtrTable cb;
    const uint8_t* pack;     // per table: [pass][chunk][hi|lo][BN x 128 B] images, pre-swizzled,
    size_t table_stride;     //            then hn[K] = cs * 0.5||e||^2, then cs, max bits
    size_t img_bytes;        // bytes of one table's images
    size_t hn_bytes;         // bytes of one table's norms (the 256 B tail {cs, max|e|, max norm^2 ...} follows)
    size_t bias_off;         // offset of one table's bias images inside its record
    float* scratch;          // per CTA: fp16 images [2 tiles] + fp32 residual rows [2 tiles]
    int S, G, K, D, Dg, T, flags;
    long long N;
    int num_tiles;
    int tiles_per_clip;      // single-product kernel: > 0 = tiles never straddle clips (tile -> clip tile / tpc,
                             // frames (tile % tpc) * 128 ..), x is streamed by tensor-map TMA; 0 = tiles are
                             // consecutive runs of 128 frames of the flattened [B*T] index
    int64_t* codes;
    float* dbg_scores;       // optional [N][K] scores of stage 0 / group 0 (tests)
    int* err;                // optional device flag set on a barrier timeout
    unsigned long long* stall;   // stall-attribution counters (ACQ_TC_DBG bit 512), after err
    int wide;                // single-product kernel: eight loader warps / one epilogue set (single-stage streamed calls)
    int guard;               // automatic kernel choice without a host round trip: 0 = always run, 1 = run only if
                             // every table of the call is fit for the single-product filter (pack tail
                             // TAIL_NSMALL), 2 = run only if one is not; the host launches both kernels
    int dbg_mode;            // perf experiments (ACQ_TC_DBG): 1 = loaders idle after their first tile,
                             // 2 = skip the B copies, 4 = skip the A copies (results are then wrong)
};

// ---- codebook pack: one record per table, so any contiguous range of tables is itself a pack ----
//   [images: (K/256) x (Dg/BK) blocks of {hi B_BYTES | lo B_BYTES}] [hn: K f32 = cs*0.5||e||^2]
//   [tail 256 B: cs f32 | max|e| bits u32 | max_k ||cs e_k||^2 | max_k ||cs e_k - fp16(cs e_k)||^2 | max_k hn
//    (the last three as fp32 bit patterns of values rounded up: the single-product kernel's error bound)]
//   [bias images: K/256 blocks of B_BYTES: row k = fp16 {-b1, -b2, -b3, 0 ...}, b1 + b2 + b3 = hn_k * bscale
//    (bscale = tail slot 5, a power of two that brings max_k hn into [2^14, 2^15)): with the matching
//    {w, w, w, 0 ...} rows on the A side, w = xs / bscale, one extra K16 MMA adds -xs * hn_k to every score]
constexpr int TAIL_CS = 0, TAIL_MAXBITS = 1, TAIL_EMAX2 = 2, TAIL_DE2MAX = 3, TAIL_HNMAX = 4, TAIL_BSCALE = 5,
              TAIL_NSMALL = 6;   // codewords whose squared norm is below 1/16 of the table's largest
__host__ __device__ inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
__host__ __device__ inline size_t images_bytes(int K, int Dg) { return (size_t)(K / BN) * (Dg / BK) * 2 * B_BYTES; }
__host__ __device__ inline size_t bias_offset_bytes(int K, int Dg) {
    return align256(images_bytes(K, Dg)) + align256((size_t)K * 4) + 256;
}
__host__ __device__ inline size_t table_stride_bytes(int K, int Dg) {
    return bias_offset_bytes(K, Dg) + (size_t)(K / BN) * B_BYTES;
}

// The single-product filter's error bound scales with the LARGEST codeword norm of a table, the score gaps with
// the norms of the codewords that compete for a frame.  A table whose codewords differ widely in norm -- an EMA
// codebook in training: a few dead codes keep their initial norm while the live ones contract towards cluster
// means -- puts dozens of codewords inside the bound for every frame, the candidate lists overflow and every
// row falls back to exact scores of all K codewords (measured: 85 ms instead of 2.1 ms for 64 000 frames, D = 512,
// n_q = 12).  The pack therefore counts, per table, the codewords with ||e||^2 < max ||e||^2 / 16; a call is fit
// for the single-product kernel iff at most K/20 of them are that small in every table it uses.  Both kernels
// are launched back to back and each one returns at once if it is not its turn (TcParams::guard).
__device__ __forceinline__ bool tables_fit_single_product(const TcParams& p) {
    bool fit = true;
    for (int t = 0; t < p.S * p.G; ++t) {
        const uint32_t* tail = reinterpret_cast<const uint32_t*>(p.pack + (size_t)t * p.table_stride + p.img_bytes + p.hn_bytes);
        fit = fit && (__ldg(tail + TAIL_NSMALL) * 20u <= (uint32_t)p.K);
    }
    return fit;
}
// true = this kernel is not the one to run for this call (uniform over the grid: every thread returns)
__device__ __forceinline__ bool guard_skips(const TcParams& p) {
    if (p.guard == 0) return false;
    __shared__ int fit_s;
    if (threadIdx.x == 0) fit_s = tables_fit_single_product(p) ? 1 : 0;
    __syncthreads();
    return (p.guard == 1) != (fit_s != 0);
}

}  // namespace tc
}  // namespace acq
