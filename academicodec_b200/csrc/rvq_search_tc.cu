// K1b: residual nearest-codeword search on the 5th-gen tensor cores (tcgen05 + TMEM + TMA bulk).
//
// Problem: per frame, argmax_k  x.e_k - 0.5||e_k||^2  over a 1024-entry codebook, for S residual
// stages x G channel groups, with the residual carried from stage to stage.  The contraction
// x.e^T is GEMM-shaped (frames x codewords x channels) and runs on tcgen05.mma; the codes must
// match the reference's fp32 arithmetic, which a single fp16/bf16/tf32 pass cannot deliver
// (0.09 % .. 0.4 % of indices flip).  Precision scheme:
//
//   every operand is scaled by an exact power of two so that its largest magnitude lies in
//   [1024, 2048) (per frame and group for x, per codebook for e) and split into two fp16 numbers
//   v = hi + lo + delta, |delta| <= 2^-22 |v|_max.  Three kind::f16 MMAs accumulate
//   hi.hi + hi.lo + lo.hi into one fp32 TMEM accumulator: the dropped lo.lo term and delta are
//   below 2^-21 relative to |x||e|, i.e. fp32-class, at 3 fp16 MMAs per product.
//
// Structure (one persistent CTA per SM, 320 threads, warp-specialised, everything mbarrier-driven):
//   warps 0-3  loaders: read the NEXT tile of x (coalesced along frames), derive the per-frame
//              scales, split to fp16 hi/lo and write the tile's K-major SWIZZLE_128B operand
//              images into per-CTA global scratch (L2 resident, double-buffered), one tile ahead
//              of the MMAs so the HBM read overlaps tensor work
//   warp 8     TMA producer: one thread streams A (residual) and B (pre-packed codebook) images,
//              already in the UMMA shared-memory layout, with cp.async.bulk into a 2-stage ring
//   warp 9     MMA issuer: one thread issues 12 tcgen05.mma per ring stage into one of two
//              256-column TMEM accumulators; tcgen05.commit frees the stage / publishes the tile
//   warps 4-7  epilogue: tcgen05.ld the accumulator (thread = frame, 32 columns at a time), add
//              the scaled -0.5||e||^2 bias, running (value, index) argmax with the lowest-index
//              tie rule, overlapping the next pass's MMAs through the second accumulator; between
//              stages they gather the winning codewords, r <- r - e[i] in fp32 exactly as the
//              reference does (core_vq.py:359 / :304), and re-split the residual for the next stage.
//
// This kernel produces codes only; quantized / loss / EMA outputs come from the fused SIMT
// kernel (rvq_search_simt.cu) or from decode.  Shapes: K % 256 == 0, (D/G) % 64 == 0.
#include "acq_common.cuh"
#include <cuda_fp16.h>
#include <stdlib.h>

namespace acq {
namespace {

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
constexpr int NSTAGE = ROWB == 64 ? 4 : 2;
constexpr int A_BYTES = BM * ROWB; // one operand image (hi or lo) of a chunk
constexpr int B_BYTES = BN * ROWB;
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // 48 / 96 KiB
constexpr int NUM_THREADS = 320;
constexpr int TMEM_COLS = 512;
constexpr int GMAX = 8;            // max channel groups
constexpr int BAR_BYTES = (2 * NSTAGE + 8 + GMAX) * 8;   // mbarriers
constexpr int CTRL_BYTES = BAR_BYTES + 16 /*tmem ptr*/ + 2 * GMAX * BM * 4 /*row scales, 2 tiles*/ +
                           GMAX * BM * 4 /*row max bits*/;
constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + (size_t)NSTAGE * STAGE_BYTES + CTRL_BYTES;

// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, N=256, M=128
//   [4,6) c_format=1(F32)  [7,10) a_format=0(F16)  [10,13) b_format=0(F16)
//   [15] a_major=0(K)  [16] b_major=0(K)  [17,23) N>>3  [24,29) M>>4
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct TcParams {
    const float* x;
    PtrTable cb;
    const uint8_t* pack;     // per table: [pass][chunk][hi|lo][BN x 128 B] images, pre-swizzled,
    size_t table_stride;     //            then hn[K] = cs * 0.5||e||^2, then cs, max bits
    size_t img_bytes;        // bytes of one table's images
    float* scratch;          // per CTA: fp16 images [2 tiles] + fp32 residual rows [2 tiles]
    int S, G, K, D, Dg, T, flags;
    long long N;
    int num_tiles;
    int64_t* codes;
    float* dbg_scores;       // optional [N][K] scores of stage 0 / group 0 (tests)
    int* err;                // optional device flag set on a barrier timeout
    int dbg_mode;            // perf experiments (ACQ_TC_DBG): 1 = loaders idle after their first tile,
                             // 2 = skip the B copies, 4 = skip the A copies (results are then wrong)
};

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
// Bounded wait: a protocol bug must not hang the GPU -- after ~4 s flag the error and trap.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* err, int code) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 8000000000LL) {
            if (err) atomicExch(err, code);
            __trap();
        }
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// L2 eviction policies: x is streamed once (evict first); the scratch images and the codebook
// pack are re-read many times and should stay resident (evict last)
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 ldg_stream(const float4* ptr, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;\n"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ void stg_keep(void* ptr, const uint4& v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;\n" ::"l"(ptr), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                              uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(src_gmem), "r"(bytes) : "memory");
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

__global__ void __launch_bounds__(NUM_THREADS, 1) rvq_search_tc_kernel(const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* ctrl = smem + NSTAGE * STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);          // [NSTAGE] TMA bytes landed
    uint64_t* empty_bar = full_bar + NSTAGE;                         // [NSTAGE] MMAs retired
    uint64_t* tfull_bar = empty_bar + NSTAGE;                        // [2] accumulator complete
    uint64_t* tempty_bar = tfull_bar + 2;                            // [2] accumulator drained
    uint64_t* t0_bar = tempty_bar + 2;                               // [2] stage-0 images of a tile ready
    uint64_t* free_bar = t0_bar + 2;                                 // [2] tile buffers reusable
    uint64_t* upd_bar = free_bar + 2;                                // [GMAX] next-stage image ready
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(ctrl + BAR_BYTES);
    float* scale_s = reinterpret_cast<float*>(ctrl + BAR_BYTES + 16);             // [2][GMAX][BM]
    uint32_t* rowmax_s = reinterpret_cast<uint32_t*>(ctrl + BAR_BYTES + 16 + 2 * GMAX * BM * 4);  // [GMAX][BM]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = p.S, G = p.G, K = p.K, D = p.D, Dg = p.Dg, T = p.T;
    const int NP = K / BN, NKC = Dg / BK;
    const bool ste = p.flags & ACQ_STE;
    const size_t tile_elems = (size_t)BM * D;
    uint8_t* Aimg = reinterpret_cast<uint8_t*>(p.scratch) + (size_t)blockIdx.x * 4 * tile_elems * 4;
    float* Rbuf = reinterpret_cast<float*>(Aimg + 2 * tile_elems * 4);
    const size_t img_tile_bytes = tile_elems * 4;      // hi + lo fp16 = 4 bytes per element

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&full_bar[i], 1);         // the TMA thread's arrive.expect_tx
            mbar_init(&empty_bar[i], 1);        // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);        // tcgen05.commit
            mbar_init(&tempty_bar[i], 128);     // epilogue threads
            mbar_init(&t0_bar[i], 128);         // loader threads
            mbar_init(&free_bar[i], 128);       // epilogue threads
        }
        for (int i = 0; i < GMAX; ++i) mbar_init(&upd_bar[i], 128);   // epilogue threads
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(tmem_ptr_s, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (warp < 4) {
        // ================= loaders: x tile -> scales, fp16 hi/lo images (and R when S > 1) ========
        // Runs one tile ahead of the MMAs (double-buffered scratch), so the HBM read of the next
        // tile overlaps the tensor work of the current one.
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t buf = it & 1;
            mbar_wait(&free_bar[buf], ((it >> 1) & 1) ^ 1, p.err, 6);
            if ((p.dbg_mode & 1) && it >= 2) { mbar_arrive(&t0_bar[buf]); continue; }
            const long long n0 = (long long)tile * BM;
            uint8_t* img = Aimg + buf * img_tile_bytes;
            float* R = Rbuf + buf * tile_elems;
            float* sc = scale_s + buf * GMAX * BM;
            for (int i = tid; i < G * BM; i += 128) rowmax_s[i] = 0u;
            named_bar_sync(2, 128);
            if ((T & 3) == 0) {
                // 4 consecutive frames per thread (one 16 B load per channel), 8 channels at a time
                const int rq = tid & 31, w4 = tid >> 5;
                const long long n = n0 + 4 * rq;
                const bool ok = n < p.N;                 // N % 4 == 0: a quad is all in or all out
                const long long b = ok ? n / T : 0, t = ok ? n % T : 0;
                const float* src = p.x + (size_t)(b * D) * T + t;
                for (int sweep = 0; sweep < 2; ++sweep) {
                    // two channel octets per iteration: 16 independent 16-byte loads in flight per
                    // thread (the loaders are latency-bound; this is what keeps them ahead of the MMAs)
                    for (int oct0 = w4; oct0 < D / 8; oct0 += 8) {
                        float4 v[2][8];
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int oct = oct0 + 4 * h;
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                v[h][i] = (ok && oct < D / 8)
                                              ? __ldg(reinterpret_cast<const float4*>(src + (size_t)(oct * 8 + i) * T))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int oct = oct0 + 4 * h;
                            if (oct >= D / 8) break;
                            const int g = (oct * 8) / Dg;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float a[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    a[i] = j == 0 ? v[h][i].x : (j == 1 ? v[h][i].y : (j == 2 ? v[h][i].z : v[h][i].w));
                                const int row = 4 * rq + j;
                                if (sweep == 0) {
                                    float m = 0.f;
#pragma unroll
                                    for (int i = 0; i < 8; ++i) m = fmaxf(m, fabsf(a[i]));
                                    atomicMax(&rowmax_s[g * BM + row], __float_as_uint(m));
                                } else {
                                    uint4 hi, lo;
                                    split8(a, sc[g * BM + row], hi, lo);
                                    uint8_t* dst = img + (size_t)(oct / CPR) * 2 * A_BYTES + sw_offset(row, oct % CPR);
                                    *reinterpret_cast<uint4*>(dst) = hi;
                                    *reinterpret_cast<uint4*>(dst + A_BYTES) = lo;
                                    if (S > 1) {
                                        float* rd = R + (size_t)row * D + oct * 8;
                                        *reinterpret_cast<float4*>(rd) = make_float4(a[0], a[1], a[2], a[3]);
                                        *reinterpret_cast<float4*>(rd + 4) = make_float4(a[4], a[5], a[6], a[7]);
                                    }
                                }
                            }
                        }
                    }
                    if (sweep == 0) {
                        named_bar_sync(2, 128);
                        for (int i = tid; i < G * BM; i += 128) sc[i] = scale_for(__uint_as_float(rowmax_s[i]));
                        named_bar_sync(2, 128);
                    }
                }
            } else {
                // general T: one frame per thread, scalar loads (still coalesced across the warp)
                const int row = tid;
                const long long n = n0 + row;
                const bool ok = n < p.N;
                const long long b = ok ? n / T : 0, t = ok ? n % T : 0;
                const float* src = p.x + (size_t)(b * D) * T + t;
                for (int sweep = 0; sweep < 2; ++sweep) {
                    for (int oct = 0; oct < D / 8; ++oct) {
                        float a[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) a[i] = ok ? __ldg(src + (size_t)(oct * 8 + i) * T) : 0.f;
                        const int g = (oct * 8) / Dg;
                        if (sweep == 0) {
                            float m = 0.f;
#pragma unroll
                            for (int i = 0; i < 8; ++i) m = fmaxf(m, fabsf(a[i]));
                            atomicMax(&rowmax_s[g * BM + row], __float_as_uint(m));
                        } else {
                            uint4 hi, lo;
                            split8(a, sc[g * BM + row], hi, lo);
                            uint8_t* dst = img + (size_t)(oct / CPR) * 2 * A_BYTES + sw_offset(row, oct % CPR);
                            *reinterpret_cast<uint4*>(dst) = hi;
                            *reinterpret_cast<uint4*>(dst + A_BYTES) = lo;
                            if (S > 1) {
                                float* rd = R + (size_t)row * D + oct * 8;
                                *reinterpret_cast<float4*>(rd) = make_float4(a[0], a[1], a[2], a[3]);
                                *reinterpret_cast<float4*>(rd + 4) = make_float4(a[4], a[5], a[6], a[7]);
                            }
                        }
                    }
                    if (sweep == 0) {
                        named_bar_sync(2, 128);
                        for (int i = tid; i < G * BM; i += 128) sc[i] = scale_for(__uint_as_float(rowmax_s[i]));
                        named_bar_sync(2, 128);
                    }
                }
            }
            fence_proxy_async_global();     // generic-proxy global writes -> TMA (async proxy) reads
            mbar_arrive(&t0_bar[buf]);
        }
    } else if (warp == 8) {
        // ================= TMA producer: one thread streams A and B operand images ================
        if (lane == 0) {
            uint32_t it = 0, ring_it = 0, upd_it[GMAX];
#pragma unroll
            for (int i = 0; i < GMAX; ++i) upd_it[i] = 0;
            const int steps_per_tile = S * G * NP * NKC;
            const int pf_per_step = (D + steps_per_tile - 1) / steps_per_tile;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const uint32_t buf = it & 1;
                const uint8_t* img = Aimg + buf * img_tile_bytes;
                // Experiment (ACQ_TC_DBG bit 8): L2 prefetch of the x tile the loaders read next-but-one.
                // Measured on B200: 1.03 -> 1.20 ms, i.e. harmful -- L2/HBM is the contended resource;
                // explicit evict_last / evict_first cache hints on the copies were also slower.
                const float* pf_base = nullptr;
                int pf_next = 0;
                {
                    const long long pt = (long long)tile + 2LL * gridDim.x;
                    if (pt < p.num_tiles && (T & 3) == 0) {
                        const long long pn = pt * BM, pb = pn / T, ptt = pn % T;
                        if (ptt + BM <= T) pf_base = p.x + (size_t)(pb * D) * T + ptt;
                    }
                }
                for (int s = 0; s < S; ++s) {
                    for (int g = 0; g < G; ++g) {
                        const uint8_t* bimg = p.pack + (size_t)(s * G + g) * p.table_stride;
                        for (int pass = 0; pass < NP; ++pass) {
                            for (int kc = 0; kc < NKC; ++kc, ++ring_it) {
                                const int st = ring_it % NSTAGE;
                                mbar_wait(&empty_bar[st], ((ring_it / NSTAGE) & 1) ^ 1, p.err, 2);
                                uint8_t* a_dst = smem + st * STAGE_BYTES;
                                uint8_t* b_dst = a_dst + 2 * A_BYTES;
                                const uint8_t* bsrc = bimg + (size_t)(pass * NKC + kc) * 2 * B_BYTES;
                                const bool skip_b = p.dbg_mode & 2, skip_a = p.dbg_mode & 4;
                                mbar_arrive_expect_tx(&full_bar[st], (skip_a ? 0 : 2 * A_BYTES) + (skip_b ? 0 : 2 * B_BYTES));
                                if (!skip_b) {
                                    bulk_g2s(b_dst, bsrc, B_BYTES, &full_bar[st]);
                                    bulk_g2s(b_dst + B_BYTES, bsrc + B_BYTES, B_BYTES, &full_bar[st]);
                                }
                                if (pf_base && (p.dbg_mode & 8)) {
                                    for (int i = 0; i < pf_per_step && pf_next < D; ++i, ++pf_next)
                                        bulk_prefetch_l2(pf_base + (size_t)pf_next * T, BM * sizeof(float));
                                }
                                if (pass == 0 && kc == 0) {
                                    // first use of this (tile, stage, group)'s residual image
                                    if (s == 0) {
                                        mbar_wait(&t0_bar[buf], (it >> 1) & 1, p.err, 7);
                                    } else {
                                        mbar_wait(&upd_bar[g], upd_it[g] & 1, p.err, 8);
                                        ++upd_it[g];
                                    }
                                    fence_proxy_async_global();
                                }
                                if (!skip_a)
                                    bulk_g2s(a_dst, img + (size_t)(g * NKC + kc) * 2 * A_BYTES, 2 * A_BYTES,
                                             &full_bar[st]);
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 9) {
        // ================= MMA issuer =================================================================
        if (lane == 0) {
            uint32_t ring_it = 0, acc_it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                for (int sg = 0; sg < S * G; ++sg) {
                    for (int pass = 0; pass < NP; ++pass, ++acc_it) {
                        const uint32_t abuf = acc_it & 1;
                        mbar_wait(&tempty_bar[abuf], ((acc_it >> 1) & 1) ^ 1, p.err, 3);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + abuf * BN;
                        for (int kc = 0; kc < NKC; ++kc, ++ring_it) {
                            const int st = ring_it % NSTAGE;
                            mbar_wait(&full_bar[st], (ring_it / NSTAGE) & 1, p.err, 4);
                            tc_fence_after();
                            const uint32_t a_hi = smem_u32(smem + st * STAGE_BYTES);
                            const uint32_t a_lo = a_hi + A_BYTES;
                            const uint32_t b_hi = a_hi + 2 * A_BYTES;
                            const uint32_t b_lo = b_hi + B_BYTES;
#pragma unroll
                            for (int kk = 0; kk < BK / UK; ++kk) {
                                const uint32_t ko = kk * UK * 2;   // bytes along K inside the swizzle atom
                                const uint64_t dah = make_desc(a_hi + ko), dal = make_desc(a_lo + ko);
                                const uint64_t dbh = make_desc(b_hi + ko), dbl = make_desc(b_lo + ko);
                                umma_f16(d_tmem, dah, dbl, IDESC, (kc | kk) != 0);   // small terms first
                                umma_f16(d_tmem, dal, dbh, IDESC, 1);
                                umma_f16(d_tmem, dah, dbh, IDESC, 1);
                            }
                            umma_commit(&empty_bar[st]);     // ring stage free once these MMAs retire
                        }
                        umma_commit(&tfull_bar[abuf]);       // accumulator complete
                    }
                }
            }
        }
    } else {
        // ================= epilogue + residual update (warps 4-7, thread = frame) ====================
        const int q = warp - 4;
        const int row = q * 32 + lane;
        uint32_t it = 0, acc_it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t buf = it & 1;
            const long long n0 = (long long)tile * BM;
            const int nf = (int)min((long long)BM, p.N - n0);
            uint8_t* img = Aimg + buf * img_tile_bytes;
            float* R = Rbuf + buf * tile_elems;
            float* sc = scale_s + buf * GMAX * BM;
            mbar_wait(&t0_bar[buf], (it >> 1) & 1, p.err, 9);    // scales of this tile are visible
            for (int s = 0; s < S; ++s) {
                for (int g = 0; g < G; ++g) {
                    const int table = s * G + g;
                    const float nxs = -sc[g * BM + row];
                    const float* hn = reinterpret_cast<const float*>(p.pack + (size_t)table * p.table_stride +
                                                                     p.img_bytes);
                    float best = -INFINITY;
                    int bidx = 0;
                    for (int pass = 0; pass < NP; ++pass, ++acc_it) {
                        const uint32_t abuf = acc_it & 1;
                        mbar_wait(&tfull_bar[abuf], (acc_it >> 1) & 1, p.err, 5);
                        tc_fence_after();
                        const uint32_t taddr = tmem_base + abuf * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
                        for (int c0 = 0; c0 < BN; c0 += 32) {
                            float acc[32];
                            tmem_ld32(taddr + c0, acc);
                            const int k0 = pass * BN + c0;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 h = __ldg(reinterpret_cast<const float4*>(hn + k0 + j));
                                const float s0 = fmaf(nxs, h.x, acc[j]), s1 = fmaf(nxs, h.y, acc[j + 1]),
                                            s2 = fmaf(nxs, h.z, acc[j + 2]), s3 = fmaf(nxs, h.w, acc[j + 3]);
                                if (s0 > best) { best = s0; bidx = k0 + j; }
                                if (s1 > best) { best = s1; bidx = k0 + j + 1; }
                                if (s2 > best) { best = s2; bidx = k0 + j + 2; }
                                if (s3 > best) { best = s3; bidx = k0 + j + 3; }
                                if (p.dbg_scores && table == 0 && row < nf) {
                                    float* o = p.dbg_scores + (size_t)(n0 + row) * K + k0 + j;
                                    const float inv = 1.0f / -nxs;      // undo the row scale
                                    o[0] = s0 * inv; o[1] = s1 * inv; o[2] = s2 * inv; o[3] = s3 * inv;
                                }
                            }
                        }
                        tc_fence_before();
                        mbar_arrive(&tempty_bar[abuf]);
                    }
                    if (row < nf) p.codes[(size_t)table * p.N + n0 + row] = bidx;
                    if (s + 1 < S) {
                        // r <- r - e[i] (exact fp32, reference order), new scale, new fp16 images;
                        // one warp per frame, lanes across channels (coalesced gathers)
                        const float* cbp = p.cb.p[table];
                        for (int rr = 0; rr < 32; ++rr) {
                            const int urow = q * 32 + rr;
                            const int idx = __shfl_sync(0xffffffffu, bidx, rr);
                            if (urow >= nf) continue;        // tail rows stay zero
                            const float* erow = cbp + (size_t)idx * Dg;
                            float* rrow = R + (size_t)urow * D + g * Dg;
                            float4 rn[4];                    // Dg <= 512: up to 4 x 128 channels per lane
                            float m = 0.f;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int d = lane * 4 + 128 * j;
                                if (d < Dg) {
                                    const float4 e = __ldg(reinterpret_cast<const float4*>(erow + d));
                                    float4 r = *reinterpret_cast<const float4*>(rrow + d);
                                    if (ste) {
                                        r.x = __fsub_rn(r.x, __fadd_rn(r.x, __fsub_rn(e.x, r.x)));
                                        r.y = __fsub_rn(r.y, __fadd_rn(r.y, __fsub_rn(e.y, r.y)));
                                        r.z = __fsub_rn(r.z, __fadd_rn(r.z, __fsub_rn(e.z, r.z)));
                                        r.w = __fsub_rn(r.w, __fadd_rn(r.w, __fsub_rn(e.w, r.w)));
                                    } else {
                                        r.x = __fsub_rn(r.x, e.x); r.y = __fsub_rn(r.y, e.y);
                                        r.z = __fsub_rn(r.z, e.z); r.w = __fsub_rn(r.w, e.w);
                                    }
                                    *reinterpret_cast<float4*>(rrow + d) = r;
                                    rn[j] = r;
                                    m = fmaxf(m, fmaxf(fmaxf(fabsf(r.x), fabsf(r.y)), fmaxf(fabsf(r.z), fabsf(r.w))));
                                }
                            }
#pragma unroll
                            for (int off = 16; off >= 1; off >>= 1)
                                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
                            const float xs = scale_for(m);
                            if (lane == 0) sc[g * BM + urow] = xs;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int d = lane * 4 + 128 * j;
                                if (d < Dg) {
                                    const float v0 = rn[j].x * xs, v1 = rn[j].y * xs, v2 = rn[j].z * xs, v3 = rn[j].w * xs;
                                    const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1),
                                                 h2 = __float2half_rn(v2), h3 = __float2half_rn(v3);
                                    const uint2 hi = make_uint2(pack_half2(h0, h1), pack_half2(h2, h3));
                                    const uint2 lo = make_uint2(
                                        pack_half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1))),
                                        pack_half2(__float2half_rn(v2 - __half2float(h2)), __float2half_rn(v3 - __half2float(h3))));
                                    const int dd = g * Dg + d;        // channel within the full latent
                                    uint8_t* dst = img + (size_t)(dd / BK) * 2 * A_BYTES +
                                                   sw_offset(urow, (dd % BK) >> 3) + ((dd & 7) >> 2) * 8;
                                    *reinterpret_cast<uint2*>(dst) = hi;
                                    *reinterpret_cast<uint2*>(dst + A_BYTES) = lo;
                                }
                            }
                        }
                        __syncwarp();
                        fence_proxy_async_global();
                        mbar_arrive(&upd_bar[g]);
                    }
                }
            }
            mbar_arrive(&free_bar[buf]);     // this tile's scratch buffers may be refilled
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------ packing
struct PackParams {
    PtrTable cb;
    int n_tables, K, Dg;
    uint8_t* pack;
    size_t table_stride, img_bytes, hn_bytes;
    __device__ uint8_t* images(int t) const { return pack + (size_t)t * table_stride; }
    __device__ float* hn(int t) const { return reinterpret_cast<float*>(images(t) + img_bytes); }
    __device__ float* cs(int t) const { return reinterpret_cast<float*>(images(t) + img_bytes + hn_bytes); }
    __device__ uint32_t* maxbits(int t) const { return reinterpret_cast<uint32_t*>(cs(t)) + 1; }
};

__global__ void pack_max_kernel(PackParams p) {
    const int t = blockIdx.y;
    const size_t n = (size_t)p.K * p.Dg;
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(__ldg(p.cb.p[t] + i)));
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) atomicMax(p.maxbits(t), __float_as_uint(m));
}

__global__ void pack_images_kernel(PackParams p) {
    // one thread per 16-byte chunk (8 channels) of one codeword
    const int t = blockIdx.y;
    const int NKC = p.Dg / BK, NP = p.K / BN;
    const int chunks_per_row = p.Dg / 8;
    const size_t total = (size_t)p.K * chunks_per_row;
    const float cs = scale_for(__uint_as_float(*p.maxbits(t)));
    if (blockIdx.x == 0 && threadIdx.x == 0) *p.cs(t) = cs;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i / chunks_per_row);
        const int ch = (int)(i % chunks_per_row);      // 8-channel chunk index within the codeword
        const float* src = p.cb.p[t] + (size_t)k * p.Dg + ch * 8;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float v0 = __ldg(src + 2 * j) * cs, v1 = __ldg(src + 2 * j + 1) * cs;
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            const __half l0 = __float2half_rn(v0 - __half2float(h0));
            const __half l1 = __float2half_rn(v1 - __half2float(h1));
            hi[j] = pack_half2(h0, h1);
            lo[j] = pack_half2(l0, l1);
        }
        const int pass = k / BN, r = k % BN;
        const int kc = ch / CPR, c = ch % CPR;
        uint8_t* blk = p.images(t) + ((size_t)(pass * NKC + kc)) * 2 * B_BYTES;
        (void)NP;
        const uint32_t off = sw_offset(r, c);
        *reinterpret_cast<uint4*>(blk + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(blk + B_BYTES + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

__global__ void pack_norms_kernel(PackParams p) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= p.n_tables * p.K) return;
    const int t = warp / p.K;
    const float* row = p.cb.p[t] + (size_t)(warp % p.K) * p.Dg;
    double acc = 0.0;
    for (int d = lane; d < p.Dg; d += 32) {
        const double v = (double)__ldg(row + d);
        acc = fma(v, v, acc);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) p.hn(t)[warp % p.K] = (float)(0.5 * acc) * scale_for(__uint_as_float(*p.maxbits(t)));
}

size_t images_bytes(int K, int Dg) { return (size_t)(K / BN) * (Dg / BK) * 2 * B_BYTES; }

__global__ void pack_clear_kernel(PackParams p) {
    if (threadIdx.x < p.n_tables) *p.maxbits(threadIdx.x) = 0u;
}

}  // namespace

// pack buffer, per table (so that any contiguous range of tables is itself a valid pack):
//   [images][hn: K f32, 256 B aligned][cs f32 | max bits u32 | pad to 256 B]
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
static size_t table_stride_bytes(int K, int Dg) {
    return align256(images_bytes(K, Dg)) + align256((size_t)K * 4) + 256;
}

size_t tc_pack_bytes(int n_tables, int K, int Dg) { return (size_t)n_tables * table_stride_bytes(K, Dg); }

bool rvq_search_tc_supported(int S, int G, int K, int D, int flags, const char** why) {
    if (G < 1 || G > GMAX || D % G) { *why = "groups"; return false; }
    if (K % BN) { *why = "codebook size must be a multiple of 256"; return false; }
    if ((D / G) % 64) { *why = "channels per group must be a multiple of 64"; return false; }
    if (D / G > 512) { *why = "channels per group must be <= 512"; return false; }
    if (S * G > ACQ_MAX_TABLE) { *why = "too many tables"; return false; }
    (void)flags;
    return true;
}

size_t tc_workspace_bytes(int D) { return (size_t)kNumSMs * 4 * BM * D * sizeof(float) + 256; }

int tc_pack_codebooks(const float* const* cb, int n_tables, int K, int Dg, void* pack,
                      cudaStream_t st) {
    if (K % BN || Dg % 64) return fail(ACQ_ESHAPE, "tc pack: K %% 256 or Dg %% 64 != 0");
    PackParams p;
    for (int i = 0; i < n_tables; ++i) {
        p.cb.p[i] = cb[i];
        if ((uintptr_t)cb[i] % 16) return fail(ACQ_EINVAL, "tc pack: codebook %d not 16-byte aligned", i);
    }
    p.n_tables = n_tables; p.K = K; p.Dg = Dg;
    p.pack = static_cast<uint8_t*>(pack);
    p.table_stride = table_stride_bytes(K, Dg);
    p.img_bytes = align256(images_bytes(K, Dg));
    p.hn_bytes = align256((size_t)K * 4);
    pack_clear_kernel<<<1, ACQ_MAX_TABLE, 0, st>>>(p);
    pack_max_kernel<<<dim3(32, n_tables), 256, 0, st>>>(p);
    pack_images_kernel<<<dim3(64, n_tables), 256, 0, st>>>(p);
    const long long warps = (long long)n_tables * K;
    pack_norms_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(p);
    return check_cuda(cudaGetLastError(), "tc pack launch");
}

int rvq_search_tc(const float* x, const float* const* cb, const void* pack, void* workspace, int S,
                  int G, int K, int D, int B, int T, int flags, int64_t* codes, float* dbg_scores,
                  cudaStream_t st) {
    const char* why = "";
    if (!rvq_search_tc_supported(S, G, K, D, flags, &why)) return fail(ACQ_ESHAPE, "tc search: %s", why);
    if (!pack || !workspace) return fail(ACQ_EINVAL, "tc search: pack/workspace missing");
    TcParams p;
    p.x = x;
    for (int i = 0; i < S * G; ++i) p.cb.p[i] = cb[i];
    const int Dg = D / G;
    p.pack = static_cast<const uint8_t*>(pack);
    p.table_stride = table_stride_bytes(K, Dg);
    p.img_bytes = align256(images_bytes(K, Dg));
    p.scratch = static_cast<float*>(workspace);
    p.S = S; p.G = G; p.K = K; p.D = D; p.Dg = Dg; p.T = T; p.flags = flags;
    p.N = (long long)B * T;
    p.num_tiles = (int)((p.N + BM - 1) / BM);
    p.codes = codes;
    p.dbg_scores = dbg_scores;
    { const char* e = getenv("ACQ_TC_DBG"); p.dbg_mode = e ? atoi(e) : 0; }
    p.err = reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + (size_t)kNumSMs * 4 * BM * D * sizeof(float));
    cudaError_t e = cudaFuncSetAttribute(rvq_search_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(rvq_search_tc)");
    const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
    rvq_search_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(p);
    return check_cuda(cudaGetLastError(), "rvq_search_tc launch");
}

}  // namespace acq
