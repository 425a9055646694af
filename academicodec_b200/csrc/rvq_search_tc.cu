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
//   warps 0-3  loaders: read upcoming tiles of x (coalesced along frames), derive the per-frame
//              scales, split to fp16 hi/lo and write the tile's K-major SWIZZLE_64B operand images
//              into per-CTA global scratch (L2), up to a tile pair ahead of the MMAs so the HBM
//              read overlaps tensor work; in multi-stage calls they spend their waits on the
//              residual-update jobs published by the epilogue (see steal_updates below)
//   warp 8     TMA producer: one thread streams A (residual) and B (pre-packed codebook) images,
//              already in the UMMA shared-memory layout, with cp.async.bulk into a 3 x 48 KiB ring
//   warp 9     MMA issuer: one thread issues 6 tcgen05.mma (M128 N256 K16) per ring stage into one
//              of two 256-column TMEM accumulators; tcgen05.commit frees the stage / publishes the
//              pass
//   warps 4-7  epilogue: tcgen05.ld the accumulator (thread = frame, 32 columns at a time), add
//              the scaled -0.5||e||^2 bias, four (value, index) argmax chains with the lowest-index
//              tie rule, overlapping the next pass's MMAs through the second accumulator; between
//              stages they publish the winning codes as an update job: r <- r - e[i] in fp32
//              exactly as the reference does (core_vq.py:359 / :304), new row scale, new images
//              -- batches of rows claimed by whichever loader / epilogue warp has nothing else to do.
//
// This kernel produces codes only; quantized / loss / EMA outputs come from the replay kernel
// (rvq_replay.cu), the fused SIMT kernel (rvq_search_simt.cu) or from decode.
// Shapes: K % 256 == 0, K <= 1024, (D/G) % 64 == 0, D/G <= 512, G <= 4.  Measured behaviour, the
// experiment log and the power-limit finding are in DESIGN.md section 4 (K1b).
#include "tc_common.cuh"
#include <stdlib.h>

namespace acq {
namespace {

using namespace tc;

constexpr int NSTAGE = ROWB == 64 ? 3 : 2;
constexpr int STAGING_BYTES = 64 * BM * 4;   // one 64-channel slice of a tile's hi+lo images (32 KiB)
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // 48 / 96 KiB
constexpr int NUM_THREADS = 320;
#ifndef ACQ_TC_NI
#define ACQ_TC_NI 2
#endif
constexpr int NI = ACQ_TC_NI;              // tiles of a CTA whose residual stages are interleaved (multi-stage calls)
constexpr int NTB = 2 * NI;        // tile buffers per CTA: NI tiles in flight + NI being loaded
constexpr int BAR_BYTES = (2 * NSTAGE + 4 + 2 * NTB + NI * GMAX) * 8;   // mbarriers
constexpr int UPD_BYTES = BM * 4 /*winning codes of the tile*/ + 80 /*job descriptor*/ + 16 /*claim, completed, all_done, job seq*/;
constexpr int XCHG_BYTES = 2 * 4 * BM * 8 /*[parity][pass owner][row] (value, index)*/ + 16 /*exchange barrier*/;
constexpr int CTRL_BYTES = BAR_BYTES + 16 /*tmem ptr*/ + NTB * GMAX * BM * 4 /*row scales per tile buffer*/ +
                           GMAX * BM * 4 /*row max bits*/ + KMAX * 4 /*scaled norms of the current table*/ + UPD_BYTES +
                           XCHG_BYTES;
constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + (size_t)NSTAGE * STAGE_BYTES + STAGING_BYTES + CTRL_BYTES;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// ---- shared residual update ---------------------------------------------------------------------
// Between two stages the 128 rows of a tile are updated in batches of RB rows (tc_common.cuh).  With
// the four epilogue warps alone this phase is as long as the tile's MMAs at D = 512 and twice as long
// at D = 128 (ACQ_TC_DBG=512: the MMA thread waited 42 % of a cfg1 launch for a drained accumulator),
// while the four loader warps sit in a buffer wait 86 % of the time.  So the batches are work items:
// the epilogue publishes a job (descriptor + winning codes in shared memory, then claim <- 0), every
// epilogue warp and every idle loader warp claims batches with an atomic counter until none is left,
// and whoever completes the last batch arrives on the image-ready barrier the TMA thread waits on.
// The epilogue does not wait for the job: it goes straight back to draining accumulators (its sweeps
// alone fit under the MMAs of a stage) and only claims batches while it would otherwise spin on an
// accumulator barrier, or when the previous job is still open at the next publication.
struct UpdJob {
    const float* cbp;
    float* R;
    uint8_t* img;
    float* sc_g;
    uint64_t* bar;           // image-ready barrier of this (tile parity, group): the last finisher arrives
    int Dg, D, g, nf, ste;
    int item0;               // split mode: this CTA owns batches item0 .. item0 + items - 1 of the tile ...
    int ncta;                //             ... and writes the images of all ncta CTAs of its cluster
};
static_assert(sizeof(UpdJob) <= 72, "job descriptor slot");

__device__ __forceinline__ int upd_items(int Dg) { return Dg <= 128 ? BM / 8 : (Dg <= 256 ? BM / 4 : BM / 2); }

// Claim and process update batches until none is left; returns immediately when no job is open.
// (`budget`: the epilogue warps take one batch at a time and look at their accumulator barrier again)
// Non-suspending, warp-uniform poll for warps that have update batches to work on while a barrier is pending
// (mbarrier.try_wait parks the thread before it reports "not yet": see rvq_search_p1.cu, warp_test_wait).
// ACQ_TC_DBG bit 268435456 restores the parking poll.
__device__ __forceinline__ bool poll_bar(uint64_t* bar, uint32_t parity, bool park) {
    return __all_sync(0xffffffffu, park ? mbar_try_wait(bar, parity) : mbar_test_wait(bar, parity));
}
template <int NDST>
__device__ __forceinline__ int steal_updates(volatile int* st, const UpdJob* job_s, const int* bidx_s, int items,
                                             int lane, int budget = 0x7fffffff) {
    int mine = 0;
    uint64_t* bar = nullptr;
    for (; budget > 0; --budget) {
        int item = 0x7fffffff;
        if (lane == 0 && st[0] < items) item = atomicAdd(const_cast<int*>(st), 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= items) break;
        __threadfence_block();                       // the job and the codes were written before claim <- 0
        const UpdJob j = *job_s;
        bar = j.bar;
        item += j.item0;
        const size_t cs = (size_t)BM * j.D * 4;      // bytes between the scratch tiles of consecutive CTAs
        if (j.Dg <= 128) {
            int idxs[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) idxs[u] = bidx_s[item * 8 + u];
            residual_update_batch<8, 1, true, false, NDST>(item * 8, lane, j.nf, idxs, j.cbp, j.Dg, j.D, j.g, j.R,
                                                           j.img, j.sc_g, nullptr, j.ste != 0, cs);
        } else if (j.Dg <= 256) {
            int idxs[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) idxs[u] = bidx_s[item * 4 + u];
            residual_update_batch<4, 2, true, false, NDST>(item * 4, lane, j.nf, idxs, j.cbp, j.Dg, j.D, j.g, j.R,
                                                           j.img, j.sc_g, nullptr, j.ste != 0, cs);
        } else {
            int idxs[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) idxs[u] = bidx_s[item * 2 + u];
            residual_update_batch<2, 4, true, false, NDST>(item * 2, lane, j.nf, idxs, j.cbp, j.Dg, j.D, j.g, j.R,
                                                           j.img, j.sc_g, nullptr, j.ste != 0, cs);
        }
        ++mine;
    }
    if (mine) {
        // one cross-proxy fence for all the batches this call finished (a job cannot be completed and
        // replaced while a claimed batch is outstanding, so they all belong to the same job)
        __syncwarp();
        if (NDST > 1) { __threadfence(); fence_acq_rel_cluster(); }   // peers read these rows (global + their smem)
        fence_proxy_async_global();                  // this warp's image writes -> the TMA threads' bulk reads
        __threadfence_block();
        if (lane == 0) {
            const int done = atomicAdd(const_cast<int*>(st + 1), mine) + mine;
            if (done == items) {
                __threadfence_block();
                if (NDST == 1) {
                    mbar_arrive(bar);                // the whole image of the next stage is in place
                } else {
                    // this CTA's share is in place everywhere: tell every CTA of the cluster
                    __threadfence();
#pragma unroll
                    for (int c = 0; c < NDST; ++c) mbar_arrive_remote(mapa_u32(smem_u32(bar), (uint32_t)c));
                }
            }
        }
    }
    return mine;
}
// Finish whatever is left of the open job (no-op when none is open).
template <int NDST>
__device__ __forceinline__ void drain_updates(volatile int* st, const UpdJob* job_s, const int* bidx_s, int items,
                                              int lane, int* err) {
    const long long tw = clock64();
    for (;;) {
        steal_updates<NDST>(st, job_s, bidx_s, items, lane);
        if (st[1] >= items) return;
        if (clock64() - tw > 8000000000LL) { if (err) atomicExch(err, 11); __trap(); }
    }
}

// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, N=256, M=128
//   [4,6) c_format=1(F32)  [7,10) a_format=0(F16)  [10,13) b_format=0(F16)
//   [15] a_major=0(K)  [16] b_major=0(K)  [17,23) N>>3  [24,29) M>>4
//
// CL > 1: the CTAs of a cluster of CL run their tile sequences in lockstep and share the codebook
// stream -- CTA r fetches slice r of every B stage and multicasts it to all CL ring slots, so the
// L2 -> SMEM traffic of B drops by CL (the operand stream, not the tensor pipe, bounds this kernel).
// A ring slot is refilled only after the MMAs of every CTA of the cluster have retired it
// (tcgen05.commit multicast on the empty barriers); a CTA whose tile list is one shorter than its
// leader's runs a dummy tile (no valid rows) to stay in step.
//
// SPLIT (small batches: every tile gets a cluster of CL = 4 or 2 CTAs): CTA r of the cluster runs only
// its share of the codebook passes of its tile, so the serial chain of a stage carries a quarter of the MMAs; the
// partial (value, index) maxima are exchanged through distributed shared memory, every CTA merges
// them to the same winner and carries its own copy of the residual (the update is replicated, not
// split -- it stays off the other CTAs' critical path and needs no further exchange).
template <int CL, bool SPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1) rvq_search_tc_kernel(const TcParams p) {
    if (guard_skips(p)) return;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* staging = smem + NSTAGE * STAGE_BYTES;                 // loaders' image slice (coalescing buffer)
    uint8_t* ctrl = staging + STAGING_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);          // [NSTAGE] TMA bytes landed
    uint64_t* empty_bar = full_bar + NSTAGE;                         // [NSTAGE] MMAs retired
    uint64_t* tfull_bar = empty_bar + NSTAGE;                        // [2] accumulator complete
    uint64_t* tempty_bar = tfull_bar + 2;                            // [2] accumulator drained
    uint64_t* t0_bar = tempty_bar + 2;                               // [NTB] stage-0 images of a tile ready
    uint64_t* free_bar = t0_bar + NTB;                               // [NTB] tile buffer reusable
    uint64_t* upd_bar = free_bar + NTB;                              // [NI][GMAX] next-stage image ready
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(ctrl + BAR_BYTES);
    float* scale_s = reinterpret_cast<float*>(ctrl + BAR_BYTES + 16);             // [NTB][GMAX][BM]
    uint32_t* rowmax_s = reinterpret_cast<uint32_t*>(ctrl + BAR_BYTES + 16 + NTB * GMAX * BM * 4);  // [GMAX][BM]
    float* hn_s = reinterpret_cast<float*>(ctrl + BAR_BYTES + 16 + (NTB + 1) * GMAX * BM * 4);       // [KMAX]
    int* bidx_s = reinterpret_cast<int*>(hn_s + KMAX);                                               // [BM]
    UpdJob* job_s = reinterpret_cast<UpdJob*>(bidx_s + BM);
    volatile int* upd_state = reinterpret_cast<volatile int*>(reinterpret_cast<uint8_t*>(job_s) + 80);  // claim, completed, all_done
    uint64_t* xchg_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(job_s) + 80 + 16);
    uint2* xchg_s = reinterpret_cast<uint2*>(xchg_bar + 2);                                         // [2][4][BM]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = p.S, G = p.G, K = p.K, D = p.D, Dg = p.Dg, T = p.T;
    const int NP = K / BN, NKC = Dg / BK;
    const bool ste = p.flags & ACQ_STE;
    // Helpers between two batches: the LOADERS poll their buffer barrier without suspending (mbarrier.try_wait parks a
    // thread before it reports "not yet", and a parked helper does not help), the EPILOGUE warps keep the parking
    // poll -- with the other one they pick up a batch of 8 rows (~11 kcycles at D = 128) just before their
    // accumulator completes and the MMA thread waits for the drain instead.  cfg1 B=4096, Mcycles per CTA / ms:
    // both park 4.15 / 2.84 (round 1), both poll 3.96 / 2.81, loaders poll 3.64 / 2.63, epilogue polls 4.20 / 2.93
    // (profiles/r05w_sweep_v3_poll.log).  ACQ_TC_DBG bits 268435456 / 1073741824 flip the epilogue / the loaders.
    const bool park = (p.dbg_mode & 268435456) == 0;
    const bool park_ld = (p.dbg_mode & 1073741824) != 0;
    const size_t tile_elems = (size_t)BM * D;
    // scratch layout is buffer-major -- [buf][CTA] images, then [buf][CTA] fp32 rows -- so that the part a
    // single-stage call touches (image buffers 0 and 1 of every CTA) is one contiguous 76 MB range
    const size_t buf_stride = (size_t)kNumSMs * tile_elems * 4;           // bytes between tile buffers
    uint8_t* Aimg = reinterpret_cast<uint8_t*>(p.scratch) + (size_t)blockIdx.x * tile_elems * 4;
    float* Rbuf = reinterpret_cast<float*>(Aimg + NTB * buf_stride);
    // tiles of this CTA are blockIdx.x + it * gridDim.x, it = 0 .. n_my-1.  They are processed in
    // groups of NI with their residual stages interleaved -- (A,s0) (B,s0) (C,s0) (A,s1) (B,s1) ... --
    // so that the epilogue / residual update of one tile overlaps the MMAs of the others.  With pairs the
    // turnaround of a tile (last sweep + update + first operand copy ~ 12 kcycles at D = 128) just about
    // equalled the other tile's MMAs (13.4 kcycles) and the MMA thread still waited 30 % of a cfg1
    // launch for next-stage images; with three tiles there are two stages of slack.
    // single-stage calls keep the scratch working set small (it must stay L2 resident): 2 tile buffers
    const uint32_t ntb = S * G == 1 ? 2u : (uint32_t)NTB;
    const uint32_t ni = S * G == 1 ? 2u : (uint32_t)NI;
    // (cluster-uniform: the tile count of the cluster's first CTA, which is the largest)
    const int lead_cta = (int)(blockIdx.x / CL) * CL;
    const uint32_t n_my = SPLIT ? 1u
                          : (p.num_tiles > lead_cta
                                 ? (uint32_t)((p.num_tiles - 1 - lead_cta) / (int)gridDim.x + 1) : 0u);
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    // tile index of this CTA's it-th tile = tile_base + it * tile_stride
    const long long tile_base = SPLIT ? (long long)(blockIdx.x / CL) : (long long)blockIdx.x;
    const long long tile_stride = SPLIT ? (long long)(gridDim.x / CL) : (long long)gridDim.x;
    constexpr int ND = SPLIT ? CL : 1;                   // CTAs whose operand images an update batch writes
    const int items_cta = upd_items(Dg) / ND;            // update batches this CTA owns per job
    const int p0 = SPLIT ? (int)crank * (NP / CL) : 0, p1 = SPLIT ? p0 + NP / CL : NP;   // codebook passes of this CTA
    constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1);

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&full_bar[i], 1);         // the TMA thread's arrive.expect_tx
            mbar_init(&empty_bar[i], SPLIT ? 1 : CL);   // tcgen05.commit of every CTA sharing the B stream
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);        // tcgen05.commit
            mbar_init(&tempty_bar[i], 128);     // epilogue threads
        }
        for (int i = 0; i < NTB; ++i) {
            mbar_init(&t0_bar[i], 128);         // loader threads
            mbar_init(&free_bar[i], 128);       // epilogue threads
        }
        for (int i = 0; i < NI * GMAX; ++i) mbar_init(&upd_bar[i], ND);   // the warp(s) that complete an update job
        upd_state[0] = 0x7fffffff;      // no update job open
        upd_state[1] = 0x7fffffff;      // ... and nothing to wait for
        upd_state[2] = 0;
        upd_state[3] = 0;
        mbar_init(xchg_bar, CL * 4);        // one arrival per epilogue warp of every CTA of the cluster
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(tmem_ptr_s, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();      // every CTA's barriers are initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (warp < 4) {
        // ================= loaders: x tile -> scales, fp16 hi/lo images (and R when S > 1) ========
        // Run up to a pair of tiles ahead of the MMAs (NTB scratch buffers), so the HBM read of the next
        // tiles overlaps the tensor work of the current ones.
        unsigned long long w_free = 0;
        const long long t_begin = clock64();
        for (uint32_t it = 0; it < n_my; ++it) {
            const long long tile = tile_base + (long long)it * tile_stride;   // may be a dummy past the end
            const uint32_t buf = it % ntb;
            if (S > 1 && !poll_bar(&free_bar[buf], ((it / ntb) & 1) ^ 1, park_ld)) {
                // no buffer to fill yet: help the epilogue with the residual updates meanwhile
                const long long tw = clock64();
                while (!poll_bar(&free_bar[buf], ((it / ntb) & 1) ^ 1, park_ld)) {
                    // (sleep when there is nothing to claim: a hot polling loop takes issue slots from the
                    //  epilogue warp that shares this scheduler)
                    if (!steal_updates<ND>(upd_state, job_s, bidx_s, items_cta, lane)) __nanosleep(128);
                    if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 6); __trap(); }
                }
                w_free += (unsigned long long)(clock64() - tw);
            }
            mbar_wait_t(&free_bar[buf], ((it / ntb) & 1) ^ 1, p.err, 6, w_free);
            if ((p.dbg_mode & 1) && it >= ntb) { mbar_arrive(&t0_bar[buf]); continue; }
            const long long n0 = tile * BM;
            uint8_t* img = Aimg + buf * buf_stride;
            float* R = Rbuf + buf * (buf_stride / 4);
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
                for (int sweep = 0; sweep < ((p.dbg_mode & 128) ? 1 : 2); ++sweep) {
                    // two ADJACENT channel octets per iteration: 16 independent 16-byte loads in flight
                    // per thread, and every image / residual store is a whole 32-byte sector
                    for (int pr = w4; pr < D / 16; pr += 4) {
                        float4 v[2][8];
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                v[h][i] = ok ? __ldg(reinterpret_cast<const float4*>(src + (size_t)(pr * 16 + h * 8 + i) * T))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
                        const int oct = 2 * pr;                 // even octet; both lie in the same group/chunk
                        const int g = (oct * 8) / Dg;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float a[2][8];
#pragma unroll
                            for (int h = 0; h < 2; ++h)
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    a[h][i] = j == 0 ? v[h][i].x : (j == 1 ? v[h][i].y : (j == 2 ? v[h][i].z : v[h][i].w));
                            const int row = 4 * rq + j;
                            if (sweep == 0) {
                                float m = 0.f;
#pragma unroll
                                for (int i = 0; i < 8; ++i) m = fmaxf(m, fmaxf(fabsf(a[0][i]), fabsf(a[1][i])));
                                atomicMax(&rowmax_s[g * BM + row], __float_as_uint(m));
                            } else {
                                uint4 hi0, lo0, hi1, lo1;
                                const float xs = sc[g * BM + row];
                                split8(a[0], xs, hi0, lo0);
                                split8(a[1], xs, hi1, lo1);
                                // The four loader warps cover one 64-channel slice per iteration; its hi/lo
                                // images form one contiguous 32 KiB range of the global image.  Scattering
                                // them from here (one sector per frame row) was measured to cost 0.29 ms of a
                                // 1.06 ms kernel, so they are assembled in shared memory and copied out with
                                // fully coalesced stores below.
                                const int oct_in = 2 * w4;      // this warp's (even) octet inside the 8-octet slice
                                uint8_t* cb_ = staging + (size_t)(oct_in / CPR) * 2 * A_BYTES;
                                {
                                    const int c = oct_in % CPR;
                                    *reinterpret_cast<uint4*>(cb_ + sw_offset(row, c)) = hi0;
                                    *reinterpret_cast<uint4*>(cb_ + sw_offset(row, c + 1)) = hi1;
                                    *reinterpret_cast<uint4*>(cb_ + A_BYTES + sw_offset(row, c)) = lo0;
                                    *reinterpret_cast<uint4*>(cb_ + A_BYTES + sw_offset(row, c + 1)) = lo1;
                                }
                                if (S > 1) {
                                    float* rd = R + (size_t)row * D + oct * 8;
                                    stg256(rd, make_uint4(__float_as_uint(a[0][0]), __float_as_uint(a[0][1]), __float_as_uint(a[0][2]), __float_as_uint(a[0][3])),
                                           make_uint4(__float_as_uint(a[0][4]), __float_as_uint(a[0][5]), __float_as_uint(a[0][6]), __float_as_uint(a[0][7])));
                                    stg256(rd + 8, make_uint4(__float_as_uint(a[1][0]), __float_as_uint(a[1][1]), __float_as_uint(a[1][2]), __float_as_uint(a[1][3])),
                                           make_uint4(__float_as_uint(a[1][4]), __float_as_uint(a[1][5]), __float_as_uint(a[1][6]), __float_as_uint(a[1][7])));
                                }
                            }
                        }
                        if (sweep == 1) {
                            // copy the assembled slice out: 32 KiB contiguous, 512 B per warp store
                            named_bar_sync(2, 128);
                            const int slice = (pr - w4) / 4;                   // 64-channel slice index
                            uint4* gdst = reinterpret_cast<uint4*>(img + (size_t)slice * STAGING_BYTES);
                            const uint4* ssrc = reinterpret_cast<const uint4*>(staging);
#pragma unroll 4
                            for (int k = tid; k < ((p.dbg_mode & 64) ? 0 : STAGING_BYTES / 16); k += 128) gdst[k] = ssrc[k];
                            named_bar_sync(2, 128);
                        }
                    }
                    if (sweep == 0) {
                        named_bar_sync(2, 128);
                        for (int i = tid; i < G * BM; i += 128) sc[i] = scale_for(__uint_as_float(rowmax_s[i]));
                        named_bar_sync(2, 128);
                    }
                }
            } else {
                // general T (e.g. HiFi-Codec's 50 frames per clip): one frame per thread, scalar loads
                // that are coalesced across the warp.  32 channels (one K-chunk of the operand image) per
                // iteration: 32 independent loads in flight per thread and whole 32-byte sectors on the
                // store side -- the first version (8 loads in flight, 16-byte stores) made the loaders
                // the bottleneck of cfg3 (214 kcycles per tile against 108 kcycles of MMAs; 64 loads in
                // flight were measured too and are no faster than 32).
                const int row = tid;
                const long long n = n0 + row;
                const bool ok = n < p.N;
                const long long b = ok ? n / T : 0, t = ok ? n % T : 0;
                const float* src = p.x + (size_t)(b * D) * T + t;
                for (int sweep = 0; sweep < 2; ++sweep) {
                    float m = 0.f;
                    for (int o4 = 0; o4 < D / 8; o4 += 4) {
                        float a[4][8];
#pragma unroll
                        for (int h = 0; h < 4; ++h)
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                a[h][i] = ok ? __ldg(src + (size_t)((o4 + h) * 8 + i) * T) : 0.f;
                        const int g = (o4 * 8) / Dg;             // Dg % 64 == 0: the 32 channels share a group
                        if (sweep == 0) {
#pragma unroll
                            for (int h = 0; h < 4; ++h)
#pragma unroll
                                for (int i = 0; i < 8; ++i) m = fmaxf(m, fabsf(a[h][i]));
                            if (((o4 + 4) * 8) % Dg == 0) {      // last chunk of the group: the row is this thread's
                                rowmax_s[g * BM + row] = __float_as_uint(m);
                                m = 0.f;
                            }
                        } else {
                            const float xs = sc[g * BM + row];
                            uint8_t* chunk = img + (size_t)(o4 / CPR) * 2 * A_BYTES;
#pragma unroll
                            for (int h = 0; h < 4; h += 2) {
                                uint4 hi0, lo0, hi1, lo1;
                                split8(a[h], xs, hi0, lo0);
                                split8(a[h + 1], xs, hi1, lo1);
                                store_chunk_pair(chunk, row, h, hi0, hi1);
                                store_chunk_pair(chunk + A_BYTES, row, h, lo0, lo1);
                            }
                            if (S > 1) {
                                float* rd = R + (size_t)row * D + o4 * 8;
#pragma unroll
                                for (int h = 0; h < 4; ++h)
                                    stg256(rd + h * 8,
                                           make_uint4(__float_as_uint(a[h][0]), __float_as_uint(a[h][1]), __float_as_uint(a[h][2]), __float_as_uint(a[h][3])),
                                           make_uint4(__float_as_uint(a[h][4]), __float_as_uint(a[h][5]), __float_as_uint(a[h][6]), __float_as_uint(a[h][7])));
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
        if ((p.dbg_mode & 512) && tid == 0) {
            atomicAdd(p.stall + 5, w_free);
            atomicAdd(p.stall + 8, (unsigned long long)(clock64() - t_begin));
        }
        if (S > 1) {
            // all tiles loaded: keep helping until the epilogue has finished its last tile
            const long long tw = clock64();
            while (upd_state[2] == 0) {
                if (!steal_updates<ND>(upd_state, job_s, bidx_s, items_cta, lane)) __nanosleep(128);
                if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 10); __trap(); }
            }
        }
    } else if (warp == 8) {
        // ================= TMA producer: one thread streams A and B operand images ================
        if (lane == 0) {
            uint32_t ring_it = 0, upd_it[NI * GMAX];
            unsigned long long w_empty = 0, w_t0 = 0;
#pragma unroll
            for (int i = 0; i < NI * GMAX; ++i) upd_it[i] = 0;
            for (uint32_t it0 = 0; it0 < n_my; it0 += ni) {
                const int npair = (int)min(ni, n_my - it0);
                for (int s = 0; s < S; ++s) {
                    for (int h = 0; h < npair; ++h) {
                        const uint32_t it = it0 + h, buf = it % ntb, par = h;
                        // (ACQ_TC_DBG bit 256: read the operand images from buffers nobody writes -- experiment)
                        const uint8_t* img = Aimg + (buf + ((p.dbg_mode & 256) ? 2 : 0)) * buf_stride;
                        for (int g = 0; g < G; ++g) {
                            const uint8_t* bimg = p.pack + (size_t)(s * G + g) * p.table_stride;
                            for (int pass = p0; pass < p1; ++pass) {
                                for (int kc = 0; kc < NKC; ++kc, ++ring_it) {
                                    const int st = ring_it % NSTAGE;
                                    mbar_wait_t(&empty_bar[st], ((ring_it / NSTAGE) & 1) ^ 1, p.err, 2, w_empty);
                                    uint8_t* a_dst = smem + st * STAGE_BYTES;
                                    uint8_t* b_dst = a_dst + 2 * A_BYTES;
                                    const uint8_t* bsrc = bimg + (size_t)(pass * NKC + kc) * 2 * B_BYTES;
                                    const bool skip_b = p.dbg_mode & 2, skip_a = p.dbg_mode & 4;
                                    mbar_arrive_expect_tx(&full_bar[st], (skip_a ? 0 : 2 * A_BYTES) + (skip_b ? 0 : 2 * B_BYTES));
                                    if (!skip_b) {
                                        if (CL == 1 || SPLIT) {
                                            bulk_g2s(b_dst, bsrc, B_BYTES, &full_bar[st]);
                                            bulk_g2s(b_dst + B_BYTES, bsrc + B_BYTES, B_BYTES, &full_bar[st]);
                                        } else {
                                            constexpr uint32_t SLICE = 2 * B_BYTES / CL;
                                            bulk_g2s_mc(b_dst + crank * SLICE, bsrc + crank * SLICE, SLICE,
                                                        &full_bar[st], CMASK);
                                        }
                                    }
                                    if (pass == p0 && kc == 0) {
                                        // first use of this (tile, stage, group)'s residual image
                                        if (s == 0) {
                                            mbar_wait_t(&t0_bar[buf], (it / ntb) & 1, p.err, 7, w_t0);
                                        } else {
                                            if (SPLIT) {
                                                const long long tw = clock64();
                                                while (!mbar_try_wait_cluster(&upd_bar[par * GMAX + g], upd_it[par * GMAX + g] & 1))
                                                    if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 8); __trap(); }
                                            } else {
                                                mbar_wait(&upd_bar[par * GMAX + g], upd_it[par * GMAX + g] & 1, p.err, 8);
                                            }
                                            ++upd_it[par * GMAX + g];
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
            if (p.dbg_mode & 512) { atomicAdd(p.stall + 3, w_empty); atomicAdd(p.stall + 4, w_t0); }
        }
    } else if (warp == 9) {
        // ================= MMA issuer =================================================================
        if (lane == 0) {
            uint32_t ring_it = 0, acc_it = 0;
            unsigned long long w_full0 = 0, w_full = 0, w_tempty = 0;
            const long long t_begin = clock64();
            for (uint32_t itm = 0; itm < n_my; ++itm) {       // (same number of items in any order)
                for (int sg = 0; sg < S * G; ++sg) {
                    for (int pass = p0; pass < p1; ++pass, ++acc_it) {
                        const uint32_t abuf = acc_it & 1;
                        mbar_wait_t(&tempty_bar[abuf], ((acc_it >> 1) & 1) ^ 1, p.err, 3, w_tempty);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + abuf * BN;
                        for (int kc = 0; kc < NKC; ++kc, ++ring_it) {
                            const int st = ring_it % NSTAGE;
                            mbar_wait_t(&full_bar[st], (ring_it / NSTAGE) & 1, p.err, 4, pass == 0 ? w_full0 : w_full);
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
                            // ring stage free once these MMAs retire (in every CTA that shares it)
                            if (CL == 1 || SPLIT) umma_commit(&empty_bar[st]);
                            else umma_commit_mc(&empty_bar[st], CMASK);
                        }
                        umma_commit(&tfull_bar[abuf]);       // accumulator complete
                    }
                }
            }
            if (p.dbg_mode & 512) {
                atomicAdd(p.stall + 0, w_full0); atomicAdd(p.stall + 1, w_full); atomicAdd(p.stall + 2, w_tempty);
                atomicAdd(p.stall + 7, (unsigned long long)(clock64() - t_begin));
            }
        }
    } else {
        // ================= epilogue + residual update (warps 4-7, thread = frame) ====================
        const int q = warp - 4;
        const int row = q * 32 + lane;
        uint32_t acc_it = 0;
        unsigned long long e_hn = 0, e_wait = 0, e_sweep = 0, e_upd = 0;     // epilogue time split (bit 512)
        int job_seq = 0;
        uint32_t xchg_it = 0;
        for (uint32_t it0 = 0; it0 < n_my; it0 += ni) {
          const int npair = (int)min(ni, n_my - it0);
          for (int s = 0; s < S; ++s) {
            for (int h = 0; h < npair; ++h) {
                const uint32_t it = it0 + h, buf = it % ntb, par = h;
                const long long n0 = (tile_base + (long long)it * tile_stride) * BM;
                const int nf = (int)min((long long)BM, p.N - n0);
                uint8_t* img = Aimg + buf * buf_stride;
                float* R = Rbuf + buf * (buf_stride / 4);
                float* sc = scale_s + buf * GMAX * BM;
                if (s == 0) mbar_wait(&t0_bar[buf], (it / ntb) & 1, p.err, 9);    // this tile's scales are visible
                for (int g = 0; g < G; ++g) {
                    const int table = s * G + g;
                    float nxs = 0.f;          // (read after the stage's first accumulator is complete: the
                                              //  row scales of stage s are written by the update job of s-1)
                    const float* hn = reinterpret_cast<const float*>(p.pack + (size_t)table * p.table_stride +
                                                                     p.img_bytes);
                    // stage this table's scaled norms in shared memory (all four epilogue warps)
                    long long tq = clock64();
                    named_bar_sync(3, 128);
                    for (int i = (tid - 128) * 4; i < K; i += 128 * 4)
                        *reinterpret_cast<float4*>(hn_s + i) = __ldg(reinterpret_cast<const float4*>(hn + i));
                    named_bar_sync(3, 128);
                    e_hn += (unsigned long long)(clock64() - tq);
                    // four independent (value, index) chains (columns mod 4) keep the compare/select
                    // dependency chain short; merged below with the lowest-index tie rule
                    float bv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                    int bi[4] = {0, 1, 2, 3};
                    for (int pass = p0; pass < p1; ++pass, ++acc_it) {
                        const uint32_t abuf = acc_it & 1;
                        if (S > 1 && !poll_bar(&tfull_bar[abuf], (acc_it >> 1) & 1, park)) {
                            // nothing to drain yet: work on the open residual-update job meanwhile
                            const long long tw = clock64();
                            while (!poll_bar(&tfull_bar[abuf], (acc_it >> 1) & 1, park)) {
                                steal_updates<ND>(upd_state, job_s, bidx_s, items_cta, lane, 1);
                                if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 5); __trap(); }
                            }
                            e_wait += (unsigned long long)(clock64() - tw);
                        }
                        mbar_wait_t(&tfull_bar[abuf], (acc_it >> 1) & 1, p.err, 5, e_wait);
                        tc_fence_after();
                        if (pass == p0) nxs = -sc[g * BM + row];
                        tq = clock64();
                        const uint32_t taddr = tmem_base + abuf * BN + ((uint32_t)(q * 32) << 16);
                        const int kbase = pass * BN;
                        if (p.dbg_scores && table == 0) {
                            float* o = p.dbg_scores + (size_t)(n0 + row) * K + kbase;
                            const float inv = 1.0f / -nxs;      // undo the row scale
                            const bool wr = row < nf;
                            for_each_score(taddr, hn_s + kbase, nxs, [&](int c, float sv) {
                                if (sv > bv[c & 3]) { bv[c & 3] = sv; bi[c & 3] = kbase + c; }
                                if (wr) o[c] = sv * inv;
                            });
                        } else {
                            for_each_score(taddr, hn_s + kbase, nxs, [&](int c, float sv) {
                                if (sv > bv[c & 3]) { bv[c & 3] = sv; bi[c & 3] = kbase + c; }
                            });
                        }
                        tc_fence_before();
                        mbar_arrive(&tempty_bar[abuf]);
                        e_sweep += (unsigned long long)(clock64() - tq);
                    }
                    float best = bv[0];
                    int bidx = bi[0];
#pragma unroll
                    for (int u = 1; u < 4; ++u)
                        if (bv[u] > best || (bv[u] == best && bi[u] < bidx)) { best = bv[u]; bidx = bi[u]; }
                    if (SPLIT) {
                        // partial maxima of the CL passes -> every CTA of the cluster, then the same merge
                        // everywhere (ascending pass = ascending index: strict > keeps the lowest index)
                        uint2* mine = xchg_s + ((size_t)xchg_it & 1) * 4 * BM + crank * BM + row;
                        const uint32_t my_addr = smem_u32(mine);
#pragma unroll
                        for (int c = 0; c < CL; ++c) st_cluster_v2(mapa_u32(my_addr, (uint32_t)c), best, bidx);
                        fence_acq_rel_cluster();
                        __syncwarp();
                        if (lane == 0) {
#pragma unroll
                            for (int c = 0; c < CL; ++c) mbar_arrive_remote(mapa_u32(smem_u32(xchg_bar), (uint32_t)c));
                        }
                        {
                            const long long tw = clock64();
                            while (!mbar_try_wait_cluster(xchg_bar, xchg_it & 1)) {
                                if (S > 1) steal_updates<ND>(upd_state, job_s, bidx_s, items_cta, lane, 1);
                                if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 13); __trap(); }
                            }
                        }
                        const uint2* all = xchg_s + ((size_t)xchg_it & 1) * 4 * BM + row;
                        best = __uint_as_float(all[0].x);
                        bidx = (int)all[0].y;
#pragma unroll
                        for (int c = 1; c < CL; ++c) {
                            const uint2 o = all[c * BM];
                            if (__uint_as_float(o.x) > best) { best = __uint_as_float(o.x); bidx = (int)o.y; }
                        }
                        ++xchg_it;
                    }
                    if (row < nf && (!SPLIT || crank == 0)) p.codes[(size_t)table * p.N + n0 + row] = bidx;
                    if (s + 1 < S) {
                        // r <- r - e[i] (exact fp32, reference order), new scale, new fp16 images;
                        // one warp per frame, lanes across channels (coalesced gathers)
                        tq = clock64();
                        const int items = items_cta;
                        drain_updates<ND>(upd_state, job_s, bidx_s, items, lane, p.err);   // previous job (other tile / group)
                        bidx_s[row] = bidx;
                        named_bar_sync(3, 128);                    // all 128 codes are in shared memory
                        if (tid == 128) {
                            job_s->cbp = p.cb.p[table]; job_s->R = R; job_s->sc_g = sc + g * BM;
                            job_s->img = img - (size_t)(SPLIT ? crank : 0) * tile_elems * 4;   // rank 0's copy
                            job_s->item0 = (SPLIT ? (int)crank : 0) * items_cta; job_s->ncta = ND;
                            job_s->bar = &upd_bar[par * GMAX + g];
                            job_s->Dg = Dg; job_s->D = D; job_s->g = g; job_s->nf = nf; job_s->ste = ste ? 1 : 0;
                            upd_state[1] = 0;
                            __threadfence_block();
                            upd_state[0] = 0;                      // opens the job: batches can be claimed
                            __threadfence_block();
                            upd_state[3] = job_seq + 1;
                        }
                        ++job_seq;
                        // If the MMA thread has no accumulator for us (single tile per CTA, or the other
                        // tiles are waiting for their own images), start on the job right away.
                        if (!poll_bar(&tfull_bar[acc_it & 1], (acc_it >> 1) & 1, park)) {
                            const long long tw = clock64();
                            while (upd_state[3] != job_seq) {
                                if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 12); __trap(); }
                            }
                            while (!poll_bar(&tfull_bar[acc_it & 1], (acc_it >> 1) & 1, park) && __any_sync(0xffffffffu, upd_state[0] < items))
                                steal_updates<ND>(upd_state, job_s, bidx_s, items, lane, 1);
                        }
                        e_upd += (unsigned long long)(clock64() - tq);
                    }
                }
                if (s == S - 1) mbar_arrive(&free_bar[buf]);     // this tile's scratch buffer may be refilled
            }
          }
        }
        if (S > 1) {
            named_bar_sync(3, 128);                              // (the last stage of a tile opens no job)
            drain_updates<ND>(upd_state, job_s, bidx_s, items_cta, lane, p.err);
            named_bar_sync(3, 128);
            if (tid == 128) upd_state[2] = 1;                    // the helping loader warps may retire
        }
        if ((p.dbg_mode & 512) && tid == 128) {
            atomicAdd(p.stall + 9, e_hn); atomicAdd(p.stall + 10, e_wait);
            atomicAdd(p.stall + 11, e_sweep); atomicAdd(p.stall + 12, e_upd);
        }
    }

    tc_fence_before();
    __syncthreads();
    // no CTA may retire while a peer can still multicast into its ring or arrive on its barriers
    if (CL > 1) cluster_sync_all();
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
    __device__ uint32_t* tail(int t, int slot) const { return reinterpret_cast<uint32_t*>(cs(t)) + slot; }
    __device__ uint8_t* bias(int t) const { return images(t) + img_bytes + hn_bytes + 256; }
    __device__ uint32_t* maxbits(int t) const { return tail(t, TAIL_MAXBITS); }
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
    const float cs = scale_for(__uint_as_float(*p.maxbits(t)));
    double acc = 0.0, del = 0.0;
    for (int d = lane; d < p.Dg; d += 32) {
        const float f = __ldg(row + d);
        const double v = (double)f;
        acc = fma(v, v, acc);
        const float sv = f * cs;                                   // exact (power-of-two scale)
        const double dv = (double)sv - (double)__half2float(__float2half_rn(sv));
        del = fma(dv, dv, del);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, off);
        del += __shfl_xor_sync(0xffffffffu, del, off);
    }
    if (lane == 0) {
        const float hn = (float)(0.5 * acc) * cs;
        p.hn(t)[warp % p.K] = hn;
        // error-bound inputs of the single-product kernel, all rounded up (non-negative floats order like
        // their bit patterns): largest squared norm of a scaled codeword, largest squared norm of its fp16
        // rounding residual, largest scaled half norm
        atomicMax(p.tail(t, TAIL_EMAX2), __float_as_uint(__double2float_ru(acc * (double)cs * (double)cs)));
        atomicMax(p.tail(t, TAIL_DE2MAX), __float_as_uint(__double2float_ru(del)));
        atomicMax(p.tail(t, TAIL_HNMAX), __float_as_uint(hn));
    }
}

// bias images of the single-product kernel (layout: tc_common.cuh); one thread per codeword
__global__ void pack_bias_kernel(PackParams p) {
    const int t = blockIdx.y;
    const uint32_t hb = *p.tail(t, TAIL_HNMAX);
    const int e = (int)((hb >> 23) & 0xFF);
    float bscale = 1.0f;
    if (e != 0 && e != 255) {                       // 2^(14 - floor(log2 hnmax))
        int se = 127 + 14 - (e - 127);
        se = se < 1 ? 1 : (se > 254 ? 254 : se);
        bscale = __uint_as_float((uint32_t)se << 23);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<float*>(p.tail(t, TAIL_BSCALE)) = bscale;
    const float hnmax = __uint_as_float(hb);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < p.K; k += gridDim.x * blockDim.x) {
        // (hn = cs * 0.5 ||e||^2: ratios of hn are ratios of squared norms) -- see tables_fit_single_product
        if (p.hn(t)[k] * 16.f < hnmax) atomicAdd(p.tail(t, TAIL_NSMALL), 1u);
        const float v = p.hn(t)[k] * bscale;
        const __half b1 = __float2half_rn(v);
        const float r1 = v - __half2float(b1);
        const __half b2 = __float2half_rn(r1);
        const float r2 = r1 - __half2float(b2);
        const __half b3 = __float2half_rn(r2);
        uint8_t* blk = p.bias(t) + (size_t)(k / BN) * B_BYTES;
        const int r = k % BN;
        *reinterpret_cast<uint4*>(blk + sw_offset(r, 0)) =
            make_uint4(pack_half2(__hneg(b1), __hneg(b2)), pack_half2(__hneg(b3), __float2half_rn(0.f)), 0u, 0u);
        *reinterpret_cast<uint4*>(blk + sw_offset(r, 1)) = make_uint4(0u, 0u, 0u, 0u);
    }
}

__global__ void pack_clear_kernel(PackParams p) {
    if (threadIdx.x < p.n_tables) {
        *p.maxbits(threadIdx.x) = 0u;
        *p.tail(threadIdx.x, TAIL_EMAX2) = 0u;
        *p.tail(threadIdx.x, TAIL_DE2MAX) = 0u;
        *p.tail(threadIdx.x, TAIL_HNMAX) = 0u;
        *p.tail(threadIdx.x, TAIL_NSMALL) = 0u;
    }
}

}  // namespace

// pack buffer, per table (so that any contiguous range of tables is itself a valid pack):
//   [images][hn: K f32, 256 B aligned][cs f32 | max bits u32 | pad to 256 B]

size_t tc_pack_bytes(int n_tables, int K, int Dg) { return (size_t)n_tables * table_stride_bytes(K, Dg); }

bool rvq_search_tc_supported(int S, int G, int K, int D, int flags, const char** why) {
    if (G < 1 || G > GMAX || D % G) { *why = "groups"; return false; }
    if (K % BN) { *why = "codebook size must be a multiple of 256"; return false; }
    if ((D / G) % 64) { *why = "channels per group must be a multiple of 64"; return false; }
    if (D / G > 512) { *why = "channels per group must be <= 512"; return false; }
    if (S * G > ACQ_MAX_TABLE) { *why = "too many tables"; return false; }
    if (K > KMAX) { *why = "codebook size must be <= 1024"; return false; }
    (void)flags;
    return true;
}

size_t tc_workspace_bytes(int D) { return (size_t)kNumSMs * 2 * NTB * BM * D * sizeof(float) + 256; }

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
    pack_bias_kernel<<<dim3((unsigned)((K + 255) / 256), n_tables), 256, 0, st>>>(p);
    return check_cuda(cudaGetLastError(), "tc pack launch");
}

namespace {
template <int CL, bool SPLIT>
int launch_tc(const TcParams& p, cudaStream_t st) {
    auto kern = rvq_search_tc_kernel<CL, SPLIT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(rvq_search_tc)");
    int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
    if (CL == 1) {
        kern<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(p);
        return check_cuda(cudaGetLastError(), "rvq_search_tc launch");
    }
    grid = SPLIT ? p.num_tiles * CL              // one cluster per tile
                 : (grid + CL - 1) / CL * CL;    // whole clusters (148 is a multiple of 2 and 4)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return check_cuda(cudaLaunchKernelEx(&cfg, kern, p), "rvq_search_tc cluster launch");
}
}  // namespace

int rvq_search_tc(const float* x, const float* const* cb, const void* pack, void* workspace, int S,
                  int G, int K, int D, int B, int T, int flags, int64_t* codes, float* dbg_scores,
                  int cluster, int guard, cudaStream_t st) {
    const char* why = "";
    if (!rvq_search_tc_supported(S, G, K, D, flags, &why)) return fail(ACQ_ESHAPE, "tc search: %s", why);
    if (!pack || !workspace) return fail(ACQ_EINVAL, "tc search: pack/workspace missing");
    TcParams p;
    p.guard = guard;
    p.x = x;
    for (int i = 0; i < S * G; ++i) p.cb.p[i] = cb[i];
    const int Dg = D / G;
    p.pack = static_cast<const uint8_t*>(pack);
    p.table_stride = table_stride_bytes(K, Dg);
    p.img_bytes = align256(images_bytes(K, Dg));
    p.hn_bytes = align256((size_t)K * 4);
    p.bias_off = bias_offset_bytes(K, Dg);
    p.scratch = static_cast<float*>(workspace);
    p.S = S; p.G = G; p.K = K; p.D = D; p.Dg = Dg; p.T = T; p.flags = flags;
    p.N = (long long)B * T;
    p.num_tiles = (int)((p.N + BM - 1) / BM);
    p.tiles_per_clip = 0;
    p.codes = codes;
    p.dbg_scores = dbg_scores;
    { const char* e = getenv("ACQ_TC_DBG"); p.dbg_mode = e ? (int)strtoul(e, nullptr, 0) : 0; }
    p.err = reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + (size_t)kNumSMs * 2 * NTB * BM * D * sizeof(float));
    p.stall = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(p.err) + 64);   // 9 counters (bit 512)
    // (Tried: pinning the scratch images in L2 with a persisting access-policy window on this launch --
    // 128 MB window / 79 MB set-aside on B200 -- no gain, see DESIGN.md experiment log.)
    // small batches: one cluster of K/256 CTAs per tile, each running one codebook pass (ACQ_TC_SPLIT=0 disables)
    const int split_ok = tc_config().split;
    const int NP = K / BN;
    if (split_ok && cluster == 1 && !dbg_scores) {
        if (NP % 4 == 0 && p.num_tiles * 4 <= kNumSMs) return launch_tc<4, true>(p, st);
        if (NP % 2 == 0 && p.num_tiles * 2 <= kNumSMs) return launch_tc<2, true>(p, st);
    }
    switch (cluster) {
        case 4: return launch_tc<4, false>(p, st);
        case 2: return launch_tc<2, false>(p, st);
        default: return launch_tc<1, false>(p, st);
    }
}

}  // namespace acq
