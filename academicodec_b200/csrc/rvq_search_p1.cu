// K1: residual nearest-codeword search, ONE fp16 tensor-core product per codeword + a rigorous filter
// + an exact re-score (tcgen05 + TMEM + TMA bulk).  Default search kernel since round 2.
//
// Problem: per frame, argmax_k  x.e_k - 0.5||e_k||^2  over a 1024-entry codebook, for S residual
// stages x G channel groups, with the residual carried from stage to stage (reference:
// EuclideanCodebook.quantize core_vq.py:175-180 inside ResidualVectorQuantization.encode :348-362;
// Quantizer.for_one_step hificodec/models.py:463-492).  The codes must match the reference's fp32
// arithmetic index for index; a single fp16 pass flips 0.1-0.4 % of them.  The
// three-product kernel (rvq_search_tc.cu) buys fp32-class scores with three MMAs per product and is
// bound by shared-memory bandwidth and board power at 23 % of the tensor peak.  This kernel issues one
// MMA per product and treats the approximate scores as a FILTER with a proven error bound:
//
//   operands are scaled by exact powers of two into [1024, 2048) (per frame and group for x, per
//   codebook for e) and rounded to fp16: x~ = x^ + dx, e~_k = e^_k + de_k.  The tensor core computes
//   x^.e^_k exactly per product and accumulates in fp32, so for every codeword
//       |s_approx(k) - s_exact(k)| <= tau
//       tau = ||dx|| E^ + (||x^|| + ||dx||) dE            Cauchy-Schwarz on the dropped cross terms,
//                                                         with the ACTUAL rounding-residual norms:
//                                                         ||dx|| per row (computed when the row is
//                                                         converted), dE = max_k ||de_k|| and
//                                                         E^ = max_k ||e^_k|| per codebook (pack tail)
//           + D 2^-23 ||x^|| E^                           fp32 accumulation of D exact products
//           + 2^-22 (xs hn_max + ||x^|| E^)               bias fma and the rounding of the norms
//   hence the true best codeword satisfies s_approx >= max_k s_approx - 2 tau.
//
// Epilogue, per 256-codeword pass: sweep 1 reads the accumulator for the pass maximum (one FFMA and half
// a 3-input FMNMX per score), sweep 2 re-reads it and records every codeword within 2 tau of the running
// maximum (typically one per frame).  A frame with a single survivor is decided; the others (3-6 % of
// random frames) are re-scored exactly -- float64 dot products of the fp32 residual row against the fp32
// codewords, one warp per frame -- and the (value, lowest index) argmax of the exact scores is the code.
// Codes thus equal the float64 argmax; they differ from the reference only where its own fp32 rounding
// decides a near-tie (counted by tests/test_gpu_scale.py on every frame of every BASELINE shape).
//
// Structure (one persistent CTA per SM, 384 threads, warp-specialised, everything mbarrier-driven):
//   warps 0-3   loaders: read upcoming tiles of x (coalesced along frames), derive the per-frame scales,
//               write the tile's K-major SWIZZLE_64B fp16 image (and fp32 rows when S > 1) to per-CTA
//               scratch, up to a tile pair ahead of the MMAs; between tiles they work on jobs (below)
//   warp 8      TMA producer: one thread streams A (residual image) and B (pre-packed codebook image)
//               chunks with cp.async.bulk into a 6 x 24 KiB ring; with CL > 1 the CTAs of a cluster share
//               one multicast codebook stream
//   warp 9      MMA issuer: one thread, 2 tcgen05.mma (M128 N256 K16) per ring stage into one of two
//               256-column TMEM accumulators
//   warps 4-7   epilogue (thread = frame): the two filter sweeps per pass, overlapped with the next pass's
//               MMAs through the second accumulator; after the last pass they publish a JOB
//   warps 10-11 workers, plus every loader / epilogue warp that would otherwise wait: claim batches of rows
//               of the open jobs -- exact re-score of the undecided rows, write the codes, and between two
//               stages r <- r - e[i] in fp32 exactly as the reference does (core_vq.py:359 / :304), new row
//               scale, new fp16 image row and its rounding-residual norms.  Whoever completes the last
//               batch of a job arrives on the barrier the TMA thread (next stage's image) or the loaders
//               (tile buffer free) wait on.
// Tiles are processed in pairs with interleaved stages -- (A,s0)(B,s0)(A,s1)(B,s1)... -- so one tile's job
// overlaps the other's MMAs.  Codes only; quantized / loss / EMA outputs come from rvq_replay.cu.
// Shapes: K % 256 == 0, K <= 1024, (D/G) % 64 == 0, D/G <= 512, G <= 4.
#include "tc_common.cuh"
#include <stdlib.h>

namespace acq {
namespace {

using namespace tc;

constexpr int NSTAGE = 6;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;       // 8 + 16 KiB: one 32-channel chunk of A and of B
constexpr int STAGING_BYTES = 2 * A_BYTES;           // a 64-channel slice of a tile's image (16 KiB)
constexpr int NUM_THREADS = 384;
constexpr int NI = 2;                                // tiles of a CTA whose stages are interleaved
constexpr int NTB = 2 * NI;                          // tile buffers per CTA
constexpr int CMAX = 8;                              // candidates kept per frame and stage
constexpr int NJOB = 2;                              // job slots (a job may still be open when the next is published)
constexpr int NBAR = 2 * NSTAGE + 4 + 2 * NTB + NI * GMAX;

struct Job {
    const float* cbp;        // fp32 codebook of this (stage, group)
    const float* x;          // latents (re-score source when the call keeps no fp32 rows: S == 1)
    float* R;                // fp32 residual rows of the tile (nullptr when S == 1)
    uint8_t* img;            // fp16 image of the tile (next stage's A operand)
    float* sc_g;             // row scales of this group            [BM]
    float* nrm_g;            // {||x^||^2, ||dx||^2} of this group  [BM][2]
    uint64_t* bar;           // next-stage image ready (count 1) or tile buffer free (count G)
    int64_t* codes;          // output row of this table, offset to the tile's first frame
    long long n0;            // first frame of the tile
    int Dg, D, g, nf, ste, last, T, K;
};
struct alignas(16) JobSlot {
    Job job;
    int state[4];                    // claim, completed, sequence number of the job in the slot, -
    int bidx[BM];                    // decided code / first candidate
    int ncand[BM];                   // 1 = decided, 2..CMAX = candidates to re-score, > CMAX = all K
    int cand_idx[CMAX][BM];
    float cand_sc[CMAX][BM];
};
static_assert(sizeof(Job) <= 104, "job descriptor");

// ---- shared memory carve-up (after the ring and the staging slice) ---------------------------------
constexpr int OFF_BAR = 0;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int OFF_SCALE = OFF_TMEM + 16;                          // [NTB][GMAX][BM] f32
constexpr int OFF_NRM = OFF_SCALE + NTB * GMAX * BM * 4;          // [NTB][GMAX][BM][2] f32
constexpr int OFF_MAX = OFF_NRM + NTB * GMAX * BM * 8;            // [GMAX][BM] u32 (loaders)
constexpr int OFF_HN = OFF_MAX + GMAX * BM * 4;                   // [KMAX] f32
constexpr int OFF_JOB = OFF_HN + KMAX * 4;                        // [NJOB] JobSlot
constexpr int OFF_DONE = OFF_JOB + NJOB * (int)sizeof(JobSlot);   // int: epilogue finished
constexpr int CTRL_BYTES = OFF_DONE + 16;
constexpr size_t SMEM_BYTES = 1024 + (size_t)NSTAGE * STAGE_BYTES + STAGING_BYTES + CTRL_BYTES;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(OFF_JOB % 16 == 0 && sizeof(JobSlot) % 16 == 0, "alignment");

__device__ __forceinline__ int job_items(int Dg) { return Dg <= 128 ? BM / 8 : (Dg <= 256 ? BM / 4 : BM / 2); }

__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// fp16 image chunk (8 channels) of scaled values + this thread's share of ||x^||^2 and ||dx||^2
__device__ __forceinline__ uint4 half8_norms(const float (&a)[8], float xs, float& qh, float& qd) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float v0 = a[2 * j] * xs, v1 = a[2 * j + 1] * xs;
        const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
        const float f0 = __half2float(h0), f1 = __half2float(h1);
        qh = fmaf(f0, f0, fmaf(f1, f1, qh));
        const float e0 = v0 - f0, e1 = v1 - f1;
        qd = fmaf(e0, e0, fmaf(e1, e1, qd));
        w[j] = pack_half2(h0, h1);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- exact re-score of one frame (one warp; lanes across channels) ---------------------------------
// r = the frame's fp32 residual of group j.g: from the tile's fp32 rows (S > 1) or straight from x (S == 1:
// channel stride T; only the undecided frames pay this gather).  Candidates in ascending index order, so
// strict > keeps the lowest index among exact ties.
template <int JN>
__device__ __forceinline__ int rescore_row(const Job& j, const JobSlot* slot, int row, int n, int lane) {
    float4 rv[JN];
    if (j.R) {
        const float* rrow = j.R + (size_t)row * j.D + j.g * j.Dg;
#pragma unroll
        for (int q = 0; q < JN; ++q) {
            const int d = lane * 4 + 128 * q;
            rv[q] = d < j.Dg ? *reinterpret_cast<const float4*>(rrow + d) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        const long long nfr = j.n0 + row;
        const long long b = nfr / j.T, t = nfr % j.T;
        const float* src = j.x + ((size_t)b * j.D + (size_t)j.g * j.Dg) * j.T + t;
#pragma unroll
        for (int q = 0; q < JN; ++q) {
            const int d = lane * 4 + 128 * q;
            if (d < j.Dg) {
                rv[q].x = __ldg(src + (size_t)d * j.T);
                rv[q].y = __ldg(src + (size_t)(d + 1) * j.T);
                rv[q].z = __ldg(src + (size_t)(d + 2) * j.T);
                rv[q].w = __ldg(src + (size_t)(d + 3) * j.T);
            } else {
                rv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    const bool full = n > CMAX;
    const int n_iter = full ? j.K : n;
    double best = -INFINITY;
    int best_k = 0;
    // two candidates per iteration: both codeword rows are in flight together
    for (int i = 0; i < n_iter; i += 2) {
        const bool two = i + 1 < n_iter;
        const int k0 = full ? i : slot->cand_idx[i][row];
        const int k1 = two ? (full ? i + 1 : slot->cand_idx[i + 1][row]) : k0;
        const float* e0 = j.cbp + (size_t)k0 * j.Dg;
        const float* e1 = j.cbp + (size_t)k1 * j.Dg;
        float4 ev0[JN], ev1[JN];
#pragma unroll
        for (int q = 0; q < JN; ++q) {
            const int d = lane * 4 + 128 * q;
            const bool in = d < j.Dg;
            ev0[q] = in ? __ldg(reinterpret_cast<const float4*>(e0 + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
            ev1[q] = in ? __ldg(reinterpret_cast<const float4*>(e1 + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        double dot0 = 0.0, nrm0 = 0.0, dot1 = 0.0, nrm1 = 0.0;
#pragma unroll
        for (int q = 0; q < JN; ++q) {
            const double rx = rv[q].x, ry = rv[q].y, rz = rv[q].z, rw = rv[q].w;
            dot0 = fma(rx, (double)ev0[q].x, dot0); nrm0 = fma((double)ev0[q].x, (double)ev0[q].x, nrm0);
            dot0 = fma(ry, (double)ev0[q].y, dot0); nrm0 = fma((double)ev0[q].y, (double)ev0[q].y, nrm0);
            dot0 = fma(rz, (double)ev0[q].z, dot0); nrm0 = fma((double)ev0[q].z, (double)ev0[q].z, nrm0);
            dot0 = fma(rw, (double)ev0[q].w, dot0); nrm0 = fma((double)ev0[q].w, (double)ev0[q].w, nrm0);
            dot1 = fma(rx, (double)ev1[q].x, dot1); nrm1 = fma((double)ev1[q].x, (double)ev1[q].x, nrm1);
            dot1 = fma(ry, (double)ev1[q].y, dot1); nrm1 = fma((double)ev1[q].y, (double)ev1[q].y, nrm1);
            dot1 = fma(rz, (double)ev1[q].z, dot1); nrm1 = fma((double)ev1[q].z, (double)ev1[q].z, nrm1);
            dot1 = fma(rw, (double)ev1[q].w, dot1); nrm1 = fma((double)ev1[q].w, (double)ev1[q].w, nrm1);
        }
        double s0 = dot0 - 0.5 * nrm0, s1 = dot1 - 0.5 * nrm1;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, off);
            s1 += __shfl_xor_sync(0xffffffffu, s1, off);
        }
        if (s0 > best) { best = s0; best_k = k0; }
        if (two && s1 > best) { best = s1; best_k = k1; }
    }
    return best_k;
}

// One batch of RB rows of a job: decide the undecided rows, write the codes, update the residual.
template <int RB, int JN>
__device__ __forceinline__ void process_batch(const Job& j, const JobSlot* slot, int item, int lane) {
    const int row0 = item * RB;
    int idxs[RB];
    int mine = 0;
#pragma unroll
    for (int u = 0; u < RB; ++u) {
        const int row = row0 + u;
        int idx = slot->bidx[row];
        const int n = slot->ncand[row];                          // (warp-uniform: shared memory)
        if (n > 1 && row < j.nf) idx = rescore_row<JN>(j, slot, row, n, lane);
        idxs[u] = idx;
        if (lane == u) mine = idx;
    }
    if (lane < RB && row0 + lane < j.nf) j.codes[row0 + lane] = (int64_t)mine;
    if (!j.last)
        residual_update_batch<RB, JN, false, true>(row0, lane, j.nf, idxs, j.cbp, j.Dg, j.D, j.g, j.R, j.img,
                                                   j.sc_g, j.nrm_g, j.ste != 0);
}

// Claim and process batches of the open jobs until none is left or `budget` batches are done; returns
// the number processed.  Slots are tried oldest job first.
// (not inlined: it is called from seven places and carries three instantiations of the batch code)
__device__ __noinline__ int steal_jobs(JobSlot* slots, int lane, int budget = 0x7fffffff) {
    int total = 0;
    const int first = (*(volatile int*)&slots[0].state[2] <= *(volatile int*)&slots[1].state[2]) ? 0 : 1;
#pragma unroll 1
    for (int k = 0; k < NJOB && total < budget; ++k) {
        JobSlot* slot = slots + ((first + k) & 1);
        volatile int* st = slot->state;
        int mine = 0;
        uint64_t* bar = nullptr;
        int items = 0;
#pragma unroll 1
        while (total + mine < budget) {
            int item = 0x7fffffff;
            if (lane == 0 && st[0] < BM) item = atomicAdd(const_cast<int*>(st), 1);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= BM) break;                               // (closed slots hold claim >= BM)
            __threadfence_block();                               // the job was written before claim <- 0
            const Job j = slot->job;
            items = job_items(j.Dg);
            if (item >= items) break;
            bar = j.bar;
            if (j.Dg <= 128) process_batch<8, 1>(j, slot, item, lane);
            else if (j.Dg <= 256) process_batch<4, 2>(j, slot, item, lane);
            else process_batch<2, 4>(j, slot, item, lane);
            ++mine;
        }
        if (mine) {
            // one cross-proxy fence for all the batches of this job this call finished (a job cannot be
            // completed and replaced while a claimed batch is outstanding)
            __syncwarp();
            fence_proxy_async_global();                          // image rows -> the TMA thread's bulk reads
            __threadfence_block();
            if (lane == 0) {
                const int done = atomicAdd(const_cast<int*>(st + 1), mine) + mine;
                if (done == items) {
                    __threadfence_block();
                    mbar_arrive(bar);
                }
            }
            total += mine;
        }
    }
    return total;
}

// (polling helpers are warp-uniform: the callers go on to warp-collective code -- shuffles in steal_jobs,
//  tcgen05.ld -- so every lane must take the same decision even if the flag flips between two lanes' reads)
__device__ __forceinline__ bool slot_done(const JobSlot* slot) {
    const volatile int* st = slot->state;
    const bool d = st[1] >= job_items(*(volatile const int*)&slot->job.Dg) || st[0] >= 0x40000000;
    return __all_sync(0xffffffffu, d);
}
__device__ __forceinline__ bool warp_try_wait(uint64_t* bar, uint32_t parity) {
    return __all_sync(0xffffffffu, mbar_try_wait(bar, parity));
}
__device__ __forceinline__ bool warp_flag_set(volatile int* flag) { return __all_sync(0xffffffffu, *flag != 0); }

template <int CL>
__global__ void __launch_bounds__(NUM_THREADS, 1) rvq_search_p1_kernel(const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* staging = smem + NSTAGE * STAGE_BYTES;
    uint8_t* ctrl = staging + STAGING_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl + OFF_BAR);   // [NSTAGE] TMA bytes landed
    uint64_t* empty_bar = full_bar + NSTAGE;                            // [NSTAGE] MMAs retired
    uint64_t* tfull_bar = empty_bar + NSTAGE;                           // [2] accumulator complete
    uint64_t* tempty_bar = tfull_bar + 2;                               // [2] accumulator drained
    uint64_t* t0_bar = tempty_bar + 2;                                  // [NTB] stage-0 image of a tile ready
    uint64_t* free_bar = t0_bar + NTB;                                  // [NTB] tile buffer reusable
    uint64_t* upd_bar = free_bar + NTB;                                 // [NI][GMAX] next-stage image ready
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(ctrl + OFF_TMEM);
    float* scale_s = reinterpret_cast<float*>(ctrl + OFF_SCALE);
    float* nrm_s = reinterpret_cast<float*>(ctrl + OFF_NRM);
    uint32_t* rowmax_s = reinterpret_cast<uint32_t*>(ctrl + OFF_MAX);
    float* hn_s = reinterpret_cast<float*>(ctrl + OFF_HN);
    JobSlot* slots = reinterpret_cast<JobSlot*>(ctrl + OFF_JOB);
    volatile int* all_done = reinterpret_cast<volatile int*>(ctrl + OFF_DONE);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = p.S, G = p.G, K = p.K, D = p.D, Dg = p.Dg, T = p.T;
    const int NP = K / BN, NKC = Dg / BK;
    const bool ste = p.flags & ACQ_STE;
    const size_t tile_elems = (size_t)BM * D;
    // scratch layout as in the three-product kernel (same workspace): [buf][CTA] images, then [buf][CTA]
    // fp32 rows.  Only the first half of an image slot is used (hi only, A_BYTES per 32-channel chunk).
    const size_t buf_stride = (size_t)kNumSMs * tile_elems * 4;
    uint8_t* Aimg = reinterpret_cast<uint8_t*>(p.scratch) + (size_t)blockIdx.x * tile_elems * 4;
    float* Rbuf = reinterpret_cast<float*>(Aimg + NTB * buf_stride);
    // single-stage calls keep the scratch working set small: 2 tile buffers
    const uint32_t ntb = S * G == 1 ? 2u : (uint32_t)NTB;
    const uint32_t ni = S * G == 1 ? 2u : (uint32_t)NI;
    // (cluster-uniform: the tile count of the cluster's first CTA, which is the largest)
    const int lead_cta = (int)(blockIdx.x / CL) * CL;
    const uint32_t n_my = p.num_tiles > lead_cta ? (uint32_t)((p.num_tiles - 1 - lead_cta) / (int)gridDim.x + 1) : 0u;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    const long long tile_base = (long long)blockIdx.x, tile_stride = (long long)gridDim.x;
    constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1);

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&full_bar[i], 1);         // the TMA thread's arrive.expect_tx
            mbar_init(&empty_bar[i], CL);       // tcgen05.commit of every CTA sharing the B stream
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);        // tcgen05.commit
            mbar_init(&tempty_bar[i], 128);     // epilogue threads
        }
        for (int i = 0; i < NTB; ++i) {
            mbar_init(&t0_bar[i], 128);         // loader threads
            mbar_init(&free_bar[i], G);         // the last-stage job of every group
        }
        for (int i = 0; i < NI * GMAX; ++i) mbar_init(&upd_bar[i], 1);   // whoever completes the update job
        for (int i = 0; i < NJOB; ++i) {
            slots[i].state[0] = 0x7fffffff;     // closed: nothing to claim ...
            slots[i].state[1] = 0x7fffffff;     // ... and nothing to wait for
            slots[i].state[2] = i - NJOB;       // sequence number
            slots[i].state[3] = 0;
            slots[i].job.Dg = Dg;
        }
        *all_done = 0;
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(tmem_ptr_s, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();      // every CTA's barriers are initialised before any multicast lands
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (warp < 4) {
        // ================= loaders: x tile -> scales, fp16 image, rounding norms (and R when S > 1) =====
        unsigned long long w_free = 0;
        const long long t_begin = clock64();
        for (uint32_t it = 0; it < n_my; ++it) {
            const long long tile = tile_base + (long long)it * tile_stride;   // may be a dummy past the end
            const uint32_t buf = it % ntb;
            if (!warp_try_wait(&free_bar[buf], ((it / ntb) & 1) ^ 1)) {
                // no buffer to fill yet: work on the open jobs meanwhile
                const long long tw = clock64();
                while (!warp_try_wait(&free_bar[buf], ((it / ntb) & 1) ^ 1)) {
                    if (!steal_jobs(slots, lane)) __nanosleep(128);
                    if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 6); __trap(); }
                }
                w_free += (unsigned long long)(clock64() - tw);
            }
            const long long n0 = tile * BM;
            uint8_t* img = Aimg + buf * buf_stride;
            float* R = Rbuf + buf * (buf_stride / 4);
            float* sc = scale_s + buf * GMAX * BM;
            float* nrm = nrm_s + buf * GMAX * BM * 2;
            for (int i = tid; i < G * BM; i += 128) {
                rowmax_s[i] = 0u;
                nrm[2 * i] = 0.f;
                nrm[2 * i + 1] = 0.f;
            }
            named_bar_sync(2, 128);
            if ((T & 3) == 0) {
                // 4 consecutive frames per thread (one 16 B load per channel), 16 channels at a time
                const int rq = tid & 31, w4 = tid >> 5;
                const long long n = n0 + 4 * rq;
                const bool ok = n < p.N;                 // N % 4 == 0: a quad is all in or all out
                const long long b = ok ? n / T : 0, t = ok ? n % T : 0;
                const float* src = p.x + (size_t)(b * D) * T + t;
                for (int sweep = 0; sweep < 2; ++sweep) {
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
                                for (int i = 0; i < 8; ++i) m = fmax3(m, fabsf(a[0][i]), fabsf(a[1][i]));
                                atomicMax(&rowmax_s[g * BM + row], __float_as_uint(m));
                            } else {
                                const float xs = sc[g * BM + row];
                                float qh = 0.f, qd = 0.f;
                                const uint4 h0 = half8_norms(a[0], xs, qh, qd);
                                const uint4 h1 = half8_norms(a[1], xs, qh, qd);
                                atomicAdd(&nrm[2 * (g * BM + row)], qh);
                                atomicAdd(&nrm[2 * (g * BM + row) + 1], qd);
                                // the four loader warps cover one 64-channel slice per iteration = one
                                // contiguous 16 KiB range of the image: assembled in shared memory, copied
                                // out with fully coalesced stores below (scattered sector stores were
                                // measured to cost 0.29 ms of a 1.06 ms launch in round 1)
                                const int oct_in = 2 * w4;      // this warp's (even) octet inside the 8-octet slice
                                uint8_t* cb_ = staging + (size_t)(oct_in / CPR) * A_BYTES;
                                const int c = oct_in % CPR;
                                *reinterpret_cast<uint4*>(cb_ + sw_offset(row, c)) = h0;
                                *reinterpret_cast<uint4*>(cb_ + sw_offset(row, c + 1)) = h1;
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
                            named_bar_sync(2, 128);
                            const int slice = (pr - w4) / 4;                   // 64-channel slice index
                            uint4* gdst = reinterpret_cast<uint4*>(img + (size_t)slice * STAGING_BYTES);
                            const uint4* ssrc = reinterpret_cast<const uint4*>(staging);
#pragma unroll 4
                            for (int k = tid; k < STAGING_BYTES / 16; k += 128) gdst[k] = ssrc[k];
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
                // general T (e.g. HiFi-Codec's 50 frames per clip): one frame per thread, scalar loads that
                // are coalesced across the warp, 32 channels (one chunk of the image) per iteration
                const int row = tid;
                const long long n = n0 + row;
                const bool ok = n < p.N;
                const long long b = ok ? n / T : 0, t = ok ? n % T : 0;
                const float* src = p.x + (size_t)(b * D) * T + t;
                for (int sweep = 0; sweep < 2; ++sweep) {
                    float m = 0.f, qh = 0.f, qd = 0.f;
                    for (int o4 = 0; o4 < D / 8; o4 += 4) {
                        float a[4][8];
#pragma unroll
                        for (int h = 0; h < 4; ++h)
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                a[h][i] = ok ? __ldg(src + (size_t)((o4 + h) * 8 + i) * T) : 0.f;
                        const int g = (o4 * 8) / Dg;             // Dg % 64 == 0: the 32 channels share a group
                        const bool group_end = ((o4 + 4) * 8) % Dg == 0;
                        if (sweep == 0) {
#pragma unroll
                            for (int h = 0; h < 4; ++h)
#pragma unroll
                                for (int i = 0; i < 8; ++i) m = fmaxf(m, fabsf(a[h][i]));
                            if (group_end) {                     // the row is this thread's alone
                                rowmax_s[g * BM + row] = __float_as_uint(m);
                                m = 0.f;
                            }
                        } else {
                            const float xs = sc[g * BM + row];
                            uint8_t* chunk = img + (size_t)(o4 / CPR) * A_BYTES;
#pragma unroll
                            for (int h = 0; h < 4; h += 2) {
                                const uint4 h0 = half8_norms(a[h], xs, qh, qd);
                                const uint4 h1 = half8_norms(a[h + 1], xs, qh, qd);
                                store_chunk_pair(chunk, row, h, h0, h1);
                            }
                            if (group_end) {
                                nrm[2 * (g * BM + row)] = qh;
                                nrm[2 * (g * BM + row) + 1] = qd;
                                qh = 0.f;
                                qd = 0.f;
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
        // all tiles loaded: keep working on jobs until the epilogue has finished its last tile
        const long long tw = clock64();
        while (!warp_flag_set(all_done)) {
            if (!steal_jobs(slots, lane)) __nanosleep(128);
            if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 10); __trap(); }
        }
    } else if (warp >= 10) {
        // ================= workers: jobs only =============================================================
        const long long tw = clock64();
        while (!warp_flag_set(all_done)) {
            if (!steal_jobs(slots, lane)) __nanosleep(64);
            if (clock64() - tw > 16000000000LL) { if (p.err) atomicExch(p.err, 14); __trap(); }
        }
    } else if (warp == 8) {
        // ================= TMA producer: one thread streams A and B image chunks ==========================
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
                        const uint8_t* img = Aimg + buf * buf_stride;
                        for (int g = 0; g < G; ++g) {
                            const uint8_t* bimg = p.pack + (size_t)(s * G + g) * p.table_stride;
                            for (int pass = 0; pass < NP; ++pass) {
                                for (int kc = 0; kc < NKC; ++kc, ++ring_it) {
                                    const int st = ring_it % NSTAGE;
                                    mbar_wait_t(&empty_bar[st], ((ring_it / NSTAGE) & 1) ^ 1, p.err, 2, w_empty);
                                    uint8_t* a_dst = smem + st * STAGE_BYTES;
                                    uint8_t* b_dst = a_dst + A_BYTES;
                                    const uint8_t* bsrc = bimg + (size_t)(pass * NKC + kc) * 2 * B_BYTES;   // hi image
                                    mbar_arrive_expect_tx(&full_bar[st], A_BYTES + B_BYTES);
                                    if (CL == 1) {
                                        bulk_g2s(b_dst, bsrc, B_BYTES, &full_bar[st]);
                                    } else {
                                        constexpr uint32_t SLICE = B_BYTES / CL;
                                        bulk_g2s_mc(b_dst + crank * SLICE, bsrc + crank * SLICE, SLICE, &full_bar[st], CMASK);
                                    }
                                    if (pass == 0 && kc == 0) {
                                        // first use of this (tile, stage, group)'s image
                                        if (s == 0) {
                                            mbar_wait_t(&t0_bar[buf], (it / ntb) & 1, p.err, 7, w_t0);
                                        } else {
                                            mbar_wait(&upd_bar[par * GMAX + g], upd_it[par * GMAX + g] & 1, p.err, 8);
                                            ++upd_it[par * GMAX + g];
                                        }
                                        fence_proxy_async_global();
                                    }
                                    bulk_g2s(a_dst, img + (size_t)(g * NKC + kc) * A_BYTES, A_BYTES, &full_bar[st]);
                                }
                            }
                        }
                    }
                }
            }
            if (p.dbg_mode & 512) { atomicAdd(p.stall + 3, w_empty); atomicAdd(p.stall + 4, w_t0); }
        }
    } else if (warp == 9) {
        // ================= MMA issuer: one product per chunk =============================================
        if (lane == 0) {
            uint32_t ring_it = 0, acc_it = 0;
            unsigned long long w_full0 = 0, w_full = 0, w_tempty = 0;
            const long long t_begin = clock64();
            for (uint32_t itm = 0; itm < n_my; ++itm) {       // (same number of items in any order)
                for (int sg = 0; sg < S * G; ++sg) {
                    for (int pass = 0; pass < NP; ++pass, ++acc_it) {
                        const uint32_t abuf = acc_it & 1;
                        mbar_wait_t(&tempty_bar[abuf], ((acc_it >> 1) & 1) ^ 1, p.err, 3, w_tempty);
                        tc_fence_after();
                        const uint32_t d_tmem = tmem_base + abuf * BN;
                        for (int kc = 0; kc < NKC; ++kc, ++ring_it) {
                            const int st = ring_it % NSTAGE;
                            mbar_wait_t(&full_bar[st], (ring_it / NSTAGE) & 1, p.err, 4, pass == 0 ? w_full0 : w_full);
                            tc_fence_after();
                            const uint32_t a_hi = smem_u32(smem + st * STAGE_BYTES);
                            const uint32_t b_hi = a_hi + A_BYTES;
#pragma unroll
                            for (int kk = 0; kk < BK / UK; ++kk) {
                                const uint32_t ko = kk * UK * 2;   // bytes along K inside the swizzle atom
                                umma_f16(d_tmem, make_desc(a_hi + ko), make_desc(b_hi + ko), IDESC, (kc | kk) != 0);
                            }
                            // ring stage free once these MMAs retire (in every CTA that shares it)
                            if (CL == 1) umma_commit(&empty_bar[st]);
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
        // ================= epilogue: filter sweeps, publish jobs (warps 4-7, thread = frame) =============
        const int q = warp - 4;
        const int row = q * 32 + lane;
        uint32_t acc_it = 0;
        unsigned long long e_wait = 0, e_sweep = 0, e_slot = 0, n_amb = 0, n_full = 0;
        int job_seq = 0;
        for (uint32_t it0 = 0; it0 < n_my; it0 += ni) {
          const int npair = (int)min(ni, n_my - it0);
          for (int s = 0; s < S; ++s) {
            for (int h = 0; h < npair; ++h) {
                const uint32_t it = it0 + h, buf = it % ntb, par = h;
                const long long n0 = (tile_base + (long long)it * tile_stride) * BM;
                const int nf = (int)max(0LL, min((long long)BM, p.N - n0));
                uint8_t* img = Aimg + buf * buf_stride;
                float* R = Rbuf + buf * (buf_stride / 4);
                float* sc = scale_s + buf * GMAX * BM;
                float* nrm = nrm_s + buf * GMAX * BM * 2;
                if (s == 0) mbar_wait(&t0_bar[buf], (it / ntb) & 1, p.err, 9);    // this tile's scales are visible
                for (int g = 0; g < G; ++g, ++job_seq) {
                    const int table = s * G + g;
                    const uint8_t* rec = p.pack + (size_t)table * p.table_stride;
                    const float* hn = reinterpret_cast<const float*>(rec + p.img_bytes);
                    const uint32_t* tail = reinterpret_cast<const uint32_t*>(rec + p.img_bytes + p.hn_bytes);
                    const float emax2 = __uint_as_float(__ldg(tail + TAIL_EMAX2));
                    const float de2max = __uint_as_float(__ldg(tail + TAIL_DE2MAX));
                    const float hnmax = __uint_as_float(__ldg(tail + TAIL_HNMAX));
                    JobSlot* slot = slots + (job_seq & 1);
                    // the slot's previous job (two publications ago) must be complete before its candidate
                    // arrays are reused; help finishing it if it is not
                    {
                        const long long tw = clock64();
                        while (!slot_done(slot)) {
                            steal_jobs(slots, lane, 1);
                            if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 11); __trap(); }
                        }
                        e_slot += (unsigned long long)(clock64() - tw);
                    }
                    // stage this table's scaled norms in shared memory (all four epilogue warps)
                    named_bar_sync(3, 128);
                    for (int i = (tid - 128) * 4; i < K; i += 128 * 4)
                        *reinterpret_cast<float4*>(hn_s + i) = __ldg(reinterpret_cast<const float4*>(hn + i));
                    named_bar_sync(3, 128);
                    float nxs = 0.f, tau2 = 0.f;
                    float gmax = -INFINITY;
                    int ncand = 0;
                    int* cidx = &slot->cand_idx[0][row];
                    float* csc = &slot->cand_sc[0][row];
                    for (int pass = 0; pass < NP; ++pass, ++acc_it) {
                        const uint32_t abuf = acc_it & 1;
                        if (!warp_try_wait(&tfull_bar[abuf], (acc_it >> 1) & 1)) {
                            // nothing to drain yet: work on the open jobs meanwhile, one batch at a time
                            const long long tw = clock64();
                            while (!warp_try_wait(&tfull_bar[abuf], (acc_it >> 1) & 1)) {
                                steal_jobs(slots, lane, 1);
                                if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 5); __trap(); }
                            }
                            e_wait += (unsigned long long)(clock64() - tw);
                        }
                        tc_fence_after();
                        if (pass == 0) {
                            // (read after the stage's first accumulator is complete: the row scale and the
                            //  rounding norms of stage s are written by the job of stage s-1)
                            const float xs = sc[g * BM + row];
                            const float xh = sqrtf(nrm[2 * (g * BM + row)]), xd = sqrtf(nrm[2 * (g * BM + row) + 1]);
                            const float dE = sqrtf(de2max), eh = sqrtf(emax2) + dE;
                            const float tau = xd * eh + (xh + xd) * dE + ((float)Dg * 1.1920929e-7f) * xh * eh +
                                              2.3841858e-7f * (xs * hnmax + xh * eh);
                            tau2 = 2.002f * tau;
                            nxs = -xs;
                        }
                        const long long tq = clock64();
                        const uint32_t taddr = tmem_base + abuf * BN + ((uint32_t)(q * 32) << 16);
                        const int kbase = pass * BN;
                        // sweep 1: maximum of this pass (two independent 3-input max chains)
                        const uint32_t hp = smem_u32(hn_s + kbase);
                        float pm0 = -INFINITY, pm1 = -INFINITY;
                        {
                            uint32_t ra[32], rb[32];
                            tmem_ld32_async(taddr, ra);
#pragma unroll 1
                            for (int c0 = 0; c0 < BN; c0 += 64) {
                                tmem_ld_wait(ra);
                                tmem_ld32_async(taddr + c0 + 32, rb);
#pragma unroll
                                for (int j = 0; j < 32; j += 4) {
                                    const float4 hh = lds128(hp + (c0 + j) * 4);
                                    pm0 = fmax3(pm0, fmaf(nxs, hh.x, __uint_as_float(ra[j])), fmaf(nxs, hh.y, __uint_as_float(ra[j + 1])));
                                    pm1 = fmax3(pm1, fmaf(nxs, hh.z, __uint_as_float(ra[j + 2])), fmaf(nxs, hh.w, __uint_as_float(ra[j + 3])));
                                }
                                tmem_ld_wait(rb);
                                if (c0 + 64 < BN) tmem_ld32_async(taddr + c0 + 64, ra);
#pragma unroll
                                for (int j = 0; j < 32; j += 4) {
                                    const float4 hh = lds128(hp + (c0 + 32 + j) * 4);
                                    pm0 = fmax3(pm0, fmaf(nxs, hh.x, __uint_as_float(rb[j])), fmaf(nxs, hh.y, __uint_as_float(rb[j + 1])));
                                    pm1 = fmax3(pm1, fmaf(nxs, hh.z, __uint_as_float(rb[j + 2])), fmaf(nxs, hh.w, __uint_as_float(rb[j + 3])));
                                }
                            }
                        }
                        const float pmax = fmaxf(pm0, pm1);
                        if (pmax - tau2 > gmax) ncand = 0;      // every earlier candidate is now out of range
                        gmax = fmaxf(gmax, pmax);
                        const float thr = gmax - tau2;
                        // sweep 2: every codeword within 2 tau of the running maximum is a candidate.  One test
                        // per four columns (a branch per column made this sweep 15 kcycles per pass): the
                        // per-column code only runs for a group that holds a candidate of some row of the warp.
                        {
                            uint32_t ra[32], rb[32];
                            auto scan32 = [&](const uint32_t (&r)[32], int cbase) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4) {
                                    const float4 hh = lds128(hp + (cbase + j) * 4);
                                    const float s0 = fmaf(nxs, hh.x, __uint_as_float(r[j]));
                                    const float s1 = fmaf(nxs, hh.y, __uint_as_float(r[j + 1]));
                                    const float s2 = fmaf(nxs, hh.z, __uint_as_float(r[j + 2]));
                                    const float s3 = fmaf(nxs, hh.w, __uint_as_float(r[j + 3]));
                                    if (fmax3(fmaxf(s0, s1), s2, s3) >= thr) {
                                        const float sv[4] = {s0, s1, s2, s3};
#pragma unroll
                                        for (int u = 0; u < 4; ++u) {
                                            if (sv[u] >= thr) {
                                                if (ncand < CMAX) {
                                                    cidx[ncand * BM] = kbase + cbase + j + u;
                                                    csc[ncand * BM] = sv[u];
                                                }
                                                ++ncand;
                                            }
                                        }
                                    }
                                }
                            };
                            tmem_ld32_async(taddr, ra);
#pragma unroll 1
                            for (int c0 = 0; c0 < BN; c0 += 64) {
                                tmem_ld_wait(ra);
                                tmem_ld32_async(taddr + c0 + 32, rb);
                                scan32(ra, c0);
                                tmem_ld_wait(rb);
                                if (c0 + 64 < BN) tmem_ld32_async(taddr + c0 + 64, ra);
                                scan32(rb, c0 + 32);
                            }
                        }
                        if (p.dbg_scores && table == 0) {      // (warp-uniform: tcgen05.ld is collective)
                            float* o = p.dbg_scores + (size_t)(n0 + row) * K + kbase;
                            const float inv = 1.0f / -nxs;
                            const bool wr = row < nf;
                            for_each_score(taddr, hn_s + kbase, nxs, [&](int c, float sv) { if (wr) o[c] = sv * inv; });
                        }
                        tc_fence_before();
                        mbar_arrive(&tempty_bar[abuf]);
                        e_sweep += (unsigned long long)(clock64() - tq);
                    }
                    // ---- final filter against the global maximum (candidates stay in ascending order)
                    int keep = 0;
                    if (ncand <= CMAX) {
                        const float thr = gmax - tau2;
                        for (int i = 0; i < ncand; ++i) {
                            if (csc[i * BM] >= thr) {
                                cidx[keep * BM] = cidx[i * BM];
                                ++keep;
                            }
                        }
                    } else {
                        keep = CMAX + 1;                        // overflow: exact scores of all K codewords
                    }
                    int bidx = (keep >= 1 && keep <= CMAX) ? cidx[0] : 0;
                    if (emax2 == 0.f || row >= nf) {            // all-zero codebook: every score ties -> index 0
                        keep = 1;
                        bidx = 0;
                    }
                    if (keep == 0) keep = 1;                     // (NaN scores: nothing compares; index 0)
                    slot->bidx[row] = bidx;
                    slot->ncand[row] = (p.dbg_mode & 32) ? 1 : keep;
                    if (p.dbg_mode & 512) { n_amb += keep > 1; n_full += keep > CMAX; }
                    named_bar_sync(3, 128);                      // all 128 rows of the job are in shared memory
                    if (tid == 128) {
                        Job& j = slot->job;
                        const bool last = s + 1 == S;
                        j.cbp = p.cb.p[table];
                        j.x = p.x;
                        j.R = S > 1 ? R : nullptr;
                        j.img = img;
                        j.sc_g = sc + g * BM;
                        j.nrm_g = nrm + 2 * g * BM;
                        j.bar = last ? &free_bar[buf] : &upd_bar[par * GMAX + g];
                        j.codes = p.codes + (size_t)table * p.N + n0;
                        j.n0 = n0;
                        j.Dg = Dg; j.D = D; j.g = g; j.nf = nf; j.ste = ste ? 1 : 0; j.last = last ? 1 : 0;
                        j.T = T; j.K = K;
                        slot->state[2] = job_seq;
                        slot->state[1] = 0;
                        __threadfence_block();
                        slot->state[0] = 0;                      // opens the job: batches can be claimed
                    }
                }
            }
          }
        }
        // wait for the open jobs (helping), then release the helper and worker warps
        {
            const long long tw = clock64();
            while (!(slot_done(slots) && slot_done(slots + 1))) {
                steal_jobs(slots, lane, 1);
                if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 12); __trap(); }
            }
        }
        named_bar_sync(3, 128);
        if (tid == 128) { __threadfence_block(); *all_done = 1; }
        if (p.dbg_mode & 512) {
            if (tid == 128) { atomicAdd(p.stall + 10, e_wait); atomicAdd(p.stall + 11, e_sweep); atomicAdd(p.stall + 12, e_slot); }
            atomicAdd(p.stall + 13, n_amb);
            atomicAdd(p.stall + 14, n_full);
        }
    }

    tc_fence_before();
    __syncthreads();
    // no CTA may retire while a peer can still multicast into its ring or arrive on its barriers
    if (CL > 1) cluster_sync_all();
    if (warp == 9) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int CL>
int launch_p1(const TcParams& p, cudaStream_t st) {
    auto kern = rvq_search_p1_kernel<CL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(rvq_search_p1)");
    int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
    if (CL == 1) {
        kern<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(p);
        return check_cuda(cudaGetLastError(), "rvq_search_p1 launch");
    }
    grid = (grid + CL - 1) / CL * CL;        // whole clusters (148 is a multiple of 2 and 4)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return check_cuda(cudaLaunchKernelEx(&cfg, kern, p), "rvq_search_p1 cluster launch");
}

}  // namespace

bool rvq_search_tc_supported(int S, int G, int K, int D, int flags, const char** why);

int rvq_search_p1(const float* x, const float* const* cb, const void* pack, void* workspace, int S,
                  int G, int K, int D, int B, int T, int flags, int64_t* codes, float* dbg_scores,
                  int cluster, cudaStream_t st) {
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
    p.hn_bytes = align256((size_t)K * 4);
    p.scratch = static_cast<float*>(workspace);
    p.S = S; p.G = G; p.K = K; p.D = D; p.Dg = Dg; p.T = T; p.flags = flags;
    p.N = (long long)B * T;
    p.num_tiles = (int)((p.N + BM - 1) / BM);
    p.codes = codes;
    p.dbg_scores = dbg_scores;
    { const char* e = getenv("ACQ_TC_DBG"); p.dbg_mode = e ? atoi(e) : 0; }
    p.err = reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + (size_t)kNumSMs * 2 * NTB * BM * D * sizeof(float));
    p.stall = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(p.err) + 64);
    switch (cluster) {
        case 4: return launch_p1<4>(p, st);
        case 2: return launch_p1<2>(p, st);
        default: return launch_p1<1>(p, st);
    }
}

}  // namespace acq
