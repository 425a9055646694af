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
//           + (D + 16) 2^-23 (||x^|| E^ + xs hn_max)      fp32 accumulation of the exact products (the
//                                                         bias enters as three more of them, see below)
//           + 2^-22 xs hn_max                             rounding of the norms and of their fp16 split
//   hence the true best codeword satisfies s_approx >= max_k s_approx - 2 tau.
//
// The bias -xs * 0.5||e_k||^2 is part of the contraction: every pass ends with one extra K16 MMA over a
// "bias chunk" (A rows {w, w, w, 0..}, B rows {-b1, -b2, -b3, 0..}, w = xs / bscale, b1 + b2 + b3 =
// hn_k * bscale split into three fp16), so the accumulator holds the final score and the epilogue needs no
// per-score arithmetic.  The row scale is capped at 2^15 * bscale to keep w representable; rows far
// smaller than the codebook are then scaled below [1024, 2048), which the bound (actual ||dx||) covers.
//
// Epilogue, per 256-codeword pass: sweep 1 reads the accumulator for the pass maximum (half a 3-input
// FMNMX per score), sweep 2 re-reads it and records every codeword within 2 tau of the running maximum
// (one test per four columns).  A frame with a single survivor is decided; the others (3-8 % of random
// frames) are re-scored exactly -- float64 dot products of the fp32 residual row against the fp32
// codewords, one warp per frame -- and the (value, lowest index) argmax of the exact scores is the code.
// Codes thus equal the float64 argmax; they differ from the reference only where its own fp32 rounding
// decides a near-tie (counted by tests/test_gpu_scale.py on every frame of every BASELINE shape).
//
// Structure (one persistent CTA per SM, warp-specialised, everything mbarrier-driven).  Two layouts:
//   standard (multi-stage calls, and single-stage calls whose x is not streamed): 384 threads at 168 registers
//     warps 0-3   loaders      warps 4-7  epilogue      warp 8  TMA producer      warp 9  MMA issuer
//     warp 10     x streamer   warp 11    worker
//   wide (single-stage, single-group calls with streamed x; NXSLOT == 8): 512 threads at 128 registers
//     warps 0-3 and 10-13 loaders, each with its own half slot of 16 channels; warps 4-7 epilogue; 8 TMA; 9 MMA;
//     14 x streamer; 15 worker
//   x streamer  one thread: tensor-map TMA boxes of the upcoming tiles of x (32 or 16 channels x 128 frames, zero
//               fill past the end of a clip) into a shared-memory slot ring -- twice per tile: once for the row
//               maxima, once for the conversion (the second pass hits L2)
//   loaders     consume the x slots (conflict-free 16-byte reads, thread = 4 frames x 16 channels), derive the
//               per-frame scales, write the tile's K-major SWIZZLE_64B fp16 image and its fp32 rows to per-CTA
//               scratch, up to a tile pair ahead of the MMAs; between tiles they work on jobs / records (below).
//               Clips that are short or whose length is not a multiple of 4 frames are read with plain loads.
//   TMA producer  one thread streams A (residual image) and B (pre-packed codebook image) chunks with
//               cp.async.bulk into a 4 x 24 KiB ring (7 stages without x slots); with CL > 1 the CTAs of a
//               cluster share one multicast codebook stream
//   MMA issuer  one elected lane, 2 tcgen05.mma (M128 N256 K16) per ring stage into one of two 256-column TMEM
//               accumulators
//   epilogue    thread = frame; warp w reads TMEM lanes 32 (w % 4)..: the two sweeps per pass, the final filter of
//               the (tile, stage, group), then it publishes a JOB (or, single-stage: appends RECORDS)
//   worker, and every loader / epilogue warp that would otherwise wait (non-suspending polls: mbarrier.try_wait
//               parks a thread before it reports "not yet"): claim batches of rows of the open jobs -- exact
//               re-score of the undecided rows, write the codes, and between two stages r <- r - e[i] in fp32
//               exactly as the reference does (core_vq.py:359 / :304), new row scale, new fp16 image row, its
//               rounding-residual norms and its bias-chunk row.  Whoever completes the last batch of a job arrives
//               on the barrier the TMA thread (next stage's image) waits on.
// Multi-stage calls process tiles in pairs with interleaved stages -- (A,s0)(B,s0)(A,s1)(B,s1)... -- so one tile's
// job overlaps the other's MMAs.  Single-stage calls settle their undecided frames through the deferred re-score
// queue (see steal_queue).  Codes only; quantized / loss / EMA outputs come from rvq_replay.cu.
// Shapes: K % 256 == 0, K <= 1024, (D/G) % 64 == 0, D/G <= 512, G <= 4.
// ACQ_TC_DBG bits (per call; measurement aids, several make the results wrong): 1 loaders skip the conversion,
// 2 no operand copies, 8 no sweeps, 16 MMA issuer free-runs, 32 no exact re-score, 128 no MMAs, 512 stall /
// phase counters, 4096 four image buffers for single-stage calls, 8192 job slots with fp32 rows, 16384 L2 evict
// hints on the x stream, 32768 two tile buffers, 65536 first version of the single-stage re-score (job slots +
// gather from x), 262144 records without scoring, 2097152 loaders never claim records while a buffer is due,
// 16777216 / 33554432 no fp32-row / image stores, 134217728 two row buffers, 268435456 / 1073741824 epilogue /
// loaders with the other kind of barrier poll, 536870912 one tile at a time, 2147483648 epilogue never helps.
#include "tc_common.cuh"
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

namespace acq {
namespace {

using namespace tc;

constexpr int NSTAGE_MAX = 7;                        // operand-ring stages: 4 next to the x slots, 7 without them
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;       // 8 + 16 KiB: one 32-channel chunk of A and of B
constexpr int XCH = 32;                              // channels per x-staging slot
constexpr int XSLOT_BYTES = XCH * BM * 4;            // [32 channels][128 frames] fp32 = 16 KiB
constexpr int NXS_MAX = 8;                           // x-staging slots: 64 KiB in flight per SM cover the HBM
                                                     // latency at full bandwidth (two slots left the loaders
                                                     // waiting for data two thirds of the time)
// Threads per CTA.  ONE epilogue set everywhere (round 2): with two sets the kernel ran 512 threads at 128 registers
// and its epilogue / publish path was riddled with spills -- several hundred local loads and stores, each an L2 round
// trip next to 224 KiB of shared memory (the publisher spent ~20 kcycles per (tile, stage) between the exchange
// barrier and the open job, against 17 kcycles of MMAs).  384 threads leave 168 registers per thread.  The wide
// layout of single-stage calls keeps 512: its four extra warps are loaders.
#ifndef ACQ_P1_THREADS
#define ACQ_P1_THREADS 384      // (512: warps 12-15 are additional job workers, at 128 registers per thread)
#endif
constexpr int threads_for(int nxslot) { return nxslot == 8 ? 512 : ACQ_P1_THREADS; }
constexpr int NI = 2;                                // tiles of a CTA whose stages are interleaved
constexpr int NTB = 2 * NI;                          // tile buffers per CTA
constexpr int CMAXS = 6;                             // candidates kept per frame, stage and epilogue set
constexpr int CG = 6;                                // 4-column groups recorded per frame, stage and set
constexpr int CG1 = 2 * CG;                          // ... of the (single) epilogue set, which owns both halves of the list area
constexpr int NJOB = 2;                              // job slots (a job may still be open when the next is published)
constexpr int NBAR = 2 * NSTAGE_MAX + 4 + 2 * NTB + NI * GMAX + 2 * NXS_MAX;

struct Job {
    const float* cbp;        // fp32 codebook of this (stage, group)
    const float* x;          // latents (re-score source when the call keeps no fp32 rows: S == 1)
    float* R;                // fp32 residual rows of the tile (nullptr when S == 1)
    uint8_t* img;            // fp16 image of the tile (next stage's A operand)
    uint8_t* bias_img;       // this group's bias chunk inside the image
    float* sc_g;             // row scales of this group            [BM]
    float* nrm_g;            // {||x^||^2, ||dx||^2} of this group  [BM][2]
    uint64_t* bar;           // next-stage image ready (count 1) or tile buffer free (count G)
    int64_t* codes;          // output row of this table, offset to the tile's first frame
    long long n0;            // first frame of the tile
    float xs_cap, inv_bscale;   // of the NEXT stage's table (row-scale cap, 1 / bscale)
    int Dg, D, g, nf, ste, last, T, K;
    unsigned long long* stall;   // stall-attribution counters (ACQ_TC_DBG bit 512) or nullptr
    int seq;                 // generation of the job (claims carry it: see steal_jobs)
    int items;               // last stage: the undecided rows (amb[0 .. items)), one re-score each (the epilogue
                             // has written the decided codes); otherwise batches of rows (job_items(Dg)): every
                             // row is decided if need be, written and carried into the next stage
};
struct alignas(16) JobSlot {
    Job job;
    int state[4];                    // claim = generation << 8 | next item, completed, generation, items
    int n[2][BM];                    // candidates of the row found by epilogue set 0 / 1 (> CMAXS: all K)
    int cand_idx[2][CMAXS][BM];
    int amb[BM];                     // last stage: rows that need the exact re-score
    int namb;
    int pad[3];
};
static_assert(sizeof(Job) <= 152, "job descriptor");

// ---- shared memory carve-up (after the operand ring and the x slots) -----------------------------------
constexpr int OFF_BAR = 0;
constexpr int OFF_TMEM = OFF_BAR + NBAR * 8;
constexpr int OFF_SCALE = OFF_TMEM + 16;                          // [NTB][GMAX][BM] f32
constexpr int OFF_NRM = OFF_SCALE + NTB * GMAX * BM * 4;          // [NTB][GMAX][BM][2] f32
constexpr int OFF_MAX = OFF_NRM + NTB * GMAX * BM * 8;            // [GMAX][BM] u32 (loaders)
constexpr int OFF_GSET = OFF_MAX + GMAX * BM * 4;                 // [2][BM] f32: each set's maximum of the stage
constexpr int OFF_GREC = OFF_GSET + 2 * BM * 4;                   // [2][CG][BM] {group max, first codeword | hits << 12}
constexpr int OFF_JOB = OFF_GREC + 2 * CG * BM * 8;               // [NJOB] JobSlot
constexpr int OFF_DONE = OFF_JOB + NJOB * (int)sizeof(JobSlot);   // int: epilogue finished
constexpr int OFF_Q = OFF_DONE + 16;                              // deferred re-score queue: 4 counters + QCtx
constexpr int NRC = 4;                                           // fp32-row buffers of a single-stage call (= NTB)
constexpr int CTRL_BYTES = OFF_Q + 16 + 16 + 80;                  // counters, records pending per tile copy, QCtx
// The operand ring is bound by its round trip (tcgen05.commit -> empty barrier -> TMA issue -> L2 -> full
// barrier, ~3000 cycles measured with ACQ_TC_DBG ablations: a launch with NO copies and NO MMAs still takes
// stages x 570 cycles with four stages), so the bytes in flight decide the rate: 4 stages next to the 64 KiB of
// x slots, 7 stages when x is read with plain loads.
// (nxs == 8: the wide layout, eight half slots of 16 channels -- the same 64 KiB as four whole ones)
constexpr size_t smem_bytes(int nst, int nxs) { return 1024 + (size_t)nst * STAGE_BYTES + (size_t)(nxs == 8 ? 4 : nxs) * XSLOT_BYTES + CTRL_BYTES; }
static_assert(smem_bytes(4, 4) <= 227 * 1024 && smem_bytes(4, 8) <= 227 * 1024 && smem_bytes(NSTAGE_MAX, 0) <= 227 * 1024, "shared memory budget");
static_assert(OFF_JOB % 16 == 0 && sizeof(JobSlot) % 16 == 0, "alignment");

__device__ __forceinline__ int job_items(int Dg) { return Dg <= 128 ? BM / 8 : (Dg <= 256 ? BM / 4 : BM / 2); }

// explicit shared-memory accesses with 32-bit addresses: with 224 KiB of shared memory there is no L1 left,
// a spilled 64-bit generic pointer costs an L2 round trip on every use (ncu: the group list's base pointers
// were spilled and every recorded group waited ~300 cycles for them)
__device__ __forceinline__ void sts64(uint32_t saddr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(saddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts64_if(uint32_t saddr, uint32_t a, uint32_t b, uint32_t pred) {
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "setp.ne.b32 P, %3, 0;\n\t"
        "@P st.shared.v2.b32 [%0], {%1, %2};\n\t"
        "}\n" ::"r"(saddr), "r"(a), "r"(b), "r"(pred)
        : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
    return v;
}
// fire-and-forget shared-memory reductions (the generic atomicAdd / atomicMax compile to ATOM with a return
// value and a generic-address check: a quarter of the loaders' conversion-sweep stall samples)
__device__ __forceinline__ void red_shared_add_f32(uint32_t saddr, float v) {
    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_shared_max_u32(uint32_t saddr, uint32_t v) {
    asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint2 lds64(uint32_t saddr) {
    uint2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr) : "memory");
    return v;
}
// one tensor-map TMA box -> shared memory (c0 = innermost coordinate)
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::
            "r"(smem_u32(dst_smem)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// same with an L2 eviction policy (createpolicy): the streamer reads every tile of x twice -- the first read is
// marked evict_last so that the second one, marked evict_first, finds it in L2
__device__ __forceinline__ void tma_load_2d_hint(void* dst_smem, const CUtensorMap* map, int c0, int c1, uint64_t* bar,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;\n" ::
            "r"(smem_u32(dst_smem)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
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
        const __half2 h = __floats2half2_rn(v0, v1);             // one packed conversion (low half = v0)
        const float2 f = __half22float2(h);
        qh = fmaf(f.x, f.x, fmaf(f.y, f.y, qh));
        const float e0 = v0 - f.x, e1 = v1 - f.y;
        qd = fmaf(e0, e0, fmaf(e1, e1, qd));
        w[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- exact re-score of one frame (one warp; lanes across channels) ---------------------------------
// r = the frame's fp32 residual of group j.g: from the tile's fp32 rows (S > 1) or straight from x (S == 1:
// channel stride T; only the undecided frames pay this gather).  The candidates of the two epilogue sets
// come from interleaved index ranges, so exact ties are broken explicitly towards the lowest index.
// exact scores of a frame's candidates (rv = its fp32 residual, lanes across channels): float64 dot products,
// (value, lowest index) argmax.  cand(i) must be warp-uniform.
template <int JN, class CandF>
__device__ __forceinline__ int rescore_core(const float4 (&rv)[JN], const float* cbp, int Dg, int n_iter, bool full,
                                            CandF cand, int lane) {
    double best = -INFINITY;
    int best_k = 0x7fffffff;
    auto score = [&](const float4 (&ev)[JN]) {
        double dot = 0.0, nrm = 0.0;
#pragma unroll
        for (int q = 0; q < JN; ++q) {
            dot = fma((double)rv[q].x, (double)ev[q].x, dot); nrm = fma((double)ev[q].x, (double)ev[q].x, nrm);
            dot = fma((double)rv[q].y, (double)ev[q].y, dot); nrm = fma((double)ev[q].y, (double)ev[q].y, nrm);
            dot = fma((double)rv[q].z, (double)ev[q].z, dot); nrm = fma((double)ev[q].z, (double)ev[q].z, nrm);
            dot = fma((double)rv[q].w, (double)ev[q].w, dot); nrm = fma((double)ev[q].w, (double)ev[q].w, nrm);
        }
        return dot - 0.5 * nrm;
    };
    auto load_e = [&](int k, float4 (&ev)[JN]) {
        const float* e = cbp + (size_t)k * Dg;
#pragma unroll
        for (int q = 0; q < JN; ++q) {
            const int d = lane * 4 + 128 * q;
            ev[q] = d < Dg ? __ldg(reinterpret_cast<const float4*>(e + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    // two candidates per iteration: both codeword rows are in flight together (and, in the first iteration,
    // together with the residual row: nothing has consumed it yet)
    for (int i = 0; i < n_iter; i += 2) {
        const bool two = i + 1 < n_iter;
        const int k0 = full ? i : cand(i);
        const int k1 = two ? (full ? i + 1 : cand(i + 1)) : k0;
        float4 ev0[JN], ev1[JN];
        load_e(k0, ev0);
        load_e(k1, ev1);
        double s0 = score(ev0), s1 = score(ev1);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, off);
            s1 += __shfl_xor_sync(0xffffffffu, s1, off);
        }
        if (s0 > best || (s0 == best && k0 < best_k)) { best = s0; best_k = k0; }
        if (two && (s1 > best || (s1 == best && k1 < best_k))) { best = s1; best_k = k1; }
    }
    return best_k == 0x7fffffff ? 0 : best_k;
}
// a frame's channels straight from x (channel stride T: 4 bytes of every sector touched)
template <int JN>
__device__ __forceinline__ void load_x_row(float4 (&rv)[JN], const float* src, int Dg, int T, int lane) {
#pragma unroll
    for (int q = 0; q < JN; ++q) {
        const int d = lane * 4 + 128 * q;
        if (d < Dg) {
            rv[q].x = __ldg(src + (size_t)d * T);
            rv[q].y = __ldg(src + (size_t)(d + 1) * T);
            rv[q].z = __ldg(src + (size_t)(d + 2) * T);
            rv[q].w = __ldg(src + (size_t)(d + 3) * T);
        } else {
            rv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}
template <int JN>
__device__ __forceinline__ int rescore_row(const Job& j, const JobSlot* slot, int row, int na, int nb, int lane) {
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
        load_x_row<JN>(rv, j.x + ((size_t)b * j.D + (size_t)j.g * j.Dg) * j.T + t, j.Dg, j.T, lane);
    }
    const bool full = na > CMAXS || nb > CMAXS;
    auto cand = [&](int i) { return i < na ? slot->cand_idx[0][i][row] : slot->cand_idx[1][i - na][row]; };
    return rescore_core<JN>(rv, j.cbp, j.Dg, full ? j.K : na + nb, full, cand, lane);
}

// Single-stage calls re-score from x itself: a frame's channels are T floats apart, i.e. 512 sectors of which 4
// bytes each are used, and by the time a tile's rows are re-scored the x stream has pushed them out of L2 --
// one HBM round trip per row, which made the re-scores 0.26 ms of a 0.86 ms launch (they are ~7 % of the
// rows).  The row after the one being processed is therefore prefetched into L2 (and so are its candidates'
// codewords, which usually are there already).
template <int JN>
__device__ __forceinline__ void prefetch_undecided(const Job& j, const JobSlot* slot, int item, int lane) {
    const int row = slot->amb[item];
    const long long nfr = j.n0 + row;
    const long long b = nfr / j.T, t = nfr % j.T;
    const float* src = j.x + ((size_t)b * j.D + (size_t)j.g * j.Dg) * j.T + t;
#pragma unroll
    for (int q = 0; q < JN; ++q) {
        const int d = lane * 4 + 128 * q;
        if (d < j.Dg) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(src + (size_t)(d + u) * j.T));
        }
    }
}

// Last-stage job: item = one undecided row; its exact re-score and its code.
template <int JN>
__device__ __forceinline__ void process_undecided(const Job& j, const JobSlot* slot, int item, int lane) {
    if (!j.R && item + 1 < j.items) prefetch_undecided<JN>(j, slot, item + 1, lane);
    const int row = slot->amb[item];
    const int idx = rescore_row<JN>(j, slot, row, slot->n[0][row], slot->n[1][row], lane);
    if (lane == 0) j.codes[row] = (int64_t)idx;
}

// One batch of RB rows of a job: decide the undecided rows, write the codes, update the residual.
template <int RB, int JN>
__device__ __forceinline__ void process_batch(const Job& j, const JobSlot* slot, int item, int lane) {
    const int row0 = item * RB;
    int idxs[RB];
    int mine = 0;
    const long long tq0 = j.stall ? clock64() : 0;
#pragma unroll
    for (int u = 0; u < RB; ++u) {
        const int row = row0 + u;
        const int na = slot->n[0][row], nb = slot->n[1][row];    // (warp-uniform: shared memory)
        int idx = na > 0 ? slot->cand_idx[0][0][row] : (nb > 0 ? slot->cand_idx[1][0][row] : 0);
        if (na + nb > 1 && row < j.nf) idx = rescore_row<JN>(j, slot, row, na, nb, lane);
        idxs[u] = idx;
        if (lane == u) mine = idx;
    }
    if (lane < RB && row0 + lane < j.nf) j.codes[row0 + lane] = (int64_t)mine;
    const long long tq1 = j.stall ? clock64() : 0;
    residual_update_batch<RB, JN, false, true, 1, true>(row0, lane, j.nf, idxs, j.cbp, j.Dg, j.D, j.g, j.R, j.img,
                                                        j.sc_g, j.nrm_g, j.ste != 0, 0, j.bias_img, j.xs_cap,
                                                        j.inv_bscale);
    if (j.stall && lane == 0) {          // (decisions incl. exact re-scores | residual update)
        atomicAdd(j.stall + 18, (unsigned long long)(tq1 - tq0));
        atomicAdd(j.stall + 19, (unsigned long long)(clock64() - tq1));
    }
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
        unsigned long long* stall = nullptr;
        uint32_t t_pub = 0;
#pragma unroll 1
        while (total + mine < budget) {
            // The claim word carries the job's generation: the number of items differs from job to job, so
            // a claim taken from a counter that belongs to the slot's previous job must not be mistaken for
            // an item of the next one (the publisher rewrites the descriptor before it resets the counter).
            int claim = -1;
            if (lane == 0 && (st[0] & 0xff) < st[3]) claim = atomicAdd(const_cast<int*>(st), 1);
            claim = __shfl_sync(0xffffffffu, claim, 0);
            if (claim < 0) break;
            __threadfence_block();                               // the job was written before the claim word
            const Job j = slot->job;
            const int item = claim & 0xff;
            if ((claim >> 8) != j.seq || item >= j.items) break;
            items = j.items;
            bar = j.bar;
            stall = j.stall;
            t_pub = (uint32_t)slot->pad[0];
            const long long t_b = stall ? clock64() : 0;
            if (j.last) {
                if (j.Dg <= 128) process_undecided<1>(j, slot, item, lane);
                else if (j.Dg <= 256) process_undecided<2>(j, slot, item, lane);
                else process_undecided<4>(j, slot, item, lane);
            } else {
                if (j.Dg <= 128) process_batch<8, 1>(j, slot, item, lane);
                else if (j.Dg <= 256) process_batch<4, 2>(j, slot, item, lane);
                else process_batch<2, 4>(j, slot, item, lane);
            }
            if (stall && lane == 0) { atomicAdd(stall + 3, (unsigned long long)(clock64() - t_b)); atomicAdd(stall + 4, 1ULL); }
            ++mine;
        }
        if (mine) {
            // one cross-proxy fence for all the batches of this job this call finished (a job cannot be
            // completed and replaced while a claimed batch is outstanding)
            __syncwarp();
            const long long tf0 = stall ? clock64() : 0;
            fence_proxy_async_global();                          // image rows -> the TMA thread's bulk reads
            __threadfence_block();
            if (stall && lane == 0) atomicAdd(stall + 20, (unsigned long long)(clock64() - tf0));
            if (lane == 0) {
                const int done = atomicAdd(const_cast<int*>(st + 1), mine) + mine;
                if (done == items && bar) {
                    __threadfence_block();
                    mbar_arrive(bar);
                }
                if (done == items && stall) {       // publication -> completion of the whole job
                    atomicAdd(stall + 0, (unsigned long long)((uint32_t)clock64() - t_pub));
                    atomicAdd(stall + 1, 1ULL);
                }
            }
            total += mine;
        }
    }
    return total;
}

// ---- deferred re-score queue (single-stage, single-group calls) ------------------------------------
// A single-stage call has nothing downstream of a frame's code, so its undecided frames need not be settled
// tile by tile.  The first version did: the job slots (above) with the exact re-score reading the frame straight
// from x -- 4 bytes out of each of D sectors, T floats apart.  Measured on cfg2 (ACQ_TC_DBG ablations, kcycles per
// CTA): 1010 without re-scores, 1040 with the re-scores reading a contiguous dummy row, 1480 with the gather from
// x -- the scattered sector requests of ~7 % of the rows slow every other global access of the SM down, the
// loaders' above all.  A bulk copy of the x slots into a tile-local [D][128] scratch copy does not help (the
// gather stays one sector per channel: 1670), the loaders writing the tile's fp32 rows does (1360-1390):
//   * the loaders store the fp32 rows [128][D] next to the image (as multi-stage calls do), into a ring of NRC
//     row buffers that is decoupled from the two image buffers;
//   * the epilogue appends one 64-byte record per undecided frame {frame, na, nb | row << 8 | row buffer << 16,
//     candidates} to a per-CTA ring in global scratch (an image buffer such a call does not use) and moves on:
//     it never waits for a job slot, and the image buffer goes back to the loaders when the tile is published;
//   * the worker warp, loaders whose next buffer is not free yet and, once their own loops have ended, all
//     warps of the CTA claim records, read the row (2 KiB contiguous), score the candidates exactly in
//     float64 and write the code; a row buffer is reused once the records that point into it are finished.
//   qc[0] = records reserved, qc[1] = published (complete), qc[2] = claimed, qc[3] = finished,
//   qc[4 + b] = unfinished records that point into row buffer b
struct QCtx {
    const float* cbp;
    int64_t* codes;
    int* ring;
    const float* rows;       // [NRC] row buffers, [128][D] fp32 each
    int rows_stride;         // floats between two row buffers
    int D, K, cap;
    int dbg;                 // (ablation: 262144 = no scoring)
};
static_assert(sizeof(QCtx) <= 80, "queue context");
constexpr int QREC = 16;                             // ints per record

__device__ __forceinline__ int ld_volatile_s32(const int* p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
template <int JN>
__device__ __forceinline__ void queue_record(const QCtx& c, int v, int lane) {
    const long long nfr = (long long)(uint32_t)__shfl_sync(0xffffffffu, v, 0) | ((long long)__shfl_sync(0xffffffffu, v, 1) << 32);
    const int na = __shfl_sync(0xffffffffu, v, 2), w3 = __shfl_sync(0xffffffffu, v, 3);
    const int nb = w3 & 0xff, trow = (w3 >> 8) & 0xff, rb = w3 >> 16;
    // (the rows were written during this launch by other threads of the CTA: ld.global.cg, no stale L1 lines)
    const float* rrow = c.rows + (size_t)rb * c.rows_stride + (size_t)trow * c.D;
    float4 rv[JN];
#pragma unroll
    for (int q = 0; q < JN; ++q) {
        const int d = lane * 4 + 128 * q;
        rv[q] = d < c.D ? __ldcg(reinterpret_cast<const float4*>(rrow + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const bool full = na > CMAXS || nb > CMAXS;
    const int nca = min(na, CMAXS);
    auto cand = [&](int i) { return __shfl_sync(0xffffffffu, v, i < nca ? 4 + i : 4 + CMAXS + (i - nca)); };
    int idx;
    if (c.dbg & 262144) idx = (int)(rv[0].x + rv[JN - 1].w) & 1;
    else idx = rescore_core<JN>(rv, c.cbp, c.D, full ? c.K : na + nb, full, cand, lane);
    if (lane == 0) c.codes[nfr] = (int64_t)idx;
}
// Claim and process up to `budget` published records; returns the number processed.
__device__ __noinline__ int steal_queue(volatile int* qc, const QCtx* ctx, int lane, int budget = 0x7fffffff) {
    int total = 0;
#pragma unroll 1
    while (total < budget) {
        int claim = -1;
        if (lane == 0) {
            int t = qc[2];
            const int pub = qc[1];
            while (t < pub) {
                const int o = atomicCAS(const_cast<int*>(qc + 2), t, t + 1);
                if (o == t) { claim = t; break; }
                t = o;
            }
        }
        claim = __shfl_sync(0xffffffffu, claim, 0);
        if (claim < 0) break;
        __threadfence_block();                               // the records were written before qc[1] moved
        const QCtx c = *ctx;
        const int* rec = c.ring + (size_t)(claim % c.cap) * QREC;
        const int v = lane < QREC ? ld_volatile_s32(rec + lane) : 0;
        if (c.D <= 128) queue_record<1>(c, v, lane);
        else if (c.D <= 256) queue_record<2>(c, v, lane);
        else queue_record<4>(c, v, lane);
        const int rb = __shfl_sync(0xffffffffu, v, 3) >> 16;
        __syncwarp();
        if (lane == 0) {
            atomicSub(const_cast<int*>(qc + 4 + rb), 1);        // one record less on its row buffer
            atomicAdd(const_cast<int*>(qc + 3), 1);
        }
        ++total;
    }
    return total;
}

// (polling helpers are warp-uniform: the callers go on to warp-collective code -- shuffles in steal_jobs,
//  tcgen05.ld -- so every lane must take the same decision even if the flag flips between two lanes' reads)
__device__ __forceinline__ bool slot_done(const JobSlot* slot) {
    const volatile int* st = slot->state;
    const bool d = st[1] >= st[3];
    return __all_sync(0xffffffffu, d);
}
// Non-suspending poll for warps that have other work (job batches, queue records) while the barrier is pending:
// mbarrier.try_wait parks the thread for a system-dependent time before it reports "not yet", and a helper that
// parks between two batches is no helper -- measured on cfg4 (64 x 10 s, D = 512, n_q = 12): a stage's job took
// 84-97 kcycles from publication to completion with ~60 batches of 6-7 kcycles, i.e. about four of the thirteen
// helper warps were effectively working (the eight epilogue warps sat in try_wait), against 17 kcycles of MMAs.
__device__ __forceinline__ bool warp_test_wait(uint64_t* bar, uint32_t parity) {
    return __all_sync(0xffffffffu, mbar_test_wait(bar, parity));
}
__device__ __forceinline__ bool warp_try_wait(uint64_t* bar, uint32_t parity) {
    return __all_sync(0xffffffffu, mbar_try_wait(bar, parity));
}
__device__ __forceinline__ bool warp_flag_set(volatile int* flag) { return __all_sync(0xffffffffu, *flag != 0); }

template <int CL, int NSTAGE, int NXSLOT>
__global__ void __launch_bounds__(threads_for(NXSLOT), 1)
rvq_search_p1_kernel(const TcParams p, const __grid_constant__ CUtensorMap xmap) {
    if (guard_skips(p)) return;
    constexpr int NXS = NXSLOT > 0 ? NXSLOT : 1;         // (slot arithmetic of the dead streaming path when NXSLOT == 0)
    // WIDE (NXSLOT == 8; single-stage, single-group calls with streamed x): EIGHT loader warps -- warps 10-13 load
    // instead of forming the second epilogue set, which such a call does not need (one set is busy ~45 % of it) --
    // each with its own half slot of 16 channels x 128 frames
    constexpr bool WIDE = NXSLOT == 8;
    constexpr int XCHk = WIDE ? 16 : XCH;               // channels per slot
    constexpr int XSB = XCHk * BM * 4;                  // bytes per slot
    constexpr int NLT = WIDE ? 256 : 128;               // loader threads
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float* xslot = reinterpret_cast<float*>(smem + NSTAGE * STAGE_BYTES);   // [NXS][XCH][BM]
    uint8_t* ctrl = smem + NSTAGE * STAGE_BYTES + NXSLOT * XSB;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl + OFF_BAR);   // [NSTAGE] TMA bytes landed
    uint64_t* empty_bar = full_bar + NSTAGE_MAX;                        // [NSTAGE] MMAs retired
    uint64_t* tfull_bar = empty_bar + NSTAGE_MAX;                       // [2] accumulator complete
    uint64_t* tempty_bar = tfull_bar + 2;                               // [2] accumulator drained
    uint64_t* t0_bar = tempty_bar + 2;                                  // [NTB] stage-0 image of a tile ready
    uint64_t* free_bar = t0_bar + NTB;                                  // [NTB] tile buffer reusable
    uint64_t* upd_bar = free_bar + NTB;                                 // [NI][GMAX] next-stage image ready
    uint64_t* xfull_bar = upd_bar + NI * GMAX;                          // [NXS] x slot filled
    uint64_t* xempty_bar = xfull_bar + NXS_MAX;                         // [NXS] x slot consumed
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(ctrl + OFF_TMEM);
    float* scale_s = reinterpret_cast<float*>(ctrl + OFF_SCALE);
    float* nrm_s = reinterpret_cast<float*>(ctrl + OFF_NRM);
    uint32_t* rowmax_s = reinterpret_cast<uint32_t*>(ctrl + OFF_MAX);
    float* gset_s = reinterpret_cast<float*>(ctrl + OFF_GSET);
    uint2* grec_s = reinterpret_cast<uint2*>(ctrl + OFF_GREC);
    JobSlot* slots = reinterpret_cast<JobSlot*>(ctrl + OFF_JOB);
    volatile int* all_done = reinterpret_cast<volatile int*>(ctrl + OFF_DONE);
    volatile int* qc = reinterpret_cast<volatile int*>(ctrl + OFF_Q);
    QCtx* qctx = reinterpret_cast<QCtx*>(ctrl + OFF_Q + 32);        // (behind the 4 counters and NRC pending counts)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = p.S, G = p.G, K = p.K, D = p.D, Dg = p.Dg, T = p.T;
    const int NP = K / BN, NKC = Dg / BK;
    const bool ste = p.flags & ACQ_STE;
    // single-stage, single-group calls settle their undecided frames through the deferred queue (steal_queue);
    // ACQ_TC_DBG bit 65536 = the job slots and the gather from x instead (the first version)
    const bool defer = S * G == 1 && !(p.dbg_mode & (65536 | 8192));
    const bool keep_rows = S > 1 || (p.dbg_mode & 8192) || defer;   // fp32 rows of the tile in scratch (re-score source)
    const size_t tile_elems = (size_t)BM * D;
    // scratch layout as in the three-product kernel (same workspace): [buf][CTA] images, then [buf][CTA]
    // fp32 rows.  An image slot holds the hi image (A_BYTES per 32-channel chunk) and, behind it, one bias
    // chunk per group.
    const size_t buf_stride = (size_t)kNumSMs * tile_elems * 4;
    uint8_t* Aimg = reinterpret_cast<uint8_t*>(p.scratch) + (size_t)blockIdx.x * tile_elems * 4;
    float* Rbuf = reinterpret_cast<float*>(Aimg + NTB * buf_stride);
    // single-stage calls keep the scratch working set small: 2 tile buffers
    const uint32_t ntb = ((S * G == 1 && !(p.dbg_mode & 4096)) || (p.dbg_mode & 32768)) ? 2u : (uint32_t)NTB;
    const uint32_t ni = (S * G == 1 || (p.dbg_mode & 536870912)) ? ((p.dbg_mode & 536870912) ? 1u : 2u) : (uint32_t)NI;   // (bit 536870912: one tile at a time)
    const uint32_t nrc = (p.dbg_mode & 134217728) ? 2u : (uint32_t)NRC;    // (experiment: two row buffers, scratch within L2)
    const bool lazy_loaders = defer && (p.dbg_mode & 2097152);   // (experiment: loaders never claim records while a buffer is due)
    auto help = [&](int budget) { return defer ? steal_queue(qc, qctx, lane, budget) : steal_jobs(slots, lane, budget); };
    // (cluster-uniform: the tile count of the cluster's first CTA, which is the largest)
    const int lead_cta = (int)(blockIdx.x / CL) * CL;
    const uint32_t n_my = p.num_tiles > lead_cta ? (uint32_t)((p.num_tiles - 1 - lead_cta) / (int)gridDim.x + 1) : 0u;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    const long long tile_base = (long long)blockIdx.x, tile_stride = (long long)gridDim.x;
    constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1);
    // x through the shared-memory ring (tensor-map TMA, one box of 32 channels x 128 frames per slot): the host
    // enables it for clips of a multiple of 4 frames that are long enough for clip-aligned tiles to waste little
    const bool stream_x = NXSLOT > 0 && p.tiles_per_clip > 0;
    // first frame (flattened b * T + t) and number of valid frames of a tile
    auto tile_frames = [&](long long tile, long long& n0, int& nf) {
        if (p.tiles_per_clip > 0) {
            const long long b = tile / p.tiles_per_clip;
            const int t0 = (int)(tile % p.tiles_per_clip) * BM;
            n0 = b * T + t0;
            nf = b < (p.N / T) ? min(BM, T - t0) : 0;
        } else {
            n0 = tile * BM;
            nf = (int)max(0LL, min((long long)BM, p.N - n0));
        }
    };
    const size_t bias_chunk0 = (size_t)G * NKC * A_BYTES;     // offset of group 0's bias chunk in an image
    auto table_tail = [&](int table) {
        return reinterpret_cast<const uint32_t*>(p.pack + (size_t)table * p.table_stride + p.img_bytes + p.hn_bytes);
    };

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
            mbar_init(&t0_bar[i], NLT);         // loader threads
            mbar_init(&free_bar[i], G);         // the last-stage job of every group
        }
        for (int i = 0; i < NI * GMAX; ++i) mbar_init(&upd_bar[i], 1);   // whoever completes the update job
        for (int i = 0; i < NXS; ++i) {
            mbar_init(&xfull_bar[i], 1);        // the streamer's arrive.expect_tx
            mbar_init(&xempty_bar[i], WIDE ? 32 : 64);   // the loader warp(s) that consume the slot
        }
        for (int i = 0; i < NJOB; ++i) {
            slots[i].state[0] = 0;              // generation 0, no items: nothing to claim ...
            slots[i].state[1] = 0;              // ... and nothing to wait for
            slots[i].state[2] = i - NJOB;       // sequence number
            slots[i].state[3] = 0;              // items
            slots[i].job.seq = -1;
        }
        *all_done = 0;
        for (int i = 0; i < 4 + NRC; ++i) qc[i] = 0;
        qctx->cbp = p.cb.p[0];
        qctx->codes = p.codes;
        // (the ring lives behind image buffer 0's fp16 image and bias chunk: this kernel uses the first half of an
        //  image slot only)
        qctx->ring = reinterpret_cast<int*>(Aimg + tile_elems * 2 + A_BYTES);
        qctx->rows = Rbuf;
        qctx->rows_stride = (int)(buf_stride / 4);
        qctx->D = D; qctx->K = K; qctx->dbg = p.dbg_mode;
        qctx->cap = (int)((tile_elems * 2 - A_BYTES) / (QREC * 4));      // >= 128 records (D >= 64)
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(tmem_ptr_s, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();      // every CTA's barriers are initialised before any multicast lands
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (warp < 4 || (WIDE && warp >= 10 && warp <= 13)) {
        // ================= loaders: x tile -> scales, fp16 image, rounding norms (and R when S > 1) =====
        const int lw = warp < 4 ? warp : warp - 6;            // loader warp 0..3 (WIDE: 0..7)
        const int ltid = lw * 32 + lane;                      // loader thread
        unsigned long long w_free = 0, w_xfull = 0;
        unsigned long long t_sw0 = 0, t_sw1 = 0, t_bar = 0;      // (phase times of the streamed path, ACQ_TC_DBG bit 512)
        const long long t_begin = clock64();
        uint32_t xit = 0;                     // x slots consumed so far
        for (uint32_t it = 0; it < n_my; ++it) {
            const long long tile = tile_base + (long long)it * tile_stride;   // may be a dummy past the end
            const uint32_t buf = it % ntb;
            const long long tw = clock64();       // (timed from before the first try_wait, which already blocks for a while)
            // (multi-stage calls poll without suspending -- their jobs are on the critical path; a single-stage call's
            //  records are not, and eight loader warps spinning cost the others issue slots: it parks in try_wait)
            const bool park = defer && !(p.dbg_mode & 1073741824);
            if (!(park ? warp_try_wait(&free_bar[buf], ((it / ntb) & 1) ^ 1) : warp_test_wait(&free_bar[buf], ((it / ntb) & 1) ^ 1))) {
                // no buffer to fill yet: work on the open jobs meanwhile
                while (!(park ? warp_try_wait(&free_bar[buf], ((it / ntb) & 1) ^ 1) : warp_test_wait(&free_bar[buf], ((it / ntb) & 1) ^ 1))) {
                    // (one batch per poll: a loader that keeps claiming re-scores while its buffer has long been
                    //  free starves the MMAs of their next tile -- 0.27 ms of a 0.92 ms single-stage launch)
                    if (lazy_loaders || !help(1)) __nanosleep(128);
                    if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 6); __trap(); }
                }
            }
            w_free += (unsigned long long)(clock64() - tw);
            long long n0;
            int nf;
            tile_frames(tile, n0, nf);
            uint8_t* img = Aimg + buf * buf_stride;
            float* R = Rbuf + (defer ? it % nrc : buf) * (buf_stride / 4);
            float* sc = scale_s + buf * GMAX * BM;
            float* nrm = nrm_s + buf * GMAX * BM * 2;
            for (int i = ltid; i < G * BM; i += NLT) {
                rowmax_s[i] = 0u;
                nrm[2 * i] = 0.f;
                nrm[2 * i + 1] = 0.f;
            }
            if (defer) {
                // the row buffer this tile's fp32 rows go to (it % nrc) still serves the re-scores of tile it - NRC
                // (published long ago: the image buffer of tile it - 2 has been handed back)
                const long long tw = clock64();
                while (!__all_sync(0xffffffffu, qc[4 + it % nrc] == 0)) {
                    if (!help(1)) __nanosleep(64);          // (must help: the loaders may be the only idle warps)
                    if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 17); __trap(); }
                }
                w_free += (unsigned long long)(clock64() - tw);
            }
            named_bar_sync(2, NLT);
            if (stream_x) {
                // x arrives in shared memory (warp 14): slot = [32 channels][128 frames] fp32.  Warps 0,1 take the
                // even slots, warps 2,3 the odd ones; thread = 4 consecutive frames x 16 channels (sixteen 16-byte
                // reads, conflict-free), which is one whole 32-byte sector of the image per frame: the image goes
                // straight to scratch, no staging, no barriers inside a sweep.  Sweep 0 folds the row maxima,
                // sweep 1 converts.
                // (WIDE: one warp per half slot of 16 channels, eight of them; otherwise a warp pair per slot of 32)
                const int rq = lane, wp = WIDE ? lw : lw >> 1, hf = WIDE ? 0 : (lw & 1);
                constexpr int CSTEP = WIDE ? 8 : 2;
                const uint32_t xslot_a = smem_u32(xslot), rowmax_a = smem_u32(rowmax_s), nrm_a = smem_u32(nrm);
                for (int sweep = 0; sweep < 2; ++sweep) {
                    const long long t_s = clock64();
                    float m[4] = {0.f, 0.f, 0.f, 0.f};
                    float qh[4] = {0.f, 0.f, 0.f, 0.f}, qd[4] = {0.f, 0.f, 0.f, 0.f};
                    float qh2[4] = {0.f, 0.f, 0.f, 0.f}, qd2[4] = {0.f, 0.f, 0.f, 0.f};
                    // group of the chunk and "this thread's last chunk of the group" by counting (Dg % 64 == 0:
                    // a group is an even number of chunks, so both warp pairs leave it together); the four row
                    // scales are read once per group
                    const int nkc_x = Dg / XCHk;
                    int g = 0, c_in_g = wp;
                    float xs4[4] = {0.f, 0.f, 0.f, 0.f};
                    bool new_group = true;
                    for (int c = wp; c < D / XCHk; c += CSTEP) {
                        const uint32_t xi = xit + (uint32_t)c;
                        const uint32_t xs_i = xi % NXS;
                        {   // (timed around the whole wait: the first try_wait already blocks for a while)
                            const long long tx = clock64();
                            mbar_wait(&xfull_bar[xs_i], (xi / NXS) & 1, p.err, 15);
                            w_xfull += (unsigned long long)(clock64() - tx);
                        }
                        if (p.dbg_mode & 1) { mbar_arrive(&xempty_bar[xs_i]); continue; }   // (ablation: slots only)
                        const uint32_t xrow = xslot_a + xs_i * (uint32_t)XSB + (uint32_t)(hf * 16 * BM * 4 + rq * 16);
                        const bool group_end = c_in_g + CSTEP >= nkc_x;   // this thread's last chunk of the group
                        if (sweep == 0) {
                            float4 v[2][8];
#pragma unroll
                            for (int h = 0; h < 2; ++h)
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    v[h][i] = lds128f(xrow + (uint32_t)((h * 8 + i) * BM * 4));   // (frames past the clip: TMA zero fill)
#pragma unroll
                            for (int h = 0; h < 2; ++h)
#pragma unroll
                                for (int i = 0; i < 8; i += 2) {
                                    m[0] = fmax3(m[0], fabsf(v[h][i].x), fabsf(v[h][i + 1].x));
                                    m[1] = fmax3(m[1], fabsf(v[h][i].y), fabsf(v[h][i + 1].y));
                                    m[2] = fmax3(m[2], fabsf(v[h][i].z), fabsf(v[h][i + 1].z));
                                    m[3] = fmax3(m[3], fabsf(v[h][i].w), fabsf(v[h][i + 1].w));
                                }
                            mbar_arrive(&xempty_bar[xs_i]);
                            if (group_end) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    red_shared_max_u32(rowmax_a + (uint32_t)(g * BM + 4 * rq + j) * 4, __float_as_uint(m[j]));
                                    m[j] = 0.f;
                                }
                            }
                        } else {
                            uint8_t* chunk = img + (size_t)(WIDE ? c >> 1 : c) * A_BYTES;   // (an image chunk = 32 channels)
                            const int oct0 = WIDE ? 2 * (c & 1) : 2 * hf;                   // first of this thread's two octets in it
                            if (new_group) {
                                const float4 s4 = *reinterpret_cast<const float4*>(sc + g * BM + 4 * rq);
                                xs4[0] = s4.x; xs4[1] = s4.y; xs4[2] = s4.z; xs4[3] = s4.w;
                                new_group = false;
                            }
                            // the thread's 16 channels in two halves of 8 (the kernel is compiled at 128 registers for
                            // all roles, and with 224 KiB of shared memory a spilled value is an L2 round trip: with all
                            // 16 float4 live, six reloads sat in this loop -- ~4 kcycles per visit for ~600 instructions)
                            // The four rows advance in lockstep, channel pair by channel pair, each with two
                            // alternating pairs of norm accumulators: eight independent dependency chains.  (A loader
                            // warp is alone on its scheduler; row after row, the serial FMUL -> F2FP -> HADD2 -> FADD ->
                            // FFMA chains left it at ~0.1 instructions per cycle, 5-6 kcycles per visit.)
                            uint4 hw0[4];
                            const bool st_rows = keep_rows && !(p.dbg_mode & 16777216), st_img = !(p.dbg_mode & 33554432);
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                float4 v8[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) v8[i] = lds128f(xrow + (uint32_t)((h * 8 + i) * BM * 4));
                                uint32_t w[4][4];
#pragma unroll
                                for (int i = 0; i < 8; i += 2) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        const float c0 = j == 0 ? v8[i].x : (j == 1 ? v8[i].y : (j == 2 ? v8[i].z : v8[i].w));
                                        const float c1 = j == 0 ? v8[i + 1].x : (j == 1 ? v8[i + 1].y : (j == 2 ? v8[i + 1].z : v8[i + 1].w));
                                        const float a0 = c0 * xs4[j], a1 = c1 * xs4[j];
                                        const __half2 hh = __floats2half2_rn(a0, a1);
                                        const float2 f = __half22float2(hh);
                                        const float e0 = a0 - f.x, e1 = a1 - f.y;
                                        if ((i >> 1) & 1) {
                                            qh2[j] = fmaf(f.x, f.x, fmaf(f.y, f.y, qh2[j]));
                                            qd2[j] = fmaf(e0, e0, fmaf(e1, e1, qd2[j]));
                                        } else {
                                            qh[j] = fmaf(f.x, f.x, fmaf(f.y, f.y, qh[j]));
                                            qd[j] = fmaf(e0, e0, fmaf(e1, e1, qd[j]));
                                        }
                                        w[j][i >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
                                    }
                                }
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const int row = 4 * rq + j;
                                    if (st_rows) {
                                        float a[8];
#pragma unroll
                                        for (int i = 0; i < 8; ++i)
                                            a[i] = j == 0 ? v8[i].x : (j == 1 ? v8[i].y : (j == 2 ? v8[i].z : v8[i].w));
                                        stg256(R + (size_t)row * D + c * XCHk + hf * 16 + h * 8,
                                               make_uint4(__float_as_uint(a[0]), __float_as_uint(a[1]), __float_as_uint(a[2]), __float_as_uint(a[3])),
                                               make_uint4(__float_as_uint(a[4]), __float_as_uint(a[5]), __float_as_uint(a[6]), __float_as_uint(a[7])));
                                    }
                                    const uint4 hh4 = make_uint4(w[j][0], w[j][1], w[j][2], w[j][3]);
                                    if (h == 0) hw0[j] = hh4;
                                    else if (st_img) store_chunk_pair(chunk, row, oct0, hw0[j], hh4);
                                    else if (hh4.x == 0x12345678u) qh[j] += 1.f;   // (ablation: keep the conversion alive)
                                }
                            }
                            mbar_arrive(&xempty_bar[xs_i]);
                            if (group_end) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    red_shared_add_f32(nrm_a + (uint32_t)(g * BM + 4 * rq + j) * 8, qh[j] + qh2[j]);
                                    red_shared_add_f32(nrm_a + (uint32_t)(g * BM + 4 * rq + j) * 8 + 4, qd[j] + qd2[j]);
                                    qh[j] = 0.f; qh2[j] = 0.f;
                                    qd[j] = 0.f; qd2[j] = 0.f;
                                }
                            }
                        }
                        c_in_g += CSTEP;
                        if (c_in_g >= nkc_x) { c_in_g -= nkc_x; ++g; new_group = true; }
                    }
                    xit += (uint32_t)(D / XCHk);
                    if (sweep == 0) t_sw0 += (unsigned long long)(clock64() - t_s); else t_sw1 += (unsigned long long)(clock64() - t_s);
                    if (sweep == 0) {
                        const long long t_b = clock64();
                        named_bar_sync(2, NLT);
                        t_bar += (unsigned long long)(clock64() - t_b);
                        for (int i = ltid; i < G * BM; i += NLT) {
                            // row scale, capped so that the bias factor w = xs / bscale fits fp16
                            const float bscale = __uint_as_float(__ldg(table_tail(i / BM) + TAIL_BSCALE));
                            const float xs = fminf(scale_for(__uint_as_float(rowmax_s[i])), 32768.f * bscale);
                            sc[i] = xs;
                            write_bias_row(img + bias_chunk0 + (size_t)(i / BM) * A_BYTES, i % BM, xs / bscale);
                        }
                        named_bar_sync(2, NLT);
                    }
                }
            } else if ((T & 3) == 0) {
                // 4 consecutive frames per thread (one 16 B load per channel), 16 channels at a time
                const int rq = tid & 31, w4 = tid >> 5;
                const long long n = n0 + 4 * rq;
                const bool ok = 4 * rq < nf;             // T % 4 == 0: a quad is all in or all out
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
                                // 16 channels of a frame = one whole 32-byte sector of the image
                                store_chunk_pair(img + (size_t)(oct / CPR) * A_BYTES, row, oct % CPR, h0, h1);
                                if (keep_rows) {
                                    float* rd = R + (size_t)row * D + oct * 8;
                                    stg256(rd, make_uint4(__float_as_uint(a[0][0]), __float_as_uint(a[0][1]), __float_as_uint(a[0][2]), __float_as_uint(a[0][3])),
                                           make_uint4(__float_as_uint(a[0][4]), __float_as_uint(a[0][5]), __float_as_uint(a[0][6]), __float_as_uint(a[0][7])));
                                    stg256(rd + 8, make_uint4(__float_as_uint(a[1][0]), __float_as_uint(a[1][1]), __float_as_uint(a[1][2]), __float_as_uint(a[1][3])),
                                           make_uint4(__float_as_uint(a[1][4]), __float_as_uint(a[1][5]), __float_as_uint(a[1][6]), __float_as_uint(a[1][7])));
                                }
                            }
                        }
                    }
                    if (sweep == 0) {
                        named_bar_sync(2, 128);
                        for (int i = tid; i < G * BM; i += 128) {
                            // row scale, capped so that the bias factor w = xs / bscale fits fp16
                            const float bscale = __uint_as_float(__ldg(table_tail(i / BM) + TAIL_BSCALE));
                            const float xs = fminf(scale_for(__uint_as_float(rowmax_s[i])), 32768.f * bscale);
                            sc[i] = xs;
                            write_bias_row(img + bias_chunk0 + (size_t)(i / BM) * A_BYTES, i % BM, xs / bscale);
                        }
                        named_bar_sync(2, 128);
                    }
                }
            } else {
                // general T (e.g. HiFi-Codec's 50 frames per clip): one frame per thread, scalar loads that
                // are coalesced across the warp, 32 channels (one chunk of the image) per iteration
                const int row = tid;
                const long long n = n0 + row;
                const bool ok = row < nf;
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
                            if (keep_rows) {
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
                        for (int i = tid; i < G * BM; i += 128) {
                            // row scale, capped so that the bias factor w = xs / bscale fits fp16
                            const float bscale = __uint_as_float(__ldg(table_tail(i / BM) + TAIL_BSCALE));
                            const float xs = fminf(scale_for(__uint_as_float(rowmax_s[i])), 32768.f * bscale);
                            sc[i] = xs;
                            write_bias_row(img + bias_chunk0 + (size_t)(i / BM) * A_BYTES, i % BM, xs / bscale);
                        }
                        named_bar_sync(2, 128);
                    }
                }
            }
            fence_proxy_async_global();     // generic-proxy global writes -> TMA (async proxy) reads
            mbar_arrive(&t0_bar[buf]);
        }
        if ((p.dbg_mode & 512) && tid == 0) {
            atomicAdd(p.stall + 5, w_free);
            atomicAdd(p.stall + 6, w_xfull);
            atomicAdd(p.stall + 8, (unsigned long long)(clock64() - t_begin));
        }
        if ((p.dbg_mode & 512) && lane == 0 && warp < 4 && S * G == 1) {
            // per loader warp: the two sweeps and the barrier between them (warps 0..3 -> slots 15.., 18.., ...)
            atomicAdd(p.stall + 15 + 2 * warp, t_sw0);
            atomicAdd(p.stall + 16 + 2 * warp, t_sw1);
            if (warp == 0) atomicAdd(p.stall + 23, t_bar);
        }
        // all tiles loaded: keep working on jobs until the epilogue has finished its last tile
        const long long tw = clock64();
        while (!warp_flag_set(all_done)) {
            if (!help(0x7fffffff)) __nanosleep(128);
            if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 10); __trap(); }
        }
    } else if (WIDE ? warp == 15 : warp >= 11) {
        // ================= worker: jobs only (the other helpers only work while they would otherwise wait) ==
        const long long tw = clock64();
        while (!warp_flag_set(all_done)) {
            if (!help(0x7fffffff)) __nanosleep(64);
            if (clock64() - tw > 16000000000LL) { if (p.err) atomicExch(p.err, 14); __trap(); }
        }
    } else if (warp == (WIDE ? 14 : 10)) {
        // ================= x streamer: tiles of x -> shared-memory slots, twice per tile =====================
        if (stream_x && lane == 0) {
            unsigned long long w_xempty = 0;
            uint32_t xit = 0;
            const uint64_t pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
            for (uint32_t it = 0; it < n_my; ++it) {
                const long long tile = tile_base + (long long)it * tile_stride;
                const long long b = tile / p.tiles_per_clip;
                const int t0 = (int)(tile % p.tiles_per_clip) * BM;
                const bool real = b < p.N / T;                      // (dummy tiles past the end: no copy)
                for (int rep = 0; rep < 2 * (D / XCHk); ++rep, ++xit) {
                    const int c = rep % (D / XCHk);
                    const uint32_t xs_i = xit % NXS;
                    mbar_wait_t(&xempty_bar[xs_i], ((xit / NXS) & 1) ^ 1, p.err, 16, w_xempty);
                    if (real) {
                        // one box: frames t0 .. t0+127 (columns past the clip are zero-filled) x 32 channel rows
                        mbar_arrive_expect_tx(&xfull_bar[xs_i], XSB);
                        if (p.dbg_mode & 16384)
                            tma_load_2d_hint(xslot + (size_t)xs_i * XCHk * BM, &xmap, t0, (int)(b * D) + c * XCHk, &xfull_bar[xs_i],
                                             rep < D / XCHk ? pol_keep : pol_drop);
                        else
                            tma_load_2d(xslot + (size_t)xs_i * XCHk * BM, &xmap, t0, (int)(b * D) + c * XCHk, &xfull_bar[xs_i]);
                    } else {
                        mbar_arrive(&xfull_bar[xs_i]);
                    }
                }
            }
            if (p.dbg_mode & 512) atomicAdd(p.stall + 9, w_xempty);
        }
        // (own loop finished: help with the deferred re-scores until the epilogue has seen the queue empty)
        __syncwarp();
        if (defer)
            while (!warp_flag_set(all_done))
                if (!help(0x7fffffff)) __nanosleep(256);
    } else if (warp == 8) {
        // ================= TMA producer: one thread streams A and B image chunks ==========================
        // ONE thread executes this loop, i.e. a dependent instruction every 4-6 cycles: the first version (64-bit
        // address products, a locally indexed phase array, timing and ablation branches per stage) ran ~140
        // instructions = 570 cycles per ring stage -- measured with no copies and no MMAs at all -- against the
        // 256 cycles the stage's two MMAs take, and bounded every shape.  Now: 32-bit shared-memory addresses,
        // source pointers advanced by constants, stage / phase counters instead of divisions.
        if (lane == 0 && !(p.dbg_mode & 16)) {
            uint32_t st = 0, ph = 1;                   // ring stage; parity its empty barrier is waited on with
            uint32_t upd_ph = 0;                       // bit (par * GMAX + g): parity of the next wait on upd_bar
            const bool nocopy = (p.dbg_mode & 2) != 0; // (ablation: barriers only)
            const uint32_t ring0 = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
            constexpr uint32_t SLICE = B_BYTES / CL;
            for (uint32_t it0 = 0; it0 < n_my; it0 += ni) {
                const int npair = (int)min(ni, n_my - it0);
                for (int s = 0; s < S; ++s) {
                    for (int h = 0; h < npair; ++h) {
                        const uint32_t it = it0 + h, buf = it % ntb, par = h;
                        const uint8_t* img = Aimg + buf * buf_stride;
                        for (int g = 0; g < G; ++g) {
                            const uint8_t* bimg = p.pack + (size_t)(s * G + g) * p.table_stride;
                            const uint8_t* bbias = bimg + p.bias_off + (CL > 1 ? crank * SLICE : 0u);
                            const uint8_t* bsrc = bimg + (CL > 1 ? crank * SLICE : 0u);
                            const uint8_t* a0 = img + (size_t)(g * NKC) * A_BYTES;
                            const uint8_t* abias = img + bias_chunk0 + (size_t)g * A_BYTES;
                            bool first = true;
                            for (int pass = 0; pass < NP; ++pass, bbias += B_BYTES) {
                                const uint8_t* asrc = a0;
                                // kc == NKC: the bias chunk (A: the tile's {w,w,w,0..} rows, B: {-b1,-b2,-b3,0..})
#pragma unroll 1
                                for (int kc = 0; kc <= NKC; ++kc) {
                                    const uint32_t fb = full0 + st * 8, a_dst = ring0 + st * STAGE_BYTES;
                                    mbar_wait_u32(empty0 + st * 8, ph, p.err, 2);
                                    const bool bias = kc == NKC;
                                    if (!nocopy) {
                                        mbar_expect_tx_u32(fb, A_BYTES + B_BYTES);
                                        if (CL == 1) bulk_g2s_u32(a_dst + A_BYTES, bias ? bbias : bsrc, B_BYTES, fb);
                                        else bulk_g2s_mc_u32(a_dst + A_BYTES + crank * SLICE, bias ? bbias : bsrc, SLICE, fb, CMASK);
                                    }
                                    if (first) {
                                        // first use of this (tile, stage, group)'s image
                                        first = false;
                                        if (s == 0) {
                                            mbar_wait(&t0_bar[buf], (it / ntb) & 1, p.err, 7);
                                        } else {
                                            const uint32_t bit = par * GMAX + g;
                                            mbar_wait(&upd_bar[bit], (upd_ph >> bit) & 1u, p.err, 8);
                                            upd_ph ^= 1u << bit;
                                        }
                                        fence_proxy_async_global();
                                    }
                                    if (!nocopy) bulk_g2s_u32(a_dst, bias ? abias : asrc, A_BYTES, fb);
                                    else mbar_arrive_u32(fb);
                                    bsrc += bias ? 0 : 2 * B_BYTES;                   // (hi images; the lo images are skipped)
                                    asrc += A_BYTES;
                                    if (++st == NSTAGE) { st = 0; ph ^= 1u; }
                                }
                            }
                        }
                    }
                }
            }
        }
        // (own loop finished: help with the deferred re-scores until the epilogue has seen the queue empty)
        __syncwarp();
        if (defer)
            while (!warp_flag_set(all_done))
                if (!help(0x7fffffff)) __nanosleep(256);
    } else if (warp == 9) {
        // ================= MMA issuer: one product per chunk =============================================
        // The whole warp runs the loop and one elected lane issues: under `if (lane == 0)` the compiler keeps
        // the ring addresses in vector registers and pays an R2UR + ELECT chain in front of every tcgen05.mma
        // (ncu: ~600 cycles of issue work per 2-MMA ring stage, more than the 256 cycles the MMAs take).  The
        // operand descriptors are built once; a ring stage adds a constant to their address field.
        {
            uint32_t st = 0, ph = 0, acc_it = 0;          // ring stage and the parity its full barrier is waited on with
            unsigned long long w_full0 = 0, w_full = 0, w_tempty = 0;
            const long long t_begin = clock64();
            const uint64_t desc_a0 = make_desc(smem_u32(smem));
            const uint64_t desc_b0 = make_desc(smem_u32(smem) + A_BYTES);
            const uint32_t full0 = smem_u32(full_bar);
            const bool leader = elect_one();
            const bool free_run = (p.dbg_mode & 16) != 0, no_mma = (p.dbg_mode & 128) != 0;
            const uint32_t n_acc = n_my * (uint32_t)(S * G * NP);
            for (; acc_it < n_acc; ++acc_it) {                // (same number of passes in any order)
                const uint32_t abuf = acc_it & 1;
                mbar_wait_t(&tempty_bar[abuf], ((acc_it >> 1) & 1) ^ 1, p.err, 3, w_tempty);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + abuf * BN;
#pragma unroll 1
                for (int kc = 0; kc <= NKC; ++kc) {
                    if (!free_run) mbar_wait_u32(full0 + st * 8, ph, p.err, 4);
                    tc_fence_after();
                    const uint64_t da = desc_a0 + (uint64_t)(st * (STAGE_BYTES >> 4));
                    const uint64_t db = desc_b0 + (uint64_t)(st * (STAGE_BYTES >> 4));
                    if (leader) {
                        if (no_mma) {
                            // (ablation: handshakes only)
                        } else if (kc < NKC) {
                            umma_f16(d_tmem, da, db, IDESC, kc != 0);
                            umma_f16(d_tmem, da + ((UK * 2) >> 4), db + ((UK * 2) >> 4), IDESC, 1);
                        } else {
                            umma_f16(d_tmem, da, db, IDESC, 1);   // bias chunk: one K16 slice
                        }
                        // ring stage free once these MMAs retire (in every CTA that shares it)
                        if (free_run) {
                        } else if (CL == 1) umma_commit(&empty_bar[st]);
                        else umma_commit_mc(&empty_bar[st], CMASK);
                    }
                    __syncwarp();
                    if (++st == NSTAGE) { st = 0; ph ^= 1u; }
                }
                if (leader) umma_commit(&tfull_bar[abuf]);       // accumulator complete
                __syncwarp();
            }
            if ((p.dbg_mode & 512) && lane == 0) {
                atomicAdd(p.stall + 0, w_full0); atomicAdd(p.stall + 1, w_full); atomicAdd(p.stall + 2, w_tempty);
                atomicAdd(p.stall + 7, (unsigned long long)(clock64() - t_begin));
            }
        }
        // (own loop finished: help with the deferred re-scores until the epilogue has seen the queue empty)
        __syncwarp();
        if (defer)
            while (!warp_flag_set(all_done))
                if (!help(0x7fffffff)) __nanosleep(256);
    } else {
        // ================= epilogue sets: filter sweeps, publish jobs (thread = frame) =====================
        // set 0 = warps 4-7 drains accumulator 0, set 1 = warps 10-13 accumulator 1; warp w reads TMEM lanes
        // 32 (w % 4) .. 32 (w % 4) + 31.
        constexpr int set = 0;                    // (one epilogue set: warps 4-7)
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const bool publisher = tid == 128;
        // this row's group list of this set (shared-memory address): record i at + i * BM * 8 =
        // {maximum of the group, first codeword | which of its four columns were candidates << 12}
        const uint32_t rec_a = smem_u32(grec_s + (size_t)set * CG * BM + row);
        uint32_t acc_it = 0;
        // (32-bit cycle counters: every register that stays live across the sweeps counts)
        uint32_t e_wait = 0, e_sweep = 0, e_slot = 0, n_amb = 0, n_full = 0, e_bar1 = 0, e_pub = 0, e_t0 = 0;
        int job_seq = 0;
        for (uint32_t it0 = 0; it0 < n_my; it0 += ni) {
          const int npair = (int)min(ni, n_my - it0);
          for (int s = 0; s < S; ++s) {
            for (int h = 0; h < npair; ++h) {
                const uint32_t it = it0 + h, buf = it % ntb, par = h;
                long long n0;
                int nf;
                tile_frames(tile_base + (long long)it * tile_stride, n0, nf);
                float* sc = scale_s + buf * GMAX * BM;
                float* nrm = nrm_s + buf * GMAX * BM * 2;
                if (s == 0) { const uint32_t t0c = (uint32_t)clock64(); mbar_wait(&t0_bar[buf], (it / ntb) & 1, p.err, 9); e_t0 += (uint32_t)clock64() - t0c; }   // this tile's scales are visible
                for (int g = 0; g < G; ++g, ++job_seq) {
                    const int table = s * G + g;
                    const uint32_t* tail = table_tail(table);
                    const float emax2 = __uint_as_float(__ldg(tail + TAIL_EMAX2));
                    const float de2max = __uint_as_float(__ldg(tail + TAIL_DE2MAX));
                    const float hnmax = __uint_as_float(__ldg(tail + TAIL_HNMAX));
                    const float bscale = __uint_as_float(__ldg(tail + TAIL_BSCALE));
                    JobSlot* slot = slots + (job_seq & 1);
                    float tau2 = 0.f;
                    float gmax = -INFINITY;
                    int ngrp = 0;
                    for (int pass = 0; pass < NP; ++pass, ++acc_it) {
                        const uint32_t abuf = acc_it & 1;
                        const bool poll = S * G > 1 && !(p.dbg_mode & 268435456);   // (bit 268435456: the parking try_wait, as before)
                        if (!(poll ? warp_test_wait(&tfull_bar[abuf], (acc_it >> 1) & 1) : warp_try_wait(&tfull_bar[abuf], (acc_it >> 1) & 1))) {
                            // nothing to drain yet: work on the open jobs meanwhile, one batch at a time
                            const long long tw = clock64();
                            // (a polling warp costs the single-lane TMA / MMA warps on its scheduler issue slots and
                            //  queue time in the shared-memory pipe: back off when there is nothing to claim; a
                            //  single-stage call has no job on the critical path, so its epilogue sleeps in the
                            //  hardware wait and leaves the re-scores to the loaders and the worker warp)
                            if (S * G == 1) {
                                mbar_wait(&tfull_bar[abuf], (acc_it >> 1) & 1, p.err, 5);
                            } else {
                                while (!(poll ? warp_test_wait(&tfull_bar[abuf], (acc_it >> 1) & 1) : warp_try_wait(&tfull_bar[abuf], (acc_it >> 1) & 1))) {
                                    if ((p.dbg_mode & 2147483648u) || !steal_jobs(slots, lane, 1)) __nanosleep(64);   // (bit 31: never help here)
                                    if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 5); __trap(); }
                                }
                            }
                            e_wait += (uint32_t)(clock64() - tw);
                        }
                        tc_fence_after();
                        const uint32_t tq = (uint32_t)clock64();
                        const uint32_t taddr = tmem_base + abuf * BN + ((uint32_t)(q * 32) << 16);
                        const int kbase = pass * BN;
                        if (tau2 == 0.f) {
                            // the bound of this (row, stage, group); read after the stage's first accumulator
                            // is complete: the row scale and the rounding norms of stage s are written by the
                            // job of stage s-1
                            const float xs = sc[g * BM + row];
                            const float xh = sqrtf(nrm[2 * (g * BM + row)]), xd = sqrtf(nrm[2 * (g * BM + row) + 1]);
                            const float dE = sqrtf(de2max), eh = sqrtf(emax2) + dE;
                            const float w = xs / bscale;                      // bias factor; < 2^-24 flushes to 0
                            const float bias = xs * hnmax;
                            const float tau = xd * eh + (xh + xd) * dE + ((float)(Dg + 16) * 1.1920929e-7f) * (xh * eh + bias) +
                                              2.3841858e-7f * bias + (w < 5.9604645e-8f ? bias : 0.f);
                            tau2 = fmaxf(2.002f * tau, 1e-30f);
                        }
                        // The sweeps are ROLLED loops over 16-column TMEM reads (double-buffered): fully unrolled
                        // 32-column versions measured 10 kcycles per pass, four fifths of it instruction-fetch
                        // stalls -- their code did not fit the instruction cache next to the other roles'.
                        // Both sweeps are BRANCH-FREE.  (The first version kept a running maximum with a vote and
                        // a branch per 4-column group: a serial FMNMX -> FSETP -> VOTE -> BRA chain per group that
                        // two warps per scheduler cannot hide -- ncu: 1200 cycles per 32 columns, 15 kcycles per
                        // pass against 4-8 kcycles of MMAs, so the tensor pipe waited for a drained accumulator a
                        // quarter of the launch.)
                        uint32_t ra[16], rb[16];
                        // sweep 1: the pass maximum (four independent 3-input max chains)
                        if (!(p.dbg_mode & 8)) {
                            float pm0 = -INFINITY, pm1 = -INFINITY, pm2 = -INFINITY, pm3 = -INFINITY;
                            auto max16 = [&](const uint32_t (&r)[16]) {
#pragma unroll
                                for (int j = 0; j < 16; j += 8) {
                                    pm0 = fmax3(pm0, __uint_as_float(r[j]), __uint_as_float(r[j + 1]));
                                    pm1 = fmax3(pm1, __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                                    pm2 = fmax3(pm2, __uint_as_float(r[j + 4]), __uint_as_float(r[j + 5]));
                                    pm3 = fmax3(pm3, __uint_as_float(r[j + 6]), __uint_as_float(r[j + 7]));
                                }
                            };
                            tmem_ld16_async(taddr, ra);
#pragma unroll 1
                            for (int c0 = 0; c0 < BN; c0 += 32) {
                                tmem_ld16_wait(ra);
                                tmem_ld16_async(taddr + c0 + 16, rb);
                                max16(ra);
                                tmem_ld16_wait(rb);
                                if (c0 + 32 < BN) tmem_ld16_async(taddr + c0 + 32, ra);
                                max16(rb);
                            }
                            const float pm = fmax3(fmaxf(pm0, pm1), pm2, pm3);
                            // the groups recorded by this set's earlier passes stay in the list unless the maximum
                            // has climbed past all of them (the final filter drops what has fallen out of range)
                            if (pm - tau2 > gmax) ngrp = 0;
                            gmax = fmaxf(gmax, pm);
                        }
                        // sweep 2: with the threshold now fixed, record every 4-column group that holds a score
                        // within 2 tau of the maximum: {group maximum, first codeword | which columns hit << 12}.
                        // Predicated stores, no votes, no branches; a full list counts on (ngrp > CG) and the row
                        // falls back to exact scores of all K codewords.
                        if (!(p.dbg_mode & 8)) {
                            const float thr = gmax - tau2;
                            int colw = kbase;
                            auto scan16 = [&](const uint32_t (&r)[16]) {
#pragma unroll
                                for (int j = 0; j < 16; j += 4) {
                                    const float s0 = __uint_as_float(r[j]), s1 = __uint_as_float(r[j + 1]);
                                    const float s2 = __uint_as_float(r[j + 2]), s3 = __uint_as_float(r[j + 3]);
                                    const float m4 = fmax3(fmaxf(s0, s1), s2, s3);
                                    const uint32_t mask = (s0 >= thr ? 1u : 0u) | (s1 >= thr ? 2u : 0u) |
                                                          (s2 >= thr ? 4u : 0u) | (s3 >= thr ? 8u : 0u);
                                    const uint32_t hit = m4 >= thr ? 1u : 0u;
                                    sts64_if(rec_a + (uint32_t)min(ngrp, CG1 - 1) * (BM * 8), __float_as_uint(m4),
                                             (uint32_t)(colw + j) | (mask << 12), hit);
                                    ngrp += (int)hit;
                                }
                                colw += 16;
                            };
                            tmem_ld16_async(taddr, ra);
#pragma unroll 1
                            for (int c0 = 0; c0 < BN; c0 += 32) {
                                tmem_ld16_wait(ra);
                                tmem_ld16_async(taddr + c0 + 16, rb);
                                scan16(ra);
                                tmem_ld16_wait(rb);
                                if (c0 + 32 < BN) tmem_ld16_async(taddr + c0 + 32, ra);
                                scan16(rb);
                            }
                        }
                        if (p.dbg_scores && table == 0) {      // (warp-uniform: tcgen05.ld is collective)
                            float* o = p.dbg_scores + (size_t)(n0 + row) * K + kbase;
                            const float inv = 1.0f / sc[g * BM + row];
                            const bool wr = row < nf;
                            for (int c0 = 0; c0 < BN; c0 += 16) {
                                tmem_ld16_async(taddr + c0, ra);
                                tmem_ld16_wait(ra);
                                if (wr)
                                    for (int j = 0; j < 16; ++j) o[c0 + j] = __uint_as_float(ra[j]) * inv;
                            }
                        }
                        tc_fence_before();
                        mbar_arrive(&tempty_bar[abuf]);
                        e_sweep += (uint32_t)clock64() - tq;
                    }
                    // ---- exchange the sets' maxima, filter both lists against the joint one, publish the job
                    gset_s[set * BM + row] = gmax;
                    {
                        // the slot's previous job (two publications ago) must be complete before its candidate
                        // arrays are reused; help finishing it if it is not
                        const long long tw = clock64();
                        while (!slot_done(slot)) {
                            steal_jobs(slots, lane, 1);
                            if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 11); __trap(); }
                        }
                        e_slot += (uint32_t)(clock64() - tw);
                    }
                    if (publisher) slot->namb = 0;
                    const uint32_t tb1 = (uint32_t)clock64();
                    named_bar_sync(3, 128);
                    const uint32_t tb2 = (uint32_t)clock64();
                    e_bar1 += tb2 - tb1;
                    {
                        const float thr = gset_s[row] - tau2;
                        int keep = 0;
                        for (int i = 0; i < min(ngrp, CG1); ++i) {
                            const uint2 rec = lds64(rec_a + i * (BM * 8));
                            if (__uint_as_float(rec.x) < thr) continue;
                            // (columns that were candidates when the group was recorded: the threshold only
                            //  rises, so this is a superset of the columns within the final range)
                            const int col = (int)(rec.y & 0xfffu);
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (rec.y & (0x1000u << u)) {
                                    // (both candidate lists of the slot, one after the other: the consumers read
                                    //  list 0 up to n[0], then list 1 up to n[1])
                                    if (keep < 2 * CMAXS) slot->cand_idx[keep / CMAXS][keep % CMAXS][row] = col + u;
                                    ++keep;
                                }
                            }
                        }
                        if (ngrp > CG1 || keep > 2 * CMAXS) keep = 3 * CMAXS;   // a list overflowed: exact scores of all K
                        if (emax2 == 0.f || row >= nf || (p.dbg_mode & 32)) {
                            // all-zero codebook: every score ties -> index 0 (also: rows past the end)
                            keep = set == 0 ? 1 : 0;
                            if (!(p.dbg_mode & 32) || keep == 0 || keep > CMAXS)
                                slot->cand_idx[set][0][row] = 0;
                        }
                        // n[0] > CMAXS marks "all K"
                        slot->n[0][row] = keep > 2 * CMAXS ? CMAXS + 1 : min(keep, CMAXS);
                        slot->n[1][row] = keep > 2 * CMAXS ? 0 : max(keep - CMAXS, 0);
                        if (p.dbg_mode & 512) { n_amb += keep > 1; n_full += keep > 2 * CMAXS; }
                    }
                    named_bar_sync(3, 128);                      // all 128 rows of both sets are in shared memory
                    const bool last = s + 1 == S;
                    if (last && set == 0) {
                        // last stage: this thread writes its row's code if it is decided; the undecided rows
                        // are collected for the job (one exact re-score each)
                        const int na = slot->n[0][row], nb = slot->n[1][row];
                        if (defer) {
                            // room for a whole tile of records (the ring only wraps on calls with very many
                            // tiles per CTA; a full ring is drained right here)
                            // (against the PUBLISHED count, which does not move while the set writes this tile's records:
                            //  with the reserved count a warp that arrives late would wait for records its neighbours
                            //  have just reserved and cannot publish before it arrives -- a deadlock when the ring holds
                            //  exactly one tile, D = 64)
                            while (!__all_sync(0xffffffffu, qc[1] + BM - qc[3] <= qctx->cap)) steal_queue(qc, qctx, lane, 1);
                        }
                        if (row < nf) {
                            if (na + nb == 1) {
                                p.codes[(size_t)table * p.N + n0 + row] =
                                    (int64_t)(na ? slot->cand_idx[0][0][row] : slot->cand_idx[1][0][row]);
                            } else if (defer) {
                                const int at = atomicAdd(const_cast<int*>(qc), 1) % qctx->cap;
                                int c[2 * CMAXS];
#pragma unroll
                                for (int i = 0; i < CMAXS; ++i) {
                                    c[i] = i < na ? slot->cand_idx[0][i][row] : 0;
                                    c[CMAXS + i] = i < nb ? slot->cand_idx[1][i][row] : 0;
                                }
                                const long long nfr = n0 + row;
                                int4* rec = reinterpret_cast<int4*>(qctx->ring + (size_t)at * QREC);
                                rec[0] = make_int4((int)(uint32_t)nfr, (int)(nfr >> 32), na, nb | (row << 8) | ((int)(it % nrc) << 16));
                                rec[1] = make_int4(c[0], c[1], c[2], c[3]);
                                rec[2] = make_int4(c[4], c[5], c[6], c[7]);
                                rec[3] = make_int4(c[8], c[9], c[10], c[11]);
                                __threadfence_block();
                            } else {
                                slot->amb[atomicAdd(&slot->namb, 1)] = row;
                            }
                        }
                        named_bar_sync(4, 128);
                    }
                    if (publisher && defer) {
                        // the tile's records are complete: open them to the helpers; the tile buffer goes back
                        // to the loaders (nothing of it is needed any more)
                        __threadfence_block();
                        const int head = qc[0];
                        atomicAdd(const_cast<int*>(qc + 4 + it % nrc), head - qc[1]);   // records pending on the row buffer
                        __threadfence_block();
                        qc[1] = head;
                        mbar_arrive(&free_bar[buf]);
                    } else if (publisher) {
                        Job& j = slot->job;
                        const int items = last ? slot->namb : job_items(Dg);
                        uint8_t* img = Aimg + buf * buf_stride;
                        // single-stage calls: the job (exact re-score from x, codes) does not touch the tile
                        // buffer, so the buffer is handed back to the loaders right here, not when the job ends
                        const bool early_free = S == 1 && !keep_rows;
                        uint64_t* bar = last ? (early_free ? nullptr : &free_bar[buf]) : &upd_bar[par * GMAX + g];
                        if (last && (early_free || items == 0)) mbar_arrive(&free_bar[buf]);
                        j.cbp = p.cb.p[table];
                        j.x = p.x;
                        j.R = keep_rows ? Rbuf + buf * (buf_stride / 4) : nullptr;
                        j.img = img;
                        j.bias_img = img + bias_chunk0 + (size_t)g * A_BYTES;
                        j.sc_g = sc + g * BM;
                        j.nrm_g = nrm + 2 * g * BM;
                        j.bar = bar;
                        j.codes = p.codes + (size_t)table * p.N + n0;
                        j.n0 = n0;
                        if (!last) {
                            const float nb = __uint_as_float(__ldg(table_tail(table + G) + TAIL_BSCALE));
                            j.xs_cap = 32768.f * nb;
                            j.inv_bscale = 1.0f / nb;
                        }
                        j.Dg = Dg; j.D = D; j.g = g; j.nf = nf; j.ste = ste ? 1 : 0; j.last = last ? 1 : 0;
                        j.T = T; j.K = K; j.items = items;
                        j.seq = job_seq + 1;
                        j.stall = (p.dbg_mode & 512) ? p.stall : nullptr;
                        slot->pad[0] = (int)(uint32_t)clock64();
                        slot->state[2] = job_seq;
                        slot->state[1] = 0;
                        slot->state[3] = items;
                        __threadfence_block();
                        slot->state[0] = (job_seq + 1) << 8;     // opens the job: items can be claimed
                    }
                    e_pub += (uint32_t)clock64() - tb2;
                }
            }
          }
        }
        // wait for the open jobs (helping), then release the helper warps
        {
            const long long tw = clock64();
            while (!(slot_done(slots) && slot_done(slots + 1))) {
                steal_jobs(slots, lane, 1);
                if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 12); __trap(); }
            }
            // (deferred queue: everything is published by now -- set 1 passed the last tile's barriers with set 0)
            while (defer && !__all_sync(0xffffffffu, qc[3] >= qc[1])) {
                if (!steal_queue(qc, qctx, lane, 1)) __nanosleep(64);
                if (clock64() - tw > 8000000000LL) { if (p.err) atomicExch(p.err, 13); __trap(); }
            }
        }
        named_bar_sync(3, 128);
        if (publisher) { __threadfence_block(); *all_done = 1; }
        if (p.dbg_mode & 512) {
            if (publisher) {
                atomicAdd(p.stall + 10, (unsigned long long)e_wait);
                atomicAdd(p.stall + 11, (unsigned long long)e_sweep);
                atomicAdd(p.stall + 12, (unsigned long long)e_slot);
                if (S * G > 1) {      // (multi-stage calls: these slots are free of the loaders' sweep counters)
                    atomicAdd(p.stall + 15, (unsigned long long)e_bar1);
                    atomicAdd(p.stall + 16, (unsigned long long)e_pub);
                    atomicAdd(p.stall + 17, (unsigned long long)e_t0);
                }
            }
            atomicAdd(p.stall + 13, (unsigned long long)n_amb);
            atomicAdd(p.stall + 14, (unsigned long long)n_full);
        }
    }

    tc_fence_before();
    __syncthreads();
    // no CTA may retire while a peer can still multicast into its ring or arrive on its barriers
    if (CL > 1) cluster_sync_all();
    if (warp == 9) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int CL, int NST, int NXSLOT>
int launch_p1(const TcParams& p, const CUtensorMap& xmap, cudaStream_t st) {
    auto kern = rvq_search_p1_kernel<CL, NST, NXSLOT>;
    constexpr size_t SMEM_BYTES = smem_bytes(NST, NXSLOT);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(rvq_search_p1)");
    int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
    if (CL == 1) {
        kern<<<grid, threads_for(NXSLOT), SMEM_BYTES, st>>>(p, xmap);
        return check_cuda(cudaGetLastError(), "rvq_search_p1 launch");
    }
    grid = (grid + CL - 1) / CL * CL;        // whole clusters (148 is a multiple of 2 and 4)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(threads_for(NXSLOT));
    cfg.dynamicSmemBytes = SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return check_cuda(cudaLaunchKernelEx(&cfg, kern, p, xmap), "rvq_search_p1 cluster launch");
}
template <int CL>
int launch_p1_cl(const TcParams& p, const CUtensorMap& xmap, cudaStream_t st) {
    // x through the shared-memory slots (4-stage ring) or by plain loads (7-stage ring)
    // single-stage, single-group calls with streamed x: the wide layout (eight loader warps, one epilogue set)
    if (p.tiles_per_clip > 0 && p.wide) return launch_p1<CL, 4, 8>(p, xmap, st);
    return p.tiles_per_clip > 0 ? launch_p1<CL, 4, 4>(p, xmap, st) : launch_p1<CL, NSTAGE_MAX, 0>(p, xmap, st);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

}  // namespace

bool rvq_search_tc_supported(int S, int G, int K, int D, int flags, const char** why);

int rvq_search_p1(const float* x, const float* const* cb, const void* pack, void* workspace, int S,
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
    // x as a 2-D tensor [B*D channel rows][T frames] for the streamer's TMA boxes (32 rows x 128 frames, zero
    // fill past the end of a clip): needs 16-byte row pitch and base, and clips long enough that tiles which
    // do not straddle clips waste little (the last tile of a clip is partial)
    CUtensorMap xmap;
    memset(&xmap, 0, sizeof(xmap));
    static const int stream_ok = [] { const char* v = getenv("ACQ_P1_STREAM"); return v ? atoi(v) : 1; }();
    static const int wide_ok = [] { const char* v = getenv("ACQ_P1_WIDE"); return v ? atoi(v) : 1; }();
    const bool wide = wide_ok && S * G == 1 && D % 128 == 0;
    p.wide = 0;
    if (stream_ok && (T & 3) == 0 && T >= 512 && ((uintptr_t)x & 15) == 0 && (long long)B * D < (1LL << 31) && encode_tiled()) {
        const cuuint64_t gdim[2] = {(cuuint64_t)T, (cuuint64_t)B * D};
        const cuuint64_t gstr[1] = {(cuuint64_t)T * 4};
        const cuuint32_t box[2] = {(cuuint32_t)BM, (cuuint32_t)(wide ? 16 : XCH)};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = encode_tiled()(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(x), gdim, gstr, box,
                                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS) {
            p.tiles_per_clip = (T + BM - 1) / BM;
            p.num_tiles = B * p.tiles_per_clip;
            p.wide = wide ? 1 : 0;
        }
    }
    p.codes = codes;
    p.dbg_scores = dbg_scores;
    { const char* e = getenv("ACQ_TC_DBG"); p.dbg_mode = e ? (int)strtoul(e, nullptr, 0) : 0; }
    p.err = reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + (size_t)kNumSMs * 2 * NTB * BM * D * sizeof(float));
    p.stall = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(p.err) + 64);
    switch (cluster) {
        case 4: return launch_p1_cl<4>(p, xmap, st);
        case 2: return launch_p1_cl<2>(p, xmap, st);
        default: return launch_p1_cl<1>(p, xmap, st);
    }
}

}  // namespace acq
