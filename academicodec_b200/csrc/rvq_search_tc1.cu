// K1b': single-product tcgen05 residual search with a rigorous filter and an exact re-score.
//
// Same warp-specialised pipeline as rvq_search_tc.cu (loaders one tile ahead -> TMA ring ->
// tcgen05.mma into two TMEM accumulators -> epilogue), but only ONE fp16 product per channel
// chunk (hi.hi) instead of three, i.e. a third of the tensor work and half of the operand
// traffic.  One fp16 pass is not index-exact, so the epilogue does not take the argmax of the
// approximate scores; it uses them as a *filter* with a proven error bound:
//
//   operands are scaled into [1024, 2048) by powers of two and rounded to fp16 (RN): every
//   element carries a relative error <= 2^-11, hence for every codeword
//       |s_approx - s_exact| <= tau,   tau = 2^-10 (1 + 2^-5) * ||x~|| * max_k ||e~_k||
//   (Cauchy-Schwarz on the dropped cross terms; the 2^-5 slack covers the fp32 accumulation in the
//   tensor core, the fp32 bias fma, fp16 subnormals and the rounding of the norms themselves).
//   The true best codeword therefore satisfies s_approx >= max_k s_approx - 2 tau.
//
// Epilogue, per 256-codeword pass: sweep 1 reads the accumulator for the pass maximum, sweep 2
// re-reads it and records every codeword within 2 tau of the running maximum (typically 1-2 per
// frame).  After the last pass, frames with a single survivor are done; the others are re-scored
// exactly -- float64 dot products of the fp32 residual row against the fp32 codewords, one warp
// per frame -- and the (value, lowest index) argmax of the exact scores is the code.  Codes thus
// equal the float64 argmax; they can differ from the reference only where its own fp32 rounding
// decides a near-tie.
//
// The fp16 images written by the loaders / the residual update contain only the hi part; the fp32
// residual rows are always kept (the re-score needs them).  Pack format and workspace are those of
// the three-product kernel (only the hi images are streamed).
#include "tc_common.cuh"
#include <stdlib.h>

namespace acq {
namespace {

using namespace tc;

constexpr int NSTAGE = ROWB == 128 ? 4 : 8;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;            // 16 + 32 KiB (SWIZZLE_128B) or 8 + 16 KiB
constexpr int NUM_THREADS = 320;
constexpr int CMAX = 8;                                   // candidates kept per frame and stage
constexpr int NBAR = 2 * NSTAGE + 8 + GMAX;
constexpr int BAR_BYTES = NBAR * 8;
constexpr int OFF_TMEM = BAR_BYTES;
constexpr int OFF_SCALE = OFF_TMEM + 16;                  // [2][GMAX][BM] f32 row scales
constexpr int OFF_SQ = OFF_SCALE + 2 * GMAX * BM * 4;     // [2][GMAX][BM] f32 row sum of squares
constexpr int OFF_MAX = OFF_SQ + 2 * GMAX * BM * 4;       // [GMAX][BM] u32 row max bits (loaders)
constexpr int OFF_CIDX = OFF_MAX + GMAX * BM * 4;         // [BM][CMAX] candidate codeword
constexpr int OFF_CSC = OFF_CIDX + BM * CMAX * 4;         // [BM][CMAX] candidate approximate score
constexpr int OFF_HN = OFF_CSC + BM * CMAX * 4;           // [KMAX] scaled norms of the current table
constexpr int CTRL_BYTES = OFF_HN + KMAX * 4;
constexpr size_t SMEM_BYTES = 1024 + (size_t)NSTAGE * STAGE_BYTES + CTRL_BYTES;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// 2 * tau / (||x~|| * max||e~||)
constexpr float TAU2_COEF = 2.0f * 0.0009765625f * (1.0f + 0.03125f);

__device__ __forceinline__ uint4 half8(const float (&a)[8], float xs) {
    return make_uint4(pack_half2(__float2half_rn(a[0] * xs), __float2half_rn(a[1] * xs)),
                      pack_half2(__float2half_rn(a[2] * xs), __float2half_rn(a[3] * xs)),
                      pack_half2(__float2half_rn(a[4] * xs), __float2half_rn(a[5] * xs)),
                      pack_half2(__float2half_rn(a[6] * xs), __float2half_rn(a[7] * xs)));
}

__global__ void __launch_bounds__(NUM_THREADS, 1) rvq_search_tc1_kernel(const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* ctrl = smem + NSTAGE * STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);          // [NSTAGE]
    uint64_t* empty_bar = full_bar + NSTAGE;                         // [NSTAGE]
    uint64_t* tfull_bar = empty_bar + NSTAGE;                        // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                            // [2]
    uint64_t* t0_bar = tempty_bar + 2;                               // [2]
    uint64_t* free_bar = t0_bar + 2;                                 // [2]
    uint64_t* upd_bar = free_bar + 2;                                // [GMAX]
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(ctrl + OFF_TMEM);
    float* scale_s = reinterpret_cast<float*>(ctrl + OFF_SCALE);
    float* sq_s = reinterpret_cast<float*>(ctrl + OFF_SQ);
    uint32_t* rowmax_s = reinterpret_cast<uint32_t*>(ctrl + OFF_MAX);
    int* cand_idx = reinterpret_cast<int*>(ctrl + OFF_CIDX);
    float* cand_sc = reinterpret_cast<float*>(ctrl + OFF_CSC);
    float* hn_s = reinterpret_cast<float*>(ctrl + OFF_HN);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = p.S, G = p.G, K = p.K, D = p.D, Dg = p.Dg, T = p.T;
    const int NP = K / BN, NKC = Dg / BK;
    const bool ste = p.flags & ACQ_STE;
    const size_t tile_elems = (size_t)BM * D;
    uint8_t* Aimg = reinterpret_cast<uint8_t*>(p.scratch) + (size_t)blockIdx.x * 4 * tile_elems * 4;
    float* Rbuf = reinterpret_cast<float*>(Aimg + 2 * tile_elems * 4);
    const size_t img_tile_bytes = tile_elems * 4;

    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 128);
            mbar_init(&t0_bar[i], 128);
            mbar_init(&free_bar[i], 128);
        }
        for (int i = 0; i < GMAX; ++i) mbar_init(&upd_bar[i], 128);
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(tmem_ptr_s, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;

    if (warp < 4) {
        // ================= loaders: scales, ||x||^2, fp16 hi image, fp32 rows ====================
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t buf = it & 1;
            mbar_wait(&free_bar[buf], ((it >> 1) & 1) ^ 1, p.err, 6);
            if ((p.dbg_mode & 1) && it >= 2) { mbar_arrive(&t0_bar[buf]); continue; }
            const long long n0 = (long long)tile * BM;
            uint8_t* img = Aimg + buf * img_tile_bytes;
            float* R = Rbuf + buf * tile_elems;
            float* sc = scale_s + buf * GMAX * BM;
            float* sq = sq_s + buf * GMAX * BM;
            for (int i = tid; i < G * BM; i += 128) {
                rowmax_s[i] = 0u;
                sq[i] = 0.f;
            }
            named_bar_sync(2, 128);
            if ((T & 3) == 0) {
                // 4 consecutive frames per thread (16-byte loads along T), two adjacent channel octets per
                // iteration: 16 loads in flight, every image / residual store is a whole 32-byte sector
                const int rq = tid & 31, w4 = tid >> 5;
                const long long n = n0 + 4 * rq;
                const bool ok = n < p.N;
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
                        const int oct = 2 * pr;
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
                                float m = 0.f, qq = 0.f;
#pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    m = fmaxf(m, fmaxf(fabsf(a[0][i]), fabsf(a[1][i])));
                                    qq = fmaf(a[0][i], a[0][i], fmaf(a[1][i], a[1][i], qq));
                                }
                                atomicMax(&rowmax_s[g * BM + row], __float_as_uint(m));
                                atomicAdd(&sq[g * BM + row], qq);
                            } else {
                                const float xs = sc[g * BM + row];
                                store_chunk_pair(img + (size_t)(oct / CPR) * 2 * A_BYTES, row, oct % CPR,
                                                 half8(a[0], xs), half8(a[1], xs));
                                float* rd = R + (size_t)row * D + oct * 8;
                                stg256(rd, make_uint4(__float_as_uint(a[0][0]), __float_as_uint(a[0][1]), __float_as_uint(a[0][2]), __float_as_uint(a[0][3])),
                                       make_uint4(__float_as_uint(a[0][4]), __float_as_uint(a[0][5]), __float_as_uint(a[0][6]), __float_as_uint(a[0][7])));
                                stg256(rd + 8, make_uint4(__float_as_uint(a[1][0]), __float_as_uint(a[1][1]), __float_as_uint(a[1][2]), __float_as_uint(a[1][3])),
                                       make_uint4(__float_as_uint(a[1][4]), __float_as_uint(a[1][5]), __float_as_uint(a[1][6]), __float_as_uint(a[1][7])));
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
                            float m = 0.f, qq = 0.f;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                m = fmaxf(m, fabsf(a[i]));
                                qq = fmaf(a[i], a[i], qq);
                            }
                            atomicMax(&rowmax_s[g * BM + row], __float_as_uint(m));
                            atomicAdd(&sq[g * BM + row], qq);
                        } else {
                            uint8_t* dst = img + (size_t)(oct / CPR) * 2 * A_BYTES + sw_offset(row, oct % CPR);
                            *reinterpret_cast<uint4*>(dst) = half8(a, sc[g * BM + row]);
                            float* rd = R + (size_t)row * D + oct * 8;
                            *reinterpret_cast<float4*>(rd) = make_float4(a[0], a[1], a[2], a[3]);
                            *reinterpret_cast<float4*>(rd + 4) = make_float4(a[4], a[5], a[6], a[7]);
                        }
                    }
                    if (sweep == 0) {
                        named_bar_sync(2, 128);
                        for (int i = tid; i < G * BM; i += 128) sc[i] = scale_for(__uint_as_float(rowmax_s[i]));
                        named_bar_sync(2, 128);
                    }
                }
            }
            fence_proxy_async_global();
            mbar_arrive(&t0_bar[buf]);
        }
    } else if (warp == 8) {
        // ================= TMA producer =============================================================
        if (lane == 0) {
            uint32_t it = 0, ring_it = 0, upd_it[GMAX];
#pragma unroll
            for (int i = 0; i < GMAX; ++i) upd_it[i] = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const uint32_t buf = it & 1;
                const uint8_t* img = Aimg + buf * img_tile_bytes;
                for (int s = 0; s < S; ++s) {
                    for (int g = 0; g < G; ++g) {
                        const uint8_t* bimg = p.pack + (size_t)(s * G + g) * p.table_stride;
                        for (int pass = 0; pass < NP; ++pass) {
                            for (int kc = 0; kc < NKC; ++kc, ++ring_it) {
                                const int st = ring_it % NSTAGE;
                                mbar_wait(&empty_bar[st], ((ring_it / NSTAGE) & 1) ^ 1, p.err, 2);
                                uint8_t* a_dst = smem + st * STAGE_BYTES;
                                const bool skip_b = p.dbg_mode & 2, skip_a = p.dbg_mode & 4;
                                mbar_arrive_expect_tx(&full_bar[st], (skip_a ? 0 : A_BYTES) + (skip_b ? 0 : B_BYTES));
                                if (!skip_b)
                                    bulk_g2s(a_dst + A_BYTES, bimg + (size_t)(pass * NKC + kc) * 2 * B_BYTES, B_BYTES,
                                             &full_bar[st]);
                                if (pass == 0 && kc == 0) {
                                    if (s == 0) {
                                        mbar_wait(&t0_bar[buf], (it >> 1) & 1, p.err, 7);
                                    } else {
                                        mbar_wait(&upd_bar[g], upd_it[g] & 1, p.err, 8);
                                        ++upd_it[g];
                                    }
                                    fence_proxy_async_global();
                                }
                                if (!skip_a)
                                    bulk_g2s(a_dst, img + (size_t)(g * NKC + kc) * 2 * A_BYTES, A_BYTES, &full_bar[st]);
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 9) {
        // ================= MMA issuer: one product per chunk ========================================
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
                            const uint32_t b_hi = a_hi + A_BYTES;
#pragma unroll
                            for (int kk = 0; kk < BK / UK; ++kk) {
                                const uint32_t ko = kk * UK * 2;
                                umma_f16(d_tmem, make_desc(a_hi + ko), make_desc(b_hi + ko), IDESC, (kc | kk) != 0);
                            }
                            umma_commit(&empty_bar[st]);
                        }
                        umma_commit(&tfull_bar[abuf]);
                    }
                }
            }
        }
    } else {
        // ================= epilogue: filter, exact re-score, residual update ======================
        const int q = warp - 4;
        const int row = q * 32 + lane;
        int* cidx = cand_idx + row * CMAX;
        float* csc = cand_sc + row * CMAX;
        uint32_t it = 0, acc_it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t buf = it & 1;
            const long long n0 = (long long)tile * BM;
            const int nf = (int)min((long long)BM, p.N - n0);
            uint8_t* img = Aimg + buf * img_tile_bytes;
            float* R = Rbuf + buf * tile_elems;
            float* sc = scale_s + buf * GMAX * BM;
            float* sq = sq_s + buf * GMAX * BM;
            mbar_wait(&t0_bar[buf], (it >> 1) & 1, p.err, 9);
            for (int s = 0; s < S; ++s) {
                for (int g = 0; g < G; ++g) {
                    const int table = s * G + g;
                    const uint8_t* rec = p.pack + (size_t)table * p.table_stride;
                    const float* hn = reinterpret_cast<const float*>(rec + p.img_bytes);
                    const float emax2 = __uint_as_float(
                        __ldg(reinterpret_cast<const uint32_t*>(rec + p.img_bytes + p.hn_bytes) + 2));
                    const float xs = sc[g * BM + row];
                    const float nxs = -xs;
                    // 2 tau, rounded up generously: ||x~|| = xs * sqrt(sum x^2)
                    const float tau2 = TAU2_COEF * xs * sqrtf(sq[g * BM + row]) * sqrtf(emax2) * 1.0001f;
                    // stage this table's scaled norms in shared memory (all four epilogue warps)
                    named_bar_sync(3, 128);
                    for (int i = (tid - 128) * 4; i < K; i += 128 * 4)
                        *reinterpret_cast<float4*>(hn_s + i) = __ldg(reinterpret_cast<const float4*>(hn + i));
                    named_bar_sync(3, 128);
                    float gmax = -INFINITY;
                    int ncand = 0;
                    for (int pass = 0; pass < NP; ++pass, ++acc_it) {
                        const uint32_t abuf = acc_it & 1;
                        mbar_wait(&tfull_bar[abuf], (acc_it >> 1) & 1, p.err, 5);
                        tc_fence_after();
                        const uint32_t taddr = tmem_base + abuf * BN + ((uint32_t)(q * 32) << 16);
                        const int kbase = pass * BN;
                        // sweep 1: maximum of this pass
                        float pmax = -INFINITY;
                        for_each_score(taddr, hn_s + kbase, nxs, [&](int, float sv) { pmax = fmaxf(pmax, sv); });
                        gmax = fmaxf(gmax, pmax);
                        const float thr = gmax - tau2;
                        // sweep 2: every codeword within 2 tau of the running maximum is a candidate
                        if (!(p.dbg_mode & 16)) {
                            for_each_score(taddr, hn_s + kbase, nxs, [&](int c, float sv) {
                                if (sv >= thr) {
                                    if (ncand < CMAX) {
                                        cidx[ncand] = kbase + c;
                                        csc[ncand] = sv;
                                    }
                                    ++ncand;
                                }
                            });
                        }
                        if (p.dbg_scores && table == 0) {      // (warp-uniform: tcgen05.ld is collective)
                            float* o = p.dbg_scores + (size_t)(n0 + row) * K + kbase;
                            const float inv = 1.0f / xs;
                            const bool wr = row < nf;
                            for_each_score(taddr, hn_s + kbase, nxs, [&](int c, float sv) { if (wr) o[c] = sv * inv; });
                        }
                        tc_fence_before();
                        mbar_arrive(&tempty_bar[abuf]);
                    }
                    // ---- final filter against the global maximum (candidates stay in ascending order)
                    const bool overflow = ncand > CMAX;
                    int keep = 0;
                    if (!overflow) {
                        const float thr = gmax - tau2;
                        for (int i = 0; i < ncand; ++i) {
                            if (csc[i] >= thr) {
                                cidx[keep] = cidx[i];
                                ++keep;
                            }
                        }
                    }
                    int bidx = keep > 0 ? cidx[0] : 0;
                    // all-zero codebook: every score is exactly 0 -> lowest index, nothing to re-score
                    bool refine = (overflow || keep > 1) && row < nf && emax2 > 0.f && !(p.dbg_mode & 32);
                    if (overflow && emax2 == 0.f) bidx = 0;
                    __syncwarp();
                    // ---- exact re-score in float64, one warp per ambiguous frame ----------------------
                    const float* cbp = p.cb.p[table];
                    unsigned need = __ballot_sync(0xffffffffu, refine);
                    while (need) {
                        const int rr = __ffs(need) - 1;
                        need &= need - 1;
                        const int urow = q * 32 + rr;
                        const int cnt = __shfl_sync(0xffffffffu, keep, rr);
                        const bool full = __shfl_sync(0xffffffffu, (int)overflow, rr) != 0;
                        const float* rrow = R + (size_t)urow * D + g * Dg;
                        float4 rv[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int d = lane * 4 + 128 * j;
                            rv[j] = d < Dg ? *reinterpret_cast<const float4*>(rrow + d) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        double best = -INFINITY;
                        int best_k = 0;
                        const int n_iter = full ? K : cnt;
                        // two candidates per iteration: both codeword rows are in flight together
                        for (int i = 0; i < n_iter; i += 2) {
                            const bool two = i + 1 < n_iter;
                            const int k0 = full ? i : cand_idx[urow * CMAX + i];
                            const int k1 = two ? (full ? i + 1 : cand_idx[urow * CMAX + i + 1]) : k0;
                            const float* e0 = cbp + (size_t)k0 * Dg;
                            const float* e1 = cbp + (size_t)k1 * Dg;
                            float4 ev0[4], ev1[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int d = lane * 4 + 128 * j;
                                const bool in = d < Dg;
                                ev0[j] = in ? __ldg(reinterpret_cast<const float4*>(e0 + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
                                ev1[j] = in ? __ldg(reinterpret_cast<const float4*>(e1 + d)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                            double dot0 = 0.0, nrm0 = 0.0, dot1 = 0.0, nrm1 = 0.0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const double rx = rv[j].x, ry = rv[j].y, rz = rv[j].z, rw = rv[j].w;
                                dot0 = fma(rx, (double)ev0[j].x, dot0); nrm0 = fma((double)ev0[j].x, (double)ev0[j].x, nrm0);
                                dot0 = fma(ry, (double)ev0[j].y, dot0); nrm0 = fma((double)ev0[j].y, (double)ev0[j].y, nrm0);
                                dot0 = fma(rz, (double)ev0[j].z, dot0); nrm0 = fma((double)ev0[j].z, (double)ev0[j].z, nrm0);
                                dot0 = fma(rw, (double)ev0[j].w, dot0); nrm0 = fma((double)ev0[j].w, (double)ev0[j].w, nrm0);
                                dot1 = fma(rx, (double)ev1[j].x, dot1); nrm1 = fma((double)ev1[j].x, (double)ev1[j].x, nrm1);
                                dot1 = fma(ry, (double)ev1[j].y, dot1); nrm1 = fma((double)ev1[j].y, (double)ev1[j].y, nrm1);
                                dot1 = fma(rz, (double)ev1[j].z, dot1); nrm1 = fma((double)ev1[j].z, (double)ev1[j].z, nrm1);
                                dot1 = fma(rw, (double)ev1[j].w, dot1); nrm1 = fma((double)ev1[j].w, (double)ev1[j].w, nrm1);
                            }
                            double s0 = dot0 - 0.5 * nrm0, s1 = dot1 - 0.5 * nrm1;
#pragma unroll
                            for (int off = 16; off >= 1; off >>= 1) {
                                s0 += __shfl_xor_sync(0xffffffffu, s0, off);
                                s1 += __shfl_xor_sync(0xffffffffu, s1, off);
                            }
                            if (s0 > best) {        // ascending k: the lowest index wins exact ties
                                best = s0;
                                best_k = k0;
                            }
                            if (two && s1 > best) {
                                best = s1;
                                best_k = k1;
                            }
                        }
                        if (lane == rr) bidx = best_k;
                    }
                    if (row < nf) p.codes[(size_t)table * p.N + n0 + row] = bidx;
                    if (s + 1 < S) {
                        // r <- r - e[i] (exact fp32, reference order), new scale / norm / fp16 image
                        residual_update<false, true>(q, lane, nf, bidx, cbp, Dg, D, g, R, img, sc + g * BM,
                                                     sq + g * BM, ste);
                        __syncwarp();
                        fence_proxy_async_global();
                        mbar_arrive(&upd_bar[g]);
                    }
                }
            }
            mbar_arrive(&free_bar[buf]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace

bool rvq_search_tc_supported(int S, int G, int K, int D, int flags, const char** why);

int rvq_search_tc1(const float* x, const float* const* cb, const void* pack, void* workspace, int S,
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
    p.hn_bytes = align256((size_t)K * 4);
    p.scratch = static_cast<float*>(workspace);
    p.S = S; p.G = G; p.K = K; p.D = D; p.Dg = Dg; p.T = T; p.flags = flags;
    p.N = (long long)B * T;
    p.num_tiles = (int)((p.N + BM - 1) / BM);
    p.codes = codes;
    p.dbg_scores = dbg_scores;
    { const char* e = getenv("ACQ_TC_DBG"); p.dbg_mode = e ? atoi(e) : 0; }
    p.err = nullptr;   // (the error flag slot of the shared workspace belongs to the three-product kernel)
    cudaError_t e = cudaFuncSetAttribute(rvq_search_tc1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(rvq_search_tc1)");
    const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
    rvq_search_tc1_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(p);
    return check_cuda(cudaGetLastError(), "rvq_search_tc1 launch");
}

}  // namespace acq
