// K1c: replay of a residual code sequence -- everything forward() returns besides the codes.
//
// Given x and the codes chosen by the search kernel, walk the residual chain exactly as the
// forward pass does (eval: r -= e[i]; straight-through: q' = r + (e[i] - r), r -= q') and emit any of
//   quantized [B, D, T]   0.0 + q'_0 + q'_1 + ...                      (core_vq.py:329,340)
//   residual  [B, D, T]   after the last stage
//   sqerr     [S] fp64    per-stage sum of squared quantization error  (core_vq.py:310, models.py:476)
//   stats                 EMA cluster sums / counts of every stage     (core_vq.py:210,218-219)
// This is what lets forward() (training, GRVQ) run its search on the tensor-core kernel, which
// writes codes only: search (tcgen05) + one memory-bound replay pass instead of the fused SIMT
// search.  Three kernels, chosen by size and alignment in rvq_replay():
//   rvq_replay_tile_kernel  >= 4096 frames: the decode tile kernel's data flow plus x and the chain
//   rvq_replay_reg_kernel   small batches: one warp per frame, r and q in registers, codes prefetched
//   rvq_replay_kernel       unaligned tables / odd widths: rows in shared memory (below)
#include "acq_common.cuh"
#include <stdlib.h>

namespace acq {
namespace {

constexpr int NT = 256;
constexpr int TM = 32;

struct ReplayParams {
    const float* x;
    const int64_t* codes;
    PtrTable cb;
    int S, G, K, D, Dg, DgP, RS, T, flags;
    long long N;
    float* quantized;
    float* residual;
    double* sqerr;
    float* sums;      // [S*G? no: S][K][D]  (stats are defined for G == 1 only)
    float* counts;    // [S][K]
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b),
                 "f"(c), "f"(d)
                 : "memory");
}

template <bool VEC>
__global__ void __launch_bounds__(NT) rvq_replay_kernel(const ReplayParams p) {
    extern __shared__ __align__(16) float smem[];
    const int RS = p.RS;
    float* r_s = smem;                                   // [TM][RS]
    float* q_s = p.quantized ? r_s + TM * RS : nullptr;  // [TM][RS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n0 = (long long)blockIdx.x * TM;
    const int nf = (int)min((long long)TM, p.N - n0);
    const int D = p.D, Dg = p.Dg, DgP = p.DgP, G = p.G, K = p.K, T = p.T;
    const bool ste = p.flags & ACQ_STE;
    const bool loss_raw = p.flags & ACQ_LOSS_RAW;
    {
        const int f = tid % TM;
        const long long n = n0 + f;
        const bool ok = f < nf;
        const long long b = ok ? n / T : 0, t = ok ? n % T : 0;
        const float* src = p.x + (size_t)(b * D) * T + t;
        for (int slot = tid / TM; slot < G * DgP; slot += NT / TM) {
            const int g = slot / DgP, dl = slot % DgP;
            r_s[f * RS + slot] = (ok && dl < Dg) ? __ldg(src + (size_t)(g * Dg + dl) * T) : 0.f;
        }
    }
    __syncthreads();
    // each warp owns its frames' rows for the whole chain: no block-level sync inside
    for (int f = warp; f < nf; f += NT / 32) {
        for (int s = 0; s < p.S; ++s) {
            float werr = 0.f;
            bool dead = false;
            for (int g = 0; g < G; ++g) {
                const int tab = s * G + g;
                const long long code = __ldg(p.codes + (size_t)tab * p.N + n0 + f);
                if (code < 0 || code >= K) { dead = true; break; }
                const float* erow = p.cb.p[tab] + (size_t)code * Dg;
                float* rrow = r_s + f * RS + g * DgP;
                float* qrow = q_s ? q_s + f * RS + g * DgP : nullptr;
                float* srow = p.sums ? p.sums + ((size_t)s * K + code) * D : nullptr;   // G == 1
                if (p.counts && lane == 0) atomicAdd(p.counts + (size_t)s * K + code, 1.0f);
                for (int d = lane * 4; d < Dg; d += 128) {
                    float e[4], r[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) r[k] = rrow[d + k];
                    if (VEC) {
                        const float4 e4 = __ldg(reinterpret_cast<const float4*>(erow + d));
                        e[0] = e4.x; e[1] = e4.y; e[2] = e4.z; e[3] = e4.w;
                        if (srow) red_add_v4(srow + d, r[0], r[1], r[2], r[3]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            e[k] = (d + k < Dg) ? __ldg(erow + d + k) : 0.f;
                            if (srow && d + k < Dg) atomicAdd(srow + d + k, r[k]);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (d + k < Dg) {
                            const float q = e[k];
                            const float qs = ste ? __fadd_rn(r[k], __fsub_rn(q, r[k])) : q;
                            const float df = loss_raw ? __fsub_rn(q, r[k]) : __fsub_rn(qs, r[k]);
                            werr = fmaf(df, df, werr);
                            rrow[d + k] = __fsub_rn(r[k], qs);
                            if (qrow) qrow[d + k] = __fadd_rn(s == 0 ? 0.f : qrow[d + k], qs);
                        }
                    }
                }
            }
            __syncwarp();
            if (dead) break;     // invalid code: drop the rest of this frame's chain
            if (p.sqerr) {
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) werr += __shfl_xor_sync(0xffffffffu, werr, off);
                if (lane == 0 && werr != 0.f) atomicAdd(p.sqerr + s, (double)werr);
            }
        }
    }
    __syncthreads();
    if (p.quantized || p.residual) {
        const int f = tid % TM;
        if (f < nf) {
            const long long n = n0 + f;
            const long long b = n / T, t = n % T;
            const size_t base = (size_t)(b * D) * T + t;
            for (int d = tid / TM; d < D; d += NT / TM) {
                const int slot = (d / Dg) * DgP + (d % Dg);
                if (p.quantized) p.quantized[base + (size_t)d * T] = q_s[f * RS + slot];
                if (p.residual) p.residual[base + (size_t)d * T] = r_s[f * RS + slot];
            }
        }
    }
}

// ------------------------------------------------------------------ register-resident variant
// The kernel above keeps the residual and the running sum of a frame in shared memory and walks the
// chain with one dependent L2 gather after another: the codes of a frame are known up front, yet
// every stage paid a code load, then a codeword gather, then the update (about 3500 cycles per stage
// with 8 warps per SM), and every frame hit the same S fp64 addresses with its loss atomics.  For
// the common case (16-byte aligned tables, D <= 1024, Dg % 4 == 0) a lane keeps its 4*NJ channels of
// r and q in registers, all S*G codes of the frame are fetched with one coalesced load and handed
// out by shuffles, the codeword rows of stage s+1 are in flight while stage s is applied, the loss
// is reduced per CTA in shared memory first, and the tile is 16 frames at D > 256 so that three
// CTAs fit an SM.  Measured on cfg5 (D=512, n_q=12, 64 000 frames): 1.97 ms -> see DESIGN.md K1c.
template <int NJ>
__global__ void __launch_bounds__(NT) rvq_replay_reg_kernel(const ReplayParams p, const int TMf) {
    extern __shared__ __align__(16) float smem[];
    const int RS = p.RS;                                  // D + 4 (G * DgP == D here)
    float* x_s = smem;                                    // [TMf][RS]  x, then the final residual
    float* q_s = p.quantized ? x_s + TMf * RS : nullptr;  // [TMf][RS]
    double* sq_s = reinterpret_cast<double*>(smem + (p.quantized ? 2 : 1) * TMf * RS);   // [S]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n0 = (long long)blockIdx.x * TMf;
    const int nf = (int)min((long long)TMf, p.N - n0);
    const int D = p.D, Dg = p.Dg, G = p.G, K = p.K, T = p.T, S = p.S;
    const int ntab = S * G;
    const bool ste = p.flags & ACQ_STE;
    const bool loss_raw = p.flags & ACQ_LOSS_RAW;
    for (int i = tid; i < S; i += NT) sq_s[i] = 0.0;
    {
        const int f = tid % TMf;
        const long long n = n0 + f;
        const bool ok = f < nf;
        const long long b = ok ? n / T : 0, t = ok ? n % T : 0;
        const float* src = p.x + (size_t)(b * D) * T + t;
        for (int d = tid / TMf; d < D; d += NT / TMf) x_s[f * RS + d] = ok ? __ldg(src + (size_t)d * T) : 0.f;
    }
    __syncthreads();
    // channel blocks of this lane: d_j = lane*4 + 128*j; group and in-group offset are fixed per block
    int gj[NJ], oj[NJ];
    bool inj[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int d = lane * 4 + 128 * j;
        inj[j] = d < D;
        gj[j] = inj[j] ? d / Dg : 0;
        oj[j] = d - gj[j] * Dg;
    }
    for (int f = warp; f < nf; f += NT / 32) {
        const long long n = n0 + f;
        // all codes of the frame: lane l holds tables l and l + 32
        long long c0 = lane < ntab ? __ldg(p.codes + (size_t)lane * p.N + n) : 0;
        long long c1 = lane + 32 < ntab ? __ldg(p.codes + (size_t)(lane + 32) * p.N + n) : 0;
        const unsigned bad0 = __ballot_sync(0xffffffffu, c0 < 0 || c0 >= K);
        const unsigned bad1 = __ballot_sync(0xffffffffu, c1 < 0 || c1 >= K);
        // an invalid code drops the rest of the chain from its stage on
        const int first_bad = bad0 ? __ffs(bad0) - 1 : (bad1 ? 32 + __ffs(bad1) - 1 : ntab);
        const int S_eff = min(S, first_bad / G);
        const int ci0 = (int)c0, ci1 = (int)c1;
        float4 r[NJ], q[NJ], e_nxt[NJ];
        int code_nxt[NJ];
        auto fetch = [&](int st) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int tab = st * G + gj[j];
                const int a = __shfl_sync(0xffffffffu, ci0, tab & 31);
                const int b2 = __shfl_sync(0xffffffffu, ci1, tab & 31);
                code_nxt[j] = tab < 32 ? a : b2;
                e_nxt[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (inj[j] && st < S_eff)
                    e_nxt[j] = __ldg(reinterpret_cast<const float4*>(p.cb.p[tab] + (size_t)code_nxt[j] * Dg + oj[j]));
            }
        };
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            r[j] = inj[j] ? *reinterpret_cast<const float4*>(x_s + f * RS + lane * 4 + 128 * j)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
            q[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        fetch(0);
        for (int st = 0; st < S_eff; ++st) {
            float4 e[NJ];
            int code[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) { e[j] = e_nxt[j]; code[j] = code_nxt[j]; }
            fetch(st + 1);                                   // next stage's rows are in flight during this one
            float werr = 0.f;
            if (p.sums) {                                    // G == 1: EMA statistics use the pre-update residual
                if (lane == 0) atomicAdd(p.counts + (size_t)st * K + code[0], 1.0f);
#pragma unroll
                for (int j = 0; j < NJ; ++j)
                    if (inj[j])
                        red_add_v4(p.sums + ((size_t)st * K + code[j]) * D + lane * 4 + 128 * j, r[j].x, r[j].y,
                                   r[j].z, r[j].w);
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                float* rr = reinterpret_cast<float*>(&r[j]);
                float* qq = reinterpret_cast<float*>(&q[j]);
                const float* ee = reinterpret_cast<const float*>(&e[j]);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (inj[j]) {
                        const float qs = ste ? __fadd_rn(rr[k], __fsub_rn(ee[k], rr[k])) : ee[k];
                        const float df = loss_raw ? __fsub_rn(ee[k], rr[k]) : __fsub_rn(qs, rr[k]);
                        werr = fmaf(df, df, werr);
                        rr[k] = __fsub_rn(rr[k], qs);
                        qq[k] = __fadd_rn(qq[k], qs);
                    }
                }
            }
            if (p.sqerr) {
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) werr += __shfl_xor_sync(0xffffffffu, werr, off);
                if (lane == 0 && werr != 0.f) atomicAdd(sq_s + st, (double)werr);
            }
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            if (inj[j]) {
                if (q_s) *reinterpret_cast<float4*>(q_s + f * RS + lane * 4 + 128 * j) = q[j];
                if (p.residual) *reinterpret_cast<float4*>(x_s + f * RS + lane * 4 + 128 * j) = r[j];
            }
        }
    }
    __syncthreads();
    if (p.sqerr)
        for (int i = tid; i < S; i += NT)
            if (sq_s[i] != 0.0) atomicAdd(p.sqerr + i, sq_s[i]);
    if (p.quantized || p.residual) {
        const int f = tid % TMf;
        if (f < nf) {
            const long long n = n0 + f;
            const long long b = n / T, t = n % T;
            const size_t base = (size_t)(b * D) * T + t;
            for (int d = tid / TMf; d < D; d += NT / TMf) {
                if (p.quantized) p.quantized[base + (size_t)d * T] = q_s[f * RS + d];
                if (p.residual) p.residual[base + (size_t)d * T] = x_s[f * RS + d];
            }
        }
    }
}

// ------------------------------------------------------------------ tile variant (T % 4 == 0)
// Same data flow as the decode tile kernel (vq_decode.cu): a CTA owns 128 channels x 64 frames, a
// warp step is 4 frames with a lane on 4 consecutive channels, codeword rows are gathered with
// 16-byte loads, and the result leaves through an XOR-swizzled shared-memory transpose as 16-byte
// streaming stores with frames contiguous.  On top of that the lane loads its 4 channels x 4 frames
// of x (one 16-byte load per channel, frames contiguous -- no transposing load), walks the residual
// chain in registers, adds the stage losses (warp shuffle -> per-CTA fp64 in shared memory -> one
// atomic per stage and CTA) and the EMA statistics (red.global.add.v4.f32).  The per-frame kernels
// above pay a transposing load and store per 16-32 frame tile and reach 0.8 TB/s on long batches;
// this one stays near the decode kernel's rate.
constexpr int RDT = 128, RFT = 64;

__device__ __forceinline__ int rtile_off(int d, int f) {
    return d * RFT + ((((f >> 2) ^ ((d >> 2) & 7)) << 2) | (f & 3));
}

__global__ void __launch_bounds__(NT, 3) rvq_replay_tile_kernel(const ReplayParams p) {
    extern __shared__ __align__(16) float smem[];
    float* tile = smem;                                                   // [RDT][RFT] swizzled
    int* code_s = reinterpret_cast<int*>(smem + RDT * RFT);               // [S*G][RFT], -1 = invalid
    double* sq_s = reinterpret_cast<double*>(code_s + p.S * p.G * RFT);   // [S]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n0 = (long long)blockIdx.x * RFT;
    const int d0 = blockIdx.y * RDT;
    const int nf = (int)min((long long)RFT, p.N - n0);
    const int nd = min(RDT, p.D - d0);
    const int S = p.S, G = p.G, K = p.K, D = p.D, Dg = p.Dg, T = p.T;
    const int ntab = S * G;
    const bool ste = p.flags & ACQ_STE;
    const bool loss_raw = p.flags & ACQ_LOSS_RAW;
    for (int i = tid; i < S; i += NT) sq_s[i] = 0.0;
    for (int u = tid; u < ntab * RFT; u += NT) {
        const int tab = u / RFT, f = u % RFT;
        int v = -1;
        if (f < nf) {
            const long long c = __ldg(p.codes + (size_t)tab * p.N + n0 + f);
            if (c >= 0 && c < K) v = (int)c;
        }
        code_s[u] = v;
    }
    const bool vt = (T & 3) == 0;             // frame quads stay inside a clip and are 16-byte aligned
    __syncthreads();

    const int dl = lane * 4;
    const bool act = dl < nd;                 // D % 4 == 0: a channel quad is all in or all out
    const int d = d0 + dl;
    const int g = act ? d / Dg : 0;
    const int dg = d - g * Dg;
    const bool count_here = p.counts && blockIdx.y == 0 && lane == 0;
    for (int pass = 0; pass < (p.residual ? 2 : 1); ++pass) {
        // pass 0 parks the quantized sum in the tile, pass 1 (only when the final residual is asked
        // for) repeats the chain and parks the residual; statistics and losses belong to pass 0
        if (!vt) {
            // any T: x enters through the same swizzled tile the results leave by (coalesced along
            // frames); a warp later overwrites exactly the cells it read, so no further barrier
            const int f = tid % RFT;
            if (f < nf) {
                const long long n = n0 + f;
                const long long b = n / T, t = n - b * T;
                const float* src = p.x + ((size_t)b * D + d0) * T + t;
#pragma unroll 8
                for (int dr = tid / RFT; dr < nd; dr += NT / RFT) tile[rtile_off(dr, f)] = __ldg(src + (size_t)dr * T);
            }
            __syncthreads();
        }
        for (int f0 = warp * 4; f0 < nf; f0 += (NT / 32) * 4) {
            float4 r[4], q[4];
            bool alive[4];
            {
                float4 xc[4];                          // channel c, frames f0..f0+3
                if (vt) {
                    const long long n = n0 + f0;       // the quad is inside one clip
                    const long long b = n / T, t = n - b * T;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        xc[c] = act ? __ldg(reinterpret_cast<const float4*>(p.x + ((size_t)b * D + d + c) * T + t))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        xc[c] = act ? *reinterpret_cast<const float4*>(tile + rtile_off(dl + c, f0))
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                r[0] = make_float4(xc[0].x, xc[1].x, xc[2].x, xc[3].x);
                r[1] = make_float4(xc[0].y, xc[1].y, xc[2].y, xc[3].y);
                r[2] = make_float4(xc[0].z, xc[1].z, xc[2].z, xc[3].z);
                r[3] = make_float4(xc[0].w, xc[1].w, xc[2].w, xc[3].w);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { q[u] = make_float4(0.f, 0.f, 0.f, 0.f); alive[u] = f0 + u < nf; }
            for (int st = 0; st < S; ++st) {
                const int tab = st * G + g;
                float4 e[4];
                int code[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    code[u] = alive[u] ? code_s[tab * RFT + f0 + u] : -1;
                    // an invalid code in any group of this stage ends the frame's chain
                    for (int gg = 0; gg < G && alive[u]; ++gg)
                        if (code_s[(st * G + gg) * RFT + f0 + u] < 0) alive[u] = false;
                    e[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (alive[u] && act)
                        e[u] = __ldg(reinterpret_cast<const float4*>(p.cb.p[tab] + (size_t)code[u] * Dg + dg));
                }
                float werr = 0.f;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (!alive[u]) continue;
                    if (pass == 0 && p.sums) {                 // G == 1: statistics use the pre-update residual
                        if (act) red_add_v4(p.sums + ((size_t)st * K + code[u]) * D + d, r[u].x, r[u].y, r[u].z, r[u].w);
                        if (count_here) atomicAdd(p.counts + (size_t)st * K + code[u], 1.0f);
                    }
                    float* rr = reinterpret_cast<float*>(&r[u]);
                    float* qq = reinterpret_cast<float*>(&q[u]);
                    const float* ee = reinterpret_cast<const float*>(&e[u]);
                    if (act) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float qs = ste ? __fadd_rn(rr[k], __fsub_rn(ee[k], rr[k])) : ee[k];
                            const float df = loss_raw ? __fsub_rn(ee[k], rr[k]) : __fsub_rn(qs, rr[k]);
                            werr = fmaf(df, df, werr);
                            rr[k] = __fsub_rn(rr[k], qs);
                            qq[k] = __fadd_rn(qq[k], qs);
                        }
                    }
                }
                if (pass == 0 && p.sqerr) {
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) werr += __shfl_xor_sync(0xffffffffu, werr, off);
                    if (lane == 0 && werr != 0.f) atomicAdd(sq_s + st, (double)werr);
                }
            }
            if (act) {
                const float4* o = pass == 0 ? q : r;
                float* base = tile + rtile_off(dl, f0);
                *reinterpret_cast<float4*>(base) = make_float4(o[0].x, o[1].x, o[2].x, o[3].x);
                *reinterpret_cast<float4*>(base + RFT) = make_float4(o[0].y, o[1].y, o[2].y, o[3].y);
                *reinterpret_cast<float4*>(base + 2 * RFT) = make_float4(o[0].z, o[1].z, o[2].z, o[3].z);
                *reinterpret_cast<float4*>(base + 3 * RFT) = make_float4(o[0].w, o[1].w, o[2].w, o[3].w);
            }
        }
        __syncthreads();
        float* outp = pass == 0 ? p.quantized : p.residual;
        if (outp && !vt) {
            const int f = tid % RFT;
            if (f < nf) {
                const long long n = n0 + f;
                const long long b = n / T, t = n - b * T;
                float* dst = outp + ((size_t)b * D + d0) * T + t;
#pragma unroll 8
                for (int dr = tid / RFT; dr < nd; dr += NT / RFT) __stcs(dst + (size_t)dr * T, tile[rtile_off(dr, f)]);
            }
        } else if (outp) {
            const int f4 = (tid & 15) * 4;
            if (f4 < nf) {
                const long long n = n0 + f4;
                const long long b = n / T, t = n - b * T;
                float* dst = outp + ((size_t)b * D + d0) * T + t;
#pragma unroll 4
                for (int dr = tid >> 4; dr < nd; dr += NT / 16)
                    __stcs(reinterpret_cast<float4*>(dst + (size_t)dr * T),
                           *reinterpret_cast<const float4*>(tile + rtile_off(dr, f4)));
            }
        }
        __syncthreads();
    }
    if (p.sqerr)
        for (int i = tid; i < S; i += NT)
            if (sq_s[i] != 0.0) atomicAdd(p.sqerr + i, sq_s[i]);
}

int launch_tile(const ReplayParams& p, cudaStream_t st) {
    const size_t smem = (size_t)RDT * RFT * 4 + (size_t)p.S * p.G * RFT * 4 + (size_t)p.S * 8 + 8;
    cudaError_t e = cudaFuncSetAttribute(rvq_replay_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(rvq_replay_tile)");
    dim3 grid((unsigned)((p.N + RFT - 1) / RFT), (unsigned)((p.D + RDT - 1) / RDT));
    rvq_replay_tile_kernel<<<grid, NT, smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "rvq_replay_tile launch");
}

template <int NJ>
int launch_reg(ReplayParams p, cudaStream_t st) {
    const int TMf = p.D > 256 ? 16 : 32;
    p.RS = p.D + 4;
    const size_t smem = (size_t)(p.quantized ? 2 : 1) * TMf * p.RS * 4 + (size_t)p.S * 8;
    auto kern = rvq_replay_reg_kernel<NJ>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(rvq_replay_reg)");
    const unsigned grid = (unsigned)((p.N + TMf - 1) / TMf);
    kern<<<grid, NT, smem, st>>>(p, TMf);
    return check_cuda(cudaGetLastError(), "rvq_replay_reg launch");
}

}  // namespace

int rvq_replay(const float* x, const int64_t* codes, const float* const* cb, int S, int G, int K, int D,
               int B, int T, int flags, float* quantized, float* residual, double* sqerr, float* stats,
               cudaStream_t st) {
    ReplayParams p;
    p.x = x; p.codes = codes;
    for (int i = 0; i < S * G; ++i) p.cb.p[i] = cb[i];
    p.S = S; p.G = G; p.K = K; p.D = D; p.Dg = D / G; p.DgP = (p.Dg + 3) & ~3;
    p.RS = G * p.DgP + 4; p.T = T; p.flags = flags; p.N = (long long)B * T;
    p.quantized = quantized; p.residual = residual; p.sqerr = sqerr;
    p.sums = stats; p.counts = stats ? stats + (size_t)S * K * D : nullptr;
    if (p.N == 0) return 0;
    if (stats && G != 1) return fail(ACQ_EINVAL, "rvq_replay: EMA statistics are defined for G == 1");
    bool vec = (p.Dg % 4 == 0) && (!stats || (uintptr_t)stats % 16 == 0);
    for (int i = 0; i < S * G && vec; ++i) vec = ((uintptr_t)cb[i] % 16 == 0);
    static const int mode = [] { const char* v = getenv("ACQ_REPLAY_KERNEL"); return v ? atoi(v) : 0; }();
    // ACQ_REPLAY_KERNEL: 1 forces the shared-memory kernel, 2 the register-resident one (A/B measurements)
    if (vec && p.N >= 4096 && mode != 1 && mode != 2) return launch_tile(p, st);
    if (vec && D <= 1024 && S * G <= 64 && mode != 1) {
        if (D <= 128) return launch_reg<1>(p, st);
        if (D <= 256) return launch_reg<2>(p, st);
        if (D <= 512) return launch_reg<4>(p, st);
        return launch_reg<8>(p, st);
    }
    const size_t smem = (size_t)(quantized ? 2 : 1) * TM * p.RS * 4;
    if (smem > 227 * 1024) return fail(ACQ_ESHAPE, "rvq_replay: D=%d too large", D);
    const unsigned grid = (unsigned)((p.N + TM - 1) / TM);
    auto kern = vec ? rvq_replay_kernel<true> : rvq_replay_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(rvq_replay)");
    kern<<<grid, NT, smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "rvq_replay launch");
}

}  // namespace acq
