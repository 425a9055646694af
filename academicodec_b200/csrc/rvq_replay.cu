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
// search.  One CTA = 32 frames staged in shared memory, one warp per frame, lanes across
// channels (coalesced 16-byte codeword gathers from L2).
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
    if (vec && D <= 1024 && S * G <= 64 && mode != 1) {     // ACQ_REPLAY_KERNEL=1 forces the shared-memory kernel
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
