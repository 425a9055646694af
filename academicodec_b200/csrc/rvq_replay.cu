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
