// K3 / K4: EMA k-means codebook update (training).
//
// K3 ema_stats: per stage, cluster counts and per-cluster sums of the residual that stage saw.
//   The reference builds a dense one-hot [N, K] and runs x.t() @ onehot (core_vq.py:210,219);
//   here each frame adds its residual row to its cluster's row with vector reductions in L2
//   (red.global.add.v4.f32), and bumps one counter.  The residual chain is recomputed from x
//   and the codes with the same arithmetic as the forward pass, so stage s sees bit-identical
//   input.  Output is one flat fp32 buffer [S*K*D sums | S*K counts] so that a single NCCL
//   all-reduce(SUM) makes the statistics global before K4.
// K4 ema_apply: cluster_size / embed_avg EMAs, Laplace smoothing, embed = embed_avg / smoothed
//   (core_vq.py:47-52,218-225), identical on every rank.
#include "acq_common.cuh"

namespace acq {
namespace {

constexpr int NT = 256;
constexpr int TM = 32;   // frames per CTA

struct StatsParams {
    const float* x;
    const int64_t* codes;
    PtrTable cb;
    int S, K, D, DP, RS, B, T, flags;
    long long N;
    float* sums;
    float* counts;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b),
                 "f"(c), "f"(d)
                 : "memory");
}

template <bool VEC>
__global__ void __launch_bounds__(NT) ema_stats_kernel(const StatsParams p) {
    extern __shared__ __align__(16) float r_s[];   // [TM][RS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n0 = (long long)blockIdx.x * TM;
    const int nf = (int)min((long long)TM, p.N - n0);
    const int D = p.D, RS = p.RS, T = p.T, K = p.K;
    const bool ste = p.flags & ACQ_STE;
    {
        const int f = tid % TM;
        const long long n = n0 + f;
        const bool ok = f < nf;
        const long long b = ok ? n / T : 0, t = ok ? n % T : 0;
        const float* src = p.x + (b * D) * (long long)T + t;
        for (int d = tid / TM; d < p.DP; d += NT / TM)
            r_s[f * RS + d] = (ok && d < D) ? __ldg(src + (long long)d * T) : 0.f;
    }
    __syncthreads();
    // every warp owns its frames' rows for all stages: no further block-level sync
    for (int f = warp; f < nf; f += NT / 32) {
        float* rrow = r_s + f * RS;
        for (int s = 0; s < p.S; ++s) {
            const long long code = __ldg(p.codes + (size_t)s * p.N + n0 + f);
            if (code < 0 || code >= K) break;   // invalid code: drop the rest of this frame
            const float* erow = p.cb.p[s] + (size_t)code * D;
            float* srow = p.sums + ((size_t)s * K + code) * D;
            if (lane == 0) atomicAdd(p.counts + (size_t)s * K + code, 1.0f);
            for (int d = lane * 4; d < D; d += 128) {
                float r[4], e[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) r[k] = rrow[d + k];
                if (VEC) {
                    red_add_v4(srow + d, r[0], r[1], r[2], r[3]);
                    const float4 e4 = __ldg(reinterpret_cast<const float4*>(erow + d));
                    e[0] = e4.x; e[1] = e4.y; e[2] = e4.z; e[3] = e4.w;
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        e[k] = 0.f;
                        if (d + k < D) {
                            atomicAdd(srow + d + k, r[k]);
                            e[k] = __ldg(erow + d + k);
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float qs = ste ? __fadd_rn(r[k], __fsub_rn(e[k], r[k])) : e[k];
                    rrow[d + k] = __fsub_rn(r[k], qs);
                }
            }
            __syncwarp();
        }
    }
}

struct ApplyParams {
    float* sums;
    float* counts;      // in: counts, out: smoothed cluster sizes
    MutPtrTable embed, embed_avg, cluster_size;
    int S, K, D;
    float decay, alpha, eps, keps;
};

// one CTA per stage: cluster_size EMA, its total, Laplace smoothing
__global__ void __launch_bounds__(1024) ema_cluster_kernel(const ApplyParams p) {
    __shared__ float part[32];
    __shared__ float total_s;
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* cs = p.cluster_size.p[s];
    float* cnt = p.counts + (size_t)s * p.K;
    // A stage no rank used in this step (bandwidth-limited forward: fewer stages than the stack has, the
    // all-reduced buffer always covers the whole stack) has all-zero counts: leave its buffers untouched
    // and mark it for ema_embed_kernel with negative "smoothed sizes" (real ones are > 0).
    float seen = 0.f;
    for (int k = tid; k < p.K; k += blockDim.x) seen += cnt[k];
    if (__syncthreads_or(seen != 0.f) == 0) {
        for (int k = tid; k < p.K; k += blockDim.x) cnt[k] = -1.f;
        return;
    }
    float local = 0.f;
    for (int k = tid; k < p.K; k += blockDim.x) {
        const float v = fmaf(p.alpha, cnt[k], __fmul_rn(cs[k], p.decay));
        cs[k] = v;
        local += v;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
    if (lane == 0) part[warp] = local;
    __syncthreads();
    if (warp == 0) {
        float v = (lane < (int)(blockDim.x >> 5)) ? part[lane] : 0.f;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) total_s = v;
    }
    __syncthreads();
    const float total = total_s;
    const float den = __fadd_rn(total, p.keps);
    for (int k = tid; k < p.K; k += blockDim.x)
        cnt[k] = __fmul_rn(__fdiv_rn(__fadd_rn(cs[k], p.eps), den), total);
}

__global__ void __launch_bounds__(256) ema_embed_kernel(const ApplyParams p) {
    const size_t per_stage = (size_t)p.K * p.D;
    const size_t total = per_stage * p.S;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int s = (int)(i / per_stage);
        const size_t j = i - (size_t)s * per_stage;
        const int k = (int)(j / p.D);
        if (p.counts[(size_t)s * p.K + k] < 0.f) continue;      // stage unused in this step
        float* ea = p.embed_avg.p[s];
        const float v = fmaf(p.alpha, p.sums[i], __fmul_rn(ea[j], p.decay));
        ea[j] = v;
        p.embed.p[s][j] = __fdiv_rn(v, p.counts[(size_t)s * p.K + k]);
    }
}

}  // namespace

int ema_stats(const float* x, const int64_t* codes, const float* const* cb, int S, int K, int D,
              int B, int T, int flags, float* stats, cudaStream_t st) {
    StatsParams p;
    p.x = x; p.codes = codes;
    for (int i = 0; i < S; ++i) p.cb.p[i] = cb[i];
    p.S = S; p.K = K; p.D = D; p.DP = (D + 3) & ~3; p.RS = p.DP + 4; p.B = B; p.T = T;
    p.flags = flags; p.N = (long long)B * T;
    p.sums = stats; p.counts = stats + (size_t)S * K * D;
    if (p.N == 0) return 0;
    bool vec = (D % 4 == 0) && ((uintptr_t)stats % 16 == 0);
    for (int i = 0; i < S && vec; ++i) vec = ((uintptr_t)cb[i] % 16 == 0);
    const size_t smem = (size_t)TM * p.RS * 4;
    if (smem > 227 * 1024) return fail(ACQ_ESHAPE, "ema_stats: D=%d too large", D);
    const unsigned grid = (unsigned)((p.N + TM - 1) / TM);
    auto kern = vec ? ema_stats_kernel<true> : ema_stats_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(ema_stats)");
    kern<<<grid, NT, smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "ema_stats launch");
}

int ema_apply(float* stats, float* const* embed, float* const* embed_avg,
              float* const* cluster_size, int S, int K, int D, double decay, double epsilon,
              cudaStream_t st) {
    ApplyParams p;
    p.sums = stats; p.counts = stats + (size_t)S * K * D;
    for (int i = 0; i < S; ++i) {
        p.embed.p[i] = embed[i];
        p.embed_avg.p[i] = embed_avg[i];
        p.cluster_size.p[i] = cluster_size[i];
    }
    p.S = S; p.K = K; p.D = D;
    // scalar conversions as torch does them: python double -> fp32 operand
    p.decay = (float)decay;
    p.alpha = (float)(1.0 - decay);
    p.eps = (float)epsilon;
    p.keps = (float)((double)K * epsilon);
    ema_cluster_kernel<<<S, 1024, 0, st>>>(p);
    int rc = check_cuda(cudaGetLastError(), "ema_cluster launch");
    if (rc) return rc;
    const size_t total = (size_t)S * K * D;
    size_t blocks = (total + 255) / 256;
    if (blocks > (size_t)kNumSMs * 16) blocks = (size_t)kNumSMs * 16;
    const unsigned grid = (unsigned)blocks;
    ema_embed_kernel<<<grid, 256, 0, st>>>(p);
    return check_cuda(cudaGetLastError(), "ema_embed launch");
}

}  // namespace acq
