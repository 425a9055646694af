// Backward of the group-residual VQ forward (HiFi-Codec `Quantizer.forward`, training).
//
// The reference builds the loss of every (stage, group) as
//     lam_cb * mean((z_q - x.detach())^2) + lam_commit * mean((z_q.detach() - x)^2)
// (hificodec/models.py:474-480) and lets autograd differentiate the embedding lookups and the
// straight-through sum.  The gradients that reach the caller are
//     d xin   = d quantized_out                                     (straight-through identity)
//             + (-2 lam_commit / numel) * dL_0 * (z_q0 - xin)       (stage 0 only: later stages see a
//                                                                    residual that is detached by the
//                                                                    straight-through arithmetic)
//     d E_sg  = scatter-add over frames of (2 lam_cb / numel) * dL_s * (z_q - r_s)[group g] into row code
// The round-1 host side computed them with ~12 eager kernels per step (a gather, a subtraction and an
// index_add_ per stage and group, transposed copies of x).  Here ONE kernel recomputes the residual
// chain from x and the codes (same arithmetic as the forward pass) and emits both: a CTA carries a
// tile of 32 frames through shared memory (coalesced along frames on the way in and out, lanes across
// channels in between), codebook gradients go out as coalesced fp32 reductions in L2.
#include "acq_common.cuh"

namespace acq {
namespace {

constexpr int NT = 256;
constexpr int TM = 32;     // frames per CTA
constexpr int DMAX = 768;   // two [32][D+1] fp32 tiles in shared memory

struct GbParams {
    const float* x;
    const int64_t* codes;
    PtrTable cb;
    MutPtrTable gw;          // per table: gradient [K, Dg] (caller zeroes) or nullptr
    const float* g_q;        // [B, D, T] gradient of the quantized sum, or nullptr
    const float* g_losses;   // [S] gradient of the per-stage losses, or nullptr
    float* grad_x;           // [B, D, T] or nullptr
    int S, G, K, D, Dg, T, RS;
    long long N;
    float c_cb, c_commit;    // 2 lam_cb / numel, -2 lam_commit / numel (rounded to fp32 as torch does)
};

__global__ void __launch_bounds__(NT) grvq_backward_kernel(const GbParams p) {
    extern __shared__ __align__(16) float r_s[];   // [TM][RS] residual rows | [TM][RS] commitment part of d xin
    float* gx_s = r_s + TM * p.RS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n0 = (long long)blockIdx.x * TM;
    const int nf = (int)min((long long)TM, p.N - n0);
    const int D = p.D, Dg = p.Dg, RS = p.RS, T = p.T;
    // frames are contiguous in x: lane = frame on the way in (and out)
    const int f_io = tid % TM;
    const long long n_io = n0 + f_io;
    const bool ok_io = f_io < nf;
    const long long b_io = ok_io ? n_io / T : 0, t_io = ok_io ? n_io % T : 0;
    {
        const float* src = p.x + (b_io * D) * (long long)T + t_io;
        for (int d = tid / TM; d < D; d += NT / TM) r_s[f_io * RS + d] = ok_io ? __ldg(src + (long long)d * T) : 0.f;
    }
    __syncthreads();
    const float gl0 = p.g_losses ? __ldg(p.g_losses) : 0.f;
    const float cx = p.c_commit * gl0;
    for (int f = warp; f < nf; f += NT / 32) {
        float* rrow = r_s + f * RS;
        float* gxrow = gx_s + f * RS;
        for (int s = 0; s < p.S; ++s) {
            const float cw = p.g_losses ? p.c_cb * __ldg(p.g_losses + s) : 0.f;
            for (int g = 0; g < p.G; ++g) {
                const int table = s * p.G + g;
                const long long code = __ldg(p.codes + (size_t)table * p.N + n0 + f);
                if (code < 0 || code >= p.K) {                      // (invalid code: no contribution)
                    if (s == 0) for (int d = lane; d < Dg; d += 32) gxrow[g * Dg + d] = 0.f;
                    continue;
                }
                const float* erow = p.cb.p[table] + (size_t)code * Dg;
                float* gwrow = p.gw.p[table] ? p.gw.p[table] + (size_t)code * Dg : nullptr;
#pragma unroll 4
                for (int d = lane; d < Dg; d += 32) {
                    const int c = g * Dg + d;
                    const float r = rrow[c], e = __ldg(erow + d);
                    const float diff = __fsub_rn(e, r);                  // z_q - r
                    if (gwrow && cw != 0.f) atomicAdd(gwrow + d, cw * diff);
                    if (s == 0) gxrow[c] = cx * diff;
                    // straight-through residual of the forward pass: r - (r + (z_q - r))  (models.py:483-485)
                    rrow[c] = __fsub_rn(r, __fadd_rn(r, diff));
                }
            }
        }
    }
    if (!p.grad_x) return;
    __syncthreads();
    {
        const long long off = (b_io * D) * (long long)T + t_io;
        if (ok_io) {
            for (int d = tid / TM; d < D; d += NT / TM) {
                const float v = gx_s[f_io * RS + d];
                p.grad_x[off + (long long)d * T] = p.g_q ? __ldg(p.g_q + off + (long long)d * T) + v : v;
            }
        }
    }
}

// D = 32 * NQ in {128, 256, 512}: the commitment part of d xin stays in registers (lane = channel c = 32 q + lane),
// one shared-memory tile instead of two, 80 registers -- three CTAs per SM instead of one (the first version ran 8 warps per SM
// and was latency-bound at a tenth of the HBM roofline: ncu 1.95 ms for 4096 x 50 frames, issue slots 12 % busy).
template <int NQ>
__global__ void __launch_bounds__(NT, 3) grvq_backward_reg_kernel(const GbParams p) {
    extern __shared__ __align__(16) float r_s[];   // [TM][RS]: residual rows, then the rows of the commitment gradient
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n0 = (long long)blockIdx.x * TM;
    const int nf = (int)min((long long)TM, p.N - n0);
    constexpr int D = 32 * NQ;
    const int Dg = p.Dg, RS = p.RS, T = p.T, G = p.G;
    const int qpg = Dg / 32;                        // lanes' channel slots per group
    const int f_io = tid % TM;
    const long long n_io = n0 + f_io;
    const bool ok_io = f_io < nf;
    const long long b_io = ok_io ? n_io / T : 0, t_io = ok_io ? n_io % T : 0;
    {
        const float* src = p.x + (b_io * D) * (long long)T + t_io;
        // (D / 8 loads per thread, 16 of them in flight: with four the tile took sixteen memory round trips on the
        //  way in and sixteen on the way out, most of the kernel's time)
#pragma unroll 16
        for (int d = tid / TM; d < D; d += NT / TM) r_s[f_io * RS + d] = ok_io ? __ldg(src + (long long)d * T) : 0.f;
    }
    __syncthreads();
    const float gl0 = p.g_losses ? __ldg(p.g_losses) : 0.f;
    const float cx = p.c_commit * gl0;
    for (int f = warp; f < nf; f += NT / 32) {
        float* rrow = r_s + f * RS;
        // (the residual lives in registers, so the shared-memory row is free for the commitment gradient right away)
        float r[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) { r[q] = rrow[q * 32 + lane]; rrow[q * 32 + lane] = 0.f; }
        for (int s = 0; s < p.S; ++s) {
            const float cw = p.g_losses ? p.c_cb * __ldg(p.g_losses + s) : 0.f;
            long long code[4];
#pragma unroll
            for (int g = 0; g < 4; ++g)
                code[g] = g < G ? __ldg(p.codes + (size_t)(s * G + g) * p.N + n0 + f) : -1;
            float e[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {        // all codeword gathers of the stage in flight together
                const int g = (q >= qpg) + (q >= 2 * qpg) + (q >= 3 * qpg);
                const long long cd = g == 0 ? code[0] : (g == 1 ? code[1] : (g == 2 ? code[2] : code[3]));
                const bool ok = cd >= 0 && cd < p.K;
                e[q] = ok ? __ldg(p.cb.p[s * G + g] + (size_t)cd * Dg + (q - g * qpg) * 32 + lane) : r[q];
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int g = (q >= qpg) + (q >= 2 * qpg) + (q >= 3 * qpg);
                const long long cd = g == 0 ? code[0] : (g == 1 ? code[1] : (g == 2 ? code[2] : code[3]));
                if (cd < 0 || cd >= p.K) continue;                  // (invalid code: no contribution)
                const float diff = __fsub_rn(e[q], r[q]);            // z_q - r
                float* gw = p.gw.p[s * G + g];
                if (gw && cw != 0.f) atomicAdd(gw + (size_t)cd * Dg + (q - g * qpg) * 32 + lane, cw * diff);
                if (s == 0) rrow[q * 32 + lane] = cx * diff;
                r[q] = __fsub_rn(r[q], __fadd_rn(r[q], diff));       // straight-through residual (models.py:483-485)
            }
        }
    }
    if (!p.grad_x) return;
    __syncthreads();
    if (ok_io) {
        const long long off = (b_io * D) * (long long)T + t_io;
        if (p.g_q) {
#pragma unroll 16
            for (int d = tid / TM; d < D; d += NT / TM)
                p.grad_x[off + (long long)d * T] = __ldg(p.g_q + off + (long long)d * T) + r_s[f_io * RS + d];
        } else {
#pragma unroll 16
            for (int d = tid / TM; d < D; d += NT / TM) p.grad_x[off + (long long)d * T] = r_s[f_io * RS + d];
        }
    }
}

template <int NQ>
int launch_reg(const GbParams& p, cudaStream_t st) {
    const size_t smem = (size_t)TM * p.RS * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(grvq_backward_reg_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(grvq_backward_reg)");
        attr_set = true;
    }
    const long long grid = (p.N + TM - 1) / TM;
    grvq_backward_reg_kernel<NQ><<<(unsigned)grid, NT, smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "grvq_backward launch");
}

}  // namespace

int grvq_backward(const float* x, const int64_t* codes, const float* const* cb, int S, int G, int K, int D, int B,
                  int T, const float* g_q, const float* g_losses, double lam_cb, double lam_commit, float* grad_x,
                  float* const* grad_w, cudaStream_t st) {
    if (D % 32 != 0 || D > DMAX || D % G != 0) return fail(ACQ_ESHAPE, "grvq backward: D %% 32 != 0 or D > %d", DMAX);
    GbParams p;
    p.x = x; p.codes = codes;
    for (int i = 0; i < S * G; ++i) {
        p.cb.p[i] = cb[i];
        p.gw.p[i] = grad_w ? grad_w[i] : nullptr;
    }
    p.g_q = g_q; p.g_losses = g_losses; p.grad_x = grad_x;
    p.S = S; p.G = G; p.K = K; p.D = D; p.Dg = D / G; p.T = T; p.RS = D + 1;
    p.N = (long long)B * T;
    const double numel = (double)B * (double)D * (double)T;
    p.c_cb = (float)(2.0 * lam_cb / numel);
    p.c_commit = (float)(-2.0 * lam_commit / numel);
    if (G <= 4 && p.Dg % 32 == 0) {
        if (D == 512) return launch_reg<16>(p, st);
        if (D == 256) return launch_reg<8>(p, st);
        if (D == 128) return launch_reg<4>(p, st);
    }
    const size_t smem = 2 * (size_t)TM * p.RS * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(grvq_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TM * (DMAX + 1) * 4);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(grvq_backward)");
        attr_set = true;
    }
    const long long grid = (p.N + TM - 1) / TM;
    grvq_backward_kernel<<<(unsigned)grid, NT, smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "grvq_backward launch");
}

}  // namespace acq
