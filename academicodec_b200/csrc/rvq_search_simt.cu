// K1a: fused residual nearest-codeword search on the fp32 CUDA cores (any shape).
//
// One CTA owns a tile of TM = 16*FR consecutive frames and carries their residual in shared
// memory through every stage and group: distance contraction (register-tiled FFMA against
// codebook chunks streamed with cp.async), -0.5||e||^2 bias, (value, index) argmax with the
// lowest-index tie rule, codeword gather, residual subtract and the running quantized sum.
// Nothing but x, the codes and the optional outputs touches HBM; the [N, K] distance matrix
// of the reference (core_vq.py:177-179) never exists.
//
// Reference semantics reproduced (see include/acq_b200.h):
//   eval  : r <- r - e[i]                               (core_vq.py:357-359)
//   STE   : q' = r + (e[i] - r);  r <- r - q'           (core_vq.py:304,339; models.py:478,502)
//   out   : 0.0 + q'_0 + q'_1 + ...  left to right      (core_vq.py:329,340)
#include "acq_common.cuh"
#include <math_constants.h>

namespace acq {

namespace {

constexpr int KC = 128;      // codewords per streamed chunk
constexpr int DC = 64;       // channels per streamed chunk
constexpr int ES = DC + 4;   // chunk row stride (floats): 17 x 16 B -> conflict-free LDS.128
constexpr int NT = 256;      // 16 (codeword lanes) x 16 (frame lanes)
constexpr int CW = KC / 16;  // codewords per thread per chunk

struct SearchParams {
    const float* x;
    PtrTable cb;
    const float* half_norms;
    int S, G, K, D, Dg, DgP, RS, B, T, flags;
    long long N;
    int64_t* codes;
    float* quantized;
    float* residual;
    double* sqerr;
};

template <bool VEC>
__device__ __forceinline__ void load_chunk(float* e_buf, const float* __restrict__ cbp, int K,
                                           int Dg, int kc, int dc, int tid) {
    if (VEC) {
#pragma unroll
        for (int k = 0; k < (KC * DC / 4) / NT; ++k) {
            int u = tid + NT * k;
            int row = u / (DC / 4);
            int q = u % (DC / 4);
            int c = kc * KC + row;
            int d = dc * DC + q * 4;
            float* dst = e_buf + row * ES + q * 4;
            if (c < K && d < Dg) {
                cp_async16(dst, cbp + (size_t)c * Dg + d);
            } else {
                *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    } else {
        for (int k = 0; k < (KC * DC) / NT; ++k) {
            int u = tid + NT * k;
            int row = u / DC;
            int dcol = u % DC;
            int c = kc * KC + row;
            int d = dc * DC + dcol;
            e_buf[row * ES + dcol] = (c < K && d < Dg) ? __ldg(cbp + (size_t)c * Dg + d) : 0.f;
        }
    }
}

template <int FR, bool VEC>
__global__ void __launch_bounds__(NT) rvq_search_simt_kernel(const SearchParams p) {
    constexpr int TM = 16 * FR;
    extern __shared__ __align__(16) float smem[];
    const int RS = p.RS;
    float* r_s = smem;                                   // [TM][RS] residual
    float* q_s = p.quantized ? r_s + TM * RS : nullptr;  // [TM][RS] running quantized sum
    float* e_s = r_s + (p.quantized ? 2 : 1) * TM * RS;  // [2][KC][ES]
    int* best_s = reinterpret_cast<int*>(e_s + 2 * KC * ES);  // [TM]

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int lane = tid & 31, warp = tid >> 5;
    const long long n0 = (long long)blockIdx.x * TM;
    const int nf = (int)min((long long)TM, p.N - n0);
    const int D = p.D, Dg = p.Dg, DgP = p.DgP, G = p.G, K = p.K, T = p.T;
    const bool ste = p.flags & ACQ_STE;
    const bool loss_raw = p.flags & ACQ_LOSS_RAW;

    // ---- load the x tile, transposing [D][T] -> [frame][channel] --------------------------
    {
        const int f = tid % TM;
        const long long n = n0 + f;
        const bool ok = f < nf;
        const long long b = ok ? n / T : 0;
        const long long t = ok ? n % T : 0;
        const float* src = p.x + (b * D) * (long long)T + t;
        for (int slot = tid / TM; slot < G * DgP; slot += NT / TM) {
            int g = slot / DgP, dl = slot % DgP;
            float v = 0.f;
            if (ok && dl < Dg) v = __ldg(src + (long long)(g * Dg + dl) * T);
            r_s[f * RS + slot] = v;
        }
        if (tid < TM) {
            for (int k = 0; k < 4; ++k) r_s[tid * RS + G * DgP + k] = 0.f;
        }
    }
    __syncthreads();

    const int nkc = (K + KC - 1) / KC;
    const int ndc = (Dg + DC - 1) / DC;
    const int total = nkc * ndc;

    for (int s = 0; s < p.S; ++s) {
        float werr = 0.f;  // per-warp squared-error partial of this stage
        for (int g = 0; g < G; ++g) {
            const float* __restrict__ cbp = p.cb.p[s * G + g];
            const float* __restrict__ hn = p.half_norms + (size_t)(s * G + g) * K;
            float best[FR];
            int bidx[FR];
#pragma unroll
            for (int i = 0; i < FR; ++i) {
                best[i] = -CUDART_INF_F;
                bidx[i] = 0;
            }
            float acc[FR][CW];

            load_chunk<VEC>(e_s, cbp, K, Dg, 0, 0, tid);
            cp_async_commit();
            for (int it = 0; it < total; ++it) {
                const int kc = it / ndc, dc = it % ndc;
                const int buf = it & 1;
                if (it + 1 < total) {
                    load_chunk<VEC>(e_s + (buf ^ 1) * KC * ES, cbp, K, Dg, (it + 1) / ndc,
                                    (it + 1) % ndc, tid);
                    cp_async_commit();
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncthreads();
                if (dc == 0) {
#pragma unroll
                    for (int i = 0; i < FR; ++i)
#pragma unroll
                        for (int j = 0; j < CW; ++j) acc[i][j] = 0.f;
                }
                const int dlen = min(DC, Dg - dc * DC);
                const int nd4 = (dlen + 3) >> 2;
                const float* rb = r_s + ty * RS + g * DgP + dc * DC;
                const float* eb = e_s + buf * KC * ES + tx * ES;
#pragma unroll 2
                for (int d4 = 0; d4 < nd4; ++d4) {
                    float4 xv[FR], ev[CW];
#pragma unroll
                    for (int i = 0; i < FR; ++i)
                        xv[i] = *reinterpret_cast<const float4*>(rb + 16 * i * RS + d4 * 4);
#pragma unroll
                    for (int j = 0; j < CW; ++j)
                        ev[j] = *reinterpret_cast<const float4*>(eb + 16 * j * ES + d4 * 4);
#pragma unroll
                    for (int i = 0; i < FR; ++i)
#pragma unroll
                        for (int j = 0; j < CW; ++j) {
                            float a = acc[i][j];
                            a = fmaf(xv[i].x, ev[j].x, a);
                            a = fmaf(xv[i].y, ev[j].y, a);
                            a = fmaf(xv[i].z, ev[j].z, a);
                            a = fmaf(xv[i].w, ev[j].w, a);
                            acc[i][j] = a;
                        }
                }
                if (dc == ndc - 1) {
#pragma unroll
                    for (int j = 0; j < CW; ++j) {
                        const int c = kc * KC + tx + 16 * j;
                        if (c < K) {
                            const float h = __ldg(hn + c);
#pragma unroll
                            for (int i = 0; i < FR; ++i) {
                                const float sc = acc[i][j] - h;   // x.e - 0.5||e||^2
                                if (sc > best[i]) {
                                    best[i] = sc;
                                    bidx[i] = c;
                                }
                            }
                        }
                    }
                }
                __syncthreads();
            }
            // ---- argmax across the 16 codeword lanes: (value desc, index asc) ---------------
#pragma unroll
            for (int i = 0; i < FR; ++i) {
                float v = best[i];
                int id = bidx[i];
#pragma unroll
                for (int off = 8; off >= 1; off >>= 1) {
                    float ov = __shfl_xor_sync(0xffffffffu, v, off);
                    int oid = __shfl_xor_sync(0xffffffffu, id, off);
                    if (ov > v || (ov == v && oid < id)) {
                        v = ov;
                        id = oid;
                    }
                }
                if (tx == 0) best_s[ty + 16 * i] = id;
            }
            __syncthreads();
            // ---- gather the winners, update residual / running sum, one warp per frame -----
            for (int f = warp; f < nf; f += NT / 32) {
                const int idx = best_s[f];
                if (lane == 0)
                    p.codes[(size_t)(s * G + g) * p.N + n0 + f] = idx;
                const float* erow = cbp + (size_t)idx * Dg;
                float* rrow = r_s + f * RS + g * DgP;
                float* qrow = q_s ? q_s + f * RS + g * DgP : nullptr;
                for (int d = lane * 4; d < Dg; d += 128) {
                    float e[4];
                    if (VEC) {
                        float4 e4 = __ldg(reinterpret_cast<const float4*>(erow + d));
                        e[0] = e4.x; e[1] = e4.y; e[2] = e4.z; e[3] = e4.w;
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) e[k] = (d + k < Dg) ? __ldg(erow + d + k) : 0.f;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (d + k < Dg) {
                            const float r = rrow[d + k];
                            const float q = e[k];
                            const float qs = ste ? __fadd_rn(r, __fsub_rn(q, r)) : q;
                            const float df = loss_raw ? __fsub_rn(q, r) : __fsub_rn(qs, r);
                            werr = fmaf(df, df, werr);
                            rrow[d + k] = __fsub_rn(r, qs);
                            if (qrow) qrow[d + k] = __fadd_rn(s == 0 ? 0.f : qrow[d + k], qs);
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (p.sqerr) {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) werr += __shfl_xor_sync(0xffffffffu, werr, off);
            if (lane == 0 && werr != 0.f) atomicAdd(p.sqerr + s, (double)werr);
        }
    }

    // ---- write [frame][channel] tiles back as [D][T] ------------------------------------------
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

__global__ void half_norms_kernel(PtrTable cb, int n_tables, int K, int Dg, float* out) {
    // one warp per codeword; fp64 accumulation, fixed tree -> deterministic
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_tables * K) return;
    const float* row = cb.p[warp / K] + (size_t)(warp % K) * Dg;
    double acc = 0.0;
    for (int d = lane; d < Dg; d += 32) {
        double v = (double)__ldg(row + d);
        acc = fma(v, v, acc);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) out[warp] = (float)(0.5 * acc);
}

template <int FR, bool VEC>
int launch(const SearchParams& p, size_t smem, cudaStream_t st) {
    auto kern = rvq_search_simt_kernel<FR, VEC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(rvq_search_simt)");
    const int TM = 16 * FR;
    const long long tiles = (p.N + TM - 1) / TM;
    kern<<<(unsigned)tiles, NT, smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "rvq_search_simt launch");
}

}  // namespace

size_t simt_smem_bytes(int FR, int RS, bool with_q) {
    const int TM = 16 * FR;
    return (size_t)(with_q ? 2 : 1) * TM * RS * 4 + 2 * KC * ES * 4 + TM * 4;
}

int rvq_search_simt(const float* x, const float* const* cb, const float* half_norms, int S, int G,
                    int K, int D, int B, int T, int flags, int64_t* codes, float* quantized,
                    float* residual, double* sqerr, cudaStream_t st) {
    SearchParams p;
    p.x = x;
    for (int i = 0; i < S * G; ++i) p.cb.p[i] = cb[i];
    p.half_norms = half_norms;
    p.S = S; p.G = G; p.K = K; p.D = D; p.Dg = D / G; p.DgP = (p.Dg + 3) & ~3;
    p.RS = G * p.DgP + 4;
    p.B = B; p.T = T; p.flags = flags; p.N = (long long)B * T;
    p.codes = codes; p.quantized = quantized; p.residual = residual; p.sqerr = sqerr;
    bool vec = (p.Dg % 4 == 0);
    for (int i = 0; i < S * G && vec; ++i) vec = ((uintptr_t)cb[i] % 16 == 0);

    const size_t cap = 227 * 1024;
    int FR = 0;
    for (int cand : {4, 2, 1}) {
        if (simt_smem_bytes(cand, p.RS, quantized != nullptr) > cap) continue;
        FR = cand;                                     // largest tile that fits ...
        const long long tiles = (p.N + 16 * cand - 1) / (16 * cand);
        if (tiles >= 2 * kNumSMs) break;               // ... unless it leaves SMs idle
    }
    if (FR == 0) return fail(ACQ_ESHAPE, "rvq_search_simt: D=%d does not fit shared memory", D);
    const size_t smem = simt_smem_bytes(FR, p.RS, quantized != nullptr);
#define ACQ_LAUNCH(F)                                                        \
    return vec ? launch<F, true>(p, smem, st) : launch<F, false>(p, smem, st)
    switch (FR) {
        case 4: ACQ_LAUNCH(4);
        case 2: ACQ_LAUNCH(2);
        default: ACQ_LAUNCH(1);
    }
#undef ACQ_LAUNCH
}

int codebook_half_norms(const float* const* cb, int n_tables, int K, int Dg, float* out,
                        cudaStream_t st) {
    PtrTable t;
    for (int i = 0; i < n_tables; ++i) t.p[i] = cb[i];
    const long long warps = (long long)n_tables * K;
    const int block = 256;
    const long long grid = (warps * 32 + block - 1) / block;
    half_norms_kernel<<<(unsigned)grid, block, 0, st>>>(t, n_tables, K, Dg, out);
    return check_cuda(cudaGetLastError(), "half_norms launch");
}

}  // namespace acq
