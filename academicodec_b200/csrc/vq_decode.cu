// K2: codebook gather-accumulate decode (RVQ decode / GRVQ embed).
//
// out[b, d, t] = 0.0 + sum_s cb[s, g(d)][code[s*G+g(d), b, t]][d mod Dg], stages added left to
// right in fp32 exactly as core_vq.py:364-370 / hificodec/models.py:510-535 do.
// HBM traffic is the codes in and the [B, D, T] latent out; the codebooks (<= 24 MiB) are
// gathered from L2.  A CTA produces a [128 channels] x [64 frames] output tile: each warp
// gathers whole codeword rows (coalesced 128 B per request), sums the stages in registers,
// parks the column in shared memory, and the tile is written out with frames contiguous.
#include "acq_common.cuh"

namespace acq {
namespace {

constexpr int DT = 128;   // channels per tile
constexpr int FT = 64;    // frames per tile
constexpr int NT = 256;

struct DecodeParams {
    const int64_t* codes;
    long long stride_table, stride_frame;
    PtrTable cb;
    int S, G, K, D, Dg, B, T;
    long long N;
    float* out;
    int* status;
};

__global__ void __launch_bounds__(NT) vq_decode_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) float dsm[];
    float (*tile)[FT + 1] = reinterpret_cast<float (*)[FT + 1]>(dsm);      // [DT][FT+1]
    int* code_s = reinterpret_cast<int*>(dsm + DT * (FT + 1));             // [S*G][FT], -1 = invalid
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n0 = (long long)blockIdx.x * FT;
    const int d0 = blockIdx.y * DT;
    const int nf = (int)min((long long)FT, p.N - n0);
    const int nd = min(DT, p.D - d0);
    const int ntab = p.S * p.G;

    // stage the tile's codes (one pass, frames fastest -> coalesced for the RVQ layout)
    bool bad = false;
    for (int u = tid; u < ntab * FT; u += NT) {
        const int tab = u / FT, f = u % FT;
        int v = -1;
        if (f < nf) {
            const long long c = __ldg(p.codes + tab * p.stride_table + (n0 + f) * p.stride_frame);
            if (c >= 0 && c < p.K) v = (int)c; else bad = true;
        }
        code_s[u] = v;
    }
    if (bad && p.status) atomicExch(p.status, 1);
    __syncthreads();

    for (int f = warp; f < nf; f += NT / 32) {
        float acc[DT / 32];
#pragma unroll
        for (int c = 0; c < DT / 32; ++c) acc[c] = 0.f;
        for (int s = 0; s < p.S; ++s) {
#pragma unroll
            for (int c = 0; c < DT / 32; ++c) {
                const int dl = lane + 32 * c;
                if (dl < nd) {
                    const int d = d0 + dl;
                    const int g = d / p.Dg;
                    const int tab = s * p.G + g;
                    const int code = code_s[tab * FT + f];
                    float e = 0.f;
                    if (code >= 0) e = __ldg(p.cb.p[tab] + (size_t)code * p.Dg + (d - g * p.Dg));
                    acc[c] = __fadd_rn(acc[c], e);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < DT / 32; ++c) tile[lane + 32 * c][f] = acc[c];
    }
    __syncthreads();
    // frames contiguous in the output: thread -> (frame f, channel rows stepping by NT/FT)
    const int f = tid % FT;
    if (f < nf) {
        const long long n = n0 + f;
        const long long b = n / p.T, t = n % p.T;
        float* dst = p.out + ((size_t)b * p.D + d0) * p.T + t;
        for (int dl = tid / FT; dl < nd; dl += NT / FT) dst[(size_t)dl * p.T] = tile[dl][f];
    }
}

}  // namespace

int vq_decode(const int64_t* codes, int64_t stride_table, int64_t stride_frame,
              const float* const* cb, int S, int G, int K, int D, int B, int T, float* out,
              int* status, cudaStream_t st) {
    DecodeParams p;
    p.codes = codes; p.stride_table = stride_table; p.stride_frame = stride_frame;
    for (int i = 0; i < S * G; ++i) p.cb.p[i] = cb[i];
    p.S = S; p.G = G; p.K = K; p.D = D; p.Dg = D / G; p.B = B; p.T = T;
    p.N = (long long)B * T; p.out = out; p.status = status;
    if (p.N == 0) return 0;
    dim3 grid((unsigned)((p.N + FT - 1) / FT), (unsigned)((D + DT - 1) / DT));
    const size_t smem = (size_t)DT * (FT + 1) * 4 + (size_t)S * G * FT * 4;
    cudaError_t e = cudaFuncSetAttribute(vq_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(vq_decode)");
    vq_decode_kernel<<<grid, NT, smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "vq_decode launch");
}

}  // namespace acq
