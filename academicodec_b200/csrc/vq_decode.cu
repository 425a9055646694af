// K2: codebook gather-accumulate decode (RVQ decode / GRVQ embed).
//
// out[b, d, t] = 0.0 + sum_s cb[s, g(d)][code[s*G+g(d), b, t]][d mod Dg], stages added left to
// right in fp32 exactly as core_vq.py:364-370 / hificodec/models.py:510-535 do.
// HBM traffic is the codes in and the [B, D, T] latent out; the codebooks (<= 24 MiB) are
// gathered from L2.  A CTA produces a [128 channels] x [64 frames] output tile: each warp
// gathers codeword rows with 16-byte loads (512 B per request, 4 frames in flight), sums the
// stages in registers, parks them in a swizzled shared tile, and the tile is written out with
// frames contiguous using streaming stores (the output is write-once; L2 is kept for codebooks).
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
    int S, G, K, D, Dg, B, T, vec;
    long long N;
    float* out;
    int* status;
};

// Shared tile [DT channels][FT frames], frames contiguous, 16-byte chunks XOR-swizzled with
// key(d) = (d >> 2) & 7.  Gather side: a lane holds a 4 channel x 4 frame block in registers (4 frames in
// flight, 4 consecutive channels per 16-byte gather), transposes it in registers and stores, per channel,
// one 16-byte vector of 4 consecutive frames -> conflict-free STS.128.  Output side: one LDS.128 + one
// 16-byte streaming store per 4 frames of a channel (at most 2-way conflicts).
__device__ __forceinline__ int tile_off(int d, int f) {
    return d * FT + ((((f >> 2) ^ ((d >> 2) & 7)) << 2) | (f & 3));
}

__global__ void __launch_bounds__(NT) vq_decode_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) float dsm[];
    float* tile = dsm;                                                     // [DT][FT] swizzled
    int* code_s = reinterpret_cast<int*>(dsm + FT * DT);                   // [S*G][FT], -1 = invalid
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

    // gather: one warp per 4 frames; lane = 4 consecutive channels
    const int dl = lane * 4;
    const bool vec = p.vec && dl < nd;        // Dg % 4 == 0 and 16-byte aligned tables
    const int d = d0 + dl;
    const int g = vec ? d / p.Dg : 0;
    const int dg = d - g * p.Dg;
    for (int f0 = warp * 4; f0 < nf; f0 += (NT / 32) * 4) {
        float4 acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = 0; s < p.S; ++s) {
            float4 e[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                e[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int f = f0 + u;
                if (f < nf) {
                    if (vec) {
                        const int tab = s * p.G + g;
                        const int code = code_s[tab * FT + f];
                        if (code >= 0)
                            e[u] = __ldg(reinterpret_cast<const float4*>(p.cb.p[tab] + (size_t)code * p.Dg + dg));
                    } else {
                        float t4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (dl + c < nd) {
                                const int dd = d0 + dl + c, gg = dd / p.Dg, tab = s * p.G + gg;
                                const int code = code_s[tab * FT + f];
                                if (code >= 0) t4[c] = __ldg(p.cb.p[tab] + (size_t)code * p.Dg + (dd - gg * p.Dg));
                            }
                        }
                        e[u] = make_float4(t4[0], t4[1], t4[2], t4[3]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc[u].x = __fadd_rn(acc[u].x, e[u].x); acc[u].y = __fadd_rn(acc[u].y, e[u].y);
                acc[u].z = __fadd_rn(acc[u].z, e[u].z); acc[u].w = __fadd_rn(acc[u].w, e[u].w);
            }
        }
        // 4x4 register transpose: per channel, the 4 consecutive frames f0..f0+3
        if (dl < DT) {
            float* base = tile + tile_off(dl, f0);          // key(d) is the same for dl..dl+3
            *reinterpret_cast<float4*>(base) = make_float4(acc[0].x, acc[1].x, acc[2].x, acc[3].x);
            *reinterpret_cast<float4*>(base + FT) = make_float4(acc[0].y, acc[1].y, acc[2].y, acc[3].y);
            *reinterpret_cast<float4*>(base + 2 * FT) = make_float4(acc[0].z, acc[1].z, acc[2].z, acc[3].z);
            *reinterpret_cast<float4*>(base + 3 * FT) = make_float4(acc[0].w, acc[1].w, acc[2].w, acc[3].w);
        }
    }
    __syncthreads();
    // frames contiguous in the output
    if ((p.T & 3) == 0) {
        // 16-byte streaming stores: thread -> (4 consecutive frames, channel rows stepping by 16).
        // T % 4 == 0 keeps a frame quad inside one clip and the address 16-byte aligned.
        const int f4 = (tid & 15) * 4;
        if (f4 < nf) {
            const long long n = n0 + f4;
            const long long b = n / p.T, t = n % p.T;
            float* dst = p.out + ((size_t)b * p.D + d0) * p.T + t;
#pragma unroll 4
            for (int dr = tid >> 4; dr < nd; dr += NT / 16) {
                const float4 v = *reinterpret_cast<const float4*>(tile + tile_off(dr, f4));
                __stcs(reinterpret_cast<float4*>(dst + (size_t)dr * p.T), v);
            }
        }
    } else {
        const int f = tid % FT;
        if (f < nf) {
            const long long n = n0 + f;
            const long long b = n / p.T, t = n % p.T;
            float* dst = p.out + ((size_t)b * p.D + d0) * p.T + t;
#pragma unroll 8
            for (int dr = tid / FT; dr < nd; dr += NT / FT) __stcs(dst + (size_t)dr * p.T, tile[tile_off(dr, f)]);
        }
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
    p.vec = (p.Dg % 4 == 0);
    for (int i = 0; i < S * G && p.vec; ++i) p.vec = ((uintptr_t)cb[i] % 16 == 0);
    if (p.N == 0) return 0;
    dim3 grid((unsigned)((p.N + FT - 1) / FT), (unsigned)((D + DT - 1) / DT));
    const size_t smem = (size_t)DT * FT * 4 + (size_t)S * G * FT * 4;
    cudaError_t e = cudaFuncSetAttribute(vq_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(vq_decode)");
    vq_decode_kernel<<<grid, NT, smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "vq_decode launch");
}

}  // namespace acq
