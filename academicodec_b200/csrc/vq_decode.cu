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
template <int FT>
__device__ __forceinline__ int tile_off(int d, int f) {
    return d * FT + ((((f >> 2) ^ ((d >> 2) & 7)) << 2) | (f & 3));
}

template <int DT, int FT>
__global__ void __launch_bounds__(NT) vq_decode_kernel(const DecodeParams p) {
    constexpr int LPF = DT / 4;          // lanes per frame (4 channels each)
    constexpr int FPW = 4 * (32 / LPF);  // frames per warp iteration
    extern __shared__ __align__(16) float dsm[];
    float* tile = dsm;                                                     // [DT][FT] swizzled
    int* code_s = reinterpret_cast<int*>(dsm + FT * DT);                   // [S*G][FT] row offsets code * Dg, -1 = invalid
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
            if (c >= 0 && c < p.K) v = (int)c * p.Dg; else bad = true;     // element offset of the codeword row
        }
        code_s[u] = v;
    }
    if (bad && p.status) atomicExch(p.status, 1);
    const bool any_bad = __syncthreads_or(bad ? 1 : 0) != 0;

    // gather: LPF lanes per frame, 4 consecutive channels per lane, 4 frames in flight per lane
    const int dl = (lane % LPF) * 4;
    const int fsub = (lane / LPF) * 4;
    const bool vec = p.vec && dl < nd;        // Dg % 4 == 0 and 16-byte aligned tables
    const int d = d0 + dl;
    const int g = vec ? d / p.Dg : 0;
    const int dg = d - g * p.Dg;
    // Fast path (full tile, every code valid, 16-byte gathers): the gather loop is issue-bound on address
    // generation -- ncu, cfg1 4096 x 100: 29 % of all instructions on the load line (table pointer from the
    // parameter bank, 64-bit multiply-add, validity and tail tests per load).  Here the table pointer is taken once
    // per stage, the codes are staged as element offsets of their rows, and the loads of two stages are
    // in flight before the first add.
    if (!any_bad && nf == FT && nd == DT && p.vec) {
        for (int f0 = warp * FPW + fsub; f0 < FT; f0 += (NT / 32) * FPW) {
            float4 acc[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            int s = 0;
            for (; s + 1 < p.S; s += 2) {
                const float* t0 = p.cb.p[s * p.G + g] + dg;
                const float* t1 = p.cb.p[(s + 1) * p.G + g] + dg;
                const int* c0 = code_s + (s * p.G + g) * FT + f0;
                const int* c1 = code_s + ((s + 1) * p.G + g) * FT + f0;
                float4 e0[4], e1[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) e0[u] = __ldg(reinterpret_cast<const float4*>(t0 + c0[u]));
#pragma unroll
                for (int u = 0; u < 4; ++u) e1[u] = __ldg(reinterpret_cast<const float4*>(t1 + c1[u]));
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    acc[u].x = __fadd_rn(__fadd_rn(acc[u].x, e0[u].x), e1[u].x);
                    acc[u].y = __fadd_rn(__fadd_rn(acc[u].y, e0[u].y), e1[u].y);
                    acc[u].z = __fadd_rn(__fadd_rn(acc[u].z, e0[u].z), e1[u].z);
                    acc[u].w = __fadd_rn(__fadd_rn(acc[u].w, e0[u].w), e1[u].w);
                }
            }
            if (s < p.S) {
                const float* t0 = p.cb.p[s * p.G + g] + dg;
                const int* c0 = code_s + (s * p.G + g) * FT + f0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 e = __ldg(reinterpret_cast<const float4*>(t0 + c0[u]));
                    acc[u].x = __fadd_rn(acc[u].x, e.x); acc[u].y = __fadd_rn(acc[u].y, e.y);
                    acc[u].z = __fadd_rn(acc[u].z, e.z); acc[u].w = __fadd_rn(acc[u].w, e.w);
                }
            }
            float* base = tile + tile_off<FT>(dl, f0);
            *reinterpret_cast<float4*>(base) = make_float4(acc[0].x, acc[1].x, acc[2].x, acc[3].x);
            *reinterpret_cast<float4*>(base + FT) = make_float4(acc[0].y, acc[1].y, acc[2].y, acc[3].y);
            *reinterpret_cast<float4*>(base + 2 * FT) = make_float4(acc[0].z, acc[1].z, acc[2].z, acc[3].z);
            *reinterpret_cast<float4*>(base + 3 * FT) = make_float4(acc[0].w, acc[1].w, acc[2].w, acc[3].w);
        }
    } else
    for (int f0 = warp * FPW + fsub; f0 < nf; f0 += (NT / 32) * FPW) {
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
                            e[u] = __ldg(reinterpret_cast<const float4*>(p.cb.p[tab] + (size_t)code + dg));
                    } else {
                        float t4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (dl + c < nd) {
                                const int dd = d0 + dl + c, gg = dd / p.Dg, tab = s * p.G + gg;
                                const int code = code_s[tab * FT + f];
                                if (code >= 0) t4[c] = __ldg(p.cb.p[tab] + (size_t)code + (dd - gg * p.Dg));
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
            float* base = tile + tile_off<FT>(dl, f0);          // key(d) is the same for dl..dl+3
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
        constexpr int QPR = FT / 4;          // frame quads per tile row
        const int f4 = (tid % QPR) * 4;
        if (f4 < nf) {
            const long long n = n0 + f4;
            const long long b = n / p.T, t = n % p.T;
            float* dst = p.out + ((size_t)b * p.D + d0) * p.T + t;
#pragma unroll 4
            for (int dr = tid / QPR; dr < nd; dr += NT / QPR) {
                const float4 v = *reinterpret_cast<const float4*>(tile + tile_off<FT>(dr, f4));
                __stcs(reinterpret_cast<float4*>(dst + (size_t)dr * p.T), v);
            }
        }
    } else if ((p.T & 1) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 7) == 0) {
        // T even (HiFi-Codec trains on 50-frame clips): frame pairs stay inside a clip and 8-byte aligned -> 8-byte
        // streaming stores, half the instructions of the scalar path (this phase is issue-bound: ncu, cfg3 4096 x 50)
        constexpr int PPR = FT / 2;          // frame pairs per tile row
        const int f2 = (tid % PPR) * 2;
        if (f2 < nf) {
            const long long n = n0 + f2;
            const long long b = n / p.T, t = n % p.T;
            float* dst = p.out + ((size_t)b * p.D + d0) * p.T + t;
#pragma unroll 4
            for (int dr = tid / PPR; dr < nd; dr += NT / PPR) {
                const float2 v = *reinterpret_cast<const float2*>(tile + tile_off<FT>(dr, f2));
                __stcs(reinterpret_cast<float2*>(dst + (size_t)dr * p.T), v);
            }
        }
    } else {
        const int f = tid % FT;
        if (f < nf) {
            const long long n = n0 + f;
            const long long b = n / p.T, t = n % p.T;
            float* dst = p.out + ((size_t)b * p.D + d0) * p.T + t;
#pragma unroll 8
            for (int dr = tid / FT; dr < nd; dr += NT / FT) __stcs(dst + (size_t)dr * p.T, tile[tile_off<FT>(dr, f)]);
        }
    }
}

// ------------------------------------------------------------------ K2b: slice-resident decode
// For long single-stage batches the tile kernel tops out at ~4.2 TB/s: the per-frame gather
// (4*D bytes, L2 -> SM) shares the L1/L2 path with the output stream.  K2b takes the gather off
// that path: a persistent CTA keeps a 32-channel slice of every table it needs in shared memory
// (S * K * 128 bytes) and streams frames through it.  A warp step is 32 frames x 32 channels;
// a lane owns 4 channels x 8 consecutive frames: per stage 8 LDS.128 from the resident slice,
// fp32 adds in stage order from 0.0, and one 32-byte store per channel (whole sectors, frames
// contiguous) -- no transpose staging.  What the measurements forced (profiles/r01i_*):
//  * lane = quad*4 + octet, so the 4 sectors of a 128-byte output line sit in adjacent lanes:
//    ~10 L1 wavefronts per store instruction instead of 41 with the octet in the high lane bits
//    (scripts/write_probe.cu); the L1 data pipe was the limiter, not DRAM;
//  * the slice rows are XOR-swizzled by code so the 4 random rows a quarter-warp reads spread
//    over the banks (~2x conflicts instead of 4x);
//  * the codes (8 bytes per frame) come through a per-lane cp.async ring 8 steps deep: behind a
//    saturated store stream an L2 hit takes thousands of cycles, a 1-step register prefetch left
//    52 % of the warp samples waiting on it;
//  * no 64-bit divisions in the loop (clip/frame are advanced incrementally).
// HBM traffic: the output once, the codes once (re-read D/32 times from L2).
constexpr int NTS = 256;
constexpr int SW = 32;                   // channels per slice
constexpr int SQL = SW / 4;              // channel quads per slice = lanes per row
constexpr int SOPW = 32 / SQL;           // frame octets per warp step
constexpr int SFPS = SOPW * 8;           // frames per warp step (= 32: one code per lane)
constexpr int RING = 8;                  // code prefetch depth, in warp steps

__device__ __forceinline__ void stg256f(float* p, const float (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"l"(p),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}

__global__ void __launch_bounds__(NTS, 1) vq_decode_slice_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) float cb_s[];                          // [S][K][SW] swizzled
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = p.S, K = p.K;
    long long* ring = reinterpret_cast<long long*>(cb_s + (size_t)S * K * SW) + tid;   // [RING][NTS]
    const int nslices = p.D / SW;
    const int slice = blockIdx.x % nslices;
    const int rank = blockIdx.x / nslices;
    const int nrank = ((int)gridDim.x - slice + nslices - 1) / nslices;
    const int d0 = slice * SW;
    const int g = d0 / p.Dg, dg0 = d0 - g * p.Dg;

    // resident slice: row k of stage s -> cb_s[(s*K + k)*SW .. +SW), quads XOR-swizzled by k
    for (int u = tid; u < S * K * SQL; u += NTS) {
        const int q = u % SQL, k = (u / SQL) % K, st = u / (SQL * K);
        cp_async16(cb_s + ((size_t)st * K + k) * SW + (q ^ (k & (SQL - 1))) * 4,
                   p.cb.p[st * p.G + g] + (size_t)k * p.Dg + dg0 + q * 4);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    const int q = lane / SOPW, o = lane % SOPW;
    const bool oct = (p.T & 7) == 0;     // an octet never straddles clips and is 32-byte aligned
    const long long T = p.T;
    const long long nblk = (p.N + SFPS - 1) / SFPS;
    // each CTA of a slice streams through its own contiguous range of 32-frame blocks, its warps
    // interleaved inside the range
    const long long per = (nblk + nrank - 1) / nrank;
    const long long blk_end = min(nblk, (rank + 1) * per);
    const long long stride = NTS / 32;
    const long long blk0 = rank * per + warp;
    // this warp's steps: i -> (block blk0 + (i / S) * stride, stage i % S)
    const long long nstep = blk0 < blk_end ? ((blk_end - 1 - blk0) / stride + 1) * S : 0;
    // lane l prefetches the code of frame blk*32 + l of one stage into its private ring slot
    long long pf_blk = blk0;
    int pf_st = 0;
    auto prefetch = [&](long long i) {
        if (i < nstep) {
            const long long n = pf_blk * SFPS + lane;
            long long* slot = ring + (size_t)(i % RING) * NTS;
            if (n < p.N) cp_async8(slot, p.codes + (pf_st * p.G + g) * p.stride_table + n * p.stride_frame);
            else *slot = 0;
            if (++pf_st == S) { pf_st = 0; pf_blk += stride; }
        }
        cp_async_commit();
    };
    for (int i = 0; i < RING - 1; ++i) prefetch(i);

    // (clip, frame-in-clip) of this lane's octet, advanced without divisions in the loop
    const long long step_b = (stride * SFPS) / T, step_t = (stride * SFPS) % T;
    long long ob = 0, ot = 0;
    if (nstep) { ob = (blk0 * SFPS + o * 8) / T; ot = (blk0 * SFPS + o * 8) % T; }
    long long blk = blk0;
    int st = 0;
    bool bad = false;
    float acc[8][4];
    for (long long i = 0; i < nstep; ++i) {
        prefetch(i + RING - 1);
        cp_async_wait<RING - 1>();
        const long long c64 = ring[(size_t)(i % RING) * NTS];
        const int mine = (c64 >= 0 && c64 < K) ? (int)c64 : -1;
        bad |= mine < 0;
        if (st == 0) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[u][c] = 0.f;
        }
        const float* base = cb_s + (size_t)st * K * SW;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int code = __shfl_sync(0xffffffffu, mine, o * 8 + u);
            float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
            if (code >= 0)
                e = *reinterpret_cast<const float4*>(base + code * SW + (q ^ (code & (SQL - 1))) * 4);
            acc[u][0] = __fadd_rn(acc[u][0], e.x); acc[u][1] = __fadd_rn(acc[u][1], e.y);
            acc[u][2] = __fadd_rn(acc[u][2], e.z); acc[u][3] = __fadd_rn(acc[u][3], e.w);
        }
        if (++st == S) {
            st = 0;
            const long long n = blk * SFPS + o * 8;
            const int vf = (int)min(8LL, p.N - n);           // 8, 4 (N % 8 == 4 tail) or <= 0
            const int d = d0 + q * 4;
            if (oct) {
                if (vf > 0) {
                    float* dst = p.out + ((size_t)ob * p.D + d) * T + ot;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float v[8] = {acc[0][c], acc[1][c], acc[2][c], acc[3][c],
                                            acc[4][c], acc[5][c], acc[6][c], acc[7][c]};
                        stg256f(dst + (size_t)c * T, v);
                    }
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h * 4 >= vf) break;
                    long long b = ob, t = ot + h * 4;        // T % 4 == 0: a quad stays in one clip
                    if (t >= T) { t -= T; ++b; }
                    float* dst = p.out + ((size_t)b * p.D + d) * T + t;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        __stcs(reinterpret_cast<float4*>(dst + (size_t)c * T),
                               make_float4(acc[h * 4][c], acc[h * 4 + 1][c], acc[h * 4 + 2][c],
                                           acc[h * 4 + 3][c]));
                }
            }
            blk += stride;
            ob += step_b; ot += step_t;
            if (ot >= T) { ot -= T; ++ob; }
        }
    }
    cp_async_wait<0>();
    if (bad && p.status) atomicExch(p.status, 1);
}

size_t slice_smem(const DecodeParams& p) { return (size_t)p.S * p.K * SW * 4 + (size_t)RING * NTS * 8; }

int launch_slice(const DecodeParams& p, cudaStream_t st) {
    const size_t smem = slice_smem(p);
    cudaError_t e = cudaFuncSetAttribute(vq_decode_slice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(vq_decode_slice)");
    const int nslices = p.D / SW;
    const int grid = (kNumSMs / nslices) * nslices;      // equal CTAs per slice, one CTA per SM
    vq_decode_slice_kernel<<<grid, NTS, smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "vq_decode_slice launch");
}

// K2b pays when the whole table set of a slice is resident and the batch is long enough to
// amortise the fill; measured against K2 it wins for S*K*128 B <= 192 KiB (S = 1 at K = 1024),
// narrower slices (more stages) lose to bank conflicts and K2 stays the kernel for those.
bool use_slice(const DecodeParams& p) {
    return p.vec && (p.T & 3) == 0 && p.N >= 16384 && p.Dg % SW == 0 && p.D / SW <= kNumSMs &&
           slice_smem(p) <= 200 * 1024;
}

template <int DT, int FT>
int launch(const DecodeParams& p, cudaStream_t st) {
    dim3 grid((unsigned)((p.N + FT - 1) / FT), (unsigned)((p.D + DT - 1) / DT));
    const size_t smem = (size_t)DT * FT * 4 + (size_t)p.S * p.G * FT * 4;
    auto kern = vq_decode_kernel<DT, FT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(vq_decode)");
    kern<<<grid, NT, smem, st>>>(p);
    return check_cuda(cudaGetLastError(), "vq_decode launch");
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
    static const int mode = [] { const char* v = getenv("ACQ_DECODE_KERNEL"); return v ? atoi(v) : 0; }();
    // ACQ_DECODE_KERNEL=1 forces the tile kernel (K2) for A/B measurements
    if (mode != 1 && use_slice(p)) return launch_slice(p, st);
    return launch<128, 64>(p, st);
}

}  // namespace acq
