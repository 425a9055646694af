// Write-bandwidth ceilings for the decode output pattern ([B, D, T] with frames contiguous).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/write_probe.cu -o /tmp/wp && /tmp/wp
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
__device__ __forceinline__ void stg256(float* p, float v) {
    unsigned r = __float_as_uint(v);
    asm volatile("st.global.v8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};\n" ::"l"(p), "r"(r) : "memory");
}
// linear: warp writes 1 KiB consecutive chunks
__global__ void k_linear(float* out, long long n) {
    long long w = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    long long nw = (long long)gridDim.x * (blockDim.x / 32);
    int lane = threadIdx.x & 31;
    for (long long c = w; c * 256 < n; c += nw) stg256(out + c * 256 + lane * 8, 1.f);
}
// rows: item = (frame block of 256, quad of 4 channels): 4 rows x 1 KiB, rows T floats apart
template <int ROWS>
__global__ void k_rows(float* out, int B, int D, int T) {
    long long w = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    long long nw = (long long)gridDim.x * (blockDim.x / 32);
    int lane = threadIdx.x & 31;
    long long N = (long long)B * T;
    int QN = D / ROWS;
    long long items = (N / 256) * QN;
    for (long long it = w; it < items; it += nw) {
        int q = it % QN; long long n = (it / QN) * 256 + lane * 8;
        long long b = n / T, t = n - b * T;
        float* dst = out + ((size_t)b * D + q * ROWS) * T + t;
#pragma unroll
        for (int c = 0; c < ROWS; ++c) stg256(dst + (size_t)c * T, 1.f);
    }
}
// slice: CTA owns 32 channels (like K2b), warps take (frame block, quad) items of that slice
__global__ void k_slice(float* out, int B, int D, int T) {
    int nsl = D / 32, slice = blockIdx.x % nsl, rank = blockIdx.x / nsl, nrank = gridDim.x / nsl;
    int lane = threadIdx.x & 31, warp = threadIdx.x / 32, nwp = blockDim.x / 32;
    long long N = (long long)B * T, items = (N / 256) * 8;
    for (long long it = (long long)rank * nwp + warp; it < items; it += (long long)nrank * nwp) {
        int q = it % 8; long long n = (it / 8) * 256 + lane * 8;
        long long b = n / T, t = n - b * T;
        float* dst = out + ((size_t)b * D + slice * 32 + q * 4) * T + t;
#pragma unroll
        for (int c = 0; c < 4; ++c) stg256(dst + (size_t)c * T, 1.f);
    }
}
// K2b-style mappings of a 32-frame x 32-channel warp step: MODE 0: lane = octet*8 + quad (8 rows,
// the row's 4 sectors 8 lanes apart); MODE 1: lane = quad*4 + octet (a row's sectors in adjacent lanes)
template <int MODE>
__global__ void k_step(float* out, int B, int D, int T) {
    int nsl = D / 32, slice = blockIdx.x % nsl, rank = blockIdx.x / nsl, nrank = gridDim.x / nsl;
    int lane = threadIdx.x & 31, warp = threadIdx.x / 32, nwp = blockDim.x / 32;
    int q = MODE == 0 ? lane % 8 : lane / 4, o = MODE == 0 ? lane / 8 : lane % 4;
    long long N = (long long)B * T, nblk = N / 32;
    for (long long blk = (long long)rank * nwp + warp; blk < nblk; blk += (long long)nrank * nwp) {
        long long n = blk * 32 + o * 8;
        long long b = n / T, t = n - b * T;
        float* dst = out + ((size_t)b * D + slice * 32 + q * 4) * T + t;
#pragma unroll
        for (int c = 0; c < 4; ++c) stg256(dst + (size_t)c * T, 1.f);
    }
}
template <class F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < 20; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 20;
}
int main() {
    const int B = 8, D = 512, T = 45000;
    const long long n = (long long)B * D * T;
    float* out; cudaMalloc(&out, n * 4);
    auto rep = [&](const char* name, float ms) { printf("%-28s %.4f ms  %.0f GB/s\n", name, ms, n * 4.0 / ms / 1e6); };
    rep("cudaMemset", timeit([&] { cudaMemsetAsync(out, 0, n * 4); }));
    for (int g : {148, 296, 592, 1184, 4736})
        for (int nt : {256, 512, 1024}) {
            char nm[64]; snprintf(nm, 64, "linear g=%d nt=%d", g, nt);
            rep(nm, timeit([&] { k_linear<<<g, nt>>>(out, n); }));
        }
    for (int g : {148, 296, 592, 2368})
        for (int nt : {256, 512, 1024}) {
            char nm[64]; snprintf(nm, 64, "rows4 g=%d nt=%d", g, nt);
            rep(nm, timeit([&] { k_rows<4><<<g, nt>>>(out, B, D, T); }));
        }
    for (int g : {148, 592})
        for (int nt : {512, 1024}) {
            char nm[64]; snprintf(nm, 64, "rows1 g=%d nt=%d", g, nt);
            rep(nm, timeit([&] { k_rows<1><<<g, nt>>>(out, B, D, T); }));
        }
    for (int nt : {256, 512, 1024}) {
        char nm[64]; snprintf(nm, 64, "slice g=144 nt=%d", nt);
        rep(nm, timeit([&] { k_slice<<<144, nt>>>(out, B, D, T); }));
    }
    rep("step lane=o*8+q g=144 nt=512", timeit([&] { k_step<0><<<144, 512>>>(out, B, D, T); }));
    rep("step lane=q*4+o g=144 nt=512", timeit([&] { k_step<1><<<144, 512>>>(out, B, D, T); }));
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
