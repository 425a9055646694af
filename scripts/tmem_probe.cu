// TMEM read throughput / latency probe (B200): how fast can the epilogue drain an accumulator?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_probe scripts/tmem_probe.cu && /tmp/tmem_probe
// Each of NW warps (warp w reads lanes 32 (w % 4)..) issues ITER tcgen05.ld of shape 32x32b.xN, either
// one at a time (issue, wait) or double-buffered (next read in flight while the previous is consumed).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void wait32(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;\n"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}

template <int MODE>   // 0: x16 serial, 1: x16 double-buffered, 2: x32 serial, 3: x32 double-buffered
__global__ void probe(long long* out, int iters) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;\n" ::"r"((uint32_t)__cvta_generic_to_shared(&tptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16) + (warp >= 4 ? 256u : 0u);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (MODE == 0) {
        uint32_t r[16];
        for (int i = 0; i < iters; ++i) { ld16(base + ((i * 16) & 255), r); wait16(r); acc += r[0] ^ r[15]; }
    } else if (MODE == 1) {
        uint32_t a[16], b[16];
        ld16(base, a);
        for (int i = 0; i < iters; i += 2) {
            wait16(a); ld16(base + (((i + 1) * 16) & 255), b); acc += a[0] ^ a[15];
            wait16(b); ld16(base + (((i + 2) * 16) & 255), a); acc += b[0] ^ b[15];
        }
        wait16(a);
    } else if (MODE == 2) {
        uint32_t r[32];
        for (int i = 0; i < iters; ++i) { ld32(base + ((i * 32) & 255), r); wait32(r); acc += r[0] ^ r[31]; }
    } else {
        uint32_t a[32], b[32];
        ld32(base, a);
        for (int i = 0; i < iters; i += 2) {
            wait32(a); ld32(base + (((i + 1) * 32) & 255), b); acc += a[0] ^ a[31];
            wait32(b); ld32(base + (((i + 2) * 32) & 255), a); acc += b[0] ^ b[31];
        }
        wait32(a);
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = acc; }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;\n" ::"r"(tptr) : "memory");
}

int main() {
    long long* d; cudaMalloc(&d, 148 * 16);
    long long h[2];
    const int iters = 4096;
    const char* names[4] = {"x16 serial", "x16 double-buffered", "x32 serial", "x32 double-buffered"};
    for (int nw = 4; nw <= 8; nw += 4) {
        for (int mode = 0; mode < 4; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) probe<0><<<148, nw * 32>>>(d, iters);
                if (mode == 1) probe<1><<<148, nw * 32>>>(d, iters);
                if (mode == 2) probe<2><<<148, nw * 32>>>(d, iters);
                if (mode == 3) probe<3><<<148, nw * 32>>>(d, iters);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            const double cols = (mode < 2 ? 16.0 : 32.0);
            const double bytes = (double)iters * cols * 32 * 4 * nw;      // per SM
            printf("%d warps, %-20s: %8lld cycles for %d reads per warp = %6.1f cycles per read, %6.1f B/clk/SM\n",
                   nw, names[mode], h[0], iters, (double)h[0] / iters, bytes / (double)h[0]);
        }
    }
    return 0;
}
