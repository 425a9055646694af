// Wire format of the codes: the reference's BitPacker / BitUnpacker
// (academicodec/binary.py:54-123) -- value i occupies bits [i*bits, (i+1)*bits) of a little-endian
// bit stream, ceil(n*bits/8) bytes.  10-bit codes instead of int64 cut the device->host traffic of
// the codes 6.4x.  One thread packs / unpacks a group of 8 values = exactly `bits` bytes.
#include "acq_common.cuh"

namespace acq {
namespace {

__global__ void pack_bits_kernel(const int64_t* __restrict__ v, long long n, int bits, uint8_t* out,
                                 long long nbytes, int* status) {
    const long long grp = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long base = grp * 8;
    if (base >= n) return;
    const unsigned long long mask = (bits >= 64) ? ~0ull : ((1ull << bits) - 1ull);
    unsigned long long lo = 0, hi = 0;      // 8 * bits <= 128 bits
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (base + j < n) {
            const long long raw = v[base + j];
            if (raw < 0 || (unsigned long long)raw > mask) bad = true;
            const unsigned long long val = (unsigned long long)raw & mask;
            const int pos = j * bits;
            if (pos < 64) {
                lo |= val << pos;
                if (pos + bits > 64) hi |= val >> (64 - pos);
            } else {
                hi |= val << (pos - 64);
            }
        }
    }
    if (bad && status) atomicExch(status, 1);
    uint8_t* dst = out + grp * bits;
    for (int b = 0; b < bits; ++b) {
        if (grp * bits + b < nbytes) dst[b] = (uint8_t)((b < 8 ? (lo >> (8 * b)) : (hi >> (8 * (b - 8)))) & 0xff);
    }
}

__global__ void unpack_bits_kernel(const uint8_t* __restrict__ in, long long n, int bits, int64_t* v,
                                   long long nbytes) {
    const long long grp = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long base = grp * 8;
    if (base >= n) return;
    unsigned long long lo = 0, hi = 0;
    const uint8_t* src = in + grp * bits;
    for (int b = 0; b < bits; ++b) {
        const unsigned long long byte = (grp * bits + b < nbytes) ? src[b] : 0ull;
        if (b < 8) lo |= byte << (8 * b); else hi |= byte << (8 * (b - 8));
    }
    const unsigned long long mask = (1ull << bits) - 1ull;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (base + j < n) {
            const int pos = j * bits;
            unsigned long long val;
            if (pos >= 64) val = hi >> (pos - 64);
            else if (pos + bits <= 64) val = lo >> pos;
            else val = (lo >> pos) | (hi << (64 - pos));
            v[base + j] = (int64_t)(val & mask);
        }
    }
}

}  // namespace

int pack_bits(const int64_t* values, long long n, int bits, uint8_t* out, int* status, cudaStream_t st) {
    if (n == 0) return 0;
    const long long nbytes = (n * bits + 7) / 8;
    const long long groups = (n + 7) / 8;
    pack_bits_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(values, n, bits, out, nbytes, status);
    return check_cuda(cudaGetLastError(), "pack_bits launch");
}

int unpack_bits(const uint8_t* in, long long n, int bits, int64_t* values, cudaStream_t st) {
    if (n == 0) return 0;
    const long long nbytes = (n * bits + 7) / 8;
    const long long groups = (n + 7) / 8;
    unpack_bits_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(in, n, bits, values, nbytes);
    return check_cuda(cudaGetLastError(), "unpack_bits launch");
}

}  // namespace acq
