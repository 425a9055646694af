// Host-buffer entry points: the same kernels fed from / drained to HOST memory in chunks, with
// host->device copy, compute and device->host copy of consecutive chunks overlapped on a ring
// of streams.  This is the end-to-end path bench.py reports as `e2e`.
#include "acq_common.cuh"
#include <algorithm>

namespace acq {
int rvq_search_dispatch(const float*, const float* const*, const float*, const void*, void*, int,
                        int, int, int, int, int, int, int, int64_t*, float*, float*, double*,
                        cudaStream_t);
size_t tc_workspace_bytes(int);
int validate_search(const float*, const float* const*, const float*, int, int, int, int, int, int,
                    const int64_t*);
int vq_decode(const int64_t*, int64_t, int64_t, const float* const*, int, int, int, int, int, int,
              float*, int*, cudaStream_t);
}  // namespace acq

using namespace acq;

struct acq_pipeline {
    static const int NBUF = 3;
    int device = 0;
    size_t chunk_bytes = 0;
    cudaStream_t stream[NBUF] = {};
    float* d_lat[NBUF] = {};
    float* d_out[NBUF] = {};      // second latent staging buffer (codec round trip), lazily allocated
    int64_t* d_codes[NBUF] = {};
    size_t codes_cap[NBUF] = {};
    void* d_work[NBUF] = {};
    size_t work_cap[NBUF] = {};
    int launches = 0;
    cudaEvent_t producer_done = nullptr;   // acq_pipeline_wait_stream
};

namespace {

int ensure_codes(acq_pipeline* p, int slot, size_t bytes) {
    if (p->codes_cap[slot] >= bytes) return 0;
    if (p->d_codes[slot]) cudaFree(p->d_codes[slot]);
    p->d_codes[slot] = nullptr;
    p->codes_cap[slot] = 0;
    int rc = check_cuda(cudaMalloc(&p->d_codes[slot], bytes), "cudaMalloc(codes staging)");
    if (rc) return rc;
    p->codes_cap[slot] = bytes;
    return 0;
}

int ensure_work(acq_pipeline* p, int slot, size_t bytes) {
    if (p->work_cap[slot] >= bytes) return 0;
    if (p->d_work[slot]) cudaFree(p->d_work[slot]);
    p->d_work[slot] = nullptr;
    p->work_cap[slot] = 0;
    int rc = check_cuda(cudaMalloc(&p->d_work[slot], bytes), "cudaMalloc(tc workspace)");
    if (rc) return rc;
    p->work_cap[slot] = bytes;
    return 0;
}

// A chunk is either `nb` whole clips starting at clip b0 (t0 = 0, nt = T) or a frame range
// [t0, t0+nt) of the single clip b0.
struct Chunk { int b0, nb, t0, nt; };

template <typename F>
int for_each_chunk(size_t chunk_bytes, int D, int B, int T, F&& fn) {
    const size_t clip_bytes = (size_t)D * T * sizeof(float);
    int idx = 0;
    if (clip_bytes <= chunk_bytes) {
        // whole clips travel as contiguous 1-D copies (the fast PCIe path; halving the first and last
        // chunk along T to shorten the pipeline's fill and drain was measured and does not pay)
        const int per = (int)std::max<size_t>(1, chunk_bytes / std::max<size_t>(clip_bytes, 1));
        for (int b = 0; b < B; b += per) {
            int rc = fn(idx++, Chunk{b, std::min(per, B - b), 0, T});
            if (rc) return rc;
        }
    } else {
        int tc = (int)(chunk_bytes / ((size_t)D * sizeof(float)));
        tc = std::max(64, tc & ~63);
        for (int b = 0; b < B; ++b)
            for (int t = 0; t < T; t += tc) {
                int rc = fn(idx++, Chunk{b, 1, t, std::min(tc, T - t)});
                if (rc) return rc;
            }
    }
    return 0;
}

int sync_all(acq_pipeline* p) {
    for (int i = 0; i < acq_pipeline::NBUF; ++i) {
        int rc = check_cuda(cudaStreamSynchronize(p->stream[i]), "pipeline stream sync");
        if (rc) return rc;
    }
    return 0;
}

}  // namespace

extern "C" {

int acq_pipeline_create(acq_pipeline** out, int device, size_t chunk_bytes) {
    if (!out) return fail(ACQ_EINVAL, "acq_pipeline_create: null out");
    int rc = check_cuda(cudaSetDevice(device), "cudaSetDevice");
    if (rc) return rc;
    acq_pipeline* p = new acq_pipeline();
    p->device = device;
    p->chunk_bytes = chunk_bytes ? chunk_bytes : ((size_t)128 << 20);
    rc = check_cuda(cudaEventCreateWithFlags(&p->producer_done, cudaEventDisableTiming), "cudaEventCreate");
    if (rc) {
        delete p;
        return rc;
    }
    for (int i = 0; i < acq_pipeline::NBUF; ++i) {
        rc = check_cuda(cudaStreamCreateWithFlags(&p->stream[i], cudaStreamNonBlocking), "cudaStreamCreate");
        if (!rc) rc = check_cuda(cudaMalloc(&p->d_lat[i], p->chunk_bytes), "cudaMalloc(latent staging)");
        if (rc) {
            acq_pipeline_destroy(p);
            return rc;
        }
    }
    *out = p;
    return 0;
}

void acq_pipeline_destroy(acq_pipeline* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    for (int i = 0; i < acq_pipeline::NBUF; ++i) {
        if (p->stream[i]) {
            cudaStreamSynchronize(p->stream[i]);
            cudaStreamDestroy(p->stream[i]);
        }
        if (p->d_lat[i]) cudaFree(p->d_lat[i]);
        if (p->d_out[i]) cudaFree(p->d_out[i]);
        if (p->d_codes[i]) cudaFree(p->d_codes[i]);
        if (p->d_work[i]) cudaFree(p->d_work[i]);
    }
    if (p->producer_done) cudaEventDestroy(p->producer_done);
    delete p;
}

// The ring streams are non-blocking: they do not order themselves after the caller's stream.  Tables the
// next *_host call reads (codebooks, half norms, tensor-core pack) are often produced asynchronously on
// that stream (acq_codebook_half_norms, acq_tc_pack_codebooks, acq_ema_apply): make every ring stream
// wait for what `producer_stream` has been given so far.
int acq_pipeline_wait_stream(acq_pipeline* p, void* producer_stream) {
    if (!p) return fail(ACQ_EINVAL, "null pipeline");
    int rc = check_cuda(cudaSetDevice(p->device), "cudaSetDevice");
    if (rc) return rc;
    rc = check_cuda(cudaEventRecord(p->producer_done, (cudaStream_t)producer_stream), "cudaEventRecord");
    for (int i = 0; i < acq_pipeline::NBUF && !rc; ++i)
        rc = check_cuda(cudaStreamWaitEvent(p->stream[i], p->producer_done, 0), "cudaStreamWaitEvent");
    return rc;
}

int acq_pipeline_last_launches(const acq_pipeline* p) { return p ? p->launches : 0; }

int acq_rvq_encode_host(acq_pipeline* p, const float* x_host, const float* const* cb,
                        const float* half_norms, const void* tc_pack, int S, int G, int K, int D,
                        int B, int T, int flags, int impl, int64_t* codes_host) {
    if (!p) return fail(ACQ_EINVAL, "null pipeline");
    int rc = validate_search(x_host, cb, half_norms, S, G, K, D, B, T, codes_host);
    if (rc) return rc;
    if ((size_t)D * 64 * sizeof(float) > p->chunk_bytes)
        return fail(ACQ_ESHAPE, "chunk_bytes too small for D=%d", D);
    rc = check_cuda(cudaSetDevice(p->device), "cudaSetDevice");
    if (rc) return rc;
    p->launches = 0;
    const int tables = S * G;
    const long long N = (long long)B * T;
    rc = for_each_chunk(p->chunk_bytes, D, B, T, [&](int idx, Chunk c) -> int {
        const int slot = idx % acq_pipeline::NBUF;
        cudaStream_t st = p->stream[slot];
        const long long frames = (long long)c.nb * c.nt;
        int r = ensure_codes(p, slot, (size_t)tables * frames * sizeof(int64_t));
        if (r) return r;
        if (tc_pack) {
            r = ensure_work(p, slot, tc_workspace_bytes(D));
            if (r) return r;
        }
        if (c.nt == T) {
            r = check_cuda(cudaMemcpyAsync(p->d_lat[slot], x_host + (size_t)c.b0 * D * T,
                                           (size_t)frames * D * sizeof(float),
                                           cudaMemcpyHostToDevice, st), "H2D latents");
        } else {
            r = check_cuda(cudaMemcpy2DAsync(p->d_lat[slot], (size_t)c.nt * sizeof(float),
                                             x_host + (size_t)c.b0 * D * T + c.t0,
                                             (size_t)T * sizeof(float), (size_t)c.nt * sizeof(float),
                                             D, cudaMemcpyHostToDevice, st), "H2D latents (2D)");
        }
        if (r) return r;
        r = rvq_search_dispatch(p->d_lat[slot], cb, half_norms, tc_pack,
                                tc_pack ? p->d_work[slot] : nullptr, S, G, K, D, c.nb, c.nt, flags,
                                impl, p->d_codes[slot], nullptr, nullptr, nullptr, st);
        if (r) return r;
        p->launches += 1;
        return check_cuda(cudaMemcpy2DAsync(codes_host + (size_t)c.b0 * T + c.t0,
                                            (size_t)N * sizeof(int64_t), p->d_codes[slot],
                                            (size_t)frames * sizeof(int64_t),
                                            (size_t)frames * sizeof(int64_t), tables,
                                            cudaMemcpyDeviceToHost, st), "D2H codes");
    });
    int rs = sync_all(p);
    return rc ? rc : rs;
}

int acq_rvq_codec_host(acq_pipeline* p, const float* x_host, const float* const* cb,
                       const float* half_norms, const void* tc_pack, int S, int G, int K, int D,
                       int B, int T, int flags, int impl, int64_t* codes_host, float* out_host) {
    if (!p) return fail(ACQ_EINVAL, "null pipeline");
    int rc = validate_search(x_host, cb, half_norms, S, G, K, D, B, T, codes_host);
    if (rc) return rc;
    if ((long long)B * T == 0) return 0;
    if (!out_host) return fail(ACQ_EINVAL, "acq_rvq_codec_host: null output");
    if ((size_t)D * 64 * sizeof(float) > p->chunk_bytes)
        return fail(ACQ_ESHAPE, "chunk_bytes too small for D=%d", D);
    rc = check_cuda(cudaSetDevice(p->device), "cudaSetDevice");
    if (rc) return rc;
    for (int i = 0; i < acq_pipeline::NBUF; ++i) {
        if (!p->d_out[i]) {
            rc = check_cuda(cudaMalloc(&p->d_out[i], p->chunk_bytes), "cudaMalloc(output staging)");
            if (rc) return rc;
        }
    }
    p->launches = 0;
    const int tables = S * G;
    const long long N = (long long)B * T;
    rc = for_each_chunk(p->chunk_bytes, D, B, T, [&](int idx, Chunk c) -> int {
        const int slot = idx % acq_pipeline::NBUF;
        cudaStream_t st = p->stream[slot];
        const long long frames = (long long)c.nb * c.nt;
        int r = ensure_codes(p, slot, (size_t)tables * frames * sizeof(int64_t));
        if (r) return r;
        if (tc_pack) {
            r = ensure_work(p, slot, tc_workspace_bytes(D));
            if (r) return r;
        }
        const size_t lat_off = (size_t)c.b0 * D * T + c.t0;
        if (c.nt == T) {
            r = check_cuda(cudaMemcpyAsync(p->d_lat[slot], x_host + lat_off, (size_t)frames * D * sizeof(float),
                                           cudaMemcpyHostToDevice, st), "H2D latents");
        } else {
            r = check_cuda(cudaMemcpy2DAsync(p->d_lat[slot], (size_t)c.nt * sizeof(float), x_host + lat_off,
                                             (size_t)T * sizeof(float), (size_t)c.nt * sizeof(float), D,
                                             cudaMemcpyHostToDevice, st), "H2D latents (2D)");
        }
        if (r) return r;
        r = rvq_search_dispatch(p->d_lat[slot], cb, half_norms, tc_pack, tc_pack ? p->d_work[slot] : nullptr,
                                S, G, K, D, c.nb, c.nt, flags, impl, p->d_codes[slot], nullptr, nullptr,
                                nullptr, st);
        if (r) return r;
        r = check_cuda(cudaMemcpy2DAsync(codes_host + (size_t)c.b0 * T + c.t0, (size_t)N * sizeof(int64_t),
                                         p->d_codes[slot], (size_t)frames * sizeof(int64_t),
                                         (size_t)frames * sizeof(int64_t), tables, cudaMemcpyDeviceToHost, st),
                       "D2H codes");
        if (r) return r;
        r = vq_decode(p->d_codes[slot], frames, 1, cb, S, G, K, D, c.nb, c.nt, p->d_out[slot], nullptr, st);
        if (r) return r;
        p->launches += 2;
        if (c.nt == T) {
            return check_cuda(cudaMemcpyAsync(out_host + lat_off, p->d_out[slot], (size_t)frames * D * sizeof(float),
                                              cudaMemcpyDeviceToHost, st), "D2H latents");
        }
        return check_cuda(cudaMemcpy2DAsync(out_host + lat_off, (size_t)T * sizeof(float), p->d_out[slot],
                                            (size_t)c.nt * sizeof(float), (size_t)c.nt * sizeof(float), D,
                                            cudaMemcpyDeviceToHost, st), "D2H latents (2D)");
    });
    int rs = sync_all(p);
    return rc ? rc : rs;
}

int acq_vq_decode_host(acq_pipeline* p, const int64_t* codes_host, int64_t stride_table,
                       int64_t stride_frame, const float* const* cb, int S, int G, int K, int D,
                       int B, int T, float* out_host) {
    if (!p) return fail(ACQ_EINVAL, "null pipeline");
    if (!cb || S < 1 || G < 1 || S * G > ACQ_MAX_TABLE || K < 1 || D < 1 || D % G != 0 || B < 0 || T < 0)
        return fail(ACQ_EINVAL, "acq_vq_decode_host: bad arguments");
    const int tables = S * G;
    const long long N = (long long)B * T;
    if (N == 0) return 0;
    if (!codes_host || !out_host) return fail(ACQ_EINVAL, "acq_vq_decode_host: null pointer");
    const bool planar = (stride_frame == 1);                       // [tables][N]
    const bool interleaved = (stride_table == 1 && stride_frame == tables);   // [N][tables]
    if (!planar && !interleaved)
        return fail(ACQ_EINVAL, "acq_vq_decode_host: unsupported code strides (%lld, %lld)",
                    (long long)stride_table, (long long)stride_frame);
    if ((size_t)D * 64 * sizeof(float) > p->chunk_bytes)
        return fail(ACQ_ESHAPE, "chunk_bytes too small for D=%d", D);
    int rc = check_cuda(cudaSetDevice(p->device), "cudaSetDevice");
    if (rc) return rc;
    p->launches = 0;
    rc = for_each_chunk(p->chunk_bytes, D, B, T, [&](int idx, Chunk c) -> int {
        const int slot = idx % acq_pipeline::NBUF;
        cudaStream_t st = p->stream[slot];
        const long long frames = (long long)c.nb * c.nt;
        const long long n_first = (long long)c.b0 * T + c.t0;
        int r = ensure_codes(p, slot, (size_t)tables * frames * sizeof(int64_t));
        if (r) return r;
        if (planar) {
            r = check_cuda(cudaMemcpy2DAsync(p->d_codes[slot], (size_t)frames * sizeof(int64_t),
                                             codes_host + n_first, (size_t)stride_table * sizeof(int64_t),
                                             (size_t)frames * sizeof(int64_t), tables,
                                             cudaMemcpyHostToDevice, st), "H2D codes");
        } else {
            r = check_cuda(cudaMemcpyAsync(p->d_codes[slot], codes_host + n_first * tables,
                                           (size_t)frames * tables * sizeof(int64_t),
                                           cudaMemcpyHostToDevice, st), "H2D codes");
        }
        if (r) return r;
        r = vq_decode(p->d_codes[slot], planar ? frames : 1, planar ? 1 : tables, cb, S, G, K, D,
                      c.nb, c.nt, p->d_lat[slot], nullptr, st);
        if (r) return r;
        p->launches += 1;
        if (c.nt == T) {
            return check_cuda(cudaMemcpyAsync(out_host + (size_t)c.b0 * D * T, p->d_lat[slot],
                                              (size_t)frames * D * sizeof(float),
                                              cudaMemcpyDeviceToHost, st), "D2H latents");
        }
        return check_cuda(cudaMemcpy2DAsync(out_host + (size_t)c.b0 * D * T + c.t0,
                                            (size_t)T * sizeof(float), p->d_lat[slot],
                                            (size_t)c.nt * sizeof(float), (size_t)c.nt * sizeof(float),
                                            D, cudaMemcpyDeviceToHost, st), "D2H latents (2D)");
    });
    int rs = sync_all(p);
    return rc ? rc : rs;
}

}  // extern "C"
