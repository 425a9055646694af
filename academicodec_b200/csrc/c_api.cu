// extern "C" surface of libacq_b200.so: argument validation, kernel selection, error strings.
#include "acq_common.cuh"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>

#ifndef ACQ_TC_DEFAULT_CLUSTER
#define ACQ_TC_DEFAULT_CLUSTER 0
#endif
#ifndef ACQ_TC_DEFAULT_VARIANT
#define ACQ_TC_DEFAULT_VARIANT 0
#endif

namespace acq {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

// kernels (defined in the other translation units)
int rvq_search_simt(const float*, const float* const*, const float*, int, int, int, int, int, int,
                    int, int64_t*, float*, float*, double*, cudaStream_t);
int rvq_search_tc(const float*, const float* const*, const void*, void*, int, int, int, int, int, int,
                  int, int64_t*, float*, int, int, cudaStream_t);
int rvq_search_p1(const float*, const float* const*, const void*, void*, int, int, int, int, int, int,
                  int, int64_t*, float*, int, int, cudaStream_t);
bool rvq_search_tc_supported(int S, int G, int K, int D, int flags, const char** why);

// Tensor-core kernel variant (TcConfig, acq_common.cuh): defaults from the environment, read once --
//   ACQ_TC_KERNEL  = 0 automatic | 3 (three fp16 products per chunk, plain argmax) | 1 (one product + filter + re-score)
//   ACQ_TC_CLUSTER = 0 automatic | 1 | 2 | 4 CTAs sharing one multicast codebook stream
//   ACQ_TC_SPLIT   = 1 | 0 small batches get one cluster per tile
// and overridden at run time by acq_tc_configure.
TcConfig& tc_config() {
    static TcConfig cfg = [] {
        TcConfig c;
        const char* e = getenv("ACQ_TC_KERNEL");
        const int v = e ? atoi(e) : ACQ_TC_DEFAULT_VARIANT;
        c.variant = (v == 1 || v == 3) ? v : 0;
        e = getenv("ACQ_TC_CLUSTER");
        const int cl = e ? atoi(e) : ACQ_TC_DEFAULT_CLUSTER;
        c.cluster = (cl == 1 || cl == 2 || cl == 4) ? cl : 0;
        e = getenv("ACQ_TC_SPLIT");
        c.split = e ? (atoi(e) != 0) : 1;
        return c;
    }();
    return cfg;
}
static int run_tc(const float* x, const float* const* cb, const void* pack, void* ws, int S, int G, int K,
                  int D, int B, int T, int flags, int64_t* codes, float* dbg, cudaStream_t st) {
    // small batches (up to 74 tiles with ACQ_TC_SPLIT) use the three-product kernel's cluster-split mode, which
    // shortens the serial chain of stages on one SM
    const long long tiles = ((long long)B * T + 127) / 128;
    const int NP = K / 256, Dg = D / G;
    const bool small = tc_config().split && !dbg &&
                       ((NP % 4 == 0 && tiles * 4 <= kNumSMs) || (NP % 2 == 0 && tiles * 2 <= kNumSMs));
    // Automatic choice for batches that fill the chip, from the B200 sweeps (profiles/r03*_search_sweep.log):
    // the single-product filter + exact re-score kernel wins where a pass is long enough to hide its epilogue
    // and job hand-offs -- D_g = 512 (cfg2 0.81-0.87 vs 1.00 ms, cfg4 64 x 10 s 2.19 vs 2.40 ms), in a 2-CTA
    // cluster sharing one multicast codebook stream (fewer L2 reads = less power = higher clock); the
    // three-product kernel stays faster at D_g <= 256 (cfg1 B=4096 2.7 vs 4.3 ms, cfg3 1.34 vs 1.53 ms).
    int variant = tc_config().variant, cluster = tc_config().cluster;
    const bool automatic = variant == 0;
    if (variant == 0) variant = (Dg >= 512 && !small) ? 1 : 3;
    const int cl1 = cluster ? cluster : 2, cl3 = cluster ? cluster : 1;
    if (variant == 1 && !small) {
        // The filter's bound degrades on tables whose codewords differ widely in norm (tc_common.cuh,
        // tables_fit_single_product); the verdict is in the pack, on the device.  In automatic mode both
        // kernels are launched and each checks it: the one whose turn it is not returns at once (~3 us).
        int rc = rvq_search_p1(x, cb, pack, ws, S, G, K, D, B, T, flags, codes, dbg, cl1, automatic ? 1 : 0, st);
        if (!rc && automatic)
            rc = rvq_search_tc(x, cb, pack, ws, S, G, K, D, B, T, flags, codes, dbg, cl3, 2, st);
        return rc;
    }
    return rvq_search_tc(x, cb, pack, ws, S, G, K, D, B, T, flags, codes, dbg, cl3, 0, st);
}
size_t tc_pack_bytes(int, int, int);
size_t tc_workspace_bytes(int);
int tc_pack_codebooks(const float* const*, int, int, int, void*, cudaStream_t);
int codebook_half_norms(const float* const*, int, int, int, float*, cudaStream_t);
int vq_decode(const int64_t*, int64_t, int64_t, const float* const*, int, int, int, int, int, int,
              float*, int*, cudaStream_t);
int ema_stats(const float*, const int64_t*, const float* const*, int, int, int, int, int, int,
              float*, cudaStream_t);
int rvq_replay(const float*, const int64_t*, const float* const*, int, int, int, int, int, int, int,
               float*, float*, double*, float*, cudaStream_t);
int pack_bits(const int64_t*, long long, int, uint8_t*, int*, cudaStream_t);
int unpack_bits(const uint8_t*, long long, int, int64_t*, cudaStream_t);
int ema_apply(float*, float* const*, float* const*, float* const*, int, int, int, double, double,
              cudaStream_t);
int peer_allreduce(float*, float* const*, int, int, size_t, cudaStream_t);
int grvq_backward(const float*, const int64_t*, const float* const*, int, int, int, int, int, int, const float*,
                  const float*, double, double, float*, float* const*, cudaStream_t);

int validate_search(const float* x, const float* const* cb, const float* hn, int S, int G, int K,
                    int D, int B, int T, const int64_t* codes) {
    (void)hn;
    if (!cb) return fail(ACQ_EINVAL, "null pointer argument");
    if ((long long)B * T > 0 && !codes) return fail(ACQ_EINVAL, "null pointer argument");
    if (S < 1 || G < 1 || S * G > ACQ_MAX_TABLE)
        return fail(ACQ_EINVAL, "stages*groups=%d outside [1, %d]", S * G, ACQ_MAX_TABLE);
    if (K < 1 || D < 1 || D % G != 0) return fail(ACQ_EINVAL, "bad K=%d D=%d G=%d", K, D, G);
    if (B < 0 || T < 0) return fail(ACQ_EINVAL, "negative batch/frames");
    if ((long long)B * T > 0 && !x) return fail(ACQ_EINVAL, "null latent pointer");
    for (int i = 0; i < S * G; ++i)
        if (!cb[i]) return fail(ACQ_EINVAL, "null codebook pointer %d", i);
    return 0;
}

int rvq_search_dispatch(const float* x, const float* const* cb, const float* hn,
                        const void* tc_pack, void* workspace, int S, int G, int K, int D, int B,
                        int T, int flags, int impl, int64_t* codes, float* quantized,
                        float* residual, double* sqerr, cudaStream_t st) {
    if ((long long)B * T == 0) return 0;
    if (impl == ACQ_IMPL_TC || impl == ACQ_IMPL_AUTO) {
        const char* why = "";
        bool ok = rvq_search_tc_supported(S, G, K, D, flags, &why);
        if (ok && (!tc_pack || !workspace)) { ok = false; why = "tc_pack / workspace not provided"; }
        const bool outputs = quantized || residual || sqerr;
        if (ok && outputs && impl == ACQ_IMPL_TC) {
            ok = false; why = "the tensor-core kernel writes codes only";
        }
        // (no minimum batch: measured with scripts/small_batch_probe.py the tensor-core kernel is 3-4x
        // faster than the SIMT kernel even for 4 frames -- both are then bound by the serial chain of
        // S stages on one SM, and a stage is shorter on the tensor pipe)
        if (ok) {
            int rc = run_tc(x, cb, tc_pack, workspace, S, G, K, D, B, T, flags, codes, nullptr, st);
            // AUTO with outputs: codes from the tensor cores, then one replay pass for quantized /
            // residual / sqerr (same arithmetic as the fused SIMT kernel, an order of magnitude faster)
            if (!rc && outputs)
                rc = rvq_replay(x, codes, cb, S, G, K, D, B, T, flags, quantized, residual, sqerr, nullptr, st);
            return rc;
        }
        if (impl == ACQ_IMPL_TC) return fail(ACQ_ESHAPE, "tensor-core search unavailable: %s", why);
    }
    if (!hn) return fail(ACQ_EINVAL, "half_norms missing for the SIMT kernel");
    return rvq_search_simt(x, cb, hn, S, G, K, D, B, T, flags, codes, quantized, residual, sqerr, st);
}

}  // namespace acq

using namespace acq;

extern "C" {

int acq_version(void) { return ACQ_VERSION; }
const char* acq_last_error(void) { return g_err; }

int acq_codebook_half_norms(const float* const* cb, int n_tables, int K, int Dg, float* out,
                            void* stream) {
    if (!cb || !out || n_tables < 1 || n_tables > ACQ_MAX_TABLE || K < 1 || Dg < 1)
        return fail(ACQ_EINVAL, "acq_codebook_half_norms: bad arguments");
    return codebook_half_norms(cb, n_tables, K, Dg, out, (cudaStream_t)stream);
}

int acq_rvq_search(const float* x, const float* const* cb, const float* half_norms,
                   const void* tc_pack, void* workspace, int S, int G, int K, int D, int B, int T,
                   int flags, int impl, int64_t* codes, float* quantized, float* residual,
                   double* sqerr, void* stream) {
    int rc = validate_search(x, cb, half_norms, S, G, K, D, B, T, codes);
    if (rc) return rc;
    return rvq_search_dispatch(x, cb, half_norms, tc_pack, workspace, S, G, K, D, B, T, flags, impl,
                               codes, quantized, residual, sqerr, (cudaStream_t)stream);
}

int acq_tc_configure(int variant, int cluster, int split) {
    TcConfig& c = tc_config();
    if (variant >= 0) {
        if (variant != 0 && variant != 1 && variant != 3) return fail(ACQ_EINVAL, "acq_tc_configure: variant must be 0, 1 or 3");
        c.variant = variant;
    }
    if (cluster >= 0) {
        if (cluster != 0 && cluster != 1 && cluster != 2 && cluster != 4)
            return fail(ACQ_EINVAL, "acq_tc_configure: cluster must be 0, 1, 2 or 4");
        c.cluster = cluster;
    }
    if (split >= 0) c.split = split != 0;
    return 0;
}

int acq_tc_query(int what) {
    const TcConfig& c = tc_config();
    return what == 0 ? c.variant : (what == 1 ? c.cluster : (what == 2 ? c.split : ACQ_EINVAL));
}

size_t acq_tc_pack_bytes(int n_tables, int K, int Dg) { return tc_pack_bytes(n_tables, K, Dg); }
size_t acq_tc_workspace_bytes(int D) { return tc_workspace_bytes(D); }

int acq_tc_pack_codebooks(const float* const* cb, int n_tables, int K, int Dg, void* pack,
                          void* stream) {
    if (!cb || !pack || n_tables < 1 || n_tables > ACQ_MAX_TABLE || K < 1 || Dg < 1)
        return fail(ACQ_EINVAL, "acq_tc_pack_codebooks: bad arguments");
    return tc_pack_codebooks(cb, n_tables, K, Dg, pack, (cudaStream_t)stream);
}

int acq_debug_tc_scores(const float* x, const float* const* cb, const void* tc_pack,
                        void* workspace, int K, int D, int B, int T, float* scores, int64_t* codes,
                        void* stream) {
    if (!x || !cb || !tc_pack || !workspace || !scores || !codes)
        return fail(ACQ_EINVAL, "acq_debug_tc_scores: null pointer");
    return run_tc(x, cb, tc_pack, workspace, 1, 1, K, D, B, T, 0, codes, scores, (cudaStream_t)stream);
}

int acq_vq_decode(const int64_t* codes, int64_t stride_table, int64_t stride_frame,
                  const float* const* cb, int S, int G, int K, int D, int B, int T, float* out,
                  int* status, void* stream) {
    if (!cb || S < 1 || G < 1 || S * G > ACQ_MAX_TABLE || K < 1 || D < 1 || D % G != 0 || B < 0 ||
        T < 0)
        return fail(ACQ_EINVAL, "acq_vq_decode: bad arguments");
    if ((long long)B * T == 0) return 0;
    if (!codes || !out) return fail(ACQ_EINVAL, "acq_vq_decode: null pointer");
    return vq_decode(codes, stride_table, stride_frame, cb, S, G, K, D, B, T, out, status,
                     (cudaStream_t)stream);
}

int acq_ema_stats(const float* x, const int64_t* codes, const float* const* cb, int S, int K, int D,
                  int B, int T, int flags, float* stats, void* stream) {
    if (!cb || !stats || S < 1 || S > ACQ_MAX_TABLE || K < 1 || D < 1 || B < 0 || T < 0)
        return fail(ACQ_EINVAL, "acq_ema_stats: bad arguments");
    if ((long long)B * T == 0) return 0;
    if (!x || !codes) return fail(ACQ_EINVAL, "acq_ema_stats: null pointer");
    return ema_stats(x, codes, cb, S, K, D, B, T, flags, stats, (cudaStream_t)stream);
}

int acq_grvq_backward(const float* x, const int64_t* codes, const float* const* cb, int S, int G, int K, int D,
                      int B, int T, const float* g_quantized, const float* g_losses, double lam_cb,
                      double lam_commit, float* grad_x, float* const* grad_cb, void* stream) {
    if (!cb || S < 1 || G < 1 || S * G > ACQ_MAX_TABLE || K < 1 || D < 1 || D % G != 0 || B < 0 || T < 0)
        return fail(ACQ_EINVAL, "acq_grvq_backward: bad arguments");
    if ((long long)B * T == 0) return 0;
    if (!x || !codes) return fail(ACQ_EINVAL, "acq_grvq_backward: null pointer");
    for (int i = 0; i < S * G; ++i)
        if (!cb[i]) return fail(ACQ_EINVAL, "acq_grvq_backward: null codebook pointer %d", i);
    return grvq_backward(x, codes, cb, S, G, K, D, B, T, g_quantized, g_losses, lam_cb, lam_commit, grad_x, grad_cb,
                         (cudaStream_t)stream);
}

int acq_rvq_replay(const float* x, const int64_t* codes, const float* const* cb, int S, int G, int K,
                   int D, int B, int T, int flags, float* quantized, float* residual, double* sqerr,
                   float* stats, void* stream) {
    if (!cb || S < 1 || G < 1 || S * G > ACQ_MAX_TABLE || K < 1 || D < 1 || D % G != 0 || B < 0 || T < 0)
        return fail(ACQ_EINVAL, "acq_rvq_replay: bad arguments");
    if ((long long)B * T == 0) return 0;
    if (!x || !codes) return fail(ACQ_EINVAL, "acq_rvq_replay: null pointer");
    return rvq_replay(x, codes, cb, S, G, K, D, B, T, flags, quantized, residual, sqerr, stats,
                      (cudaStream_t)stream);
}

int64_t acq_packed_bytes(int64_t n, int bits) { return (n * bits + 7) / 8; }

int acq_pack_codes(const int64_t* values, int64_t n, int bits, uint8_t* packed, int* status, void* stream) {
    if (n < 0 || bits < 1 || bits > 16) return fail(ACQ_EINVAL, "acq_pack_codes: need n >= 0 and 1 <= bits <= 16");
    if (n == 0) return 0;
    if (!values || !packed) return fail(ACQ_EINVAL, "acq_pack_codes: null pointer");
    return pack_bits(values, n, bits, packed, status, (cudaStream_t)stream);
}

int acq_unpack_codes(const uint8_t* packed, int64_t n, int bits, int64_t* values, void* stream) {
    if (n < 0 || bits < 1 || bits > 16) return fail(ACQ_EINVAL, "acq_unpack_codes: need n >= 0 and 1 <= bits <= 16");
    if (n == 0) return 0;
    if (!values || !packed) return fail(ACQ_EINVAL, "acq_unpack_codes: null pointer");
    return unpack_bits(packed, n, bits, values, (cudaStream_t)stream);
}

int acq_ema_apply(float* stats, float* const* embed, float* const* embed_avg,
                  float* const* cluster_size, int S, int K, int D, double decay, double epsilon,
                  void* stream) {
    if (!stats || !embed || !embed_avg || !cluster_size || S < 1 || S > ACQ_MAX_TABLE || K < 1 ||
        D < 1)
        return fail(ACQ_EINVAL, "acq_ema_apply: bad arguments");
    return ema_apply(stats, embed, embed_avg, cluster_size, S, K, D, decay, epsilon,
                     (cudaStream_t)stream);
}

int acq_peer_allreduce(float* multicast, float* const* peers, int world, int rank, size_t n, void* stream) {
    return peer_allreduce(multicast, peers, world, rank, n, (cudaStream_t)stream);
}

}  // extern "C"
