/* acq_b200.h -- C ABI of the B200-native RVQ / GRVQ quantize-codec path.
 *
 * Drop-in boundary.  The reference (jacquelm/AcademiCodec) has no FFI
 * layer: its boundary is the Python nn.Module surface.  These entry points are what a binding
 * for that surface calls; each one names the reference code it replaces (path:line relative
 * to the reference checkout).  The Python shim in academicodec_b200/ binds them with ctypes.
 *
 * Conventions
 *   - every pointer named *_dev / x / codes / out is a DEVICE pointer unless the function name
 *     ends in _host; `cb`, `embed` ... tables are HOST arrays of device pointers;
 *   - latents are [B, D, T] fp32 with T contiguous (what SEANet / the HiFi encoder emit,
 *     reference net3.py:39-43, hificodec/train.py:214-216); codes are int64;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - return value 0 = success, otherwise a negative ACQ_E* code or a positive cudaError_t;
 *     acq_last_error() gives the message of the calling thread's last failure;
 *   - no function allocates device memory except the acq_pipeline_* family.
 */
#ifndef ACQ_B200_H
#define ACQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACQ_VERSION 100            /* 0.1.0 */
#define ACQ_MAX_TABLE 64           /* max stages * groups per call */

/* error codes */
#define ACQ_EINVAL   (-1)          /* bad argument (shape, alignment, null pointer) */
#define ACQ_ESHAPE   (-2)          /* shape not supported by the selected kernel */

/* flags for acq_rvq_search */
#define ACQ_STE        1           /* straight-through arithmetic: q' = r + (q - r); r -= q'
                                      (core_vq.py:304 in training, hificodec/models.py:478,490 always) */
#define ACQ_LOSS_RAW   2           /* sqerr accumulates (q - r)^2 (hificodec/models.py:476-477);
                                      default accumulates (q' - r)^2 (core_vq.py:310) */
/* kernel selection for acq_rvq_search */
#define ACQ_IMPL_AUTO  0
#define ACQ_IMPL_SIMT  1           /* fp32 CUDA-core kernel (any shape)            */
#define ACQ_IMPL_TC    2           /* tcgen05 tensor-core kernel   */

int acq_version(void);
const char* acq_last_error(void);

/* 0.5*||e_k||^2 per codeword, accumulated in fp64, rounded once to fp32 (the search
 * maximises x.e_k - 0.5||e_k||^2, which orders codewords exactly as the squared distance does).
 * Replaces the per-call `embed.pow(2).sum(0)` of core_vq.py:178 and
 * `torch.sum(weight**2, 1)` of hificodec/models.py:438.  Must be re-run when a codebook
 * changes (EMA step, load_state_dict).
 *   cb[i]  -> [K, Dg] fp32 row-major, i in [0, n_tables)
 *   out    -> [n_tables, K] fp32                                                    */
int acq_codebook_half_norms(const float* const* cb, int n_tables, int K, int Dg,
                            float* out, void* stream);

/* Fused residual nearest-codeword search over S stages x G channel groups.
 * Replaces ResidualVectorQuantization.encode / .forward (core_vq.py:328-362) with
 * EuclideanCodebook.quantize / dequantize inside (core_vq.py:175-187), and
 * Quantizer.forward / for_one_step / Quantizer_module.forward (hificodec/models.py:436-508).
 *   x          [B, D, T]           latents
 *   cb[s*G+g]  [K, D/G]            codebook of stage s, group g (group g = channels [g*D/G, (g+1)*D/G))
 *   half_norms [S*G, K]            from acq_codebook_half_norms
 *   codes      [S*G, B*T] int64    codes[(s*G+g)*B*T + b*T + t]   (RVQ: [S,B,T]; GRVQ: the list
 *                                  idx_s0g0, idx_s0g1, .., idx_s1g0, .. of models.py:504)
 *   quantized  [B, D, T] or NULL   sum over stages of the (straight-through) quantized latent,
 *                                  accumulated left to right from 0.0 (core_vq.py:340)
 *   residual   [B, D, T] or NULL   residual after the last stage
 *   sqerr      [S] fp64 or NULL    += sum over elements of the squared quantization error per
 *                                  stage (caller zeroes; divide by B*D*T for the mse)
 *   Tie rule: lowest index among equal distances (torch max/argmin).                */
int acq_rvq_search(const float* x, const float* const* cb, const float* half_norms,
                   const void* tc_pack, void* workspace,
                   int S, int G, int K, int D, int B, int T, int flags, int impl,
                   int64_t* codes, float* quantized, float* residual, double* sqerr,
                   void* stream);

/* Tensor-core operands (ACQ_IMPL_TC).  The tcgen05 kernel consumes the codebooks as pre-scaled,
 * fp16 hi/lo split, SWIZZLE_64B K-major shared-memory images that TMA bulk copies stream
 * straight into the MMA ring; acq_tc_pack_codebooks builds them (plus the scaled half norms)
 * for `n_tables` = S*G codebooks into `pack` (acq_tc_pack_bytes bytes, 256 B aligned).  Like
 * the half norms it must be re-run when a codebook changes.  `workspace` is per-call scratch of
 * acq_tc_workspace_bytes(D) bytes (residual rows of the tiles in flight, L2 resident); calls
 * that may run concurrently need distinct workspaces.  With tc_pack == NULL or workspace == NULL
 * acq_rvq_search uses the SIMT kernel.  The tensor-core kernel writes codes only: under
 * ACQ_IMPL_AUTO a call that also asks for quantized / residual / sqerr runs the tensor-core search
 * followed by the replay pass (acq_rvq_replay) on the same stream; ACQ_IMPL_TC rejects it.  Under
 * ACQ_IMPL_AUTO the tensor-core kernel is used for every batch size: it is 3-4x faster than
 * the SIMT kernel even for a handful of frames.
 * Requirements: K % 256 == 0, K <= 1024, (D/G) % 64 == 0, D/G <= 512, G <= 4.          */
size_t acq_tc_pack_bytes(int n_tables, int K, int Dg);
size_t acq_tc_workspace_bytes(int D);
int acq_tc_pack_codebooks(const float* const* cb, int n_tables, int K, int Dg, void* pack,
                          void* stream);
/* Test hook: single codebook search that also dumps cs*(x.e_k - 0.5||e_k||^2) for every
 * frame and codeword ([B*T, K] fp32; cs = the pack's power-of-two codebook scale).    */
int acq_debug_tc_scores(const float* x, const float* const* cb, const void* tc_pack,
                        void* workspace, int K, int D, int B, int T, float* scores,
                        int64_t* codes, void* stream);
/* Process-wide choice of the tensor-core search variant (defaults: environment ACQ_TC_KERNEL,
 * ACQ_TC_CLUSTER, ACQ_TC_SPLIT).  A negative argument leaves that setting unchanged.
 *   variant  0 = automatic (by shape: single product for D/G >= 512, three products otherwise),
 *            1 = one fp16 product + rigorous filter + exact re-score, 3 = three-product fp16 split
 *   cluster  0 = automatic | 1 | 2 | 4 CTAs share one multicast codebook stream
 *   split    0 | 1   small batches: one cluster per tile, codebook passes split across its CTAs */
int acq_tc_configure(int variant, int cluster, int split);
/* Current setting: what = 0 variant, 1 cluster, 2 split. */
int acq_tc_query(int what);

/* Codebook gather-accumulate.  Replaces ResidualVectorQuantization.decode
 * (core_vq.py:364-370, F.embedding + 'b n d -> b d n' per stage) and Quantizer.embed
 * (hificodec/models.py:510-535).
 *   code of (table i = s*G+g, frame n = b*T+t) is codes[i*stride_table + n*stride_frame]
 *     RVQ  [S,B,T]  : stride_table = B*T, stride_frame = 1
 *     GRVQ [B,T,2G] : stride_table = 1,   stride_frame = 2G
 *   out    [B, D, T] = 0.0 + stage0 + stage1 + ... (fp32, left to right)
 *   status [1] int32 device or NULL: set to 1 if any code is outside [0, K) (such codes
 *          contribute zeros; the reference raises IndexError from F.embedding)      */
int acq_vq_decode(const int64_t* codes, int64_t stride_table, int64_t stride_frame,
                  const float* const* cb, int S, int G, int K, int D, int B, int T,
                  float* out, int* status, void* stream);

/* EMA k-means statistics (training).  Replaces F.one_hot + onehot.sum(0) + x.t() @ onehot
 * (core_vq.py:210,218-219) for every stage of the residual stack at once; the residual fed
 * to stage s is recomputed from x and the codes exactly as the forward pass does.
 *   stats  [S*K*D sums][S*K counts] fp32, caller zeroes; sums[s][k][:] += r_s[n] and
 *          counts[s][k] += 1 for every frame n with codes[s][n] == k
 *   The buffer is flat so that one NCCL all-reduce(SUM) covers all stages.           */
int acq_ema_stats(const float* x, const int64_t* codes, const float* const* cb,
                  int S, int K, int D, int B, int T, int flags, float* stats, void* stream);

/* Replay of a code sequence: everything forward() returns besides the codes, from x and the
 * codes (same arithmetic and order as acq_rvq_search, which it complements when the search ran
 * on the codes-only tensor-core kernel).  Any of quantized / residual / sqerr / stats may be
 * NULL; stats (layout of acq_ema_stats, caller zeroes) requires G == 1.              */
int acq_rvq_replay(const float* x, const int64_t* codes, const float* const* cb,
                   int S, int G, int K, int D, int B, int T, int flags,
                   float* quantized, float* residual, double* sqerr, float* stats, void* stream);

/* Backward of the group-residual VQ training forward (HiFi-Codec Quantizer.forward / for_one_step,
 * hificodec/models.py:463-508: what autograd derives from the per-(stage, group) losses
 * lam_cb * mse(z_q, x.detach()) + lam_commit * mse(z_q.detach(), x) and the straight-through sum).
 * One kernel recomputes the residual chain from x and the codes and writes
 *   grad_x   [B, D, T] = g_quantized + (-2 lam_commit / numel) * g_losses[0] * (z_q0 - x)      (or NULL)
 *   grad_cb  HOST array of S*G device pointers [K, D/G], caller zeroes (entries / the array may be NULL):
 *            row code += (2 lam_cb / numel) * g_losses[s] * (z_q - r_s)[group]   for every frame
 * g_quantized [B, D, T] and g_losses [S] (device) may be NULL (no gradient from that output).
 * Requirements: D % 32 == 0, D <= 768.                                                           */
int acq_grvq_backward(const float* x, const int64_t* codes, const float* const* cb, int S, int G, int K, int D,
                      int B, int T, const float* g_quantized, const float* g_losses, double lam_cb,
                      double lam_commit, float* grad_x, float* const* grad_cb, void* stream);

/* EMA apply.  Replaces ema_inplace x2, laplace_smoothing and embed.copy_
 * (core_vq.py:47-52,218-225).  In place on the module buffers; consumes `stats`
 * (the counts part is overwritten with the smoothed cluster sizes).                 */
int acq_ema_apply(float* stats, float* const* embed, float* const* embed_avg,
                  float* const* cluster_size, int S, int K, int D, double decay, double epsilon,
                  void* stream);

/* EMA statistics exchange over NVLink peer memory (multi-GPU training; the only collective of the path).
 * The reference has none: every rank updates from its local batch and DDP re-broadcasts rank 0's buffers
 * (core_vq.py:214-225, main_launch.py:199-204).  In-place two-shot all-reduce(SUM) of `n` floats (n % 4 == 0)
 * that live in symmetric memory: rank r reduces elements [r*n/W, (r+1)*n/W) and writes the sums to every rank.
 *   multicast  NVLS multicast address of the buffer (multimem.ld_reduce / multimem.st through the switch), or NULL
 *   peers      HOST array of `world` device pointers: rank q's copy mapped into this process (P2P path when
 *              multicast is NULL; may be NULL otherwise)
 * The caller brackets the call with cross-rank barriers on the same stream (all ranks' statistics written
 * before, all slices stored after).  Every rank receives bit-identical sums.                                */
int acq_peer_allreduce(float* multicast, float* const* peers, int world, int rank, size_t n, void* stream);

/* Code wire format: the reference's BitPacker / BitUnpacker (academicodec/binary.py:54-123).
 * Value i occupies bits [i*bits, (i+1)*bits) of a little-endian bit stream; 1 <= bits <= 16.
 *   values [n] int64 device  <->  packed [acq_packed_bytes(n, bits)] uint8 device
 *   status [1] int32 device or NULL: set to 1 if a value does not fit in `bits` bits.   */
int64_t acq_packed_bytes(int64_t n, int bits);
int acq_pack_codes(const int64_t* values, int64_t n, int bits, uint8_t* packed, int* status,
                   void* stream);
int acq_unpack_codes(const uint8_t* packed, int64_t n, int bits, int64_t* values, void* stream);

/* ---- host-buffer pipeline (the end-to-end path: H2D, kernels, D2H overlapped in chunks) ---- */
typedef struct acq_pipeline acq_pipeline;

/* device = CUDA ordinal; chunk_bytes = staging size per in-flight chunk (0 = 128 MiB).
 * Clips that fit a chunk travel as contiguous 1-D copies (48-50 GB/s per direction with both directions
 * busy on PCIe 5 x16); longer clips are cut along T into 2-D copies, which reach ~39 GB/s. */
int acq_pipeline_create(acq_pipeline** out, int device, size_t chunk_bytes);
void acq_pipeline_destroy(acq_pipeline* p);
/* The pipeline's streams do not order themselves after the caller's: call this with the stream that
 * produced the tables (codebooks, half norms, tensor-core pack, EMA updates) before a *_host call. */
int acq_pipeline_wait_stream(acq_pipeline* p, void* producer_stream);

/* Same operations as above on HOST buffers (pinned memory makes the copies asynchronous).
 * The codebook tables are still device pointers (codebooks live on the GPU).        */
int acq_rvq_encode_host(acq_pipeline* p, const float* x_host, const float* const* cb,
                        const float* half_norms, const void* tc_pack, int S, int G, int K, int D,
                        int B, int T, int flags, int impl, int64_t* codes_host);
int acq_vq_decode_host(acq_pipeline* p, const int64_t* codes_host, int64_t stride_table,
                       int64_t stride_frame, const float* const* cb, int S, int G, int K, int D,
                       int B, int T, float* out_host);
/* Encode followed by decode of the same batch (what the reference's inference CLI does on the
 * GPU, models/encodec/test.py:116-119: codes never leave the device between the two): per
 * chunk H2D latents -> search -> D2H codes, decode -> D2H reconstructed latents; with the
 * chunk ring the upload of chunk i+1 overlaps the download of chunk i (full-duplex PCIe).
 *   codes_host [S*G, B*T] int64, out_host [B, D, T] fp32                            */
int acq_rvq_codec_host(acq_pipeline* p, const float* x_host, const float* const* cb,
                       const float* half_norms, const void* tc_pack, int S, int G, int K, int D,
                       int B, int T, int flags, int impl, int64_t* codes_host, float* out_host);
/* kernels launched by the last *_host call (for bench.py's gpu_launches)             */
int acq_pipeline_last_launches(const acq_pipeline* p);

#ifdef __cplusplus
}
#endif
#endif /* ACQ_B200_H */
