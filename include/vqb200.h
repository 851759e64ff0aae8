/* vqb200.h -- C ABI of the B200-native vector-quantisation bottleneck (libvqb200.so, sm_100a only).
 *
 * Drop-in boundary for ONE path of vliu15/speech-masters-thesis: models/vqvae/bottleneck.py.
 * The reference has no FFI layer (it is pure PyTorch); every entry point below cites the reference
 * lines whose work it replaces.  The Python module speech-masters-thesis_b200/quantizer.py (via _lib.py) binds these
 * with ctypes (tensor.data_ptr(), torch.cuda.current_stream().cuda_stream) and keeps the reference's
 * nn.Module surface; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; vq_last_error() gives the message
 *     (thread-local, valid until the next failing call on that thread).  No exception crosses the ABI.
 *   - all pointers are DEVICE pointers unless the name ends in _host; nothing is allocated or freed
 *     behind the caller's back except inside a vq_host_ctx.
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work, they never synchronise
 *     (the *_host entry points are the exception: they return when the host buffers are filled).
 *   - latents x / x_q / grad tensors are [N, D, T] fp32, T contiguous ("NCT", the layout the reference's
 *     encoder emits, bottleneck.py:92-95); indices are [N, T] int64; mask is [N, 1, T] fp32 (== [N*T]);
 *     the codebook is [K, D] fp32 row-major (buffer `k`, bottleneck.py:24).
 *   - there is no CPU fallback.  On a machine without an sm_100 GPU the compute calls fail with an error.
 */
#ifndef VQB200_H
#define VQB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQB200_VERSION 100

/* ---- algorithm selector for vq_assign --------------------------------------------------------- */
enum {
    VQ_ALGO_AUTO = 0,   /* tcgen05 path when the shape allows it, else the SIMT path                   */
    VQ_ALGO_SIMT = 1,   /* exact-FP32 register-tiled CUDA-core kernel (any shape)                      */
    VQ_ALGO_TC   = 2,   /* FP16 tcgen05/TMEM distance GEMM + (best, runner-up) scan + provable check (+ exact fallback) */
    /* OR-able flag: the workspace already holds this codebook's prepared operands (FP16 image, norms) from an earlier
     * vq_assign on the SAME workspace with the SAME k / k_bins / emb_width and unchanged codebook values: skip the
     * per-call preparation.  What the generate_vq_dataset loop wants (frozen codebook, many batches). */
    VQ_ALGO_PREPARED = 256
};

/* ---- scalar slots (fp64 accumulators in caller-owned device memory, VQ_NUM_SCALARS doubles) --- */
enum {
    VQ_S_SUM_MIN_D   = 0,   /* sum over ALL rows of the winning distance  (fit * K, bottleneck.py:140)  */
    VQ_S_COMMIT_SQ   = 1,   /* sum over valid rows of ||e - x||^2         (bottleneck.py:194)           */
    VQ_S_MASK_SUM    = 2,   /* sum of the mask                            (bottleneck.py:194)           */
    VQ_S_UNSAFE_ROWS = 3,   /* rows the tcgen05 path handed to the exact fallback (diagnostic)          */
    VQ_S_COUNT_TOTAL = 4,   /* sum of the (all-reduced) per-code counts   (bottleneck.py:85)            */
    VQ_S_ENTROPY     = 5,
    VQ_S_USED_CURR   = 6,
    VQ_S_USAGE       = 7,
    VQ_S_DK_SQ       = 8,
    VQ_S_TICKET      = 9,   /* last-block tickets */
    VQ_S_ELEM_TOTAL  = 10,  /* sum of the updated cluster sizes k_elem (the n of the optional Laplace smoothing) */
    VQ_NUM_SCALARS   = 16
};

/* ---- float outputs written by the finishing blocks (VQ_NUM_RESULTS floats) -------------------- */
enum {
    VQ_R_FIT = 0,        /* bottleneck.py:140 */
    VQ_R_COMMIT = 1,     /* bottleneck.py:194 */
    VQ_R_ENTROPY = 2,    /* bottleneck.py:86  */
    VQ_R_USAGE = 3,      /* bottleneck.py:88  */
    VQ_R_DK = 4,         /* bottleneck.py:89  */
    VQ_R_USED_CURR = 5,  /* bottleneck.py:87 (exact small integer stored as float; also as int64, see finalize) */
    VQ_NUM_RESULTS = 8
};

int         vq_version(void);
const char* vq_last_error(void);
/* 1 when the current device can run the tcgen05 kernels (compute capability 10.x), else 0. */
int         vq_device_supported(void);

/* Bytes of scratch vq_assign needs for this shape (FP16 codebook image, norms, fallback worklist). */
size_t vq_workspace_bytes(int64_t n_utt, int64_t t_frames, int k_bins, int emb_width);

/* K1 -- replaces BottleneckBlock.preprocess + quantize (bottleneck.py:92-100,126-141) and, for the
 * generate script, BottleneckBlock.encode (bottleneck.py:147-158; caller scripts/generate_vq_dataset.py:69).
 *   x [N,D,T] fp32, k [K,D] fp32  ->  idx [N*T] int64 (argmin, lowest index on ties),
 *   min_d [N*T] fp32 or NULL, scalars[VQ_S_SUM_MIN_D] += sum(min_d) (scalars may be NULL).
 * The tcgen05 path uses FP16 operands only to SHORTLIST (best, runner-up) per frame; a frame keeps the shortlisted code
 * only when its margin exceeds a rigorous bound on the FP16 error, every other frame is re-scanned by the exact FP32
 * kernel (lowest index on ties, like torch.min).  min_d / sum(min_d), when requested, are ||x||^2 - 2 x.e + ||e||^2
 * evaluated in FP32 for the winner exactly as the reference expression does. */
int vq_assign(const float* x, int64_t n_utt, int64_t emb_width, int64_t t_frames,
              const float* k, int k_bins,
              int64_t* idx, float* min_d, double* scalars,
              void* workspace, size_t workspace_bytes, int algo, void* stream);

/* vq_assign for BF16 latents (bf16 autocast training / inference: the encoder hands over [N,D,T] bfloat16, T contiguous).
 * Same outputs and the same exactness contract as vq_assign on x.float() -- BF16 -> FP16 is exact in the normal range, and
 * any frame that is not provably safe is re-scanned from the BF16 values in FP32 -- without the extra pass that an up-cast
 * would cost (6D bytes per frame).  The tcgen05 path needs T % 8 == 0 and a 16-byte aligned tensor; other shapes take the
 * exact CUDA-core kernel. */
int vq_assign_bf16(const void* x_bf16, int64_t n_utt, int64_t emb_width, int64_t t_frames,
                   const float* k, int k_bins,
                   int64_t* idx, float* min_d, double* scalars,
                   void* workspace, size_t workspace_bytes, int algo, void* stream);

/* Grouped (phoneme-conditioned) K1 -- replaces the per-frame codebook gather + bmm + min of the TTS quantiser
 * (models/vqtts/bottleneck.py:38-58): the codebook k is [n_vocab * l_bins, D]; frame j only competes among the l_bins
 * codes of its token tok[j] (int64 [N*T], the aligned token ids of :28; values are clamped to [0, n_vocab)).
 *   q_rel [N*T] int64: index inside the group (what the reference returns, :52);  q_abs = tok * l_bins + q_rel (:58), the
 *   index dequantize / update_k use with the other entry points;  min_d [N*T] or NULL;  scalars[VQ_S_SUM_MIN_D] += sum(min_d).
 * Exact FP32 on the CUDA cores (one warp per frame walks its token's code rows in place); same workspace as vq_assign,
 * sized by vq_workspace_bytes(N, T, n_vocab * l_bins, D). */
int vq_assign_grouped(const float* x, int64_t n_utt, int64_t emb_width, int64_t t_frames,
                      const float* k, int n_vocab, int l_bins, const int64_t* tok,
                      int64_t* q_rel, int64_t* q_abs, float* min_d, double* scalars,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Audit variant of vq_assign (tcgen05 path only): additionally writes, per frame, the four floats
 * {approximate best score, bound on every other code's approximate score, exact FP32 score of the shortlisted
 * code, rigorous FP16 error bound} into shortlist4 [N*T*4] (16-byte aligned), so tests can check that the
 * bound really dominates the observed FP16 error and count how many frames took the exact fallback
 * (scalars[VQ_S_UNSAFE_ROWS]).  trace (may be NULL) receives a clock64 timeline of the first thread block,
 * [10 events][trace_tiles], used by tools/tc_timeline.py to see which pipeline stage a tile waits on;
 * with shortlist4 == NULL and trace != NULL the production (non-rescoring) kernel variant is traced. */
int vq_assign_debug(const float* x, int64_t n_utt, int64_t emb_width, int64_t t_frames,
                    const float* k, int k_bins, int64_t* idx, float* shortlist4, double* scalars,
                    void* workspace, size_t workspace_bytes, void* stream,
                    int64_t* trace, int trace_tiles);

/* K2 forward -- replaces dequantize + commit loss + straight-through + postprocess + mask multiply
 * (bottleneck.py:143-145,194,197,118-124,201).
 *   x_q[n,d,t] = (x + (k[idx] - x)) * mask      (same two FP32 roundings as the reference expression)
 *   scalars[VQ_S_COMMIT_SQ] += sum_{mask!=0} ||k[idx]-x||^2 ; scalars[VQ_S_MASK_SUM] += sum(mask)
 *   scalars[VQ_S_SUM_MIN_D] += sum over ALL frames of ||k[idx]-x||^2 (the winning distance in its direct form;
 *   so call vq_assign with scalars == NULL when this kernel follows, or the numerator is counted twice)
 *   the last block writes results[VQ_R_COMMIT] = COMMIT_SQ / (MASK_SUM * D) and
 *   results[VQ_R_FIT] = SUM_MIN_D / K.  mask may be NULL (all ones). */
int vq_gather_st_fwd(const float* x, const int64_t* idx, const float* mask, const float* k,
                     int64_t n_utt, int64_t emb_width, int64_t t_frames, int k_bins,
                     float* x_q, double* scalars, float* results, void* stream);

/* K2 forward + K3a in one pass over x (training forward): everything vq_gather_st_fwd does, plus the per-code sums and
 * counts of the valid frames ADDED into stats [K*D + K] (what vq_ema_accumulate computes; bottleneck.py:64-68), so x is read
 * from HBM once instead of twice.  Each thread block keeps a private [K][64] slab of per-code sums in shared memory, hence
 * K <= 512 (any D <= 512, in 64-deep slices); also T % 4 == 0 and 16-byte aligned x / x_q.  vq_gather_st_fwd_ema_supported
 * returns 1 when the shape qualifies (the caller otherwise uses vq_gather_st_fwd + vq_ema_accumulate). */
int vq_gather_st_fwd_ema_supported(int64_t emb_width, int64_t t_frames, int k_bins);
int vq_gather_st_fwd_ema(const float* x, const int64_t* idx, const float* mask, const float* k,
                         int64_t n_utt, int64_t emb_width, int64_t t_frames, int k_bins,
                         float* x_q, double* scalars, float* results, float* stats, void* stream);

/* K2 backward -- the autograd contract of bottleneck.py:194-201: only x receives gradient,
 *   grad_x = mask * grad_xq + [mask!=0] * grad_commit * 2 (x - k[idx]) / (MASK_SUM * D).
 * grad_commit is a device scalar; scalars[VQ_S_MASK_SUM] must still hold the forward's value.
 * grad_xq may be NULL (treated as zero). */
int vq_gather_st_bwd(const float* x, const int64_t* idx, const float* mask, const float* k,
                     const float* grad_xq, const float* grad_commit, const double* scalars,
                     int64_t n_utt, int64_t emb_width, int64_t t_frames, int k_bins,
                     float* grad_x, void* stream);

/* Decode -- replaces BottleneckBlock.decode (bottleneck.py:160-169; callers
 * scripts/generate_vq_dataset.py:75, models/transformer_lm/transformer_lm.py:103): idx [N,T] -> [N,D,T]. */
int vq_decode(const int64_t* idx, const float* k, int64_t n_utt, int64_t emb_width, int64_t t_frames,
              int k_bins, float* x_d, void* stream);

/* K3a -- replaces the dense one-hot scatter + GEMM + row-sum of update_k (bottleneck.py:64-68).
 *   stats is [K*D + K] fp32: per-code sums of the valid rows followed by per-code counts.
 *   The call ADDS into stats (zero it first); it is the buffer the caller all-reduces over NCCL
 *   (bottleneck.py:74-75) before vq_ema_finalize.  scratch (may be NULL) is ceil(T/64)*N bytes of device memory;
 *   when given together with a mask, 64-frame tiles without a valid frame are never read. */
int vq_ema_accumulate(const float* x, const int64_t* idx, const float* mask,
                      int64_t n_utt, int64_t emb_width, int64_t t_frames, int k_bins,
                      float* stats, void* scratch, void* stream);

/* K3b -- replaces the EMA lerp, usage threshold, dead-code revival and the four metrics of update_k
 * (bottleneck.py:78-89).  k_sum, k_elem are updated in place; the new codebook is written to k, which may
 * alias k_old (the reference rebinds self.k to a fresh tensor every step, and autograd may still hold the
 * old one, so the Python module passes a fresh buffer); k_rand [K,D] holds the restart rows
 * (bottleneck.py:69-70,73).  results[VQ_R_ENTROPY..VQ_R_USED_CURR] and *used_curr (int64, may be NULL)
 * are written by the last block.  mu and threshold are the Python floats of the reference (the kernel
 * rounds mu and (1 - mu) to FP32 separately, as `mu * t + (1. - mu) * s` does).  laplace_eps = 0 reproduces the reference (it has no smoothing);
 * a positive value divides k_sum by (k_elem + eps) / (n + K eps) * n with n = sum_c k_elem[c] (the updated cluster sizes);
 * k_elem itself is stored unsmoothed. */
int vq_ema_finalize(const float* stats, const float* k_rand, const float* k_old, float* k, float* k_sum, float* k_elem,
                    int k_bins, int emb_width, double mu, double threshold, double laplace_eps,
                    double* scalars, float* results, int64_t* used_curr, void* stream);

/* ---- the one exchange step of the path over NVLink peer memory (one process per GPU, one node) ------------------------
 * Replaces distributed.broadcast(_k_rand, 0) + all_reduce(_k_sum) + all_reduce(_k_elem) (bottleneck.py:72-75).  Every rank
 * allocates a REGION with vq_p2p_alloc, sends the 64-byte handle to its peers (any transport: the Python module uses
 * torch.distributed.all_gather_object once, at set-up) and maps theirs with vq_p2p_open.  Per training step `step` (1, 2, ...,
 * the same on all ranks) a rank lets vq_ema_accumulate add into vq_p2p_stats_slot(own region, step) (zeroed first) and
 * writes its restart rows to vq_p2p_krand_slot(own region, step); vq_p2p_exchange then publishes a flag in every peer's
 * region, waits for all peers' flags, and writes the sum over ranks of the statistics (summed in rank order: bit-identical
 * on every rank) to stats_out [K*D + K] and rank 0's restart rows to k_rand_out [K*D] -- the inputs of vq_ema_finalize.
 * `regions` is a HOST array of n_ranks device pointers (this rank's own region at index `rank`, the peers' mappings
 * elsewhere), n_ranks <= 16.  Everything is enqueued on `stream`; a peer that never arrives traps after seconds. */
size_t vq_p2p_region_bytes(int k_bins, int emb_width);
int    vq_p2p_alloc(size_t bytes, void** region_out, unsigned char* handle64_out);
int    vq_p2p_open(const unsigned char* handle64, void** region_out);
int    vq_p2p_close(void* peer_region);
int    vq_p2p_free(void* region);
float* vq_p2p_stats_slot(void* region, unsigned int step, int k_bins, int emb_width);
float* vq_p2p_krand_slot(void* region, unsigned int step, int k_bins, int emb_width);
int    vq_p2p_exchange(void* const* regions, int n_ranks, int rank, unsigned int step, int k_bins, int emb_width,
                       float* stats_out, float* k_rand_out, void* stream);
/* The two halves of vq_p2p_exchange, so that other work (K2) can be enqueued between them and hide the peers' latency:
 * publish as soon as this rank's slots of `step` are written, collect (wait for all peers + rank-ordered sum) when needed. */
int    vq_p2p_publish(void* const* regions, int n_ranks, int rank, unsigned int step, void* stream);
int    vq_p2p_collect(void* const* regions, int n_ranks, int rank, unsigned int step, int k_bins, int emb_width,
                      float* stats_out, float* k_rand_out, void* stream);

/* Gather K rows of the flattened [N*T, D] view of an NCT tensor: out[j,:] = x[n_j, :, t_j] with
 * row = n*T + t.  Used for the restart rows y[randperm][:K] (bottleneck.py:40,70) so that only K rows
 * are touched instead of a full permuted copy. */
int vq_gather_rows(const float* x, const int64_t* rows, int64_t n_rows, int64_t n_utt, int64_t emb_width,
                   int64_t t_frames, float* out, void* stream);

/* Device-side variant of the restart-row draw (no host round trip, different random stream than the reference:
 * BottleneckBlock(rng_parity=False)): out[j,:] = x[n_j,:,t_j] for k_bins frames drawn uniformly, with replacement, among
 * the frames with mask != 0 (all frames when mask is NULL) -- the distribution of y[randperm(M)][:K] (bottleneck.py:40,70)
 * up to repeats.  When there are fewer valid frames than codes the rows get the N(0, (0.01/sqrt(D))^2) jitter of `_tile`
 * (bottleneck.py:26-33); with no valid frame at all they are zero.  The counter-based stream is selected by
 * seed ^ *seed_dev (seed_dev: optional device word, e.g. a call counter the caller increments on the device so that a
 * captured CUDA graph draws fresh rows on every replay); scratch: (n_utt + 1) * 8 bytes of device memory. */
int vq_restart_rows_device(const float* x, const float* mask, int64_t n_utt, int64_t emb_width, int64_t t_frames,
                           int k_bins, uint64_t seed, const uint64_t* seed_dev, float* out, void* scratch, void* stream);

/* Optional per-kernel timing of vq_assign with CUDA events recorded on the launching stream (what bench.py's
 * roofline uses).  vq_profile_enable(1) clears the ring and starts recording the next (up to 64) calls;
 * vq_profile_read synchronises on the last recorded event and writes the AVERAGE milliseconds per call of
 * {codebook prepare, main assign kernel, exact fallback kernel, number of calls averaged}. */
int vq_profile_enable(int on);
int vq_profile_read(float* ms4);

/* ---- host-buffer convenience path (what bench.py's e2e figure and a non-PyTorch caller use) ---- */
typedef struct vq_host_ctx vq_host_ctx;

/* Creates device buffers, pinned staging and two streams for encode jobs of up to max_rows frames.
 * Returns NULL on failure (see vq_last_error). */
vq_host_ctx* vq_host_ctx_create(int device, int64_t max_rows, int k_bins, int emb_width);
void         vq_host_ctx_destroy(vq_host_ctx* ctx);
/* Pinned staging buffers owned by the context (so callers can fill them without an extra copy). */
float*       vq_host_ctx_x_staging(vq_host_ctx* ctx);      /* max_rows * D floats */
int64_t*     vq_host_ctx_idx_staging(vq_host_ctx* ctx);    /* max_rows int64      */
int          vq_host_ctx_set_codebook(vq_host_ctx* ctx, const float* k_host);
/* Encode x_host [N,D,T] (any host memory; pinned is faster) into idx_host [N,T]: H2D copy, K1, D2H copy,
 * chunked by utterance and double-buffered over two streams; returns after idx_host is complete. */
int          vq_encode_host(vq_host_ctx* ctx, const float* x_host, int64_t n_utt, int64_t t_frames,
                            int64_t* idx_host, double* sum_min_d_host);

/* Compact variant for the code dump (scripts/generate_vq_dataset.py:83-90,105-121 keeps q[:ql] of every utterance): same
 * pipeline, but the indices are packed on the device into RAGGED uint16 codes -- utterance n contributes its first
 * lengths_host[n] frames (NULL: all t_frames), utterance-major, no padding -- so 2 bytes per valid frame come back instead of
 * 8 per padded one.  codes_host must hold sum(lengths) uint16 (<= n_utt * t_frames; vq_host_ctx_codes_staging gives a pinned
 * buffer of max_rows); *total_codes_out (may be NULL) receives sum(lengths).  Needs k_bins <= 65536. */
uint16_t*    vq_host_ctx_codes_staging(vq_host_ctx* ctx);   /* max_rows uint16, pinned, allocated on first use */
int          vq_encode_host_u16(vq_host_ctx* ctx, const float* x_host, int64_t n_utt, int64_t t_frames,
                                const int32_t* lengths_host, uint16_t* codes_host, int64_t* total_codes_out);

#ifdef __cplusplus
}
#endif
#endif /* VQB200_H */
