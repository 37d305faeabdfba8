/*
 * hvb.h -- batched C-ABI of the B200-native Turing-codec pixel hot path.
 *
 * This is the drop-in boundary underneath the reference encoder's havoc function tables
 * (turing/StateFunctionTables.h:37-102).  The reference calls one primitive on one <=64x64 block
 * through a C function pointer (havoc/sad.h:56-120, ssd.h:32-52, hadamard.h:31-53,
 * pred_inter.h:33-106, pred_intra.h:29-60, transform.h:33-146, quantize.h:40-99); this ABI takes a
 * BATCH of such calls against device-resident pictures, so that a CTU's (or a wavefront's) worth of
 * PU/TU candidates is one kernel launch.  The literal per-block tables are provided on top of it by
 * include/havoc_b200.h.
 *
 * Conventions
 *  - extern "C", opaque handle, plain pointers and sizes; no CUDA or torch types.
 *  - every function returns 0 on success or a negative hvb_status; hvb_last_error() has the text.
 *  - a context owns one CUDA stream (replaceable with hvb_set_stream) and is used by ONE host
 *    thread at a time; the reference's pool threads (turing/ThreadPool.cpp:87-103) each own one.
 *  - `mem` says where the task / result arrays live: HVB_HOST (pageable or pinned host memory, the
 *    call stages them through pinned buffers and returns after the results have landed) or
 *    HVB_DEVICE (device pointers; the call only enqueues work on the context's stream).
 *  - a "picture" is a device-resident planar 4:2:0 picture (Y, Cb, Cr) whose planes are padded on
 *    all sides like the reference's (turing/StatePictures.h:154-156), so motion vectors may point
 *    outside the picture exactly as far as the reference allows.
 *  - all results are bit-exact with the reference's C path (`--asm 0`), including its quirks:
 *    16-bit SAD >> 2, SSD >> 4 (mod 2^32), SATD >> 2, forward-transform wrap to int16.
 */
#ifndef HVB_H
#define HVB_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hvb_context hvb_context;

typedef enum
{
    HVB_OK = 0,
    HVB_ERR_INVALID = -1,  /* bad argument */
    HVB_ERR_CUDA = -2,     /* CUDA runtime error (see hvb_last_error) */
    HVB_ERR_NOMEM = -3,
    HVB_ERR_NO_DEVICE = -4 /* no usable sm_100 device: there is NO CPU fallback */
} hvb_status;

typedef enum
{
    HVB_HOST = 0,
    HVB_DEVICE = 1
} hvb_mem;

/* ---- context ------------------------------------------------------------------------ */

/* bytes_per_sample: 1 (uint8_t pictures, bit_depth 8) or 2 (uint16_t pictures, bit_depth 8..10);
 * replaces havoc_new_code()/havoc_delete_code() (havoc/havoc.h:143-149). */
int hvb_create(int device, int bytes_per_sample, int bit_depth, hvb_context **out);
void hvb_destroy(hvb_context *ctx);
const char *hvb_last_error(hvb_context *ctx);
/* use an existing cudaStream_t (passed as void*) for all subsequent work; NULL = own stream */
int hvb_set_stream(hvb_context *ctx, void *cuda_stream);
int hvb_sync(hvb_context *ctx);
/* 1 when everything enqueued on the context so far has completed, 0 when work is still in flight, < 0 on error; never
 * blocks and may be called from any thread (a completion thread that serves several contexts without parking a core
 * per context inside the driver's wait). */
int hvb_poll(hvb_context *ctx);
/* Completion without driver calls: after everything enqueued on the context so far has completed (its effects visible to the
 * host), `value` is stored to *flag, which must lie in memory from hvb_host_alloc.  A completion thread that watches many
 * contexts reads flags instead of querying streams (each query takes the driver's lock, which the launching threads need). */
int hvb_signal(hvb_context *ctx, int32_t *flag, int32_t value);
/* Device-side stopwatch: hvb_mark records a timestamp on the context's stream (slot 0..15, events created on first use);
 * hvb_elapsed_ms gives the device time between two recorded slots once the work between them has completed. */
int hvb_mark(hvb_context *ctx, int slot);
int hvb_elapsed_ms(hvb_context *ctx, int from, int to, float *ms);
/* Pipelined host mode.  Off (default): every HVB_HOST call returns after its results have landed.  On: a batch call
 * whose task and result arrays are page-locked (cudaHostAlloc / cudaHostRegister -- an encoder's arena) only enqueues
 * work -- tasks go up on a copy-in stream, the kernels run on the context's stream, the results come back on a copy-out
 * stream, through a ring of four staging slots -- and returns; picture / pool uploads from page-locked memory go on the
 * copy-in stream too: batches issued AFTER an upload see it, batches issued BEFORE it are not waited for, so upload into
 * pictures / pool regions that no batch in flight reads (double-buffer).  Results and source buffers belong to the
 * library until hvb_sync() returns.  Calls with pageable buffers keep the blocking behaviour. */
int hvb_set_pipelined(hvb_context *ctx, int on);
/* Page-locked host memory that the device can address directly (cudaHostAlloc, portable + mapped): an encoder's task /
 * result arena.  With unified addressing the returned pointer is valid on both sides, so arrays in it may be passed
 * with HVB_HOST (copied from / to without staging) or with HVB_DEVICE (the kernels read the tasks and write the results
 * across the bus themselves: no copy is enqueued at all -- the cheapest form for the small, latency-bound batches of a
 * submission queue; results are valid after hvb_sync). */
int hvb_host_alloc(hvb_context *ctx, size_t bytes, void **out);
int hvb_host_free(hvb_context *ctx, void *ptr);
/* hvb_sad_batch / hvb_sad4_batch with the blocks staged into shared memory by the Tensor Memory Accelerator
 * (cp.async.bulk.tensor over per-plane tensor maps) instead of the load/store path; same results.  Default: off, or the
 * value of HVB_TMA in the environment.  Pictures must own device memory (not hvb_picture_wrap). */
int hvb_set_tma(hvb_context *ctx, int on);
/* hvb_tu_chain_batch has two forms with identical results: staged kernels sized for whole frames, and one launch with a warp per
 * block for callers that wait for a few blocks (the batched encoder).  Batches of at most `blocks` blocks take the second
 * (default 256, HVB_TU_FUSED_MAX in the environment; 0: always the staged form). */
int hvb_set_tu_fused_max(hvb_context *ctx, int blocks);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
int64_t hvb_launch_count(hvb_context *ctx);
/* 1 when the device is present and kernels for it are in this binary */
int hvb_device_ok(int device);

/* ---- pictures ------------------------------------------------------------------------ */

/* luma width x height; chroma planes are (w/2) x (h/2).  pad = luma padding in samples on every
 * side (chroma gets pad/2); the reference uses 96 (StatePictures.h:154-156).  */
int hvb_picture_create(hvb_context *ctx, int width, int height, int pad, int *pic);
/* One device allocation for the planes of the next `count` pictures of this geometry (hvb_picture_create carves from it while it
 * lasts; such a picture's memory returns with hvb_destroy, not hvb_picture_destroy): a session that creates hundreds of pictures
 * spares as many trips through the allocator.  Once per context. */
int hvb_picture_reserve(hvb_context *ctx, int width, int height, int pad, int count);
int hvb_picture_destroy(hvb_context *ctx, int pic);
/* copy rows [y0, y0+rows) of plane cIdx from host (stride in samples) */
int hvb_picture_upload(hvb_context *ctx, int pic, int cIdx, const void *host, intptr_t stride,
                       int y0, int rows);
int hvb_picture_download(hvb_context *ctx, int pic, int cIdx, void *host, intptr_t stride,
                         int y0, int rows);
/* rectangle variants (x0,y0,w,h in samples of that plane; the rectangle may lie in the padding):
 * what the per-block table shim and the encoder's per-CTU reconstruction commit (turing/Write.h:828-830) use.
 * In pipelined mode an upload from page-locked memory is only enqueued (as hvb_picture_upload). */
int hvb_picture_upload_rect(hvb_context *ctx, int pic, int cIdx, const void *host, intptr_t stride,
                            int x0, int y0, int w, int h);
int hvb_picture_download_rect(hvb_context *ctx, int pic, int cIdx, void *host, intptr_t stride,
                              int x0, int y0, int w, int h);
/* replicate edge samples into the padding of all three planes (turing/Padding.h) */
int hvb_picture_pad(hvb_context *ctx, int pic);
/* device-to-device copy of all three planes incl. their padding (same geometry required): the copy of the deblocked
 * picture that SAO filters from (turing/TaskSao.cpp:96-121) */
int hvb_picture_copy(hvb_context *ctx, int dst_pic, int src_pic);
/* A one-plane picture (cIdx 0, no padding) over caller memory from hvb_host_alloc: the kernels read and write it across
 * the bus, nothing is copied.  For blocks a caller supplies or wants back per task (a CU's prediction and reconstruction,
 * hvbenc_tu_chain) when a copy per block would cost more than the bytes.  hvb_picture_destroy releases the id only. */
int hvb_picture_wrap(hvb_context *ctx, void *host, intptr_t stride, int width, int height, int *pic);
/* Make picture `owner_pic` of context `owner` visible in `ctx` under the id *pic: the same device memory, not a copy (both
 * contexts must be on the same device).  The importing context never frees it; the owner must outlive the import.  For
 * several contexts (one per dispatcher thread, each with its own stream) working on one set of pictures. */
int hvb_picture_import(hvb_context *ctx, hvb_context *owner, int owner_pic, int *pic);
/* raw device view of a plane: pointer to sample (0,0) and stride in samples (for zero-copy fills) */
int hvb_picture_plane(hvb_context *ctx, int pic, int cIdx, void **dev_ptr, intptr_t *stride);

/* ---- block metrics (hot loops A, B, C) ------------------------------------------------ */

/* A block of a picture plane.  x,y in samples of that plane; may lie in the padding. */
typedef struct
{
    int16_t pic;
    int16_t cIdx;
    int16_t x, y;
} hvb_block;

/* one candidate of havoc_sad / havoc_ssd / measureSatd: metric(a, b) over w x h */
typedef struct
{
    hvb_block a, b;
    int16_t w, h;
    int32_t reserved;
} hvb_metric_task; /* 24 bytes */

/* havoc_sad<Sample> (havoc/sad.h:58): out[i] = int32 SAD; 16-bit samples: >> 2 */
int hvb_sad_batch(hvb_context *ctx, const hvb_metric_task *tasks, int n, int32_t *out, hvb_mem mem);
/* havoc_ssd<Sample> (havoc/ssd.h:33), any w x h: out[i] = uint32 SSD (mod 2^32; 16-bit >> 4) */
int hvb_ssd_batch(hvb_context *ctx, const hvb_metric_task *tasks, int n, uint32_t *out, hvb_mem mem);
/* measureSatd (turing/Measure.h:96-135) over havoc_hadamard_satd tiles (havoc/hadamard.h:32) */
int hvb_satd_batch(hvb_context *ctx, const hvb_metric_task *tasks, int n, int32_t *out, hvb_mem mem);

/* havoc_sad_multiref<Sample> (havoc/sad.h:100): four references, one source block */
typedef struct
{
    hvb_block src;
    int16_t ref_pic, ref_cIdx;
    int16_t w, h;
    int16_t rx[4], ry[4]; /* top-left of each reference block in the reference plane */
} hvb_sad4_task; /* 32 bytes */
int hvb_sad4_batch(hvb_context *ctx, const hvb_sad4_task *tasks, int n, int32_t *out /* [n][4] */,
                   hvb_mem mem);

/* ---- inter prediction (hot loop B) ----------------------------------------------------- */

/* HavocPredUni (havoc/pred_inter.h:35) / HavocPredBi (:63).  Luma (cIdx 0): 8-tap, mv in
 * quarter-samples; chroma: 4-tap, mv in eighth-samples of the chroma plane (the caller passes
 * the luma mv unchanged, as turing/Dsp.h:789-812 does).  dst receives w x h samples. */
typedef struct
{
    hvb_block dst;
    int16_t ref_pic[2]; /* ref_pic[1] < 0 -> uni-prediction */
    int16_t x, y;       /* block position in the reference plane(s) before displacement */
    int16_t w, h;
    int16_t mvx[2], mvy[2];
    int32_t reserved;
} hvb_pred_task; /* 32 bytes */
int hvb_pred_batch(hvb_context *ctx, const hvb_pred_task *tasks, int n, hvb_mem mem);

/* havoc::SubtractBi (havoc/pred_inter.h:87): dst = clip(2*src - pred) */
typedef struct
{
    hvb_block dst, pred, src;
    int16_t w, h;
    int32_t reserved;
} hvb_subtract_bi_task; /* 32 bytes */
int hvb_subtract_bi_batch(hvb_context *ctx, const hvb_subtract_bi_task *tasks, int n, hvb_mem mem);

/* costDistortionMv (turing/Search.hpp:1965-2000) without the lambda product: interpolate the luma
 * block at quarter-pel mv and return measureSatd(src, prediction).  Nothing is written back. */
typedef struct
{
    hvb_block src;
    int16_t ref_pic, reserved0;
    int16_t w, h;
    int16_t mvx, mvy; /* quarter-pel */
} hvb_interp_satd_task; /* 20 bytes */
int hvb_interp_satd_batch(hvb_context *ctx, const hvb_interp_satd_task *tasks, int n, int32_t *out,
                          hvb_mem mem);

/* ---- intra prediction (hot loop D) ------------------------------------------------------ */

/* havoc::intra::Function (havoc/pred_intra.h:32).  Neighbours live in a device-resident sample
 * pool (hvb_pool_upload); `nb` is the index of the sample holding p(-1,-1) so that
 * p(x,-1) = pool[nb + 1 + x] and p(-1,y) = pool[nb - 1 - y]  (4n+1 samples). */
int hvb_pool_upload(hvb_context *ctx, const void *samples, size_t count, size_t offset);

typedef struct
{
    hvb_block dst;
    int32_t nb;
    int8_t log2n, mode, edge_flag, reserved;
} hvb_intra_task; /* 16 bytes */
int hvb_intra_pred_batch(hvb_context *ctx, const hvb_intra_task *tasks, int n, hvb_mem mem);

/* 35-mode SATD sweep of searchIntraPartition (turing/Search.hpp:39-267 via
 * Reconstruct.cpp:630-712): out[i][m] = SATD(src, intra(m)), m = 0..34.  nb_unfiltered /
 * nb_filtered select the reference array per mode by the filterFlag rule (turing/Dsp.h:57-70);
 * nb_filtered < 0 asks the kernel to derive the filtered array itself
 * (turing/IntraReferenceSamples.h:373-419). */
typedef struct
{
    hvb_block src;
    int32_t nb_unfiltered, nb_filtered;
    int8_t log2n, cIdx;
    int8_t strong_intra_smoothing; /* sps flag, used when nb_filtered < 0 (filtered array derived on device) */
    int8_t reserved[5];
} hvb_intra_sweep_task; /* 24 bytes */
int hvb_intra_satd35_batch(hvb_context *ctx, const hvb_intra_sweep_task *tasks, int n,
                           int32_t *out /* [n][35] */, hvb_mem mem);

/* ---- transform / quantisation (hot loop C) ------------------------------------------------ */

/* raw coefficient-domain primitives over a device-resident int16 pool (hvb_coeff_upload /
 * hvb_coeff_download); offsets are in int16 elements, blocks are n x n contiguous. */
int hvb_coeff_upload(hvb_context *ctx, const int16_t *data, size_t count, size_t offset);
/* (pipelined mode, page-locked destination: the copy is only enqueued; the data is valid after hvb_sync / hvb_poll) */
int hvb_coeff_download(hvb_context *ctx, int16_t *data, size_t count, size_t offset);

typedef struct
{
    int32_t src, dst;     /* pool offsets */
    int32_t src_stride;   /* forward transform only: residual stride in elements */
    int8_t log2n, trType; /* trType 1 = 4x4 DST */
    int16_t reserved;
} hvb_transform_task; /* 16 bytes */
/* havoc::Transform (havoc/transform.h:117) */
int hvb_transform_fwd_batch(hvb_context *ctx, const hvb_transform_task *tasks, int n, hvb_mem mem);
/* havoc::inverse_transform (havoc/transform.h:33): residual only */
int hvb_transform_inv_batch(hvb_context *ctx, const hvb_transform_task *tasks, int n, hvb_mem mem);

typedef struct
{
    int32_t src, dst;
    int32_t n;
    int32_t scale, shift, offset;
} hvb_quant_task; /* 24 bytes */
/* havoc_quantize (havoc/quantize.h:63): cbf[i] = 1 when any output is non-zero, else 0 */
int hvb_quantize_batch(hvb_context *ctx, const hvb_quant_task *tasks, int n, int32_t *cbf, hvb_mem mem);
/* havoc_quantize_inverse (havoc/quantize.h:42) */
int hvb_quantize_inverse_batch(hvb_context *ctx, const hvb_quant_task *tasks, int n, hvb_mem mem);

/* havoc::inverse_transform_add (havoc/transform.h:61): rec = clip(pred + IT(coeffs)) */
typedef struct
{
    hvb_block dst, pred;
    int32_t coeffs; /* pool offset */
    int8_t log2n, trType;
    int16_t reserved;
} hvb_ita_task; /* 24 bytes */
int hvb_inverse_transform_add_batch(hvb_context *ctx, const hvb_ita_task *tasks, int n, hvb_mem mem);

/* The whole TU pipeline of ReconstructInterBlock / the intra flavour
 * (turing/Reconstruct.cpp:180-356, :731-857) for one transform block:
 *   residual = src - pred; coeffs = T(residual); levels = Q(coeffs) [plain or RDOQ];
 *   if cbf: rec = clip(pred + IT(IQ(levels))) else rec = pred;  ssd = SSD(src, rec)
 * levels are written to the coefficient pool at `levels` (n*n int16). */
typedef struct
{
    hvb_block src, pred, rec;
    int32_t levels;
    int8_t log2n, trType, cIdx, flags; /* flags bit0: use RDOQ, bit1: isIntra, bit2: SDH */
    int32_t qscale, qshift, qoffset;   /* forward quantiser (QpState.h) */
    int32_t iqscale, iqshift;          /* inverse quantiser */
    int8_t scanIdx, reserved[3];
    int32_t rdoq_ctx;                  /* index of the RDOQ context snapshot (hvb_rdoq_contexts_upload) */
} hvb_tu_task; /* 60 bytes */

typedef struct
{
    uint32_t ssd;     /* havoc_ssd(src, rec)  (Reconstruct.cpp:849-853) */
    uint32_t ssdPred; /* havoc_ssd(src, pred) (Reconstruct.cpp:856) */
    int32_t cbf;
    int32_t status;      /* 0; -1: the task was rejected (levels outside the coefficient pool, log2n outside 2..5, or an RDOQ
                            task whose rdoq_ctx was never uploaded) and nothing was computed for it */
    uint32_t sadQuad[4]; /* sum |src - pred| over the block's quadrants, [2 * (y >= n/2) + (x >= n/2)]: what
                            reconstructInter accumulates into Candidate::sadResidueQuad (Reconstruct.cpp:1268-1287, read by
                            Aps::analyseResidueEnergy, turing/Aps.h:62-76); a caller whose CU spans several blocks adds them up */
} hvb_tu_result; /* 32 bytes */
int hvb_tu_chain_batch(hvb_context *ctx, const hvb_tu_task *tasks, int n, hvb_tu_result *out, hvb_mem mem);


/* RDOQ inputs that are not pixels: a snapshot of the CABAC context states the reference's
 * Rdoq reads through estimateBits (turing/Rdoq.cpp:26-31) plus lambda.  One snapshot serves many
 * TUs (the reference snapshots per CU).  Layout: see hvb_rdoq_ctx below. */
typedef struct
{
    uint8_t sig_coeff_flag[44];
    uint8_t greater1_flag[24];
    uint8_t greater2_flag[6];
    uint8_t coded_sub_block_flag[4];
    uint8_t last_x_prefix[18];
    uint8_t last_y_prefix[18];
    uint8_t cbf_luma[2];
    uint8_t cbf_cbcr[5];
    uint8_t rqt_root_cbf[1];
    uint8_t reserved[6];
    double lambda; /* as passed to Rdoq::Rdoq (turing/Rdoq.h:170) */
} hvb_rdoq_ctx; /* 136 bytes */
int hvb_rdoq_contexts_upload(hvb_context *ctx, const hvb_rdoq_ctx *snapshots, int count, int first);
/* A caller that waits for every batch (the submission queue of hvb_encoder.h) can spare the copies around the one-launch form of
 * hvb_tu_chain_batch: levels are written to, and context snapshots read from, the caller's own page-locked arrays (hvb_host_alloc)
 * across the bus.  `levels`: the pool hvb_tu_task.levels indexes; `snapshots`: the array hvb_tu_task.rdoq_ctx indexes -- both must
 * stay untouched while a batch is in flight.  NULL / 0 returns to the context's device arrays.  While either is set, batches larger
 * than hvb_set_tu_fused_max are refused. */
int hvb_coeff_pool_wrap(hvb_context *ctx, int16_t *levels, size_t count);
int hvb_rdoq_contexts_wrap(hvb_context *ctx, const hvb_rdoq_ctx *snapshots, int count);

/* Rdoq::runQuantisation alone (turing/Rdoq.cpp:35-450) on pool coefficients */
typedef struct
{
    int32_t src, dst;
    int32_t qscale, qshift;
    int32_t iqscale;
    int8_t log2n, cIdx, scanIdx, flags; /* flags bit1: isIntra, bit2: SDH */
    int32_t rdoq_ctx;
} hvb_rdoq_task; /* 28 bytes */
int hvb_rdoq_batch(hvb_context *ctx, const hvb_rdoq_task *tasks, int n, int32_t *cbf, hvb_mem mem);

/* The distortion half of measurePuCost (turing/Search.hpp:1668-1682): predictInter without weighted prediction
 * (turing/Dsp.h:866-915 -> predictUni :769-806 / predictBi :808-864; luma origin clamped by clipMvLumaComponent
 * :723-731, chroma origin = luma origin >> 1, chroma phase = mv & 7) followed by measureSatd of Y, Cb and Cr against
 * the source picture (turing/Measure.h:96-176; a chroma block that is not a multiple of 4 contributes 0).
 * out[i] = {satdY, satdCb, satdCr}; the caller adds the PU's rate and multiplies by lambda (:1703).  When dst_pic >= 0
 * the three predicted blocks are also stored there (what predictInter leaves in the reconstructed picture). */
typedef struct
{
    int16_t src_pic, dst_pic;
    int16_t ref_pic[2];          /* L0, L1; < 0: list not used (uni-prediction from the other) */
    int16_t x0, y0, w, h;        /* PU in luma samples */
    int16_t mvx[2], mvy[2];      /* quarter-pel luma vectors of L0, L1 */
} hvb_pu_cost_task; /* 24 bytes */
int hvb_pu_cost_batch(hvb_context *ctx, const hvb_pu_cost_task *tasks, int n, int32_t *out /* [n][3] */, hvb_mem mem);

/* ---- motion search (hot loops A + B with their control flow) ------------------------------ */

typedef struct
{
    int16_t x, y;
} hvb_mv;

/* One uni-directional PU search: fullPelMotionEstimation (turing/Search.hpp:2064-2336) followed
 * by subPelRefinement (:2339-2357), as chained by searchMotionUni (:1315-1352).  Everything the
 * reference reads from encoder state is an explicit input. */
typedef struct
{
    int16_t src_pic, ref_pic;
    int16_t x0, y0, w, h;        /* PU in luma samples */
    hvb_mv mvp[2];               /* AMVP predictors, quarter-pel */
    int64_t rateMvpFlag[2];      /* Cost (Q16) of mvp_lX_flag = 0 / 1 */
    int32_t lambda;              /* Lambda (Q16): FixedPoint<int32,16>::set(reciprocal sqrt lambda) */
    hvb_mv limitMin, limitMax;   /* LimitFullPelMv (Search.hpp:1366-1407), full-pel, inclusive */
    hvb_mv prev2Nx2N;            /* mvPreviousInteger2Nx2N[refList], quarter-pel units */
    uint8_t smallSearchWindow, met, log2CbSize, usePrev2Nx2N;
    uint8_t halfPel, quarterPel, reserved[2];
} hvb_me_task; /* 64 bytes */

typedef struct
{
    hvb_mv mv, mvd;              /* after sub-pel refinement when enabled */
    hvb_mv mvInteger;            /* best integer vector (mvPreviousInteger2Nx2N update) */
    int32_t mvpFlag;
    int64_t cost;                /* best integer-search cost (Q16) */
    int64_t costMvdZero[2];
    int64_t subpelCost;          /* bestCost after subPelRefinement (valid when halfPel) */
    int32_t nSad;                /* SAD evaluations performed (statistics; equals the reference's call count x4) */
    int32_t flags;               /* bit 0: the integer search returned through MET (Search.hpp:2125), so the
                                    caller must NOT update mvPreviousInteger2Nx2N and costMvdZero may be partial */
} hvb_me_result; /* 56 bytes */
int hvb_me_search_batch(hvb_context *ctx, const hvb_me_task *tasks, int n, hvb_me_result *out, hvb_mem mem);

/* One searchMotionBi call (turing/Search.hpp:1498-1653): refine list X's vector of a bi-predicted PU against the
 * "ideal" block 2*src - pred(other list) -- uni prediction from the other list (:1518-1534), SubtractBi (:1541-1548),
 * the exhaustive (2r+1)^2 integer grid with its SAD4 grouping (:1551-1625) and the 3x3 half / quarter rounds
 * (:1628-1650).  The caller chains L0 then L1 as Search<prediction_unit>::searchBi does (:1805-1823), feeding the
 * refined vector of one call into `mvOther` of the next.  Shares its first 56 bytes with hvb_me_task. */
typedef struct
{
    int16_t src_pic, ref_pic;    /* ref_pic: list X's reference picture */
    int16_t x0, y0, w, h;
    hvb_mv mvp[2];               /* list X's AMVP predictors */
    int16_t other_pic, reserved0;/* the other list's reference picture */
    int64_t rateMvpFlag[2];
    int32_t lambda;              /* Lambda (Q16) of getReciprocalSqrtLambda * 0.5 (:1568) */
    hvb_mv limitMin, limitMax;   /* LimitFullPelMv of the PU */
    hvb_mv mvStart;              /* puData.mv(X): the uni-directional result for this list, quarter-pel */
    hvb_mv mvOther;              /* puData.mv(1 - X) */
    uint8_t smallWindow;         /* Speed::useBiSmallSearchWindow(): range 1 instead of 5 */
    uint8_t halfPel, quarterPel, reserved1;
} hvb_me_bi_task; /* 64 bytes */

typedef struct
{
    hvb_mv mv, mvd;              /* refined vector and its mvd against the cheaper predictor */
    hvb_mv mvInteger;            /* best of the integer grid */
    int32_t mvpFlag;
    int64_t cost;                /* cost of the winner of the last round (Q16) */
    int32_t nSad, reserved;
} hvb_me_bi_result; /* 32 bytes */
int hvb_me_bi_search_batch(hvb_context *ctx, const hvb_me_bi_task *tasks, int n, hvb_me_bi_result *out, hvb_mem mem);

/* ---- in-loop deblocking, pixel pass (first GPU verification pending) ---- */

/* LoopFilter::Block (turing/LoopFilter.h:50-90), one per 8x8 luma block as the encoder's process*() calls leave it:
 * data = QpY << 1 | (pcm-with-loop-filter-disabled or cu_transquant_bypass), packedBs = four 2-bit boundary strengths,
 * bits 4*edgeType + 2*position (edgeType 0: the block's left edge, 1: its top edge; position: which 4-sample half). */
typedef struct
{
    int8_t data;
    uint8_t packedBs;
} hvb_deblock_block; /* 2 bytes */
/* the two members of LoopFilter::Ctu (turing/LoopFilter.h:92-96) the deblocking filter reads, per CTU in raster order */
typedef struct
{
    int8_t tc_offset_div2, beta_offset_div2;
} hvb_deblock_ctu; /* 2 bytes */
/* Side information of picture `pic` (host arrays; kept on the device until replaced or the picture is destroyed).
 * blockStride x blockRows is the reference's grid: ((PicWidthInCtbsY << CtbLog2SizeY) >> 3) + 1 records per row
 * (turing/LoopFilter.h:436-443).  Edges on the picture boundary must carry strength 0, as the encoder leaves them. */
int hvb_deblock_info_upload(hvb_context *ctx, int pic, const hvb_deblock_block *blocks, int blockStride, int blockRows,
                            const hvb_deblock_ctu *ctus, int picWidthInCtbs, int picHeightInCtbs, int ctbLog2);
/* One call of LoopFilter::Picture::deblock<edgeType>(h, recL, recCb, recCr, xBegin, yBegin, xEnd, yEnd)
 * (turing/LoopFilter.h:739-777): the edges of that type of the 8x8 blocks [xBegin/8, xEnd/8) x [yBegin/8, yEnd/8),
 * luma and 4:2:0 chroma, filtered in place.  The tasks of one batch run concurrently and must not share samples: all
 * vertical-edge regions of a picture (disjoint regions, e.g. TaskDeblock's per-CTU ones, turing/TaskDeblock.cpp:104-114)
 * in one call, the horizontal-edge regions in the next (H.265 8.7.2: horizontal edges see the vertical pass's output). */
typedef struct
{
    int16_t pic;
    int16_t edgeType; /* 0: vertical edges, 1: horizontal edges */
    int16_t xBegin, yBegin, xEnd, yEnd; /* luma samples */
    int16_t cbQpOffset, crQpOffset;     /* pps_cb_qp_offset, pps_cr_qp_offset */
} hvb_deblock_task; /* 16 bytes */
int hvb_deblock_batch(hvb_context *ctx, const hvb_deblock_task *tasks, int n, hvb_mem mem);

/* ---- sample adaptive offset, application (first GPU verification pending) ---- */

/* One CTU's SAO parameters and neighbourhood as LoopFilter::Ctu holds them (turing/LoopFilter.h:92-163, :476-534):
 * left/top/right/bottom = luma-sample limits of the area whose samples may be read across the CTU's borders (picture
 * limits, or the CTU's own border where a slice / tile boundary forbids it), the four corner CTUs' availability, and per
 * component SaoTypeIdx (0 off, 1 band, 2 edge), SaoEoClass or sao_band_position, SaoOffsetVal[1..4] (already scaled
 * by 1 << (bitDepth - min(bitDepth, 10)), LoopFilter.h:157-160). */
typedef struct
{
    int16_t left, top, right, bottom;
    uint8_t topLeft, topRight, bottomLeft, bottomRight;
    struct
    {
        int8_t typeIdx, classOrBand;
        int16_t offset[4];
    } plane[3];
} hvb_sao_ctu; /* 42 bytes */
/* SAO records of every CTU of picture `pic`, raster order (host array).  The disabled-block bits (pcm / transquant bypass,
 * restoreUnfilteredRegions, turing/LoopFilter.h:849-877) and the CTB geometry are those of hvb_deblock_info_upload, which
 * must have been called for `pic`. */
int hvb_sao_info_upload(hvb_context *ctx, int pic, const hvb_sao_ctu *ctus, int n);
/* LoopFilter::Picture::applySaoCTU (turing/LoopFilter.h:794-811 -> filterBlockSao :885-1017, sao_filter_edge / _band of
 * turing/sao.cpp) for the CTUs [ctuBegin, ctuEnd) of dst_pic in raster order: SAO of src_pic (a copy of the deblocked
 * picture, turing/TaskSao.cpp:96-121) into dst_pic, whose side information is used; visible samples only.  src_pic and
 * dst_pic must be different pictures (the edge classes read the unfiltered neighbours). */
typedef struct
{
    int16_t src_pic, dst_pic;
    int16_t ctuBegin, ctuEnd;
    uint8_t lumaFlag, chromaFlag; /* slice_sao_luma_flag, slice_sao_chroma_flag */
    int16_t reserved;
} hvb_sao_task; /* 12 bytes */
int hvb_sao_batch(hvb_context *ctx, const hvb_sao_task *tasks, int n, hvb_mem mem);

/* ---- SAO statistics, encoder side (first GPU verification pending) ---- */

/* EncSao::edge_offset_stats_class0..3 and band_offset_luma_stats (turing/EncSao.h:111-284) over one block of one
 * component -- the CTU clipped to the picture, as saoRdEstimateLuma / saoRdEstimateChroma cut it (:290-297, :535-539);
 * the block's first and last row and column are left out, as there.  org_pic: the source picture, rec_pic: the
 * reconstruction the statistics are taken on (deblocked, or not, by --sao-slow: :303-305). */
typedef struct
{
    int16_t org_pic, rec_pic;
    int16_t cIdx;
    int16_t x0, y0, w, h; /* in samples of that component */
    int16_t reserved;
} hvb_sao_stats_task; /* 16 bytes */
/* per edge class and category: sum of (original - reconstructed) and number of samples (category 0 of class 0 also
 * carries the reference's second visit of column 1, EncSao.h:167-181); the same per band of 8 << (bitDepth - 8) levels.
 * The chroma band statistics of the reference (:62-109) are the U and V results added; startBand (:137-148) and the
 * offset / type decisions (:328-526) stay with the caller. */
typedef struct
{
    int32_t edgeE[4][5], edgeCount[4][5];
    int32_t bandE[32], bandCount[32];
} hvb_sao_stats; /* 416 bytes */
int hvb_sao_stats_batch(hvb_context *ctx, const hvb_sao_stats_task *tasks, int n, hvb_sao_stats *out, hvb_mem mem);

/* ---- coded-data feed (first GPU verification pending) ---- */

/* CodedData::storeResidual (turing/CodedData.h:457-517): the levels of a transform block (raster (1 << log2n)^2 int16 at
 * coefficient-pool offset `levels`, e.g. where hvb_tu_chain_batch left them) serialised into the encoder's coded-data
 * record -- the transform-skip word (0), the coded_sub_block flag word(s), then per significant 4x4 sub-block from the last
 * in scan order: significance, greater-than-1 and sign masks (bit 15 - n for scan position n) and the magnitudes above 1.
 * Records are packed into the pool region [recordsBase, recordsBase + capacityWords) (uint16 words; fetch it with
 * hvb_coeff_download); out[i] = where block i's record starts and its length (0: all-zero block, nothing written;
 * -1: the region was full), out[n] = {end of the used part, words that did not fit}.  The order of the records in the
 * region is not defined (blocks reserve their room concurrently). */
typedef struct
{
    int32_t levels;
    int8_t log2n, scanIdx;
    int16_t reserved;
} hvb_coded_residual_task; /* 8 bytes */
typedef struct
{
    int32_t offset, words;
} hvb_coded_residual; /* 8 bytes */
int hvb_coded_residual_batch(hvb_context *ctx, const hvb_coded_residual_task *tasks, int n, int32_t recordsBase, int32_t capacityWords,
                             hvb_coded_residual *out /* [n + 1] */, hvb_mem mem);

/* ---- pre-analysis ---- */

/* EstimateIntraComplexity::preAnalysis (turing/EstimateIntraComplexity.h:159-176) over a region of a source picture's
 * luma plane: for each 8x8 block of the wBlocks x hBlocks blocks starting at (x0, y0) (multiples of 8), computeSatd8x8
 * (:55-157), the sum of the absolute Hadamard coefficients of the samples without the DC term, (s + 2) >> 2 (16-bit
 * samples: >> 2 more); out[task.out + by * wBlocks + bx].  A task must not overlap another's output.  The caller sums
 * (m_satdSum, getSatdCtu). */
typedef struct
{
    int16_t pic, reserved;
    int16_t x0, y0, wBlocks, hBlocks;
    int32_t out;
} hvb_intra_complexity_task; /* 16 bytes */
int hvb_intra_complexity_batch(hvb_context *ctx, const hvb_intra_complexity_task *tasks, int n, int32_t *out, int outCount, hvb_mem mem);

/* AdaptiveQuantisation::preAnalysis (turing/AdaptiveQuantisation.h:172-246), one layer per task: the luma plane of a source
 * picture is cut into `unit` x `unit` squares (unit = maxCuSize >> depth, a power of two, 4..64; the last row / column of
 * units is clipped to the picture), each unit into four quadrants at (w >> 1, h >> 1), and for each quadrant
 * variance = sumSquare / N - (sum / N)^2 in INTEGER quotients with N = (w * h) >> 2 -- with the reference's two quirks kept:
 * quadrant 0's "sum of squares" is the square of its last sample only (:197) and quadrant 1's is its plain sum (:208).
 * out[task.out + row * unitsPerRow + col] = the minimum of the four (an integer; the reference holds it in a double);
 * the unit's activity is 1.0 + that (:239), the layer's average activity the truncated mean of the activities (:244-245),
 * which the caller forms (integer-valued doubles: any order gives the same sum).  getAqOffset (:147-169) stays with the caller. */
typedef struct
{
    int16_t pic, unit;
    int32_t out;
} hvb_aq_layer_task; /* 8 bytes */
int hvb_aq_activity_batch(hvb_context *ctx, const hvb_aq_layer_task *tasks, int n, int64_t *out, int outCount, hvb_mem mem);

/* ShotChangeDetection::processSeq's per-picture pixel pass (turing/SCDetection.h:237-262, :299-304, :402-407): the 64-bin
 * histogram of the luma plane, sample -> (unsigned char)(sample >> 2) for 16-bit samples (:244, :284), then >> SHIFT_DOWN (2).
 * out[64 * i + bin] for picture pics[i].  Black / white / fade tests and the histogram differences stay with the caller. */
int hvb_scd_histogram_batch(hvb_context *ctx, const int16_t *pics, int n, int32_t *out /* [64 n] */, hvb_mem mem);

/* ShotChangeDetection::getLikelihood's block statistics (turing/SCDetection.h:71-147): the picture's luma plane, reduced to
 * 8 bits as above and taken as the reference's packed width x height byte vector, is cut into blocks of (width >> 3) x
 * (height >> 3); for the blocks at j = margin * bh, (margin + 1) * bh, .. < height - margin * bh (rows) and likewise i over
 * the columns, avg = sum / (bh * bw) and var = (sum over the block in raster order of (elem - avg)^2) / (bh * bw), IEEE
 * doubles, each operation rounded once, additions in the reference's order.  The reference addresses element (h, w) of a
 * block as block + h * HEIGHT + w (:87, :103 -- the picture's height where the row pitch was meant); that addressing is
 * kept, so a "block" is a sheared set of samples, and a picture for which it would leave the plane is refused
 * (HVB_ERR_INVALID; the reference reads past its vector there).  margin 1: the previous picture's 6 x 6 grid, margin 2:
 * the current picture's 4 x 4.  out[task.out + 2 k] = avg, out[task.out + 2 k + 1] = var of block k (raster
 * order).  calc_likelihood and the 3 x 3 neighbourhood minimum (:40-47, :149-176) stay with the caller. */
typedef struct
{
    int16_t pic, margin;
    int32_t out; /* index of the task's first double in `out` */
} hvb_scd_stats_task; /* 8 bytes */
int hvb_scd_block_stats_batch(hvb_context *ctx, const hvb_scd_stats_task *tasks, int n, double *out, int outCount, hvb_mem mem);

#ifdef __cplusplus
}
#endif

#endif
