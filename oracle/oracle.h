/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C CPU restatement of the Turing codec's pixel hot path (the havoc primitive
 * library plus the turing/Search + Rdoq loops that drive it).  It exists so that the
 * CUDA path can be checked bit-for-bit.  Nothing in the product (turingcodec_b200/,
 * include/) may include, link or call this; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do.
 *
 * Parity status: PINNED.  Every function below is checked in tests/test_oracle_pin.py
 * against oracle/_ref/libhavoc_ref.so (the unmodified reference havoc library built by
 * oracle/Makefile from /root/reference/havoc) on the reference self-test's own input
 * recipes, and against committed golden vectors in tests/golden/
 * that were generated from that reference build (tests/golden/make_golden.py).
 *
 * Conventions: strides are in SAMPLES (as in the reference), `bps` is bytes per sample
 * (1 -> uint8_t planes, 2 -> uint16_t planes).  Each function cites the reference body
 * it restates.
 */
#ifndef ORACLE_H
#define ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- distortion metrics ------------------------------------------------------------ */

/* havoc/sad.cpp:432-449 (havoc_sad_c_ref): sum |src-ref| over w x h; 16-bit result >>= 2. */
int orc_sad(const void *src, intptr_t stride_src, const void *ref, intptr_t stride_ref,
            int w, int h, int bps);

/* havoc/sad.cpp:513-542 (havoc_sad_multiref_4_c_ref): four SADs sharing one source block. */
void orc_sad_multiref4(const void *src, intptr_t stride_src, const void *const ref[4],
                       intptr_t stride_ref, int sad[4], int w, int h, int bps);

/* havoc/ssd.cpp:28-43 (havoc_ssd_c_ref): sum (a-b)^2 mod 2^32; 16-bit result >>= 4. */
uint32_t orc_ssd(const void *a, intptr_t stride_a, const void *b, intptr_t stride_b,
                 int w, int h, int bps);

/* havoc/diff.cpp:29-38 (havoc_ssd_linear_c_ref). */
int orc_ssd_linear(const uint8_t *a, const uint8_t *b, int n);

/* havoc/hadamard.cpp:58-98 (compute_satd_c_ref<n>): n = 1<<log2n in {2,4,8}. */
int orc_hadamard_satd(const void *a, intptr_t stride_a, const void *b, intptr_t stride_b,
                      int log2n, int bps);

/* turing/Measure.h:96-135 (measureSatd): tiles a w x h block with 8x8, 4x4 or 2x2 SATDs. */
int32_t orc_measure_satd(const void *a, intptr_t stride_a, const void *b, intptr_t stride_b,
                         int w, int h, int bps);

/* ---- inter prediction --------------------------------------------------------------- */

/* havoc/pred_inter.cpp:39-69 (havoc_pred_coefficient). */
int orc_pred_coefficient(int taps, int frac, int k);

/* havoc/pred_inter.cpp:76-202: copy / h / v / hv uni-prediction, taps = 8 (luma) or 4 (chroma). */
void orc_pred_uni(void *dst, intptr_t stride_dst, const void *ref, intptr_t stride_ref,
                  int w, int h, int xFrac, int yFrac, int bitDepth, int taps, int bps);

/* havoc/pred_inter.cpp:1207-1252 (havocPredBi_c_ref). */
void orc_pred_bi(void *dst, intptr_t stride_dst, const void *ref0, const void *ref1,
                 intptr_t stride_ref, int w, int h, int xFrac0, int yFrac0, int xFrac1, int yFrac1,
                 int bitDepth, int taps, int bps);

/* havoc/pred_inter.cpp:2063-2080 (subtractBi_c_ref): dst = clip(2*src - pred). */
void orc_subtract_bi(void *dst, intptr_t stride_dst, const void *pred, intptr_t stride_pred,
                     const void *src, intptr_t stride_src, int w, int h, int bitDepth, int bps);

/* ---- intra prediction ---------------------------------------------------------------- */

/* havoc/pred_intra.cpp:43-51,76-98,20282-20401: planar(0) / DC(1) / angular(2..34).
 * `neighbours` points at the sample above-left+1, i.e. p(x,y) = neighbours[x - y - 1].
 * edge_flag selects the cIdx==0 && log2<5 filtered variants (pred_intra.h:41-49). */
void orc_pred_intra(void *dst, intptr_t stride_dst, const void *neighbours, int mode,
                    int log2n, int bitDepth, int edge_flag, int bps);

/* ---- transform / quantisation -------------------------------------------------------- */

/* havoc/transform.cpp:3071-3397: forward DCT-II (trType 0, log2n 2..5) or DST-VII (trType 1, 4x4).
 * Each pass result is truncated to int16 with wrap-around (shiftRight, :3071-3084). */
void orc_transform_fwd(int16_t *coeffs, const int16_t *src, intptr_t stride_src,
                       int trType, int log2n, int bitDepth);

/* havoc/transform.cpp:50-401 + transform.h:104-114: inverse transform, add to pred, clip. */
void orc_inverse_transform_add(void *dst, intptr_t stride_dst, const void *pred, intptr_t stride_pred,
                               const int16_t *coeffs, int trType, int log2n, int bitDepth, int bps);

/* residual-only inverse (havoc::inverse_transform, transform.cpp:2861-2868 table) */
void orc_inverse_transform(int16_t *res, const int16_t *coeffs, int trType, int log2n, int bitDepth);

/* havoc/quantize.cpp:278-304 (havoc_quantize_c_ref); returns OR of outputs (only ==0 is meaningful). */
int orc_quantize(int16_t *dst, const int16_t *src, int scale, int shift, int offset, int n);

/* havoc/quantize.cpp:37-46 (havoc_quantize_inverse_c_ref). */
void orc_quantize_inverse(int16_t *dst, const int16_t *src, int scale, int shift, int n);

/* havoc/quantize.cpp:538-548 (havoc_quantize_reconstruct_c_ref). */
void orc_quantize_reconstruct(uint8_t *rec, intptr_t stride_rec, const uint8_t *pred,
                              intptr_t stride_pred, const int16_t *res, int n);

/* ---- motion search (turing/Search.hpp) ------------------------------------------------ */

typedef struct
{
    int16_t x, y;
} orc_mv;

typedef struct
{
    /* block, in luma samples relative to the plane origin */
    int x0, y0, w, h;
    /* AMVP predictors (quarter-pel) and the rate of mvp_lX_flag = 0 / 1 as Cost (Q16, int64) */
    orc_mv mvp[2];
    int64_t rateMvpFlag[2];
    /* Lambda (Q16 int32) = FixedPoint<int32,16>::set(getReciprocalSqrtLambda) */
    int32_t lambda;
    /* LimitFullPelMv (Search.hpp:1366-1407): inclusive full-pel limits */
    orc_mv limitMin, limitMax;
    /* switches (Search.hpp:2088-2094, :2116-2126) */
    int smallSearchWindow; /* speed->useSmallSearchWindow() */
    int met;               /* stateEncode->met */
    int log2CbSize;
    int usePrev2Nx2N;      /* PartMode != 2Nx2N || cqtDepth != 0 */
    orc_mv prev2Nx2N;      /* mvPreviousInteger2Nx2N[refList] */
    int halfPel, quarterPel; /* Speed::doHalfPelRefinement / doQuarterPelRefinement */
    int bitDepth;
} orc_me_task;

typedef struct
{
    orc_mv mv, mvd;      /* after sub-pel refinement when enabled */
    orc_mv mvInteger;    /* best integer vector (what mvPreviousInteger2Nx2N would receive) */
    int earlyExit;       /* 1: fullPelMotionEstimation returned through MET (Search.hpp:2125) */
    int64_t cost;        /* best integer candidate cost (Q16) */
    int mvpFlag;
    int64_t costMvdZero[2];
    int64_t subpelCost;  /* bestCost after subPelRefinement (valid when halfPel) */
    int nSad;            /* number of SAD evaluations performed (statistics only) */
} orc_me_result;

/* turing/Search.hpp:2064-2336 (fullPelMotionEstimation) followed by :2339-2357 (subPelRefinement)
 * as chained in searchMotionUni (:1315-1352).  src/ref are plane origins (sample (0,0)). */
void orc_me_search(const void *srcPlane, intptr_t stride_src, const void *refPlane, intptr_t stride_ref,
                   const orc_me_task *task, orc_me_result *out, int bps);

/* one searchMotionBi call (turing/Search.hpp:1498-1653): refine the vector of list X of a bi-predicted PU */
typedef struct
{
    int x0, y0, w, h;
    orc_mv mvOther;  /* puData.mv(1 - X): the other list's vector, quarter-pel */
    orc_mv mvStart;  /* puData.mv(X): where the refinement starts (the uni-directional result) */
    orc_mv mvp[2];
    int64_t rateMvpFlag[2];
    int32_t lambda;  /* Lambda::set(getReciprocalSqrtLambda * 0.5) */
    orc_mv limitMin, limitMax;
    int smallWindow; /* speed->useBiSmallSearchWindow(): range 1 instead of 5 */
    int halfPel, quarterPel, bitDepth;
} orc_me_bi_task;

typedef struct
{
    orc_mv mv, mvd, mvInteger;
    int mvpFlag;
    int64_t cost;
    int nSad;
} orc_me_bi_result;

/* ref = list X's reference plane, other = list (1-X)'s; all three are plane origins */
void orc_me_bi_search(const void *srcPlane, intptr_t stride_src, const void *refPlane, intptr_t stride_ref,
                      const void *otherPlane, intptr_t stride_other, const orc_me_bi_task *task,
                      orc_me_bi_result *out, int bps);

/* the distortion half of measurePuCost (turing/Search.hpp:1668-1682): predictInter + SATD of Y, Cb, Cr */
typedef struct
{
    const void *p;   /* sample (0,0) of a padded plane */
    intptr_t stride; /* in samples */
} orc_plane;

typedef struct
{
    int x0, y0, w, h;       /* PU, luma samples */
    int predFlag[2];
    orc_mv mv[2];           /* quarter-pel luma vectors */
    int picWidth, picHeight;
    int bitDepthY, bitDepthC;
} orc_pu_cost_task;

/* predOut (optional, may be NULL / hold NULLs): receives the predicted w x h (w/2 x h/2) blocks, contiguous */
void orc_pu_cost(const orc_plane src[3], const orc_plane ref0[3], const orc_plane ref1[3], const orc_pu_cost_task *task,
                 int32_t satd[3], void *predOut[3], int bps);

/* turing/Measure.h:177-220: rateOf(mvd) as a Q16 Cost. */
int64_t orc_rate_of_mvd(int dx, int dy);

#ifdef __cplusplus
}
#endif

/* ---- RDOQ (turing/Rdoq.cpp) ------------------------------------------------------------ */
#ifdef __cplusplus
extern "C" {
#endif

/* CABAC context states (ContextModel::state bytes, turing/ContextModel.h:30) that
 * Rdoq::runQuantisation reads through estimateBits (Rdoq.cpp:26-31), and lambda as passed to the
 * Rdoq constructor (Rdoq.h:170).  Same byte layout as hvb_rdoq_ctx in include/hvb.h. */
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
    double lambda;
} orc_rdoq_ctx;

/* turing/Rdoq.cpp:35-450 (runQuantisation) incl. sign-data hiding (:889-1023), with the constructor
 * arithmetic of Rdoq.h:170-188.  Returns the OR of the coded levels (0 = no coded coefficient). */
int orc_rdoq(int16_t *dst, const int16_t *src, const orc_rdoq_ctx *ctx, int quantiserScale,
             int quantiserShift, int invQuantScale, int log2n, int cIdx, int scanIdx, int isIntra,
             int sdh, int bitDepth);
/* the same function with stage 1 computed group by group for every (carry, right, below) and selected afterwards: must equal
 * orc_rdoq bit for bit -- the decomposition a warp-parallel RDOQ walk rests on (oracle_rdoq.c, DESIGN.md roadmap 0(a)) */
int orc_rdoq_grouped(int16_t *dst, const int16_t *src, const orc_rdoq_ctx *ctx, int quantiserScale, int quantiserShift, int inverseScale,
                     int log2TrafoSize, int cIdx, int scanIdx, int isIntra, int sdh, int bitDepth);

/* turing/ScanOrder.h: x (comp 0) / y (comp 1) of scan position `pos` in a (1<<log2)^2 block */
int orc_scan_order(int log2, int scanIdx, int pos, int comp);

/* ---- in-loop deblocking, pixel pass ---------------------------------- */

/* turing/LoopFilter.h:739-777 (Picture::deblock<edgeType>) with :229-423 (Luma/ChromaBlockEdge): filters the edges of
 * type `edgeType` (0 vertical, 1 horizontal) of the 8x8 blocks [xBegin/8, xEnd/8) x [yBegin/8, yEnd/8) in place, 4:2:0.
 * blocks = (data, packedBs) byte pairs of LoopFilter::Block (:50-90), blockStride records per row; ctuOffsets =
 * (slice_tc_offset_div2, slice_beta_offset_div2) per CTU, raster order.  Pinned: tests/test_oracle_pin_loopfilter.py. */
void orc_deblock(void *const planes[3], const intptr_t strides[3], int bps, int bitDepthY, int bitDepthC, const uint8_t *blocks,
                 int blockStride, const int8_t *ctuOffsets, int picWidthInCtbs, int ctbLog2, int cbQpOffset, int crQpOffset,
                 int edgeType, int xBegin, int yBegin, int xEnd, int yEnd);

/* One CTU's SAO parameters and neighbourhood as LoopFilter::Ctu holds them (turing/LoopFilter.h:92-163, :476-534):
 * left/top/right/bottom in luma samples, the four corner availabilities, per component SaoTypeIdx, SaoEoClass or
 * sao_band_position, and SaoOffsetVal[1..4] (already scaled by 1 << (bitDepth - min(bitDepth, 10))). */
typedef struct
{
    int16_t left, top, right, bottom;
    uint8_t topLeft, topRight, bottomLeft, bottomRight;
    struct
    {
        int8_t typeIdx, classOrBand;
        int16_t offset[4];
    } plane[3];
} orc_sao_ctu; /* 42 bytes */

/* turing/LoopFilter.h:885-1017 (filterBlockSao) over every CTU of the picture, in applySaoCTU's order (:794-811): SAO of
 * `src` (the deblocked picture) into `dst`, visible samples only.  Pinned: tests/test_oracle_pin_loopfilter.py. */
void orc_sao(void *const dst[3], void *const src[3], const intptr_t strides[3], int bps, int bitDepthY, int bitDepthC, int picWidth,
             int picHeight, int ctbLog2, const uint8_t *blocks, int blockStride, const orc_sao_ctu *ctus, int lumaFlag, int chromaFlag);

/* turing/EncSao.h:111-284: SAO statistics of the interior of a w x h block (the four edge classes' per-category counts and
 * sums of original - reconstructed, the same per band); out = [class][E[5], count[5]], band E[32], band count[32]; returns
 * startBand.  Pinned: tests/test_oracle_pin_loopfilter.py. */
int orc_sao_stats(const void *org, intptr_t strideOrg, const void *rec, intptr_t strideRec, int w, int h, int shift, int bps, int64_t out[104]);

/* ---- coded-data feed ------------------------------------------------------------- */

/* turing/CodedData.h:457-517 (storeResidual): the uint16 record of a (1 << log2n)^2 block of quantised levels (raster
 * order) in the encoder's coded-data stream; returns its length in words, 0 for an all-zero block (nothing written).
 * `out` needs 5 + 19 * (number of 4x4 sub-blocks) words at most.  Pinned: tests/test_oracle_pin_codeddata.py. */
int orc_coded_residual(const int16_t *levels, int log2n, int scanIdx, uint16_t *out);

/* ---- pre-analysis ------------------------------------------------------------------ */

/* turing/EstimateIntraComplexity.h:55-157 (computeSatd8x8): AC Hadamard energy of an 8x8 block of source samples. */
int orc_intra_complexity_8x8(const void *p, intptr_t stride, int bps);
/* turing/EstimateIntraComplexity.h:159-176 (preAnalysis): the same for every whole 8x8 block of a plane, raster order;
 * returns the sum.  Pinned: tests/test_oracle_pin_preanalysis.py. */
int orc_intra_complexity(const void *plane, intptr_t stride, int width, int height, int bps, int32_t *out);
/* turing/AdaptiveQuantisation.h:172-246 (preAnalysis), one layer: per unit the smallest quadrant variance (activity - 1);
 * returns the layer's truncated average activity. */
int orc_aq_layer(const void *plane, intptr_t stride, int width, int height, int bps, int unit, int64_t *out);
/* turing/SCDetection.h:237-262: 64-bin luma histogram; :71-147: block statistics of getLikelihood; :40-47,:149-178: the ratio. */
void orc_scd_histogram(const void *plane, intptr_t stride, int width, int height, int bps, int32_t *hist);
int orc_scd_block_stats(const void *plane, intptr_t stride, int width, int height, int bps, int margin, double *out);
double orc_scd_likelihood(const double *prev, const double *cur);

#ifdef __cplusplus
}
#endif

/* turing/Dsp.h:57-70 (filterFlag) as a rule, and turing/IntraReferenceSamples.h:373-419 (filter) on a
 * 4n+1 neighbour array whose corner sits at index 2n (left(y) at 2n-1-y, top(x) at 2n+1+x), samples widened to u16. */
int orc_intra_filter_flag(int cIdx, int mode, int n);
void orc_intra_filter_neighbours(uint16_t *f, const uint16_t *u, int n, int bitDepth, int strongEnabled);

#endif /* ORACLE_H */
