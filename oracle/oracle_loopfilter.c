/*
 * oracle/oracle_loopfilter.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * CPU restatement of the pixel pass of the HEVC deblocking filter as the Turing encoder applies it to a reconstructed
 * picture (the first row "next" after the hot path):
 *   LoopFilter::Picture::deblock<edgeType>        turing/LoopFilter.h:739-777   (region walk, 4:2:0 chroma rule)
 *   LumaBlockEdge  decisions + filter             turing/LoopFilter.h:229-357   (H.265 8.7.2.5.3, 8.7.2.5.6, 8.7.2.5.7)
 *   ChromaBlockEdge                               turing/LoopFilter.h:359-423   (H.265 8.7.2.5.5, 8.7.2.5.8)
 *   Block (QP, disable bit, packed bS)            turing/LoopFilter.h:50-90
 *   betaTable / tCTable / QpC                     turing/LoopFilter.h:219-227, turing/Global.h:1417-1423
 *
 * Written from the algorithm, edge-segment-centric: an edge segment is four lines crossing a block boundary; a line is
 * addressed by a pointer to its q0 sample and the step `across` the edge, so vertical and horizontal edges share all
 * code.  Pinned against the unmodified reference templates by tests/test_oracle_pin_loopfilter.py through
 * oracle/ref_shim_loopfilter.cpp.
 */
#include "oracle.h"
#include <stdlib.h>

static const uint8_t kBeta[52] = {0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15,
                                  16, 17, 18, 20, 22, 24, 26, 28, 30, 32, 34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64};
static const uint8_t kTc[54] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1,  1,  1,  1,
                                2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24};

static int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

static int chroma_qp(int qPi)
{
    static const uint8_t mid[13] = {29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37};
    return qPi < 30 ? qPi : (qPi > 42 ? qPi - 6 : mid[qPi - 30]);
}

/* a picture plane seen through one edge direction */
typedef struct
{
    uint8_t *base; /* sample (0,0) */
    intptr_t stride;
    int bps;
} plane_t;

static int get(const plane_t *pl, intptr_t at) { return pl->bps == 1 ? pl->base[at] : ((const uint16_t *)pl->base)[at]; }
static void put(const plane_t *pl, intptr_t at, int v)
{
    if (pl->bps == 1)
        pl->base[at] = (uint8_t)v;
    else
        ((uint16_t *)pl->base)[at] = (uint16_t)v;
}

typedef struct
{
    int qp, enabled;
} side_t;

static side_t side_of(const uint8_t *block) { return (side_t){(int8_t)block[0] >> 1, !(block[0] & 1)}; }
static int bs_of(const uint8_t *block, int edgeType, int position) { return (block[1] >> (4 * edgeType + 2 * position)) & 3; }

/* one luma edge segment: q0 of line k is at `at + k * along`, sample i beyond the edge at `+ i * across` */
static void luma_segment(const plane_t *pl, intptr_t at, intptr_t across, intptr_t along, int bS, side_t P, side_t Q, int tcOffsetDiv2,
                         int betaOffsetDiv2, int bitDepth)
{
    if (!bS) return;
    const int qPL = (Q.qp + P.qp + 1) >> 1;
    const int beta = kBeta[clip3(0, 51, qPL + 2 * betaOffsetDiv2)] << (bitDepth - 8);
    const int tC = kTc[clip3(0, 53, qPL + 2 * (bS - 1) + 2 * tcOffsetDiv2)] << (bitDepth - 8);
    const int maxv = (1 << bitDepth) - 1;

    int s[4][8]; /* [line][p3 p2 p1 p0 q0 q1 q2 q3] */
    for (int k = 0; k < 4; ++k)
        for (int i = 0; i < 8; ++i) s[k][i] = get(pl, at + k * along + (i - 4) * across);
#define P_(k, i) s[k][3 - (i)]
#define Q_(k, i) s[k][4 + (i)]
    const int dp0 = abs(P_(0, 2) - 2 * P_(0, 1) + P_(0, 0)), dp3 = abs(P_(3, 2) - 2 * P_(3, 1) + P_(3, 0));
    const int dq0 = abs(Q_(0, 2) - 2 * Q_(0, 1) + Q_(0, 0)), dq3 = abs(Q_(3, 2) - 2 * Q_(3, 1) + Q_(3, 0));
    if (dp0 + dq0 + dp3 + dq3 >= beta) return;
    int strong = 1;
    for (int k = 0; k < 4; k += 3)
    {
        const int dpq = 2 * ((k ? dp3 : dp0) + (k ? dq3 : dq0));
        strong &= dpq < (beta >> 2) && abs(P_(k, 3) - P_(k, 0)) + abs(Q_(k, 0) - Q_(k, 3)) < (beta >> 3) &&
                  abs(P_(k, 0) - Q_(k, 0)) < ((5 * tC + 1) >> 1);
    }
    const int sideThreshold = (beta + (beta >> 1)) >> 3;
    const int dEp = dp0 + dp3 < sideThreshold, dEq = dq0 + dq3 < sideThreshold;

    for (int k = 0; k < 4; ++k)
    {
        const int p0 = P_(k, 0), p1 = P_(k, 1), p2 = P_(k, 2), p3 = P_(k, 3), q0 = Q_(k, 0), q1 = Q_(k, 1), q2 = Q_(k, 2), q3 = Q_(k, 3);
        const intptr_t line = at + k * along;
        if (strong)
        {
            if (P.enabled)
            {
                put(pl, line - 1 * across, clip3(p0 - 2 * tC, p0 + 2 * tC, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3));
                put(pl, line - 2 * across, clip3(p1 - 2 * tC, p1 + 2 * tC, (p2 + p1 + p0 + q0 + 2) >> 2));
                put(pl, line - 3 * across, clip3(p2 - 2 * tC, p2 + 2 * tC, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3));
            }
            if (Q.enabled)
            {
                put(pl, line, clip3(q0 - 2 * tC, q0 + 2 * tC, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3));
                put(pl, line + across, clip3(q1 - 2 * tC, q1 + 2 * tC, (p0 + q0 + q1 + q2 + 2) >> 2));
                put(pl, line + 2 * across, clip3(q2 - 2 * tC, q2 + 2 * tC, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3));
            }
        }
        else
        {
            int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
            if (abs(delta) >= tC * 10) continue;
            delta = clip3(-tC, tC, delta);
            if (P.enabled) put(pl, line - across, clip3(0, maxv, p0 + delta));
            if (Q.enabled) put(pl, line, clip3(0, maxv, q0 - delta));
            if (dEp && P.enabled) put(pl, line - 2 * across, clip3(0, maxv, p1 + clip3(-(tC >> 1), tC >> 1, (((p2 + p0 + 1) >> 1) - p1 + delta) >> 1)));
            if (dEq && Q.enabled) put(pl, line + across, clip3(0, maxv, q1 + clip3(-(tC >> 1), tC >> 1, (((q2 + q0 + 1) >> 1) - q1 - delta) >> 1)));
        }
    }
#undef P_
#undef Q_
}

/* one chroma edge segment of four lines (only bS == 2 edges are filtered) */
static void chroma_segment(const plane_t *pl, intptr_t at, intptr_t across, intptr_t along, side_t P, side_t Q, int cQpPicOffset,
                           int tcOffsetDiv2, int bitDepth)
{
    const int QpC = chroma_qp(((Q.qp + P.qp + 1) >> 1) + cQpPicOffset);
    const int tC = kTc[clip3(0, 53, QpC + 2 + 2 * tcOffsetDiv2)] << (bitDepth - 8);
    const int maxv = (1 << bitDepth) - 1;
    for (int k = 0; k < 4; ++k)
    {
        const intptr_t line = at + k * along;
        const int p1 = get(pl, line - 2 * across), p0 = get(pl, line - across), q0 = get(pl, line), q1 = get(pl, line + across);
        const int delta = clip3(-tC, tC, (((q0 - p0) << 2) + p1 - q1 + 4) >> 3);
        if (P.enabled) put(pl, line - across, clip3(0, maxv, p0 + delta));
        if (Q.enabled) put(pl, line, clip3(0, maxv, q0 - delta));
    }
}

/* LoopFilter::Picture::deblock<edgeType> over the 8x8 blocks [xBegin/8, xEnd/8) x [yBegin/8, yEnd/8), 4:2:0.
 * blocks: (data, packedBs) byte pairs, blockStride records per row; ctuOffsets: (slice_tc_offset_div2,
 * slice_beta_offset_div2) per CTU in raster order. */
void orc_deblock(void *const planes[3], const intptr_t strides[3], int bps, int bitDepthY, int bitDepthC, const uint8_t *blocks,
                 int blockStride, const int8_t *ctuOffsets, int picWidthInCtbs, int ctbLog2, int cbQpOffset, int crQpOffset,
                 int edgeType, int xBegin, int yBegin, int xEnd, int yEnd)
{
    const plane_t pl[3] = {{planes[0], strides[0], bps}, {planes[1], strides[1], bps}, {planes[2], strides[2], bps}};
    for (int by = yBegin / 8; by < yEnd / 8; ++by)
        for (int bx = xBegin / 8; bx < xEnd / 8; ++bx)
        {
            const uint8_t *q = blocks + 2 * ((intptr_t)by * blockStride + bx);
            const uint8_t *p = edgeType == 0 ? q - 2 : q - 2 * (intptr_t)blockStride; /* left / upper neighbour; only read when bS != 0 */
            const int8_t *ctu = ctuOffsets + 2 * (picWidthInCtbs * ((by << 3) >> ctbLog2) + ((bx << 3) >> ctbLog2));
            for (int position = 0; position < 2; ++position)
            {
                const int bS = bs_of(q, edgeType, position);
                if (!bS) continue;
                const intptr_t across = edgeType == 0 ? 1 : pl[0].stride, along = edgeType == 0 ? pl[0].stride : 1;
                luma_segment(&pl[0], (intptr_t)(8 * by) * pl[0].stride + 8 * bx + 4 * position * along, across, along, bS, side_of(p), side_of(q),
                             ctu[0], ctu[1], bitDepthY);
            }
            /* 4:2:0: chroma edges lie on the 8-sample chroma grid = every second luma block edge; one segment of four
             * chroma lines covers the whole 8-sample luma edge and takes the strength of its first half */
            if ((edgeType == 0 ? bx : by) & 1) continue;
            if (bs_of(q, edgeType, 0) != 2) continue;
            for (int c = 1; c < 3; ++c)
            {
                const intptr_t across = edgeType == 0 ? 1 : pl[c].stride, along = edgeType == 0 ? pl[c].stride : 1;
                chroma_segment(&pl[c], (intptr_t)(4 * by) * pl[c].stride + 4 * bx, across, along, side_of(p), side_of(q),
                               c == 1 ? cbQpOffset : crQpOffset, ctu[0], bitDepthC);
            }
        }
}

/* ---- sample adaptive offset, application ------------------------------------------------------------------------------
 *   LoopFilter::Picture::filterBlockSao          turing/LoopFilter.h:885-1017  (per CTU and component)
 *   sao_filter_edge / sao_filter_band            turing/sao.cpp:35-92          (H.265 8.7.3)
 *   restoreUnfilteredRegions                     turing/LoopFilter.h:849-877   (pcm / transquant-bypass blocks)
 *   Ctu neighbourhood (left/top/right/bottom, corner flags)   turing/LoopFilter.h:92-98, :476-534
 *
 * The reference filters a whole CTU, copies the disabled 8x8 blocks back and then "undoes" runs of samples on the CTU's
 * top row / left column / last column / last row according to which neighbouring CTUs are available.  Restated here per
 * SAMPLE of the visible picture: a sample keeps its deblocked value when its CTU's type is 0, when its 8x8 block is
 * disabled, or (edge offset only) when it lies in one of those four runs; otherwise it gets the band or edge offset.
 * Samples outside the picture (the reference's full-CTB copies and its luma-sized clip of chroma CTUs,
 * LoopFilter.h:894-896, reach there) are not produced: nothing reads them before the padding pass overwrites them.  */
static int sao_sign(int v) { return (v > 0) - (v < 0); }

void orc_sao(void *const dst[3], void *const src[3], const intptr_t strides[3], int bps, int bitDepthY, int bitDepthC, int picWidth,
             int picHeight, int ctbLog2, const uint8_t *blocks, int blockStride, const orc_sao_ctu *ctus, int lumaFlag, int chromaFlag)
{
    static const int8_t hOff[4] = {-1, 0, -1, 1}, vOff[4] = {0, -1, -1, -1};
    const int widthInCtbs = (picWidth + (1 << ctbLog2) - 1) >> ctbLog2, heightInCtbs = (picHeight + (1 << ctbLog2) - 1) >> ctbLog2;
    for (int c = 0; c < 3; ++c)
    {
        if (!(c ? chromaFlag : lumaFlag)) continue;
        const int sh = c ? 1 : 0, n = (1 << ctbLog2) >> sh, bitDepth = c ? bitDepthC : bitDepthY, maxv = (1 << bitDepth) - 1;
        const int w = picWidth >> sh, h = picHeight >> sh;
        const plane_t d = {dst[c], strides[c], bps}, s = {src[c], strides[c], bps};
        for (int ry = 0; ry < heightInCtbs; ++ry)
            for (int rx = 0; rx < widthInCtbs; ++rx)
            {
                const orc_sao_ctu *ctu = &ctus[ry * widthInCtbs + rx];
                const int type = ctu->plane[c].typeIdx, x0 = rx * n, y0 = ry * n;
                /* the four undo runs of an edge-offset CTU (LoopFilter.h:917-992) */
                int undoT = 0, undoL = 0, undoR = 0, undoB = 0, right = ctu->right >> sh, bottom = ctu->bottom >> sh;
                const int eo = ctu->plane[c].classOrBand;
                if (type == 2)
                {
                    const int availL = (ctu->left >> sh) < x0, availR = right > x0 + n, availT = (ctu->top >> sh) < y0, availB = bottom > y0 + n;
                    if (eo == 2)
                    {
                        if (!ctu->topLeft) ++undoT, ++undoL;
                        if (!ctu->bottomRight) ++undoR, ++undoB;
                    }
                    if (eo != 1)
                    {
                        if (!availL) undoL = n;
                        if (!availR) undoR = n;
                    }
                    if (eo != 0)
                    {
                        if (!availT) undoT = n;
                        if (!availB) undoB = n;
                    }
                    if (eo == 3)
                    {
                        if (ctu->topRight) --undoT, --undoR;
                        if (ctu->bottomLeft) --undoL, --undoB;
                    }
                    if (right > x0 + n) right = x0 + n;
                    if (bottom > y0 + n) bottom = y0 + n;
                }
                for (int y = y0; y < y0 + n && y < h; ++y)
                    for (int x = x0; x < x0 + n && x < w; ++x)
                    {
                        const intptr_t at = (intptr_t)y * s.stride + x;
                        const int v = get(&s, at);
                        const uint8_t *block = blocks + 2 * ((intptr_t)((y << sh) >> 3) * blockStride + ((x << sh) >> 3));
                        int out = v;
                        if (type && !(block[0] & 1))
                        {
                            if (type == 1)
                            {
                                const int k = ((v >> (bitDepth - 5)) - eo) & 31; /* band index relative to sao_band_position */
                                if (k < 4) out = clip3(0, maxv, v + ctu->plane[c].offset[k]);
                            }
                            else
                            {
                                const int lx = x - x0, ly = y - y0;
                                const int undone = (ly == 0 && lx < undoT) || (lx == 0 && ly < undoL) || (x == right - 1 && ly >= n - undoR) ||
                                                   (y == bottom - 1 && lx >= n - undoB);
                                if (!undone)
                                {
                                    const intptr_t step = (intptr_t)vOff[eo] * s.stride + hOff[eo];
                                    const int idx = 2 + sao_sign(v - get(&s, at + step)) + sao_sign(v - get(&s, at - step));
                                    /* edgeIdx 0,1,2 -> 1,2,0 (H.265 8.7.3.2): categories 1..4 index SaoOffsetVal, a flat sample gets none */
                                    static const int8_t category[5] = {1, 2, 0, 3, 4};
                                    if (category[idx]) out = clip3(0, maxv, v + ctu->plane[c].offset[category[idx] - 1]);
                                }
                            }
                        }
                        put(&d, at, out);
                    }
            }
    }
}

/* ---- sample adaptive offset, encoder-side statistics ---------------------------------------------------------------------
 *   EncSao::edge_offset_stats_class0..3     turing/EncSao.h:151-284
 *   EncSao::band_offset_luma_stats          turing/EncSao.h:111-149 (the chroma variant :62-109 is this one on U and on V, added)
 * For the interior of a block (the CTU clipped to the picture, first and last row and column left out) and each of the
 * four edge classes: per category the number of samples and the sum of (original - reconstructed); the same per band of
 * the reconstructed sample.  One pass over the block computes all five.  The reference's class-0 loop visits x = 1 of
 * every row twice, the second time with signA = -signB, i.e. as category 0 (EncSao.h:167-181): reproduced.
 * out: [class][E[5], count[5]] (40 values), band E[32], band count[32].  Returns startBand (:137-148). */
int orc_sao_stats(const void *org, intptr_t strideOrg, const void *rec, intptr_t strideRec, int w, int h, int shift, int bps, int64_t out[104])
{
    static const int8_t hOff[4] = {-1, 0, -1, 1}, vOff[4] = {0, -1, -1, -1}, category[5] = {1, 2, 0, 3, 4};
    const plane_t o = {(uint8_t *)org, strideOrg, bps}, r = {(uint8_t *)rec, strideRec, bps};
    for (int i = 0; i < 104; ++i) out[i] = 0;
    for (int y = 1; y < h - 1; ++y)
        for (int x = 1; x < w - 1; ++x)
        {
            const intptr_t at = (intptr_t)y * strideRec + x;
            const int v = get(&r, at), diff = get(&o, (intptr_t)y * strideOrg + x) - v;
            for (int c = 0; c < 4; ++c)
            {
                const intptr_t step = (intptr_t)vOff[c] * strideRec + hOff[c];
                const int cat = category[2 + sao_sign(v - get(&r, at + step)) + sao_sign(v - get(&r, at - step))];
                out[c * 10 + cat] += diff;
                out[c * 10 + 5 + cat] += 1;
            }
            if (x == 1) out[0] += diff, out[5] += 1; /* the class-0 double visit */
            const int band = v >> (3 + shift);
            out[40 + band] += diff;
            out[72 + band] += 1;
        }
    /* the densest run of four bands; first maximum wins */
    int start = 0;
    int64_t best = 0;
    for (int b = 0; b < 29; ++b)
    {
        const int64_t run = out[72 + b] + out[73 + b] + out[74 + b] + out[75 + b];
        if (run > best) best = run, start = b;
    }
    return start + 1 < 2 ? 2 : start + 1;
}
