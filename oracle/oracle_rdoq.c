/*
 * oracle/oracle_rdoq.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * CPU restatement of the reference's rate-distortion optimised quantisation
 * (turing/Rdoq.cpp:35-1023, constructor arithmetic turing/Rdoq.h:170-188) with its fixed-point
 * cost algebra (turing/FixedPoint.h, Cost.h: Cost = Q16 in int64, Lambda = Q16 in int32) and the
 * scan tables of turing/ScanOrder.h.  Pinned in tests/test_oracle_pin.py against the reference's
 * own Rdoq.cpp compiled into oracle/_ref/libhavoc_ref.so (ref_shim_rdoq.cpp).
 *
 * Three stages, as in the reference:
 *   1. walk the coefficients in reverse scan order; from the first non-zero rounding level on,
 *      pick for each coefficient the level in {q, q-1} (or {1,0} / {2,1,0}) of least D + lambda*R,
 *      tracking the CABAC level-coding state (greater1/greater2 counters, Rice parameter, context set);
 *      after each 4x4 coefficient group decide whether zeroing the whole group is cheaper;
 *   2. choose the last significant position of least total cost;
 *   3. optional sign-data hiding: fix the parity of each group's level sum at least cost.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <float.h>

typedef int64_t cost_t; /* Q16 */

/* Bit costs per CABAC state, Q15 (1 bit = 0x8000): the HM table as carried by the reference
 * (turing/Write.h:413-423, `entropyBitsHm`).  The reference indexes it with
 * ContextModel::getState() ^ bin, i.e. (state >> 1) ^ bin (Rdoq.cpp:26-31, ContextModel.h:58-61). */
static const int32_t entropy_bits[128] = {
    0x07b23, 0x085f9, 0x074a0, 0x08cbc, 0x06ee4, 0x09354, 0x067f4, 0x09c1b, 0x060b0, 0x0a62a, 0x05a9c, 0x0af5b,
    0x0548d, 0x0b955, 0x04f56, 0x0c2a9, 0x04a87, 0x0cbf7, 0x045d6, 0x0d5c3, 0x04144, 0x0e01b, 0x03d88, 0x0e937,
    0x039e0, 0x0f2cd, 0x03663, 0x0fc9e, 0x03347, 0x10600, 0x03050, 0x10f95, 0x02d4d, 0x11a02, 0x02ad3, 0x12333,
    0x0286e, 0x12cad, 0x02604, 0x136df, 0x02425, 0x13f48, 0x021f4, 0x149c4, 0x0203e, 0x1527b, 0x01e4d, 0x15d00,
    0x01c99, 0x166de, 0x01b18, 0x17017, 0x019a5, 0x17988, 0x01841, 0x18327, 0x016df, 0x18d50, 0x015d9, 0x19547,
    0x0147c, 0x1a083, 0x0138e, 0x1a8a3, 0x01251, 0x1b418, 0x01166, 0x1bd27, 0x01068, 0x1c77b, 0x00f7f, 0x1d18e,
    0x00eda, 0x1d91a, 0x00e19, 0x1e254, 0x00d4f, 0x1ec9a, 0x00c90, 0x1f6e0, 0x00c01, 0x1fef8, 0x00b5f, 0x208b1,
    0x00ab6, 0x21362, 0x00a15, 0x21e46, 0x00988, 0x2285d, 0x00934, 0x22ea8, 0x008a8, 0x239b2, 0x0081d, 0x24577,
    0x007c9, 0x24ce6, 0x00763, 0x25663, 0x00710, 0x25e8f, 0x006a0, 0x26a26, 0x00672, 0x26f23, 0x005e8, 0x27ef8,
    0x005ba, 0x284b5, 0x0055e, 0x29057, 0x0050c, 0x29bab, 0x004c1, 0x2a674, 0x004a7, 0x2aa5e, 0x0046f, 0x2b32f,
    0x0041f, 0x2c0ad, 0x003e7, 0x2ca8d, 0x003ba, 0x2d323, 0x0010c, 0x3bfbb};

static inline int32_t bits(int bin, uint8_t state) { return entropy_bits[(state >> 1) ^ bin]; }

/* ---- scan order (turing/ScanOrder.h:32-101) ------------------------------------------- */

static uint8_t scan_tab[4][3][64][2]; /* [log2 (0..3)][scanIdx][pos][x,y], blocks up to 8x8 */
static int scan_ready;

static void build_scans(void)
{
    for (int lg = 0; lg <= 3; ++lg)
    {
        const int n = 1 << lg;
        int i = 0;
        /* up-right diagonal: walk anti-diagonals from bottom-left to top-right */
        for (int d = 0; d < 2 * n - 1; ++d)
            for (int x = 0; x <= d; ++x)
            {
                int y = d - x;
                if (x < n && y < n)
                {
                    scan_tab[lg][0][i][0] = (uint8_t)x;
                    scan_tab[lg][0][i][1] = (uint8_t)y;
                    ++i;
                }
            }
        for (i = 0; i < n * n; ++i)
        {
            scan_tab[lg][1][i][0] = (uint8_t)(i % n); /* horizontal: row by row */
            scan_tab[lg][1][i][1] = (uint8_t)(i / n);
            scan_tab[lg][2][i][0] = (uint8_t)(i / n); /* vertical: column by column */
            scan_tab[lg][2][i][1] = (uint8_t)(i % n);
        }
    }
    scan_ready = 1;
}

int orc_scan_order(int log2, int scanIdx, int pos, int comp)
{
    if (!scan_ready) build_scans();
    if (log2 < 1 || log2 > 3) return 0; /* ScanOrder() returns 0 for log2 0 (:206-217); > 3 unused here */
    return scan_tab[log2][scanIdx][pos][comp];
}

/* ---- engine state ------------------------------------------------------------------------ */

typedef struct
{
    const orc_rdoq_ctx *cx;
    int32_t lambda;    /* Q16 */
    int32_t distScale; /* Q16 */
    int shdFactor;
    int iqScale, iqShift, iqOffset;
    cost_t rdCostCoeff[1024], rateCostSig[1024], distCoeff0[1024];
    cost_t rateCostCgSig[64];
    int csbf[64];
} engine;

static inline cost_t lam(const engine *e, int32_t rate) { return (cost_t)e->lambda * rate; }

static inline cost_t dist(const engine *e, int32_t err)
{
    /* int32 product first (as `err * err * scale` associates), wrapping like the reference's imul */
    int32_t sq = (int32_t)((uint32_t)err * (uint32_t)err);
    return (cost_t)sq * e->distScale;
}

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

/* Rdoq.cpp:512-598 */
static int sig_ctx(int prevCsbf, int scanIdx, int xC, int yC, int log2, int cIdx)
{
    static const uint8_t map4x4[16] = {0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8};
    int inc;
    if (log2 == 2)
        inc = map4x4[(yC << 2) + xC];
    else if (xC + yC == 0)
        inc = 0;
    else
    {
        const int xP = xC & 3, yP = yC & 3;
        if (prevCsbf == 0) inc = (xP + yP == 0) ? 2 : (xP + yP < 3) ? 1 : 0;
        else if (prevCsbf == 1) inc = (yP == 0) ? 2 : (yP == 1) ? 1 : 0;
        else if (prevCsbf == 2) inc = (xP == 0) ? 2 : (xP == 1) ? 1 : 0;
        else inc = 2;
        if (cIdx == 0)
        {
            if ((xC >> 2) + (yC >> 2) > 0) inc += 3;
            inc += log2 == 3 ? (scanIdx == 0 ? 9 : 15) : 21;
        }
        else
            inc += log2 == 3 ? 9 : 12;
    }
    return cIdx == 0 ? inc : 27 + inc;
}

/* Rdoq.cpp:601-617 / :675-697: neighbours right (bit 0) and below (bit 1) */
static int prev_csbf(const int *csbf, int xS, int yS, int log2)
{
    const int wcg = 1 << (log2 - 2);
    int v = 0;
    if (xS < wcg - 1) v += csbf[yS * wcg + xS + 1];
    if (yS < wcg - 1) v += csbf[(yS + 1) * wcg + xS] << 1;
    return v;
}

static int cg_sig_ctx(const int *csbf, int xS, int yS, int log2, int cIdx)
{
    const int wcg = 1 << (log2 - 2);
    int v = 0;
    if (xS < wcg - 1) v += csbf[yS * wcg + xS + 1];
    if (yS < wcg - 1) v += csbf[(yS + 1) * wcg + xS];
    if (v > 1) v = 1;
    return cIdx == 0 ? v : 2 + v;
}

static inline int base_level(int g1Cnt, int g2Cnt) { return g1Cnt < 8 ? 2 + (g2Cnt < 1) : 1; }

/* Rdoq.cpp:619-673: lambda * (sign bit + level bins), Q15 rate */
static cost_t level_rate_cost(const engine *e, int level, int g1Ctx, int g2Ctx, int rice, int g1Cnt, int g2Cnt)
{
    int32_t rate = 32768;
    const int base = base_level(g1Cnt, g2Cnt);
    if (level >= base)
    {
        int symbol = level - base, length;
        if (symbol < (3 << rice))
        {
            length = symbol >> rice;
            rate += (length + 1 + rice) << 15;
        }
        else
        {
            length = rice;
            symbol -= 3 << rice;
            while (symbol >= (1 << length)) symbol -= 1 << (length++);
            rate += (3 + length + 1 - rice + length) << 15;
        }
        if (g1Cnt < 8)
        {
            rate += bits(1, e->cx->greater1_flag[g1Ctx]);
            if (g2Cnt < 1) rate += bits(1, e->cx->greater2_flag[g2Ctx]);
        }
    }
    else if (level == 1)
        rate += bits(0, e->cx->greater1_flag[g1Ctx]);
    else if (level == 2)
    {
        rate += bits(1, e->cx->greater1_flag[g1Ctx]);
        rate += bits(0, e->cx->greater2_flag[g2Ctx]);
    }
    return lam(e, rate);
}

/* Rdoq.cpp:805-870: level bins only (no sign), used for the sign-hiding deltas */
static int level_rate(const engine *e, int level, int g1Ctx, int g2Ctx, int rice, int g1Cnt, int g2Cnt)
{
    static const int riceRange[5] = {7, 14, 26, 46, 78};
    static const int ricePrefix[5] = {8, 7, 6, 5, 4};
    int rate = 0;
    const int base = base_level(g1Cnt, g2Cnt);
    if (level >= base)
    {
        int symbol = level - base;
        const int maxVlc = riceRange[rice];
        if (symbol > maxVlc)
        {
            int rest = symbol - maxVlc, egs = 1;
            for (int m = 2; rest >= m; m <<= 1) egs += 2;
            rate += egs << 15;
            if (symbol > maxVlc + 1) symbol = maxVlc + 1;
        }
        int prefix = symbol >> (rice + 1);
        if (prefix > ricePrefix[rice]) prefix = ricePrefix[rice];
        rate += (prefix + rice) << 15;
        if (g1Cnt < 8)
        {
            rate += bits(1, e->cx->greater1_flag[g1Ctx]);
            if (g2Cnt < 1) rate += bits(1, e->cx->greater2_flag[g2Ctx]);
        }
    }
    else if (level == 1)
        rate += bits(0, e->cx->greater1_flag[g1Ctx]);
    else if (level == 2)
    {
        rate += bits(1, e->cx->greater1_flag[g1Ctx]);
        rate += bits(0, e->cx->greater2_flag[g2Ctx]);
    }
    return rate;
}

/* Rdoq.h:133-138 */
static int recon_level(const engine *e, int level)
{
    int v = (clip3(-32768, 32767, level) * e->iqScale + e->iqOffset) >> e->iqShift;
    return clip3(-32768, 32767, v);
}

/* Rdoq.cpp:452-510 */
static int adjust_level(engine *e, int sp, int absCoeff, int q, int sigCtx, int g1Ctx, int g2Ctx, int rice,
                        int g1Cnt, int g2Cnt, int isLast)
{
    cost_t sigCost = 0;
    int best = 0;
    if (!isLast && q < 3)
    {
        e->rateCostSig[sp] = lam(e, bits(0, e->cx->sig_coeff_flag[sigCtx]));
        e->rdCostCoeff[sp] = e->distCoeff0[sp] + e->rateCostSig[sp];
        if (q == 0) return 0;
    }
    else
        e->rdCostCoeff[sp] = INT64_MAX;
    if (!isLast) sigCost = lam(e, bits(1, e->cx->sig_coeff_flag[sigCtx]));
    const int lowest = q > 1 ? q - 1 : 1;
    for (int level = q; level >= lowest; --level)
    {
        const int32_t err = absCoeff - recon_level(e, level);
        cost_t c = dist(e, err) + level_rate_cost(e, level, g1Ctx, g2Ctx, rice, g1Cnt, g2Cnt) + sigCost;
        if (c < e->rdCostCoeff[sp])
        {
            best = level;
            e->rdCostCoeff[sp] = c;
            e->rateCostSig[sp] = sigCost;
        }
    }
    return best;
}

/* Rdoq.cpp:742-760 */
static int last_prefix_ctx(int binIdx, int cIdx, int log2)
{
    const int off = cIdx ? 15 : 3 * (log2 - 2) + ((log2 - 1) >> 2);
    const int sh = cIdx ? log2 - 2 : (log2 + 1) >> 2;
    return clip3(0, 17, (binIdx >> sh) + off);
}

/* Rdoq.cpp:699-740 */
static cost_t last_pos_cost(const engine *e, int xC, int yC, int cIdx, int log2)
{
    static const uint8_t len[32] = {0, 1, 2, 3, 4, 4, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7,
                                    8, 8, 8, 8, 8, 8, 8, 8, 9, 9, 9, 9, 9, 9, 9, 9};
    const int lx = len[xC], ly = len[yC];
    int32_t rate = 0;
    for (int i = 0; i < lx; ++i) rate += bits(1, e->cx->last_x_prefix[last_prefix_ctx(i, cIdx, log2)]);
    if (lx < 9) rate += bits(0, e->cx->last_x_prefix[last_prefix_ctx(lx, cIdx, log2)]);
    for (int i = 0; i < ly; ++i) rate += bits(1, e->cx->last_y_prefix[last_prefix_ctx(i, cIdx, log2)]);
    if (ly < 9) rate += bits(0, e->cx->last_y_prefix[last_prefix_ctx(ly, cIdx, log2)]);
    if (lx > 3) rate += 32768 * ((lx - 2) >> 1);
    if (ly > 3) rate += 32768 * ((ly - 2) >> 1);
    return lam(e, rate);
}

static void sign_data_hiding(const engine *e, int totalCg, int16_t *dst, const int16_t *src, const int *scan,
                             const int *rateUp, const int *rateDown, const int *sigDelta, const int *deltaU);

static int32_t fixed16(double d) { return (int32_t)(d * 65536 + 0.5); } /* FixedPoint<int32,16>::set(double) */

/* stage 2 (last significant position), signs, sign-data hiding: everything after the level decisions of stage 1 */
static int rdoq_finish(engine *e, int16_t *dst, const int16_t *src, const orc_rdoq_ctx *ctx, const int *scan, const int *rateUp,
                       const int *rateDown, const int *sigDelta, const int *deltaU, int log2, int cIdx, int scanIdx, int isIntra, int sdh,
                       int lastSp, int lastCg, cost_t totalDist0, cost_t rdCostTu)
{
    const int n = 1 << (2 * log2), totalCg = n >> 4, log2Cg = log2 - 2;
    int cbf = 0;
    if (lastSp >= 0)
    {
        /* ---- stage 2: last significant position (Rdoq.cpp:313-397) ---- */
        cost_t best;
        int lastIdx = 0;
        if (!isIntra && cIdx == 0)
        {
            best = totalDist0 + lam(e, bits(0, ctx->rqt_root_cbf[0]));
            rdCostTu += lam(e, bits(1, ctx->rqt_root_cbf[0]));
        }
        else
        {
            /* getCbfCtxIdx(isLuma, rqtDepth 0): luma 1, chroma 0 (Rdoq.cpp:687-697) */
            const uint8_t st = cIdx == 0 ? ctx->cbf_luma[1] : ctx->cbf_cbcr[0];
            best = totalDist0 + lam(e, bits(0, st));
            rdCostTu += lam(e, bits(1, st));
        }
        int found = 0;
        for (int cg = lastCg; cg >= 0 && !found; --cg)
        {
            const int cgX = orc_scan_order(log2Cg, scanIdx, cg, 0), cgY = orc_scan_order(log2Cg, scanIdx, cg, 1);
            const int cgPos = cgY * (1 << log2Cg) + cgX;
            rdCostTu -= e->rateCostCgSig[cg];
            if (!e->csbf[cgPos]) continue;
            for (int k = 15; k >= 0; --k)
            {
                const int sp = cg * 16 + k;
                if (sp > lastSp) continue;
                const int pos = scan[sp];
                if (dst[pos])
                {
                    const int x = pos & ((1 << log2) - 1), y = pos >> log2;
                    const cost_t lastCost = scanIdx == 2 ? last_pos_cost(e, y, x, cIdx, log2) : last_pos_cost(e, x, y, cIdx, log2);
                    const cost_t total = rdCostTu + lastCost - e->rateCostSig[sp];
                    if (total < best)
                    {
                        lastIdx = sp + 1;
                        best = total;
                    }
                    if (dst[pos] > 1)
                    {
                        found = 1;
                        break;
                    }
                    rdCostTu -= e->rdCostCoeff[sp];
                    rdCostTu += e->distCoeff0[sp];
                }
                else
                    rdCostTu -= e->rateCostSig[sp];
            }
        }

        /* signs back, uncoded tail to zero (Rdoq.cpp:414-431) */
        int absSum = 0;
        for (int sp = 0; sp < lastIdx; ++sp)
        {
            const int pos = scan[sp], level = dst[pos];
            absSum += level;
            dst[pos] = (int16_t)(src[pos] < 0 ? -level : level);
            cbf |= level;
        }
        for (int sp = lastIdx; sp <= lastSp; ++sp) dst[scan[sp]] = 0;

        if (sdh && absSum >= 2) sign_data_hiding(e, totalCg, dst, src, scan, rateUp, rateDown, sigDelta, deltaU);
    }

    return cbf;
}

int orc_rdoq(int16_t *dst, const int16_t *src, const orc_rdoq_ctx *ctx, int qScale, int qShift, int iqScale,
             int log2, int cIdx, int scanIdx, int isIntra, int sdh, int bitDepth)
{
    const int n = 1 << (2 * log2), totalCg = n >> 4, log2Cg = log2 - 2;
    engine *e = (engine *)calloc(1, sizeof(engine)); /* fresh Rdoq object per TU: all member arrays zero */
    int *rateUp = (int *)calloc(n, sizeof(int)), *rateDown = (int *)calloc(n, sizeof(int));
    int *sigDelta = (int *)calloc(n, sizeof(int)), *deltaU = (int *)calloc(n, sizeof(int));
    int *scan = (int *)malloc(n * sizeof(int));

    /* Rdoq::Rdoq (Rdoq.h:170-188) */
    e->cx = ctx;
    e->lambda = fixed16(ctx->lambda);
    e->shdFactor = (int)(iqScale * iqScale / ctx->lambda / 16 + 0.5);
    {
        const int transformShift = 15 - bitDepth - log2;
        const int distShift = 15 - 2 * transformShift - 2 * (bitDepth - 8);
        e->distScale = fixed16((double)(1 << distShift));
        e->iqScale = iqScale;
        e->iqShift = 20 - 14 - transformShift;
        e->iqOffset = 1 << (e->iqShift - 1);
    }

    for (int cg = 0, i = 0; cg < totalCg; ++cg)
        for (int k = 0; k < 16; ++k)
        {
            const int x = (orc_scan_order(log2Cg, scanIdx, cg, 0) << 2) + orc_scan_order(2, scanIdx, k, 0);
            const int y = (orc_scan_order(log2Cg, scanIdx, cg, 1) << 2) + orc_scan_order(2, scanIdx, k, 1);
            scan[i++] = (y << log2) + x;
        }

    cost_t totalDist0 = 0, rdCostTu = 0;
    int lastSp = -1, lastCg = -1;
    int ctxSet = 0, g1Idx = 1, g1Cnt = 0, g2Cnt = 0, rice = 0;
    const int g1Off = cIdx > 0 ? 16 : 0, g2Off = cIdx > 0 ? 4 : 0;

    /* ---- stage 1 (Rdoq.cpp:89-305) ---- */
    for (int cg = totalCg - 1; cg >= 0; --cg)
    {
        const int cgX = orc_scan_order(log2Cg, scanIdx, cg, 0), cgY = orc_scan_order(log2Cg, scanIdx, cg, 1);
        const int cgPos = cgY * (1 << log2Cg) + cgX;
        int nzBeforePos0 = 0;
        cost_t cgDist0 = 0, cgRateSig = 0, cgRateSigPos0 = 0, cgRdCoeff = 0;
        const int prev = prev_csbf(e->csbf, cgX, cgY, log2);

        for (int k = 15; k >= 0; --k)
        {
            const int sp = cg * 16 + k, pos = scan[sp];
            const int x = pos & ((1 << log2) - 1), y = pos >> log2;
            const int a = abs(src[pos]);
            const int scaled = a * qScale;
            const int q = (scaled + (1 << (qShift - 1))) >> qShift;
            e->distCoeff0[sp] = dist(e, a);
            totalDist0 += e->distCoeff0[sp];
            dst[pos] = (int16_t)q;

            if (q > 0 && lastSp < 0)
            {
                lastSp = sp;
                ctxSet = (sp < 16 || cIdx != 0) ? 0 : 2;
                lastCg = cg;
            }
            if (lastSp >= 0)
            {
                const int g1Ctx = 4 * ctxSet + g1Idx + g1Off, g2Ctx = ctxSet + g2Off;
                const int sc = sig_ctx(prev, scanIdx, x, y, log2, cIdx);
                const int level = adjust_level(e, sp, a, q, sc, g1Ctx, g2Ctx, rice, g1Cnt, g2Cnt, sp == lastSp);
                deltaU[pos] = (scaled - (level << qShift)) >> (qShift - 8);
                if (sp != lastSp)
                    sigDelta[pos] = bits(1, ctx->sig_coeff_flag[sc]) - bits(0, ctx->sig_coeff_flag[sc]);
                if (level > 0)
                {
                    const int now = level_rate(e, level, g1Ctx, g2Ctx, rice, g1Cnt, g2Cnt);
                    rateUp[pos] = level_rate(e, level + 1, g1Ctx, g2Ctx, rice, g1Cnt, g2Cnt) - now;
                    rateDown[pos] = level_rate(e, level - 1, g1Ctx, g2Ctx, rice, g1Cnt, g2Cnt) - now;
                }
                else
                    rateUp[pos] = bits(0, ctx->greater1_flag[g1Ctx]);
                dst[pos] = (int16_t)level;
                rdCostTu += e->rdCostCoeff[sp];

                /* updateEntropyCodingEngine (Rdoq.cpp:762-803) */
                if (level >= base_level(g1Cnt, g2Cnt) && level > 3 * (1 << rice)) rice = rice + 1 > 4 ? 4 : rice + 1;
                if (level >= 1) g1Cnt++;
                if (level > 1)
                {
                    g1Idx = 0;
                    g2Cnt++;
                }
                else if (g1Idx < 3 && g1Idx > 0 && level)
                    g1Idx++;
                if ((sp % 16 == 0) && sp > 0)
                {
                    rice = 0;
                    g1Cnt = 0;
                    g2Cnt = 0;
                    ctxSet = (sp == 16 || cIdx != 0) ? 0 : 2;
                    if (g1Idx == 0) ctxSet++;
                    g1Idx = 1;
                }
            }
            else
                rdCostTu += e->distCoeff0[sp];

            cgRateSig += e->rateCostSig[sp];
            if (k == 0) cgRateSigPos0 = e->rateCostSig[sp];
            if (dst[pos])
            {
                e->csbf[cgPos] = 1;
                cgRdCoeff += e->rdCostCoeff[sp] - e->rateCostSig[sp];
                cgDist0 += e->distCoeff0[sp];
                if (k != 0) nzBeforePos0++;
            }
        }

        /* coefficient-group zeroing (Rdoq.cpp:200-304) */
        if (lastCg >= 0)
        {
            if (cg)
            {
                if (e->csbf[cgPos] == 0)
                {
                    const int c = cg_sig_ctx(e->csbf, cgX, cgY, log2, cIdx);
                    const cost_t zero = lam(e, bits(0, ctx->coded_sub_block_flag[c]));
                    rdCostTu += zero - cgRateSig;
                    e->rateCostCgSig[cg] = zero;
                }
                else if (cg < lastCg)
                {
                    if (nzBeforePos0 == 0)
                    {
                        rdCostTu -= cgRateSigPos0;
                        cgRateSig -= cgRateSigPos0;
                    }
                    const int c = cg_sig_ctx(e->csbf, cgX, cgY, log2, cIdx);
                    const cost_t zero = lam(e, bits(0, ctx->coded_sub_block_flag[c]));
                    const cost_t one = lam(e, bits(1, ctx->coded_sub_block_flag[c]));
                    cost_t allZero = rdCostTu;
                    rdCostTu += one;
                    allZero += zero;
                    e->rateCostCgSig[cg] = one;
                    allZero += cgDist0;
                    allZero -= cgRdCoeff;
                    allZero -= cgRateSig;
                    if (allZero < rdCostTu)
                    {
                        e->csbf[cgPos] = 0;
                        rdCostTu = allZero;
                        e->rateCostCgSig[cg] = zero;
                        for (int k = 15; k >= 0; --k)
                        {
                            const int sp = cg * 16 + k, pos = scan[sp];
                            if (dst[pos])
                            {
                                dst[pos] = 0;
                                e->rdCostCoeff[sp] = e->distCoeff0[sp];
                                e->rateCostSig[sp] = 0;
                            }
                        }
                    }
                }
            }
            else
                e->csbf[cgPos] = 1;
        }
    }

    const int cbf = rdoq_finish(e, dst, src, ctx, scan, rateUp, rateDown, sigDelta, deltaU, log2, cIdx, scanIdx, isIntra, sdh, lastSp, lastCg,
                                totalDist0, rdCostTu);

    free(scan);
    free(deltaU);
    free(sigDelta);
    free(rateDown);
    free(rateUp);
    free(e);
    return cbf;
}

/* ---- the same function with stage 1 restructured by coefficient group ---------------------------------------------------
 * Claim under test (DESIGN.md 1b / roadmap 0(a)): the level decisions of a 4x4 group depend on the rest of the block only through
 * (1) one bit carried from the previous group in scan order -- whether it ended with a level above 1 (g1Idx == 0 at its last
 * coefficient, which raises the next group's context set by one) -- and (2) the coded flags of its right and below neighbours
 * (the significance-context pattern and the coded_sub_block_flag context).  The Rice parameter and the greater-1 / greater-2
 * counters restart at every group; the running cost only enters comparisons whose two sides contain it alike; all sums are
 * integers.  So every group can be walked for the 2 x 4 combinations of (carry, neighbour pattern) WITHOUT knowing the others, and
 * a cheap pass in scan order picks each group's variant and commits it.  orc_rdoq_grouped does exactly that and must equal
 * orc_rdoq bit for bit (tests/test_oracle_rdoq.py); a device kernel can then spread the walks over the lanes of a warp. */
typedef struct
{
    int16_t level[16];
    cost_t rdCostCoeff[16], rateCostSig[16];
    int deltaU[16], sigDelta[16], rateUp[16], rateDown[16];
    cost_t rdDelta, rateCostCgSig;
    int coded, carryOut;
} group_out;

static void group_walk(const engine *base, const int16_t *src, const int *scan, int qScale, int qShift, int log2, int cIdx, int scanIdx,
                       int cg, int lastSp, int lastCg, int carry, int right, int below, group_out *o)
{
    engine *e = (engine *)malloc(sizeof(engine)); /* private copy: adjust_level writes per-position members */
    memcpy(e, base, sizeof(engine));
    const orc_rdoq_ctx *ctx = e->cx;
    const int g1Off = cIdx > 0 ? 16 : 0, g2Off = cIdx > 0 ? 4 : 0;
    const int prev = right + (below << 1);
    int ctxSet = ((cg == 0 || cIdx != 0) ? 0 : 2) + carry, g1Idx = 1, g1Cnt = 0, g2Cnt = 0, rice = 0;
    int nzBeforePos0 = 0, coded = 0;
    cost_t cgDist0 = 0, cgRateSig = 0, cgRateSigPos0 = 0, cgRdCoeff = 0, rd = 0;
    memset(o, 0, sizeof(*o));
    for (int k = 15; k >= 0; --k)
    {
        const int sp = cg * 16 + k, pos = scan[sp];
        if (sp > lastSp) continue; /* the tail of the last group: accounted for before stage 1 */
        const int x = pos & ((1 << log2) - 1), y = pos >> log2;
        const int a = abs(src[pos]);
        const int scaled = a * qScale;
        const int q = (scaled + (1 << (qShift - 1))) >> qShift;
        const int g1Ctx = 4 * ctxSet + g1Idx + g1Off, g2Ctx = ctxSet + g2Off;
        const int sc = sig_ctx(prev, scanIdx, x, y, log2, cIdx);
        const int level = adjust_level(e, sp, a, q, sc, g1Ctx, g2Ctx, rice, g1Cnt, g2Cnt, sp == lastSp);
        o->deltaU[k] = (scaled - (level << qShift)) >> (qShift - 8);
        if (sp != lastSp) o->sigDelta[k] = bits(1, ctx->sig_coeff_flag[sc]) - bits(0, ctx->sig_coeff_flag[sc]);
        if (level > 0)
        {
            const int now = level_rate(e, level, g1Ctx, g2Ctx, rice, g1Cnt, g2Cnt);
            o->rateUp[k] = level_rate(e, level + 1, g1Ctx, g2Ctx, rice, g1Cnt, g2Cnt) - now;
            o->rateDown[k] = level_rate(e, level - 1, g1Ctx, g2Ctx, rice, g1Cnt, g2Cnt) - now;
        }
        else
            o->rateUp[k] = bits(0, ctx->greater1_flag[g1Ctx]);
        o->level[k] = (int16_t)level;
        rd += e->rdCostCoeff[sp];
        if (level >= base_level(g1Cnt, g2Cnt) && level > 3 * (1 << rice)) rice = rice + 1 > 4 ? 4 : rice + 1;
        if (level >= 1) g1Cnt++;
        if (level > 1)
        {
            g1Idx = 0;
            g2Cnt++;
        }
        else if (g1Idx < 3 && g1Idx > 0 && level)
            g1Idx++;
        if (k == 0) o->carryOut = g1Idx == 0; /* what the reset at a group's last coefficient hands to the next group */
        cgRateSig += e->rateCostSig[sp];
        if (k == 0) cgRateSigPos0 = e->rateCostSig[sp];
        if (level)
        {
            coded = 1;
            cgRdCoeff += e->rdCostCoeff[sp] - e->rateCostSig[sp];
            cgDist0 += e->distCoeff0[sp];
            if (k != 0) nzBeforePos0++;
        }
    }
    /* coefficient-group zeroing (Rdoq.cpp:200-304): both sides of its comparison carry the running cost, so it is local */
    if (cg)
    {
        int c = right + below;
        if (c > 1) c = 1;
        if (cIdx) c += 2;
        const cost_t zero = lam(e, bits(0, ctx->coded_sub_block_flag[c]));
        if (!coded)
        {
            rd += zero - cgRateSig;
            o->rateCostCgSig = zero;
        }
        else if (cg < lastCg)
        {
            if (nzBeforePos0 == 0)
            {
                rd -= cgRateSigPos0;
                cgRateSig -= cgRateSigPos0;
            }
            const cost_t one = lam(e, bits(1, ctx->coded_sub_block_flag[c]));
            const cost_t allZero = rd + zero + cgDist0 - cgRdCoeff - cgRateSig;
            rd += one;
            o->rateCostCgSig = one;
            if (allZero < rd)
            {
                coded = 0;
                rd = allZero;
                o->rateCostCgSig = zero;
                for (int k = 0; k < 16; ++k)
                    if (o->level[k])
                    {
                        const int sp = cg * 16 + k;
                        o->level[k] = 0;
                        e->rdCostCoeff[sp] = e->distCoeff0[sp];
                        e->rateCostSig[sp] = 0;
                    }
            }
        }
    }
    else
        coded = 1; /* the DC group is always treated as coded */
    for (int k = 0; k < 16; ++k)
    {
        o->rdCostCoeff[k] = e->rdCostCoeff[cg * 16 + k];
        o->rateCostSig[k] = e->rateCostSig[cg * 16 + k];
    }
    o->rdDelta = rd;
    o->coded = coded;
    free(e);
}

int orc_rdoq_grouped(int16_t *dst, const int16_t *src, const orc_rdoq_ctx *ctx, int qScale, int qShift, int iqScale,
                     int log2, int cIdx, int scanIdx, int isIntra, int sdh, int bitDepth)
{
    const int n = 1 << (2 * log2), totalCg = n >> 4, log2Cg = log2 - 2;
    engine *e = (engine *)calloc(1, sizeof(engine));
    int *rateUp = (int *)calloc(n, sizeof(int)), *rateDown = (int *)calloc(n, sizeof(int));
    int *sigDelta = (int *)calloc(n, sizeof(int)), *deltaU = (int *)calloc(n, sizeof(int));
    int *scan = (int *)malloc(n * sizeof(int));
    e->cx = ctx;
    e->lambda = fixed16(ctx->lambda);
    e->shdFactor = (int)(iqScale * iqScale / ctx->lambda / 16 + 0.5);
    {
        const int transformShift = 15 - bitDepth - log2;
        const int distShift = 15 - 2 * transformShift - 2 * (bitDepth - 8);
        e->distScale = fixed16((double)(1 << distShift));
        e->iqScale = iqScale;
        e->iqShift = 20 - 14 - transformShift;
        e->iqOffset = 1 << (e->iqShift - 1);
    }
    for (int cg = 0, i = 0; cg < totalCg; ++cg)
        for (int k = 0; k < 16; ++k)
        {
            const int x = (orc_scan_order(log2Cg, scanIdx, cg, 0) << 2) + orc_scan_order(2, scanIdx, k, 0);
            const int y = (orc_scan_order(log2Cg, scanIdx, cg, 1) << 2) + orc_scan_order(2, scanIdx, k, 1);
            scan[i++] = (y << log2) + x;
        }
    /* before stage 1: rounding levels, the first coded position of the reverse scan, the "all zero" distortion sums */
    cost_t totalDist0 = 0, rdCostTu = 0;
    int lastSp = -1;
    for (int sp = n - 1; sp >= 0; --sp)
    {
        const int pos = scan[sp], a = abs(src[pos]);
        const int q = (a * qScale + (1 << (qShift - 1))) >> qShift;
        e->distCoeff0[sp] = dist(e, a);
        totalDist0 += e->distCoeff0[sp];
        dst[pos] = 0;
        if (q > 0 && lastSp < 0) lastSp = sp;
        if (lastSp < 0) rdCostTu += e->distCoeff0[sp];
    }
    const int lastCg = lastSp >= 0 ? lastSp >> 4 : -1;
    /* every group, every (carry, right, below): independent of each other -- this is what a warp would do in parallel */
    group_out *var = (group_out *)malloc(sizeof(group_out) * 8 * (size_t)(lastCg + 1 > 0 ? lastCg + 1 : 1));
    for (int cg = 0; cg <= lastCg; ++cg)
        for (int v = 0; v < 8; ++v)
            group_walk(e, src, scan, qScale, qShift, log2, cIdx, scanIdx, cg, lastSp, lastCg, v & 1, (v >> 1) & 1, (v >> 2) & 1, &var[cg * 8 + v]);
    /* in scan order: pick each group's variant by what the groups before it left behind, and commit it */
    int carry = 0;
    const int wcg = 1 << log2Cg;
    for (int cg = lastCg; cg >= 0; --cg)
    {
        const int cgX = orc_scan_order(log2Cg, scanIdx, cg, 0), cgY = orc_scan_order(log2Cg, scanIdx, cg, 1);
        const int right = cgX < wcg - 1 ? e->csbf[cgY * wcg + cgX + 1] : 0, below = cgY < wcg - 1 ? e->csbf[(cgY + 1) * wcg + cgX] : 0;
        const group_out *o = &var[cg * 8 + (carry | right << 1 | below << 2)];
        for (int k = 0; k < 16; ++k)
        {
            const int sp = cg * 16 + k, pos = scan[sp];
            if (sp > lastSp) continue;
            dst[pos] = o->level[k];
            e->rdCostCoeff[sp] = o->rdCostCoeff[k];
            e->rateCostSig[sp] = o->rateCostSig[k];
            deltaU[pos] = o->deltaU[k];
            sigDelta[pos] = o->sigDelta[k];
            rateUp[pos] = o->rateUp[k];
            rateDown[pos] = o->rateDown[k];
        }
        rdCostTu += o->rdDelta;
        e->rateCostCgSig[cg] = o->rateCostCgSig;
        e->csbf[cgY * wcg + cgX] = o->coded;
        carry = o->carryOut;
    }
    free(var);
    const int cbf = rdoq_finish(e, dst, src, ctx, scan, rateUp, rateDown, sigDelta, deltaU, log2, cIdx, scanIdx, isIntra, sdh, lastSp, lastCg,
                                totalDist0, rdCostTu);
    free(scan);
    free(deltaU);
    free(sigDelta);
    free(rateDown);
    free(rateUp);
    free(e);
    return cbf;
}

/* Rdoq.cpp:889-1023 */
static void sign_data_hiding(const engine *e, int totalCg, int16_t *dst, const int16_t *src, const int *scan,
                             const int *rateUp, const int *rateDown, const int *sigDelta, const int *deltaU)
{
    int lastCG = -1;
    for (int cg = totalCg - 1; cg >= 0; --cg)
    {
        const int *sc = scan + (cg << 4);
        int firstNZ = 16, lastNZ = -1, absSum = 0;
        for (int k = 15; k >= 0; --k)
            if (dst[sc[k]])
            {
                lastNZ = k;
                break;
            }
        for (int k = 0; k < 16; ++k)
            if (dst[sc[k]])
            {
                firstNZ = k;
                break;
            }
        for (int k = firstNZ; k <= lastNZ; ++k) absSum += dst[sc[k]];
        if (lastNZ >= 0 && lastCG == -1) lastCG = 1;

        if (lastNZ - firstNZ >= 4)
        {
            const int signbit = dst[sc[firstNZ]] > 0 ? 0 : 1;
            if (signbit != (absSum & 1))
            {
                int minCost = INT_MAX, minPos = -1, finalChange = 0;
                for (int k = (lastCG == 1 ? lastNZ : 15); k >= 0; --k)
                {
                    const int pos = sc[k];
                    int cost, change;
                    if (dst[pos] != 0)
                    {
                        const int up = e->shdFactor * (-deltaU[pos]) + rateUp[pos];
                        int down = e->shdFactor * deltaU[pos] + rateDown[pos] -
                                   (abs(dst[pos]) == 1 ? ((1 << 15) + sigDelta[pos]) : 0);
                        if (lastCG == 1 && lastNZ == k && abs(dst[pos]) == 1) down -= 4 << 15;
                        if (up < down)
                        {
                            cost = up;
                            change = 1;
                        }
                        else
                        {
                            change = -1;
                            cost = (k == firstNZ && abs(dst[pos]) == 1) ? INT_MAX : down;
                        }
                    }
                    else
                    {
                        cost = e->shdFactor * (-abs(deltaU[pos])) + (1 << 15) + rateUp[pos] + sigDelta[pos];
                        change = 1;
                        if (k < firstNZ && (src[pos] >= 0 ? 0 : 1) != signbit) cost = INT_MAX;
                    }
                    if (cost < minCost)
                    {
                        minCost = cost;
                        finalChange = change;
                        minPos = pos;
                    }
                }
                if (dst[minPos] == 32767 || dst[minPos] == -32768) finalChange = -1;
                if (src[minPos] >= 0) dst[minPos] = (int16_t)(dst[minPos] + finalChange);
                else dst[minPos] = (int16_t)(dst[minPos] - finalChange);
            }
        }
        if (lastCG == 1) lastCG = 0;
    }
}
