/*
 * oracle/oracle_search.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * CPU restatement of the reference's uni-directional motion search for one prediction unit:
 *   fullPelMotionEstimation   turing/Search.hpp:2064-2336
 *   StateMeFullPel::considerPattern / LimitFullPelMv / MvCandidate   :1254-1312, :1366-1495
 *   subPelRefinement / patternSearch / costMv / costDistortionMv     :1965-2060, :2339-2357
 *   rateOf / estimateRateOfMvdComponent                              turing/Measure.h:177-220
 * with the fixed-point cost algebra of turing/FixedPoint.h (Cost = Q16 int64, Lambda = Q16 int32)
 * and int16 motion-vector arithmetic of turing/MotionVector.h.
 *
 * Everything the reference reads from encoder state is an explicit field of orc_me_task.
 *
 * PARITY STATUS OF THIS FILE: pinned.  tests/test_oracle_search_pin.py runs the reference's own templates
 * (oracle/ref_shim_search.cpp instantiates them unmodified with a state-only stand-in handler) on the same
 * pictures and per-PU state and requires every output to be equal.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#include <limits.h>

typedef struct
{
    orc_mv mv, mvd;
    int64_t cost;
    int mvpFlag;
} cand;

typedef struct
{
    const orc_me_task *t;
    const char *src, *ref; /* block origins: sample (x0, y0) of each plane */
    intptr_t ss, sr;
    int bps;
    cand best;
    int nSad;
} search;

/* Measure.h:177-206: bit length of |d| (0 for d == 0) */
static unsigned mvd_component_rate(int d)
{
    unsigned u = (unsigned)abs(d), r = 0;
    while (u)
    {
        ++r;
        u >>= 1;
    }
    return r;
}

/* Measure.h:210-216: Cost::make(rx + ry + 1, -1) = (rx + ry + 1) << 17 */
int64_t orc_rate_of_mvd(int dx, int dy)
{
    return (int64_t)(mvd_component_rate(dx) + mvd_component_rate(dy) + 1) << 17;
}

static const void *ref_at(const search *s, int mvx, int mvy) /* full-pel displacement */
{
    return s->ref + ((intptr_t)mvy * s->sr + mvx) * s->bps;
}

static void limit(const search *s, orc_mv *mv)
{
    const orc_me_task *t = s->t;
    if (mv->x < t->limitMin.x) mv->x = t->limitMin.x;
    if (mv->y < t->limitMin.y) mv->y = t->limitMin.y;
    if (mv->x > t->limitMax.x) mv->x = t->limitMax.x;
    if (mv->y > t->limitMax.y) mv->y = t->limitMax.y;
}

/* MvCandidate(h, refList, mv, predictors) (Search.hpp:1262-1298): cheaper of the two predictors */
static cand make_candidate(const search *s, orc_mv mv)
{
    const orc_me_task *t = s->t;
    cand c;
    c.mvpFlag = 0;
    c.mvd.x = (int16_t)(mv.x - t->mvp[0].x);
    c.mvd.y = (int16_t)(mv.y - t->mvp[0].y);
    c.cost = orc_rate_of_mvd(c.mvd.x, c.mvd.y) + t->rateMvpFlag[0];
    {
        orc_mv d1 = {(int16_t)(mv.x - t->mvp[1].x), (int16_t)(mv.y - t->mvp[1].y)};
        int64_t c1 = orc_rate_of_mvd(d1.x, d1.y) + t->rateMvpFlag[1];
        if (c1 < c.cost)
        {
            c.mvpFlag = 1;
            c.mvd = d1;
            c.cost = c1;
        }
    }
    c.mv = mv;
    return c;
}

static int consider(search *s, const cand *c)
{
    if (c->cost < s->best.cost)
    {
        s->best = *c;
        return 1;
    }
    return 0;
}

/* bench only (oracle_bench.c): route the pixel primitives through the reference's own tables */
int orc_hook_active(void);
int orc_hook_sad(const void *a, intptr_t sa, const void *b, intptr_t sb, int w, int h, int bps);
void orc_hook_pred_uni(void *dst, intptr_t sd, const void *ref, intptr_t sr, int w, int h, int xf, int yf, int bd, int bps);
int orc_hook_satd_tile(const void *a, intptr_t sa, const void *b, intptr_t sb, int log2n, int bps);

static int sad_at(search *s, int mvx, int mvy)
{
    s->nSad++;
    if (orc_hook_active()) return orc_hook_sad(s->src, s->ss, ref_at(s, mvx, mvy), s->sr, s->t->w, s->t->h, s->bps);
    return orc_sad(s->src, s->ss, ref_at(s, mvx, mvy), s->sr, s->t->w, s->t->h, s->bps);
}

/* StateMeFullPel::considerPattern (Search.hpp:1447-1482).  origin is in quarter-samples, pattern
 * entries are scaled by dist and divided by 4 with C truncation; groups of four share one SAD4 call. */
static int consider_pattern(search *s, orc_mv origin, const int8_t (*pattern)[2], int n, int step, int dist)
{
    int improved = 0;
    for (int j = 0; j < n; j += 4 * step)
        for (int i = 0; i < 4; ++i, pattern += step)
        {
            orc_mv mv;
            mv.x = (int16_t)((origin.x + dist * (*pattern)[0]) / 4);
            mv.y = (int16_t)((origin.y + dist * (*pattern)[1]) / 4);
            limit(s, &mv);
            const int sad = sad_at(s, mv.x, mv.y);
            mv.x = (int16_t)(mv.x * 4);
            mv.y = (int16_t)(mv.y * 4);
            cand c = make_candidate(s, mv);
            c.cost += (int64_t)s->t->lambda * sad;
            improved |= consider(s, &c);
        }
    return improved;
}

static const int8_t kDiamond4[4][2] = {{-4, 0}, {0, 4}, {4, 0}, {0, -4}};
static const int8_t kHexagon8[8][2] = {{0, -8}, {8, -4}, {8, 4}, {0, 8}, {-8, 4}, {-8, -4}, {-8, 4}, {-8, -4}};
static const int8_t kDiamond16[16][2] = {{0, -4}, {1, -3}, {2, -2}, {3, -1}, {4, 0},  {3, 1},   {2, 2},   {1, 3},
                                         {0, 4},  {-1, 3}, {-2, 2}, {-3, 1}, {-4, 0}, {-3, -1}, {-2, -2}, {-1, -3}};
static const int8_t kSquare4[4][2] = {{-4, -4}, {-4, 4}, {4, 4}, {4, -4}};
static const int8_t kLine4[4][2] = {{0, 0}, {1, 0}, {2, 0}, {3, 0}};

/* the early-termination probe after an improving start candidate (Search.hpp:2110-2126 and twice more) */
static int met_terminates(search *s)
{
    int trigger = !consider_pattern(s, s->best.mv, kDiamond4, 4, 1, 1);
    if (trigger && s->t->log2CbSize >= 5) trigger = !consider_pattern(s, s->best.mv, kHexagon8, 8, 1, 1);
    return trigger;
}

/* returns 1 when the search returned early through MET (the caller's mvPreviousInteger2Nx2N is then
 * NOT updated, Search.hpp:2125 `return`) */
static int full_pel(search *s, orc_me_result *out)
{
    const orc_me_task *t = s->t;
    const int window = t->smallSearchWindow ? 32 : 64;
    const int maxCounter = t->smallSearchWindow ? 2 : 3;
    const int raster = t->smallSearchWindow ? 120 : 240;

    /* zero vector (:2103-2129) -- not clamped */
    {
        orc_mv zero = {0, 0};
        cand c = make_candidate(s, zero);
        c.cost += (int64_t)t->lambda * sad_at(s, 0, 0);
        if (consider(s, &c) && t->met && met_terminates(s)) return 1;
    }
    /* the two predictors rounded to full-pel (:2131-2171) */
    for (int flag = 0; flag < 2; ++flag)
    {
        cand c;
        c.mvpFlag = flag;
        c.mv.x = (int16_t)((int16_t)(t->mvp[flag].x + 1) >> 2);
        c.mv.y = (int16_t)((int16_t)(t->mvp[flag].y + 1) >> 2);
        limit(s, &c.mv);
        c.mv.x = (int16_t)(c.mv.x << 2);
        c.mv.y = (int16_t)(c.mv.y << 2);
        c.mvd.x = (int16_t)(c.mv.x - t->mvp[flag].x);
        c.mvd.y = (int16_t)(c.mv.y - t->mvp[flag].y);
        c.cost = orc_rate_of_mvd(c.mvd.x, c.mvd.y) + t->rateMvpFlag[flag];
        c.cost += (int64_t)t->lambda * sad_at(s, c.mv.x >> 2, c.mv.y >> 2);
        out->costMvdZero[flag] = c.cost;
        if (consider(s, &c) && t->met && met_terminates(s)) return 1;
    }
    /* previous 2Nx2N integer vector (:2173-2198) */
    if (t->usePrev2Nx2N)
    {
        orc_mv mv = {(int16_t)(t->prev2Nx2N.x >> 2), (int16_t)(t->prev2Nx2N.y >> 2)};
        limit(s, &mv);
        mv.x = (int16_t)(mv.x << 2);
        mv.y = (int16_t)(mv.y << 2);
        cand c = make_candidate(s, mv);
        c.cost += (int64_t)t->lambda * sad_at(s, c.mv.x >> 2, c.mv.y >> 2);
        if (consider(s, &c) && t->met && met_terminates(s)) return 1;
    }

    /* star search (:2202-2247) */
    orc_mv start = s->best.mv;
    int distBest = 0, counter = 0, step = 4;
    for (int dist = 1; dist <= window && counter < maxCounter; dist <<= 1)
    {
        if (dist == 2 || dist == 8) step >>= 1;
        if (consider_pattern(s, start, kDiamond16, 16, step, dist))
        {
            distBest = dist;
            counter = 0;
        }
        else
            ++counter;
    }
    if (distBest == 1)
    {
        distBest = 0;
        consider_pattern(s, s->best.mv, kSquare4, 4, 1, 1);
    }
    /* raster (:2258-2273): absolute displacements, every 5th sample */
    if (distBest > 5)
    {
        orc_mv mv;
        for (mv.y = (int16_t)-raster; mv.y <= raster; mv.y = (int16_t)(mv.y + 20))
            for (mv.x = (int16_t)-raster; mv.x <= raster; mv.x = (int16_t)(mv.x + 80))
                consider_pattern(s, mv, kLine4, 4, 1, 20);
        distBest = 5;
    }
    /* star refinement (:2276-2302) */
    while (distBest > 0)
    {
        start = s->best.mv;
        distBest = 0;
        step = 4;
        for (int dist = 1; dist <= window; dist <<= 1)
        {
            if (dist == 2 || dist == 8) step >>= 1;
            if (consider_pattern(s, start, kDiamond16, 16, step, dist)) distBest = dist;
        }
        if (distBest == 1)
        {
            consider_pattern(s, start, kSquare4, 4, 1, 1);
            distBest = 0;
        }
    }
    /* one-sample diamond until no improvement (:2303-2334) */
    if (!t->smallSearchWindow)
    {
        static const int8_t d1[4][2] = {{0, -1}, {-1, 0}, {0, 1}, {1, 0}};
        int j;
        do
        {
            cand c[4];
            int sad[4];
            for (int i = 0; i < 4; ++i)
            {
                orc_mv mv = {(int16_t)(s->best.mv.x / 4 + d1[i][0]), (int16_t)(s->best.mv.y / 4 + d1[i][1])};
                limit(s, &mv);
                sad[i] = sad_at(s, mv.x, mv.y);
                mv.x = (int16_t)(mv.x * 4);
                mv.y = (int16_t)(mv.y * 4);
                c[i].mv = mv;
            }
            j = -1;
            for (int i = 0; i < 4; ++i)
            {
                cand full = make_candidate(s, c[i].mv);
                full.cost += (int64_t)t->lambda * sad[i];
                if (consider(s, &full)) j = i;
            }
        } while (j >= 0);
    }
    return 0;
}

/* costMv (Search.hpp:2003-2008): rateOf(mvd) + lambda * SATD(src, interpolated prediction) */
static int64_t cost_mv(search *s, orc_mv mv, orc_mv mvd)
{
    const orc_me_task *t = s->t;
    uint16_t pred[64 * 64 + 64] __attribute__((aligned(32))); /* slack: the reference's JIT may write right of the block */
    int satd;
    if (orc_hook_active())
    {
        orc_hook_pred_uni(pred, 64, ref_at(s, mv.x >> 2, mv.y >> 2), s->sr, t->w, t->h, mv.x & 3, mv.y & 3, t->bitDepth, s->bps);
        /* measureSatd tiling (turing/Measure.h:96-135) over the reference's hadamard table */
        const int lg = ((t->w | t->h) & 3) ? 1 : (((t->w | t->h) & 7) ? 2 : 3), n = 1 << lg;
        satd = 0;
        for (int y = 0; y < t->h; y += n)
            for (int x = 0; x < t->w; x += n)
                satd += orc_hook_satd_tile(s->src + ((intptr_t)y * s->ss + x) * s->bps, s->ss, (const char *)pred + (y * 64 + x) * s->bps,
                                           64, lg, s->bps);
    }
    else
    {
        orc_pred_uni(pred, 64, ref_at(s, mv.x >> 2, mv.y >> 2), s->sr, t->w, t->h, mv.x & 3, mv.y & 3, t->bitDepth, 8, s->bps);
        satd = orc_measure_satd(s->src, s->ss, pred, 64, t->w, t->h, s->bps);
    }
    return orc_rate_of_mvd(mvd.x, mvd.y) + (int64_t)t->lambda * satd;
}

/* patternSearch with maxIterations = 1 (Search.hpp:2011-2060) */
static void pattern_search(search *s, const int8_t (*pattern)[2], int tryOrigin, orc_mv *mv, orc_mv *mvd, int64_t *bestCost)
{
    if (tryOrigin) *bestCost = cost_mv(s, *mv, *mvd);
    int best = -1;
    for (int i = 0; i < 8; ++i)
    {
        orc_mv m = {(int16_t)(mv->x + pattern[i][0]), (int16_t)(mv->y + pattern[i][1])};
        orc_mv d = {(int16_t)(mvd->x + pattern[i][0]), (int16_t)(mvd->y + pattern[i][1])};
        const int64_t c = cost_mv(s, m, d);
        if (c < *bestCost)
        {
            best = i;
            *bestCost = c;
        }
    }
    if (best >= 0)
    {
        mv->x = (int16_t)(mv->x + pattern[best][0]);
        mv->y = (int16_t)(mv->y + pattern[best][1]);
        mvd->x = (int16_t)(mvd->x + pattern[best][0]);
        mvd->y = (int16_t)(mvd->y + pattern[best][1]);
    }
}

void orc_me_search(const void *srcPlane, intptr_t ss, const void *refPlane, intptr_t sr, const orc_me_task *t,
                   orc_me_result *out, int bps)
{
    static const int8_t half[8][2] = {{-2, -2}, {0, -2}, {2, -2}, {-2, 0}, {2, 0}, {-2, 2}, {0, 2}, {2, 2}};
    static const int8_t quarter[8][2] = {{-1, -1}, {0, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};
    search s;
    s.t = t;
    s.bps = bps;
    s.ss = ss;
    s.sr = sr;
    s.src = (const char *)srcPlane + ((intptr_t)t->y0 * ss + t->x0) * bps;
    s.ref = (const char *)refPlane + ((intptr_t)t->y0 * sr + t->x0) * bps;
    s.best.cost = INT64_MAX;
    s.best.mv.x = s.best.mv.y = s.best.mvd.x = s.best.mvd.y = 0;
    s.best.mvpFlag = 0;
    s.nSad = 0;
    out->costMvdZero[0] = out->costMvdZero[1] = 0;
    out->subpelCost = 0;

    const int early = full_pel(&s, out);
    out->cost = s.best.cost;
    out->mvpFlag = s.best.mvpFlag;
    out->mvInteger = s.best.mv;
    out->earlyExit = early;

    orc_mv mv = s.best.mv, mvd = s.best.mvd;
    if (t->halfPel)
    {
        int64_t bestCost = 0;
        pattern_search(&s, half, 1, &mv, &mvd, &bestCost);
        if (t->quarterPel) pattern_search(&s, quarter, 0, &mv, &mvd, &bestCost);
        out->subpelCost = bestCost;
    }
    out->mv = mv;
    out->mvd = mvd;
    out->nSad = s.nSad;
}

/* ---- searchMotionBi (turing/Search.hpp:1498-1653) ------------------------------------------------
 * Refines the vector of one list of a bi-predicted PU against the "ideal" block 2*src - predOther:
 *   :1518-1534  uni prediction from the OTHER list (integer part clamped by LimitFullPelMv)
 *   :1541-1548  SubtractBi at bit depth 6 + 2*sizeof(Sample)  (8 for u8, 10 for u16 whatever BitDepthY is)
 *   :1551-1625  exhaustive (2r+1)^2 integer grid, r = 1 (small window) or 5; SADs come from SAD4 calls on
 *               groups of four columns whose members are clamped individually AFTER adding the offset to
 *               the already clamped group base (so a clamped base and its candidate vector can disagree)
 *   :1628-1650  3x3 half- then 3x3 quarter-sample refinement, best cost reset before each, SATD against the
 *               ideal block, lambda = Lambda(reciprocalSqrtLambda * 0.5) throughout */
void orc_me_bi_search(const void *srcPlane, intptr_t ss, const void *refPlane, intptr_t sr, const void *otherPlane,
                      intptr_t so, const orc_me_bi_task *b, orc_me_bi_result *out, int bps)
{
    orc_me_task t;
    search s;
    uint16_t other[64 * 64 + 64] __attribute__((aligned(32))), ideal[64 * 64 + 64] __attribute__((aligned(32)));
    uint16_t pred[64 * 64 + 64] __attribute__((aligned(32)));

    t.x0 = b->x0, t.y0 = b->y0, t.w = b->w, t.h = b->h;
    t.mvp[0] = b->mvp[0], t.mvp[1] = b->mvp[1];
    t.rateMvpFlag[0] = b->rateMvpFlag[0], t.rateMvpFlag[1] = b->rateMvpFlag[1];
    t.lambda = b->lambda;
    t.limitMin = b->limitMin, t.limitMax = b->limitMax;
    t.bitDepth = b->bitDepth;
    s.t = &t;
    s.bps = bps;
    s.ss = 64;
    s.sr = sr;
    s.src = (const char *)ideal;
    s.ref = (const char *)refPlane + ((intptr_t)b->y0 * sr + b->x0) * bps;
    s.nSad = 0;

    /* prediction from the other list */
    {
        orc_mv mv = {(int16_t)(b->mvOther.x >> 2), (int16_t)(b->mvOther.y >> 2)};
        limit(&s, &mv);
        const char *p = (const char *)otherPlane + ((intptr_t)(b->y0 + mv.y) * so + b->x0 + mv.x) * bps;
        orc_pred_uni(other, 64, p, so, b->w, b->h, b->mvOther.x & 3, b->mvOther.y & 3, b->bitDepth, 8, bps);
    }
    orc_subtract_bi(ideal, 64, other, 64, (const char *)srcPlane + ((intptr_t)b->y0 * ss + b->x0) * bps, ss, b->w, b->h,
                    6 + 2 * bps, bps);

    orc_mv start = {(int16_t)((int16_t)(b->mvStart.x + 1) >> 2), (int16_t)((int16_t)(b->mvStart.y + 1) >> 2)};
    limit(&s, &start);
    orc_mv group[4] = {start, start, start, start};
    const orc_mv origin = {(int16_t)(start.x << 2), (int16_t)(start.y << 2)};
    s.best.mv = origin;
    s.best.mvd.x = s.best.mvd.y = 0;
    s.best.mvpFlag = 0;
    s.best.cost = INT64_MAX;

    const int range = b->smallWindow ? 1 : 5;
    for (int y = -range; y <= range; ++y)
    {
        int sads[4] = {0, 0, 0, 0};
        for (int x = -range; x <= range; ++x)
        {
            orc_mv mv = {(int16_t)((int16_t)(origin.x + 4 * x) >> 2), (int16_t)((int16_t)(origin.y + 4 * y) >> 2)};
            limit(&s, &mv);
            const int i = (x + range) % 4;
            if (i == 0)
            {
                group[0] = mv;
                for (int k = 1; k < 4; ++k)
                {
                    group[k] = mv;
                    group[k].x = (int16_t)(group[k].x + k);
                    limit(&s, &group[k]);
                }
                for (int k = 0; k < 4; ++k) sads[k] = sad_at(&s, group[k].x, group[k].y);
            }
            mv.x = (int16_t)(mv.x << 2);
            mv.y = (int16_t)(mv.y << 2);
            cand c = make_candidate(&s, mv);
            c.cost += (int64_t)t.lambda * sads[i];
            consider(&s, &c);
        }
    }
    out->mvInteger = s.best.mv;

    if (b->halfPel)
    {
        const int refinement = b->quarterPel ? 1 : 2;
        for (int step = 2; step; step -= refinement)
        {
            const orc_mv o = s.best.mv;
            s.best.cost = INT64_MAX;
            for (int y = -step; y <= step; y += step)
                for (int x = -step; x <= step; x += step)
                {
                    orc_mv mv = {(int16_t)(o.x + x), (int16_t)(o.y + y)};
                    cand c = make_candidate(&s, mv);
                    orc_pred_uni(pred, 64, ref_at(&s, mv.x >> 2, mv.y >> 2), sr, b->w, b->h, mv.x & 3, mv.y & 3, b->bitDepth, 8, bps);
                    c.cost += (int64_t)t.lambda * orc_measure_satd(ideal, 64, pred, 64, b->w, b->h, bps);
                    consider(&s, &c);
                }
        }
    }
    out->mv = s.best.mv;
    out->mvd = s.best.mvd;
    out->mvpFlag = s.best.mvpFlag;
    out->cost = s.best.cost;
    out->nSad = s.nSad;
}

/* ---- the distortion half of measurePuCost (turing/Search.hpp:1668-1682) ----------------------------
 * predictInter without weighted prediction (turing/Dsp.h:866-915):
 *   predictUni (:769-806) / predictBi (:808-864): luma block origin = clipMvLumaComponent(xPb + (mv >> 2)) (:723-731),
 *   chroma origin = luma origin >> 1, chroma vector = the luma vector read in eighth-samples (mv*2/SubWidthC),
 *   8-tap luma, 4-tap chroma;
 * then Compute<Satd, prediction_unit> (turing/Measure.h:141-176): measureSatd of the PU in Y, Cb, Cr, where a
 * chroma block whose width|height is not a multiple of 4 contributes 0. */
static int clip_mv_luma_component(int component, int nPbSize, int pictureSize)
{
    if (component + nPbSize + 4 < 0) return -nPbSize - 4;
    if (component > pictureSize + 2) return pictureSize + 2;
    return component;
}

void orc_pu_cost(const orc_plane src[3], const orc_plane ref0[3], const orc_plane ref1[3], const orc_pu_cost_task *t,
                 int32_t satd[3], void *predOut[3], int bps)
{
    uint16_t pred[64 * 64 + 64] __attribute__((aligned(32)));
    const orc_plane *refs[2] = {ref0, ref1};
    const int lists = (t->predFlag[0] ? 1 : 0) + (t->predFlag[1] ? 1 : 0);
    int bx[2], by[2];
    for (int l = 0; l < 2; ++l)
    {
        bx[l] = clip_mv_luma_component(t->x0 + (t->mv[l].x >> 2), t->w, t->picWidth);
        by[l] = clip_mv_luma_component(t->y0 + (t->mv[l].y >> 2), t->h, t->picHeight);
    }
    for (int c = 0; c < 3; ++c)
    {
        const int sh = c ? 1 : 0, w = t->w >> sh, h = t->h >> sh, taps = c ? 4 : 8, mask = c ? 7 : 3;
        const int bd = c ? t->bitDepthC : t->bitDepthY;
        const char *p[2] = {0, 0};
        for (int l = 0; l < 2; ++l)
            if (t->predFlag[l])
                p[l] = (const char *)refs[l][c].p + ((intptr_t)(by[l] >> sh) * refs[l][c].stride + (bx[l] >> sh)) * bps;
        if (lists == 2)
        {
            /* havocPredBi takes ONE stride for both references (pred_inter.h:63); pictures share their geometry */
            orc_pred_bi(pred, 64, p[0], p[1], refs[0][c].stride, w, h, t->mv[0].x & mask, t->mv[0].y & mask, t->mv[1].x & mask,
                        t->mv[1].y & mask, bd, taps, bps);
        }
        else
        {
            const int l = t->predFlag[0] ? 0 : 1;
            orc_pred_uni(pred, 64, p[l], refs[l][c].stride, w, h, t->mv[l].x & mask, t->mv[l].y & mask, bd, taps, bps);
        }
        const char *s = (const char *)src[c].p + ((intptr_t)(t->y0 >> sh) * src[c].stride + (t->x0 >> sh)) * bps;
        satd[c] = (c && ((w | h) & 3)) ? 0 : orc_measure_satd(s, src[c].stride, pred, 64, w, h, bps);
        if (predOut && predOut[c])
            for (int y = 0; y < h; ++y) memcpy((char *)predOut[c] + (size_t)y * w * bps, (const char *)pred + (size_t)y * 64 * bps, (size_t)w * bps);
    }
}
