/*
 * oracle/oracle_bench.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Runs the bench's frame-pass workload (turingcodec_b200/workload.py) on the host CPU cores so that
 * bench.py can report a CPU baseline next to the GPU number.  The loops and their control flow are
 * the oracle's; the pixel primitives underneath are called through a small table that is either
 *   - the oracle's own C restatement (kind "port"), or
 *   - the UNMODIFIED reference havoc library's function tables, AVX2/xbyak JIT when the CPU has it
 *     (kind "reference"; installed from Python with orc_bench_use_reference(), pointers taken from
 *     oracle/_ref/libhavoc_ref.so).
 * Task structs are the byte layouts of include/hvb.h, re-declared here so that the oracle does not
 * include product headers.  Parallelised over tasks with a pthread work-sharing loop (one task = one reference call chain).
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include <unistd.h>

typedef struct { int16_t pic, cIdx, x, y; } b_block;
typedef struct { int16_t x, y; } b_mv;

typedef struct
{
    int16_t src_pic, ref_pic, x0, y0, w, h;
    b_mv mvp[2];
    int64_t rateMvpFlag[2];
    int32_t lambda;
    b_mv limitMin, limitMax, prev2Nx2N;
    uint8_t smallSearchWindow, met, log2CbSize, usePrev2Nx2N, halfPel, quarterPel, reserved[2];
} b_me_task; /* 64 bytes */

typedef struct
{
    b_mv mv, mvd, mvInteger;
    int32_t mvpFlag;
    int64_t cost, costMvdZero[2], subpelCost;
    int32_t nSad, flags;
} b_me_result; /* 56 bytes */

typedef struct
{
    b_block src;
    int32_t nb_unfiltered, nb_filtered;
    int8_t log2n, cIdx, strong, reserved[5];
} b_intra_task; /* 24 bytes */

typedef struct
{
    b_block src, pred, rec;
    int32_t levels;
    int8_t log2n, trType, cIdx, flags;
    int32_t qscale, qshift, qoffset, iqscale, iqshift;
    int8_t scanIdx, reserved[3];
    int32_t rdoq_ctx;
} b_tu_task; /* 60 bytes */

typedef struct { uint32_t ssd, ssdPred; int32_t cbf, status; uint32_t sadQuad[4]; } b_tu_result;

/* a picture plane on the host: pointer to sample (0,0), stride in samples */
typedef struct { const void *base; intptr_t stride; } b_plane;

/* ---- primitive table ---------------------------------------------------------------------- */

typedef int (*fn_sad)(void *, const void *, intptr_t, const void *, intptr_t, int, int, int);
typedef int (*fn_satd)(void *, const void *, intptr_t, const void *, intptr_t, int, int);
typedef int (*fn_pred_uni)(void *, void *, intptr_t, const void *, intptr_t, int, int, int, int, int, int, int);
typedef int (*fn_pred_intra)(void *, void *, intptr_t, const void *, int, int, int, int, int);
typedef void (*fn_fwd)(void *, int16_t *, const int16_t *, intptr_t, int, int, int);
typedef void (*fn_ita)(void *, void *, intptr_t, const void *, intptr_t, const int16_t *, int, int, int, int);
typedef void (*fn_iq)(void *, int16_t *, const int16_t *, int, int, int);
typedef uint32_t (*fn_ssd)(void *, const void *, intptr_t, const void *, intptr_t, int, int);

static struct
{
    void *handle; /* reference tables (ref_create), or NULL for the port */
    fn_sad sad;
    fn_satd satd;
    fn_pred_uni pred_uni;
    fn_pred_intra pred_intra;
    fn_fwd fwd;
    fn_ita ita;
    fn_iq iq;
    fn_ssd ssd;
} prim;

void orc_bench_use_port(void) { memset(&prim, 0, sizeof(prim)); }

void orc_bench_use_reference(void *handle, void *sad, void *satd, void *pred_uni, void *pred_intra, void *fwd, void *ita,
                             void *iq, void *ssd)
{
    prim.handle = handle;
    prim.sad = (fn_sad)sad;
    prim.satd = (fn_satd)satd;
    prim.pred_uni = (fn_pred_uni)pred_uni;
    prim.pred_intra = (fn_pred_intra)pred_intra;
    prim.fwd = (fn_fwd)fwd;
    prim.ita = (fn_ita)ita;
    prim.iq = (fn_iq)iq;
    prim.ssd = (fn_ssd)ssd;
}

/* hooks used by oracle_search.c when a reference table is installed */
int orc_hook_active(void) { return prim.handle != NULL; }
int orc_hook_sad(const void *a, intptr_t sa, const void *b, intptr_t sb, int w, int h, int bps)
{
    return prim.sad(prim.handle, a, sa, b, sb, w, h, bps);
}
void orc_hook_pred_uni(void *dst, intptr_t sd, const void *ref, intptr_t sr, int w, int h, int xf, int yf, int bd, int bps)
{
    prim.pred_uni(prim.handle, dst, sd, ref, sr, w, h, xf, yf, bd, 8, bps);
}
int orc_hook_satd_tile(const void *a, intptr_t sa, const void *b, intptr_t sb, int log2n, int bps)
{
    return prim.satd(prim.handle, a, sa, b, sb, log2n, bps);
}

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

/* ---- a minimal work-sharing loop over host threads (no OpenMP runtime in this image) ------- */

static int g_threads = 0;

int orc_bench_threads(void)
{
    if (g_threads <= 0)
    {
        long n = sysconf(_SC_NPROCESSORS_ONLN);
        g_threads = n > 0 ? (int)n : 1;
    }
    return g_threads;
}

void orc_bench_set_threads(int n) { g_threads = n; }

typedef void (*body_fn)(void *args, int i);
typedef struct
{
    body_fn body;
    void *args;
    int n, chunk;
    int next;
} pf_state;

static void *pf_worker(void *p)
{
    pf_state *st = (pf_state *)p;
    for (;;)
    {
        const int begin = __atomic_fetch_add(&st->next, st->chunk, __ATOMIC_RELAXED);
        if (begin >= st->n) break;
        const int end = begin + st->chunk < st->n ? begin + st->chunk : st->n;
        for (int i = begin; i < end; ++i) st->body(st->args, i);
    }
    return NULL;
}

static void parallel_for(int n, int chunk, body_fn body, void *args)
{
    pf_state st = {body, args, n, chunk, 0};
    const int nt = orc_bench_threads();
    pthread_t th[256];
    const int spawn = nt > 256 ? 256 : nt;
    for (int k = 1; k < spawn; ++k) pthread_create(&th[k], NULL, pf_worker, &st);
    pf_worker(&st);
    for (int k = 1; k < spawn; ++k) pthread_join(th[k], NULL);
}

/* ---- loops A+B ---------------------------------------------------------------------------- */

typedef struct
{
    const b_plane *planes;
    const b_me_task *tasks;
    b_me_result *out;
    int bps, bitDepth;
} me_args;

static void me_body(void *a, int i)
{
    const me_args *A = (const me_args *)a;
    const b_plane *planes = A->planes;
    b_me_result *out = A->out;
    const int bps = A->bps, bitDepth = A->bitDepth;
    {
        const b_me_task *t = &A->tasks[i];
        orc_me_task ot;
        orc_me_result r;
        ot.x0 = t->x0; ot.y0 = t->y0; ot.w = t->w; ot.h = t->h;
        for (int k = 0; k < 2; ++k)
        {
            ot.mvp[k].x = t->mvp[k].x; ot.mvp[k].y = t->mvp[k].y;
            ot.rateMvpFlag[k] = t->rateMvpFlag[k];
        }
        ot.lambda = t->lambda;
        ot.limitMin.x = t->limitMin.x; ot.limitMin.y = t->limitMin.y;
        ot.limitMax.x = t->limitMax.x; ot.limitMax.y = t->limitMax.y;
        ot.smallSearchWindow = t->smallSearchWindow; ot.met = t->met; ot.log2CbSize = t->log2CbSize;
        ot.usePrev2Nx2N = t->usePrev2Nx2N;
        ot.prev2Nx2N.x = t->prev2Nx2N.x; ot.prev2Nx2N.y = t->prev2Nx2N.y;
        ot.halfPel = t->halfPel; ot.quarterPel = t->quarterPel; ot.bitDepth = bitDepth;
        const b_plane *sp = &planes[t->src_pic * 3], *rp = &planes[t->ref_pic * 3];
        orc_me_search(sp->base, sp->stride, rp->base, rp->stride, &ot, &r, bps);
        b_me_result *o = &out[i];
        o->mv.x = r.mv.x; o->mv.y = r.mv.y; o->mvd.x = r.mvd.x; o->mvd.y = r.mvd.y;
        o->mvInteger.x = r.mvInteger.x; o->mvInteger.y = r.mvInteger.y;
        o->mvpFlag = r.mvpFlag; o->cost = r.cost;
        o->costMvdZero[0] = r.costMvdZero[0]; o->costMvdZero[1] = r.costMvdZero[1];
        o->subpelCost = r.subpelCost; o->nSad = r.nSad; o->flags = r.earlyExit;
    }
}

double orc_bench_me(const b_plane *planes /* [pic*3 + cIdx] */, const b_me_task *tasks, int n, b_me_result *out, int bps,
                    int bitDepth)
{
    me_args A = {planes, tasks, out, bps, bitDepth};
    const double t0 = now();
    parallel_for(n, 16, me_body, &A);
    return now() - t0;
}

/* ---- loop D ------------------------------------------------------------------------------- */

int orc_intra_filter_flag(int cIdx, int mode, int n)
{
    if (cIdx != 0 || mode == 1 || n == 4) return 0;
    if (mode == 0) return 1;
    int d1 = abs(mode - 26), d2 = abs(mode - 10), d = d1 < d2 ? d1 : d2;
    return d > (n == 8 ? 7 : (n == 16 ? 1 : 0));
}

/* turing/IntraReferenceSamples.h:373-419 on an array whose corner sits at index 2n */
void orc_intra_filter_neighbours(uint16_t *f, const uint16_t *u, int n, int bitDepth, int strongEnabled)
{
    const int c = 2 * n;
    int strong = 0;
    if (strongEnabled && n == 32)
        strong = abs(u[c] + u[c + 64] - 2 * u[c + 32]) < (1 << (bitDepth - 5)) && abs(u[c] + u[c - 64] - 2 * u[c - 32]) < (1 << (bitDepth - 5));
    for (int i = 0; i <= 4 * n; ++i)
    {
        if (i == 0 || i == 4 * n) f[i] = u[i];
        else if (strong)
        {
            if (i == c) f[i] = u[c];
            else if (i > c) { int x = i - c - 1; f[i] = (uint16_t)(((63 - x) * u[c] + (x + 1) * u[c + 64] + 32) >> 6); }
            else { int y = c - 1 - i; f[i] = (uint16_t)(((63 - y) * u[c] + (y + 1) * u[c - 64] + 32) >> 6); }
        }
        else
            f[i] = (uint16_t)((u[i - 1] + 2 * u[i] + u[i + 1] + 2) >> 2);
    }
}

typedef struct
{
    const b_plane *planes;
    const void *pool;
    const b_intra_task *tasks;
    int32_t *out;
    int bps, bitDepth;
} intra_args;

static void intra_body(void *a, int i)
{
    const intra_args *A = (const intra_args *)a;
    const b_plane *planes = A->planes;
    const void *pool = A->pool;
    int32_t *out = A->out;
    const int bps = A->bps, bitDepth = A->bitDepth;
    {
        const b_intra_task *t = &A->tasks[i];
        const int nn = 1 << t->log2n, c = 2 * nn;
        uint16_t u16[129], f16[129];
        uint8_t u8[129 + 1], f8[129 + 1];
        for (int k = 0; k <= 4 * nn; ++k)
            u16[k] = bps == 1 ? ((const uint8_t *)pool)[t->nb_unfiltered - c + k] : ((const uint16_t *)pool)[t->nb_unfiltered - c + k];
        orc_intra_filter_neighbours(f16, u16, nn, bitDepth, t->strong);
        for (int k = 0; k <= 4 * nn; ++k) { u8[k] = (uint8_t)u16[k]; f8[k] = (uint8_t)f16[k]; }
        const b_plane *sp = &planes[t->src.pic * 3 + t->src.cIdx];
        const char *src = (const char *)sp->base + ((intptr_t)t->src.y * sp->stride + t->src.x) * bps;
        uint16_t pred[32 * 32];
        const int lt = t->log2n == 2 ? 2 : 3, tn = 1 << lt;
        const int edge = t->cIdx == 0 && t->log2n < 5;
        for (int mode = 0; mode < 35; ++mode)
        {
            const int ff = orc_intra_filter_flag(t->cIdx, mode, nn);
            const void *nb = bps == 1 ? (const void *)((ff ? f8 : u8) + c + 1) : (const void *)((ff ? f16 : u16) + c + 1);
            if (prim.handle) prim.pred_intra(prim.handle, pred, nn, nb, mode, t->log2n, bitDepth, t->cIdx, bps);
            else orc_pred_intra(pred, nn, nb, mode, t->log2n, bitDepth, edge, bps);
            int acc = 0;
            for (int y = 0; y < nn; y += tn)
                for (int x = 0; x < nn; x += tn)
                {
                    const void *a = src + ((intptr_t)y * sp->stride + x) * bps;
                    const void *b = (const char *)pred + (y * nn + x) * bps;
                    acc += prim.handle ? prim.satd(prim.handle, a, sp->stride, b, nn, lt, bps) : orc_hadamard_satd(a, sp->stride, b, nn, lt, bps);
                }
            out[i * 35 + mode] = acc;
        }
    }
}

double orc_bench_intra(const b_plane *planes, const void *pool, const b_intra_task *tasks, int n, int32_t *out, int bps, int bitDepth)
{
    intra_args A = {planes, pool, tasks, out, bps, bitDepth};
    const double t0 = now();
    parallel_for(n, 64, intra_body, &A);
    return now() - t0;
}

/* ---- loop C ------------------------------------------------------------------------------- */

typedef struct
{
    const b_plane *planes, *recPlanes;
    const orc_rdoq_ctx *ctx;
    const b_tu_task *tasks;
    int16_t *levelPool;
    b_tu_result *out;
    int bps, bitDepth;
} tu_args;

static void tu_body(void *a, int i)
{
    const tu_args *A = (const tu_args *)a;
    const b_plane *planes = A->planes, *recPlanes = A->recPlanes;
    const orc_rdoq_ctx *ctx = A->ctx;
    int16_t *levelPool = A->levelPool;
    b_tu_result *out = A->out;
    const int bps = A->bps, bitDepth = A->bitDepth;
    {
        const b_tu_task *t = &A->tasks[i];
        const int nn = 1 << t->log2n, count = nn * nn;
        const b_plane *sp = &planes[t->src.pic * 3 + t->src.cIdx], *pp = &planes[t->pred.pic * 3 + t->pred.cIdx];
        const b_plane *rp = &recPlanes[t->rec.pic * 3 + t->rec.cIdx];
        const char *src = (const char *)sp->base + ((intptr_t)t->src.y * sp->stride + t->src.x) * bps;
        const char *predPic = (const char *)pp->base + ((intptr_t)t->pred.y * pp->stride + t->pred.x) * bps;
        char *rec = (char *)rp->base + ((intptr_t)t->rec.y * rp->stride + t->rec.x) * bps;
        /* the encoder keeps predictions in 32-byte aligned cache pieces (turing/ReconstructionCache.h) and the
         * reference's SIMD bodies rely on it: stage the (motion-compensated, hence unaligned) block the same way */
        uint16_t predBuf[32 * 32] __attribute__((aligned(32)));
        for (int y = 0; y < nn; ++y) memcpy((char *)predBuf + (size_t)y * nn * bps, predPic + (size_t)y * pp->stride * bps, (size_t)nn * bps);
        const char *pred = (const char *)predBuf;
        const b_plane predPlane = {predBuf, nn};
        pp = &predPlane;
        int16_t res[1024] __attribute__((aligned(32))), coeffs[1024] __attribute__((aligned(32)));
        int16_t deq[1024] __attribute__((aligned(32)));
        int16_t *levels = levelPool + t->levels;
        uint32_t sadQuad[4] = {0, 0, 0, 0};
        for (int y = 0; y < nn; ++y)
            for (int x = 0; x < nn; ++x)
            {
                int s = bps == 1 ? ((const uint8_t *)src)[y * sp->stride + x] : ((const uint16_t *)src)[y * sp->stride + x];
                int p = bps == 1 ? ((const uint8_t *)pred)[y * pp->stride + x] : ((const uint16_t *)pred)[y * pp->stride + x];
                res[y * nn + x] = (int16_t)(s - p);
                /* Candidate::sadResidueQuad (turing/Reconstruct.cpp:1268-1287) */
                sadQuad[2 * (y >> (t->log2n - 1)) + (x >> (t->log2n - 1))] += (uint32_t)abs(s - p);
            }
        if (prim.handle) prim.fwd(prim.handle, coeffs, res, nn, t->trType, t->log2n, bitDepth);
        else orc_transform_fwd(coeffs, res, nn, t->trType, t->log2n, bitDepth);
        int cbf;
        if (t->flags & 1)
            cbf = orc_rdoq(levels, coeffs, &ctx[t->rdoq_ctx], t->qscale, t->qshift, t->iqscale, t->log2n, t->cIdx, t->scanIdx,
                           (t->flags >> 1) & 1, (t->flags >> 2) & 1, bitDepth);
        else
            cbf = orc_quantize(levels, coeffs, t->qscale, t->qshift, t->qoffset, count);
        if (prim.handle)
        {
            prim.iq(prim.handle, deq, levels, t->iqscale, t->iqshift, count);
            prim.ita(prim.handle, rec, rp->stride, pred, pp->stride, deq, t->trType, t->log2n, bitDepth, bps);
        }
        else
        {
            orc_quantize_inverse(deq, levels, t->iqscale, t->iqshift, count);
            orc_inverse_transform_add(rec, rp->stride, pred, pp->stride, deq, t->trType, t->log2n, bitDepth, bps);
        }
        if (prim.handle && nn >= 16) /* the JIT SSD uses aligned 16-byte loads (havoc/ssd.cpp:122-123) */
        {
            out[i].ssd = prim.ssd(prim.handle, src, sp->stride, rec, rp->stride, t->log2n, bps);
            out[i].ssdPred = prim.ssd(prim.handle, src, sp->stride, pred, pp->stride, t->log2n, bps);
        }
        else
        {
            out[i].ssd = orc_ssd(src, sp->stride, rec, rp->stride, nn, nn, bps);
            out[i].ssdPred = orc_ssd(src, sp->stride, pred, pp->stride, nn, nn, bps);
        }
        out[i].cbf = cbf != 0;
        out[i].status = 0;
        memcpy(out[i].sadQuad, sadQuad, sizeof sadQuad);
    }
}

double orc_bench_tu(const b_plane *planes, b_plane *recPlanes /* writable views, same indexing */, const orc_rdoq_ctx *ctx,
                    const b_tu_task *tasks, int n, int16_t *levelPool, b_tu_result *out, int bps, int bitDepth)
{
    tu_args A = {planes, recPlanes, ctx, tasks, levelPool, out, bps, bitDepth};
    const double t0 = now();
    parallel_for(n, 32, tu_body, &A);
    return now() - t0;
}
