// hvb_me.cu -- batched uni-directional motion search: the integer pattern search and the sub-pel
// refinement WITH their control flow, one warp per prediction unit.
//
// Reference semantics (bit-exact decisions):
//   fullPelMotionEstimation   turing/Search.hpp:2064-2336
//   considerPattern / LimitFullPelMv / MvCandidate   turing/Search.hpp:1254-1312, :1366-1495
//   subPelRefinement / patternSearch / costMv / costDistortionMv   :1965-2060, :2339-2357
//   rateOf                    turing/Measure.h:177-220
//   Cost / Lambda             turing/FixedPoint.h, Cost.h (Q16: int64 / int32)
//
// In the reference every pattern step is a host round trip: 4 candidate vectors -> one
// havoc_sad_multiref call -> compare -> next origin.  A launch per call is hopeless (SURVEY.md
// section 7), so the data-dependent loop itself runs on the device: a warp walks the reference's
// exact candidate order for its PU, evaluating the candidates of one pattern call with the whole
// warp and applying the reference's ordered `consider` (strict <, first minimum wins).  The source
// block is held in registers for the life of the search (it is compared against 100-300
// candidates), candidate blocks stream from the L2-resident reference window.  Per PU the
// algorithmic traffic is w*h*B (source, once) + nSad * w*h*B (candidates) + 17 (w+7)(h+7)B (sub-pel).
#include "hvb_internal.cuh"
#include "hvb_satd.cuh"
#include "hvb_interp.cuh"

namespace {

using namespace hvb_interp;
constexpr int kWarps = 4;

struct Cand
{
    hvb_mv mv, mvd;
    long long cost;
    int mvpFlag;
};

// Measure.h:177-216: Cost::make(bits(|dx|) + bits(|dy|) + 1, -1) = (...) << 17
__device__ __forceinline__ long long rateOfMvd(int dx, int dy)
{
    const int rx = 32 - __clz(abs(dx)), ry = 32 - __clz(abs(dy)); // __clz(0) == 32
    return (long long)(rx + ry + 1) << 17;
}

template <typename Sample>
struct Search
{
    const hvb_me_task &t;
    const Sample *src, *ref; // sample (x0, y0) of the source / reference plane
    int ss, sr;
    int lane;
    Cand best;
    int nSad;
    uint32_t srcw[32]; // u8 fast path: this lane's words of the source block (w*h/4 words over 32 lanes)
    bool cached;

    __device__ Search(const hvb_me_task &task, const HvbPlane *planes, int lane_) : t(task), lane(lane_)
    {
        const HvbPlane &sp = planes[t.src_pic * 3], &rp = planes[t.ref_pic * 3];
        ss = sp.stride;
        sr = rp.stride;
        src = reinterpret_cast<const Sample *>(sp.base) + (intptr_t)t.y0 * ss + t.x0;
        ref = reinterpret_cast<const Sample *>(rp.base) + (intptr_t)t.y0 * sr + t.x0;
        best.cost = 0x7fffffffffffffffLL;
        best.mv = best.mvd = hvb_mv{0, 0};
        best.mvpFlag = 0;
        nSad = 0;
        cached = sizeof(Sample) == 1 && !(t.w & 3);
        if (cached)
        {
            const int wq = t.w >> 2, total = wq * t.h;
#pragma unroll
            for (int k = 0; k < 32; ++k)
            {
                const int i = lane + 32 * k;
                if (i < total)
                {
                    const int y = i / wq, x = (i - y * wq) << 2;
                    srcw[k] = hvbLoad4u8(reinterpret_cast<const uint8_t *>(src) + y * ss + x);
                }
            }
        }
    }

    __device__ __forceinline__ void limit(hvb_mv &mv) const
    {
        mv.x = max(mv.x, t.limitMin.x);
        mv.y = max(mv.y, t.limitMin.y);
        mv.x = min(mv.x, t.limitMax.x);
        mv.y = min(mv.y, t.limitMax.y);
    }

    // warp-cooperative havoc_sad of the PU against the reference displaced by a full-pel vector
    __device__ int sadAt(int mvx, int mvy)
    {
        ++nSad;
        const Sample *r = ref + (intptr_t)mvy * sr + mvx;
        int acc = 0;
        if (cached)
        {
            const int wq = t.w >> 2, total = wq * t.h;
#pragma unroll
            for (int k = 0; k < 32; ++k)
            {
                const int i = lane + 32 * k;
                if (i < total)
                {
                    const int y = i / wq, x = (i - y * wq) << 2;
                    acc = __vsadu4(srcw[k], hvbLoad4u8(reinterpret_cast<const uint8_t *>(r) + y * sr + x)) + acc;
                }
            }
        }
        else
        {
            const int w = t.w, total = t.w * t.h;
            for (int i = lane; i < total; i += 32)
            {
                const int y = i / w, x = i - y * w;
                acc += abs((int)src[y * ss + x] - (int)r[y * sr + x]);
            }
        }
        acc = hvbWarpSum(acc);
        return sizeof(Sample) == 2 ? acc >> 2 : acc;
    }

    // MvCandidate(h, refList, mv, predictors) (Search.hpp:1262-1298)
    __device__ __forceinline__ Cand makeCandidate(hvb_mv mv) const
    {
        Cand c;
        c.mvpFlag = 0;
        c.mvd.x = (int16_t)(mv.x - t.mvp[0].x);
        c.mvd.y = (int16_t)(mv.y - t.mvp[0].y);
        c.cost = rateOfMvd(c.mvd.x, c.mvd.y) + t.rateMvpFlag[0];
        const hvb_mv d1{(int16_t)(mv.x - t.mvp[1].x), (int16_t)(mv.y - t.mvp[1].y)};
        const long long c1 = rateOfMvd(d1.x, d1.y) + t.rateMvpFlag[1];
        if (c1 < c.cost)
        {
            c.mvpFlag = 1;
            c.mvd = d1;
            c.cost = c1;
        }
        c.mv = mv;
        return c;
    }

    __device__ __forceinline__ bool consider(const Cand &c)
    {
        if (c.cost < best.cost)
        {
            best = c;
            return true;
        }
        return false;
    }

    // StateMeFullPel::considerPattern (Search.hpp:1447-1482): `pattern` holds (x, y) pairs
    __device__ bool considerPattern(hvb_mv origin, const int8_t *pattern, int n, int step, int dist)
    {
        bool improved = false;
        for (int j = 0; j < n; j += step)
        {
            const int8_t *p = pattern + 2 * j;
            hvb_mv mv;
            mv.x = (int16_t)((origin.x + dist * p[0]) / 4);
            mv.y = (int16_t)((origin.y + dist * p[1]) / 4);
            limit(mv);
            const int sad = sadAt(mv.x, mv.y);
            mv.x = (int16_t)(mv.x * 4);
            mv.y = (int16_t)(mv.y * 4);
            Cand c = makeCandidate(mv);
            c.cost += (long long)t.lambda * sad;
            improved |= consider(c);
        }
        return improved;
    }
};

__device__ __constant__ int8_t kDiamond4[8] = {-4, 0, 0, 4, 4, 0, 0, -4};
__device__ __constant__ int8_t kHexagon8[16] = {0, -8, 8, -4, 8, 4, 0, 8, -8, 4, -8, -4, -8, 4, -8, -4};
__device__ __constant__ int8_t kDiamond16[32] = {0,  -4, 1,  -3, 2,  -2, 3,  -1, 4,  0, 3,  1,  2,  2,  1,  3,
                                                 0,  4,  -1, 3,  -2, 2,  -3, 1,  -4, 0, -3, -1, -2, -2, -1, -3};
__device__ __constant__ int8_t kSquare4[8] = {-4, -4, -4, 4, 4, 4, 4, -4};
__device__ __constant__ int8_t kLine4[8] = {0, 0, 1, 0, 2, 0, 3, 0};
__device__ __constant__ int8_t kDiamond1[8] = {0, -1, -1, 0, 0, 1, 1, 0};
__device__ __constant__ int8_t kHalf[16] = {-2, -2, 0, -2, 2, -2, -2, 0, 2, 0, -2, 2, 0, 2, 2, 2};
__device__ __constant__ int8_t kQuarter[16] = {-1, -1, 0, -1, 1, -1, -1, 0, 1, 0, -1, 1, 0, 1, 1, 1};

template <typename Sample>
__device__ bool metTerminates(Search<Sample> &s)
{
    bool trigger = !s.considerPattern(s.best.mv, kDiamond4, 4, 1, 1);
    if (trigger && s.t.log2CbSize >= 5) trigger = !s.considerPattern(s.best.mv, kHexagon8, 8, 1, 1);
    return trigger;
}

// returns true when the reference would have returned early through MET (Search.hpp:2125)
template <typename Sample>
__device__ bool fullPel(Search<Sample> &s, long long (&costMvdZero)[2])
{
    const hvb_me_task &t = s.t;
    const int window = t.smallSearchWindow ? 32 : 64;
    const int maxCounter = t.smallSearchWindow ? 2 : 3;
    const int raster = t.smallSearchWindow ? 120 : 240;

    { // zero vector, not clamped (:2103-2129)
        Cand c = s.makeCandidate(hvb_mv{0, 0});
        c.cost += (long long)t.lambda * s.sadAt(0, 0);
        if (s.consider(c) && t.met && metTerminates(s)) return true;
    }
    for (int flag = 0; flag < 2; ++flag) // the predictors rounded to full-pel (:2131-2171)
    {
        Cand c;
        c.mvpFlag = flag;
        c.mv.x = (int16_t)((int16_t)(t.mvp[flag].x + 1) >> 2);
        c.mv.y = (int16_t)((int16_t)(t.mvp[flag].y + 1) >> 2);
        s.limit(c.mv);
        c.mv.x = (int16_t)(c.mv.x << 2);
        c.mv.y = (int16_t)(c.mv.y << 2);
        c.mvd.x = (int16_t)(c.mv.x - t.mvp[flag].x);
        c.mvd.y = (int16_t)(c.mv.y - t.mvp[flag].y);
        c.cost = rateOfMvd(c.mvd.x, c.mvd.y) + t.rateMvpFlag[flag];
        c.cost += (long long)t.lambda * s.sadAt(c.mv.x >> 2, c.mv.y >> 2);
        costMvdZero[flag] = c.cost;
        if (s.consider(c) && t.met && metTerminates(s)) return true;
    }
    if (t.usePrev2Nx2N) // previous 2Nx2N integer vector (:2173-2198)
    {
        hvb_mv mv{(int16_t)(t.prev2Nx2N.x >> 2), (int16_t)(t.prev2Nx2N.y >> 2)};
        s.limit(mv);
        mv.x = (int16_t)(mv.x << 2);
        mv.y = (int16_t)(mv.y << 2);
        Cand c = s.makeCandidate(mv);
        c.cost += (long long)t.lambda * s.sadAt(mv.x >> 2, mv.y >> 2);
        if (s.consider(c) && t.met && metTerminates(s)) return true;
    }

    // star search (:2202-2247)
    hvb_mv start = s.best.mv;
    int distBest = 0, counter = 0, step = 4;
    for (int dist = 1; dist <= window && counter < maxCounter; dist <<= 1)
    {
        if (dist == 2 || dist == 8) step >>= 1;
        if (s.considerPattern(start, kDiamond16, 16, step, dist))
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
        s.considerPattern(s.best.mv, kSquare4, 4, 1, 1);
    }
    if (distBest > 5) // raster: absolute displacements on a 5-sample grid (:2258-2273)
    {
        for (int my = -raster; my <= raster; my += 20)
            for (int mx = -raster; mx <= raster; mx += 80) s.considerPattern(hvb_mv{(int16_t)mx, (int16_t)my}, kLine4, 4, 1, 20);
        distBest = 5;
    }
    while (distBest > 0) // star refinement (:2276-2302)
    {
        start = s.best.mv;
        distBest = 0;
        step = 4;
        for (int dist = 1; dist <= window; dist <<= 1)
        {
            if (dist == 2 || dist == 8) step >>= 1;
            if (s.considerPattern(start, kDiamond16, 16, step, dist)) distBest = dist;
        }
        if (distBest == 1)
        {
            s.considerPattern(start, kSquare4, 4, 1, 1);
            distBest = 0;
        }
    }
    if (!t.smallSearchWindow) // one-sample diamond until no improvement (:2303-2334)
    {
        int j;
        do
        {
            hvb_mv mv[4];
            int sad[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
            {
                mv[i].x = (int16_t)(s.best.mv.x / 4 + kDiamond1[2 * i]);
                mv[i].y = (int16_t)(s.best.mv.y / 4 + kDiamond1[2 * i + 1]);
                s.limit(mv[i]);
                sad[i] = s.sadAt(mv[i].x, mv[i].y);
                mv[i].x = (int16_t)(mv[i].x * 4);
                mv[i].y = (int16_t)(mv[i].y * 4);
            }
            j = -1;
#pragma unroll
            for (int i = 0; i < 4; ++i)
            {
                Cand c = s.makeCandidate(mv[i]);
                c.cost += (long long)t.lambda * sad[i];
                if (s.consider(c)) j = i;
            }
        } while (j >= 0);
    }
    return false;
}

// costMv (Search.hpp:2003-2008): rateOf(mvd) + lambda * SATD(src, 8-tap prediction at quarter-pel mv)
template <typename Sample>
__device__ long long costMv(Search<Sample> &s, int16_t *mid, int16_t *pred, hvb_mv mv, hvb_mv mvd, int bitDepth)
{
    const hvb_me_task &t = s.t;
    const int w = t.w, h = t.h;
    const int shift1 = min(4, bitDepth - 8), shift3 = max(2, 14 - bitDepth);
    const Sample *r = s.ref + (intptr_t)(mv.y >> 2) * s.sr + (mv.x >> 2);
    passH<Sample, 8>(mid, r, s.sr, w, h, mv.x & 3, shift1, s.lane);
    __syncwarp();
    int cy[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) cy[k] = coef<8>(mv.y & 3, k);
    const int total = w * h, maxv = (1 << bitDepth) - 1;
    for (int j = s.lane; j < total; j += 32)
    {
        const int y = j / w, x = j - y * w;
        pred[j] = (int16_t)hvbClip3(0, maxv, (passV<8>(mid, w, x, y, cy) + (1 << (5 + shift3))) >> (6 + shift3));
    }
    __syncwarp();
    int satd = hvbMeasureSatdLanes<Sample, int16_t>(s.src, s.ss, pred, w, w, h, s.lane, 32, sizeof(Sample) == 2 ? 2 : 0);
    satd = hvbWarpSum(satd);
    __syncwarp();
    return rateOfMvd(mvd.x, mvd.y) + (long long)t.lambda * satd;
}

// patternSearch with maxIterations = 1 (Search.hpp:2011-2060)
template <typename Sample>
__device__ void patternSearch(Search<Sample> &s, int16_t *mid, int16_t *pred, const int8_t *pattern, bool tryOrigin, hvb_mv &mv,
                              hvb_mv &mvd, long long &bestCost, int bitDepth)
{
    if (tryOrigin) bestCost = costMv(s, mid, pred, mv, mvd, bitDepth);
    int best = -1;
    for (int i = 0; i < 8; ++i)
    {
        const hvb_mv m{(int16_t)(mv.x + pattern[2 * i]), (int16_t)(mv.y + pattern[2 * i + 1])};
        const hvb_mv d{(int16_t)(mvd.x + pattern[2 * i]), (int16_t)(mvd.y + pattern[2 * i + 1])};
        const long long c = costMv(s, mid, pred, m, d, bitDepth);
        if (c < bestCost)
        {
            best = i;
            bestCost = c;
        }
    }
    if (best >= 0)
    {
        mv.x = (int16_t)(mv.x + pattern[2 * best]);
        mv.y = (int16_t)(mv.y + pattern[2 * best + 1]);
        mvd.x = (int16_t)(mvd.x + pattern[2 * best]);
        mvd.y = (int16_t)(mvd.y + pattern[2 * best + 1]);
    }
}

template <typename Sample>
__global__ void __launch_bounds__(kWarps * 32)
    meSearchKernel(const HvbPlane *__restrict__ planes, const hvb_me_task *__restrict__ tasks, int n, hvb_me_result *__restrict__ out,
                   int bitDepth)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int16_t *mid = reinterpret_cast<int16_t *>(smemRaw + warp * kSmemPerWarp);
    int16_t *pred = mid + kMidElems;
    const int warpsTotal = gridDim.x * kWarps;
    for (int i = blockIdx.x * kWarps + warp; i < n; i += warpsTotal)
    {
        const hvb_me_task t = tasks[i];
        Search<Sample> s(t, planes, lane);
        long long costMvdZero[2] = {0, 0};
        const bool early = fullPel(s, costMvdZero);

        hvb_me_result r;
        r.mvInteger = s.best.mv;
        r.mvpFlag = s.best.mvpFlag;
        r.cost = s.best.cost;
        r.costMvdZero[0] = costMvdZero[0];
        r.costMvdZero[1] = costMvdZero[1];
        r.subpelCost = 0;
        r.flags = early ? 1 : 0; // bit 0: returned through MET -> mvPreviousInteger2Nx2N is not updated
        hvb_mv mv = s.best.mv, mvd = s.best.mvd;
        if (t.halfPel) // searchMotionUni (Search.hpp:1335-1347)
        {
            long long bestCost = 0;
            patternSearch(s, mid, pred, kHalf, true, mv, mvd, bestCost, bitDepth);
            if (t.quarterPel) patternSearch(s, mid, pred, kQuarter, false, mv, mvd, bestCost, bitDepth);
            r.subpelCost = bestCost;
        }
        r.mv = mv;
        r.mvd = mvd;
        r.nSad = s.nSad;
        if (lane == 0) out[i] = r;
        __syncwarp();
    }
}

} // namespace

extern "C" int hvb_me_search_batch(hvb_context *ctx, const hvb_me_task *tasks, int n, hvb_me_result *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(hvb_me_result) * n, mem, &st);
    if (rc) return rc;
    const int smem = kWarps * kSmemPerWarp;
    int blocks = (n + kWarps - 1) / kWarps;
    const int cap = ctx->smCount * 3;
    if (blocks > cap) blocks = cap;
    const auto *dT = static_cast<const hvb_me_task *>(st.dTasks);
    auto *dO = static_cast<hvb_me_result *>(st.dOut);
    if (ctx->bps == 1)
    {
        cudaFuncSetAttribute(meSearchKernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        meSearchKernel<uint8_t><<<blocks, kWarps * 32, smem, ctx->stream>>>(ctx->dPlanes, dT, n, dO, ctx->bitDepth);
    }
    else
    {
        cudaFuncSetAttribute(meSearchKernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        meSearchKernel<uint16_t><<<blocks, kWarps * 32, smem, ctx->stream>>>(ctx->dPlanes, dT, n, dO, ctx->bitDepth);
    }
    HVB_LAUNCH_CHECK(ctx, "meSearchKernel");
    return hvbStageOut(ctx, out, sizeof(hvb_me_result) * n, mem, st);
}
