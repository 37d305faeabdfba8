// hvb_me_small.cu -- the integer motion search for PUs of at most 8x8 samples (8x8, 8x4, 4x8): four searches per warp, one
// lane per candidate.  Template on the sample type (8-bit: source block in 16 registers; 16-bit: in shared memory).
//
// Reference semantics (bit-exact decisions, same path through the search, same SAD count):
//   fullPelMotionEstimation                          turing/Search.hpp:2064-2336
//   considerPattern / LimitFullPelMv / MvCandidate   turing/Search.hpp:1254-1312, :1366-1495
//   rateOf                                           turing/Measure.h:177-220
//   havoc_sad / havoc_sad_multiref                   havoc/sad.cpp:432-449, :513-542
//
// Three quarters of a frame's PUs are this small, and for them the warp-per-PU kernel (hvb_me.cu) spends its time in
// the bookkeeping around 4..16 SADs of 16 words each.  Here a search owns 8 lanes.  The reference's control flow
// (zero vector, predictors, MET early exit, star, raster, star refinement, one-sample diamond) is written as a state
// machine whose step is one considerPattern call, so the four searches of a warp -- each at its own point of its own
// walk -- execute the same instruction stream: generate the call's candidates (a lane takes candidates l and l + 8),
// SAD of the whole block per lane (the source block lives in 16 registers; each candidate row is three aligned
// loads and two funnel shifts, 24 independent loads in flight per lane), cost, ordered arg-min across the 8 lanes
// (strict <, lowest candidate index wins ties: the reference's sequential `consider`), transition.
#include "hvb_internal.cuh"

namespace {

constexpr int kWarps = 4;
constexpr int kGroups = 4; // searches per warp

enum Phase : int
{
    PH_ZERO,
    PH_MVP0,
    PH_MVP1,
    PH_PREV,
    PH_MET_D4,
    PH_MET_H8,
    PH_STAR,
    PH_STAR_SQUARE,
    PH_RASTER,
    PH_REFINE,
    PH_REFINE_SQUARE,
    PH_DIAMOND1,
    PH_DONE
};

// pattern tables (Search.hpp:2114-2119, :2208-2222, :2255, :2303), concatenated; offsets below
__device__ __constant__ int8_t kPatterns[72] = {
    -4, 0,  0,  4,  4,  0,  0,  -4,                                                                         // diamond4      @0
    0,  -8, 8,  -4, 8,  4,  0,  8,  -8, 4,  -8, -4, -8, 4,  -8, -4,                                         // hexagon8      @8
    0,  -4, 1,  -3, 2,  -2, 3,  -1, 4,  0,  3,  1,  2,  2,  1,  3,  0,  4,  -1, 3,  -2, 2,  -3, 1,  -4, 0, -3, -1, -2, -2, -1, -3, // diamond16 @24
    -4, -4, -4, 4,  4,  4,  4,  -4,                                                                         // square4       @56
    0,  -4, -4, 0,  0,  4,  4,  0};                                                                         // one-sample diamond, x4 @64
constexpr int kPatD4 = 0, kPatH8 = 8, kPatD16 = 24, kPatSq4 = 56, kPatD1 = 64;

__device__ __forceinline__ long long rateOfMvd(int dx, int dy)
{
    const int rx = 32 - __clz(abs(dx)), ry = 32 - __clz(abs(dy)); // __clz(0) == 32
    return (long long)(rx + ry + 1) << 17;
}

template <typename Sample>
struct SmallSmem
{
    hvb_me_task task[kWarps][kGroups];
    uint32_t src[sizeof(Sample) == 2 ? kWarps * kGroups : 1][8][4]; // 16-bit samples: the groups' source blocks, 16 bytes per row
    int8_t patterns[72];
};

template <typename Sample>
__global__ void __launch_bounds__(kWarps * 32, 6)
    meSearchSmallKernel(const HvbPlane *__restrict__ planes, const hvb_me_task *__restrict__ tasks, int n, hvb_me_result *__restrict__ out,
                        int *__restrict__ cursor)
{
    __shared__ __align__(16) SmallSmem<Sample> sm;
    constexpr bool k16 = sizeof(Sample) == 2;
    uint32_t (*sSrc)[4] = sm.src[k16 ? (threadIdx.x >> 5) * kGroups + ((threadIdx.x & 31) >> 3) : 0];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int grp = lane >> 3, l8 = lane & 7, base8 = lane & 24;
    if (threadIdx.x < 72) sm.patterns[threadIdx.x] = kPatterns[threadIdx.x];
    __syncthreads();
    // Searches differ a lot in length (half of them leave through MET after two steps), so a group does not wait for
    // its three neighbours: when its search ends it writes the result and takes the next PU from a device-wide
    // cursor (kBatch indices per atomic), and the warp keeps stepping whatever mix of searches it holds.
    constexpr int kBatch = 8;
    const unsigned groupMask = 0xffu << base8;
    const hvb_me_task &t = sm.task[warp][grp];
    int i = -1, cur = 0, end = 0;
    bool mine = false;
    int phase = PH_DONE;
    uint32_t src[8][2]; // 8-bit samples only
    const Sample *ref = nullptr;
    int sr = 0, w = 8, h = 8, lambda = 0, window = 64, maxCounter = 3, raster = 240, rasterCols = 28, rasterTotal = 700;
    // search state (identical in the 8 lanes of a group)
    long long bestCost = 0x7fffffffffffffffLL, costMvdZero0 = 0, costMvdZero1 = 0;
    int bestMv = 0, bestMvd = 0, bestFlag = 0; // vectors packed x | y << 16
    int nSad = 0, early = 0;
    int ret = PH_DONE, start = 0, dist = 1, distBest = 0, counter = 0, stepP = 4, rasterBase = 0;
    bool exhausted = false;

    for (;;)
    {
        if (phase == PH_DONE && !exhausted)
        {
            if (mine && l8 == 0)
            {
                hvb_me_result r;
                const hvb_mv mv{(int16_t)bestMv, (int16_t)(bestMv >> 16)}, mvd{(int16_t)bestMvd, (int16_t)(bestMvd >> 16)};
                r.mv = mv; // the sub-pel kernel starts from here
                r.mvd = mvd;
                r.mvInteger = mv;
                r.mvpFlag = bestFlag;
                r.cost = bestCost;
                r.costMvdZero[0] = costMvdZero0;
                r.costMvdZero[1] = costMvdZero1;
                r.subpelCost = 0;
                r.nSad = nSad;
                r.flags = early;
                out[i] = r;
            }
            mine = false;
            for (;;) // next PU of at most 8x8 samples (the larger ones belong to the warp-per-PU kernel)
            {
                if (cur == end)
                {
                    int v = 0;
                    if (l8 == 0) v = atomicAdd(cursor, kBatch);
                    cur = __shfl_sync(groupMask, v, base8);
                    end = min(cur + kBatch, n);
                    if (cur >= n)
                    {
                        exhausted = true;
                        break;
                    }
                }
                i = cur++;
                __syncwarp(groupMask);
                reinterpret_cast<uint2 *>(&sm.task[warp][grp])[l8] = __ldg(reinterpret_cast<const uint2 *>(tasks + i) + l8);
                __syncwarp(groupMask);
                if (t.w <= 8 && t.h <= 8)
                {
                    mine = true;
                    break;
                }
            }
            if (mine)
            {
                w = t.w;
                h = t.h;
                const HvbPlane &sp = planes[t.src_pic * 3], &rp = planes[t.ref_pic * 3];
                sr = rp.stride;
                ref = reinterpret_cast<const Sample *>(rp.base) + (intptr_t)t.y0 * sr + t.x0;
                // the source block: x0 is a multiple of 4 and the plane rows are 256-byte aligned
                const Sample *s0 = reinterpret_cast<const Sample *>(sp.base) + (intptr_t)t.y0 * sp.stride + t.x0;
                if (!k16)
                {
#pragma unroll
                    for (int y = 0; y < 8; ++y)
                    {
                        src[y][0] = src[y][1] = 0;
                        if (y < h)
                        {
                            const uint32_t *q = reinterpret_cast<const uint32_t *>(s0 + (intptr_t)y * sp.stride);
                            src[y][0] = __ldg(q);
                            if (w == 8) src[y][1] = __ldg(q + 1);
                        }
                    }
                }
                else
                {
                    // lane l stages row l (zero beyond the block; those words are masked on the candidate side too)
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (l8 < h)
                    {
                        const uint32_t *q = reinterpret_cast<const uint32_t *>(s0 + (intptr_t)l8 * sp.stride);
                        v.x = __ldg(q);
                        v.y = __ldg(q + 1);
                        if (w == 8)
                        {
                            v.z = __ldg(q + 2);
                            v.w = __ldg(q + 3);
                        }
                    }
                    *reinterpret_cast<uint4 *>(sSrc[l8]) = v;
                    __syncwarp(groupMask);
                }
                lambda = t.lambda;
                window = t.smallSearchWindow ? 32 : 64;
                maxCounter = t.smallSearchWindow ? 2 : 3;
                raster = t.smallSearchWindow ? 120 : 240;
                rasterCols = 4 * ((2 * raster) / 80 + 1);
                rasterTotal = ((2 * raster) / 20 + 1) * rasterCols;
                bestCost = 0x7fffffffffffffffLL;
                costMvdZero0 = costMvdZero1 = 0;
                bestMv = bestMvd = bestFlag = 0;
                nSad = early = 0;
                phase = PH_ZERO;
            }
        }
        if (!__any_sync(0xffffffffu, phase != PH_DONE)) break;
        {
            // ---- the call of this step -------------------------------------------------------------------------
            // pattern calls: candidate ci = origin + dist * pattern[ci * stepP], ci < nc; single: one given vector
            int nc = 0, pat = 0, origin = bestMv, pdist = 1, pstep = 1;
            int singleX = 0, singleY = 0, fixedFlag = -1; // fixedFlag >= 0: the candidate is tied to predictor fixedFlag
            bool single = false, clampSingle = true, isRaster = false;
            switch (phase)
            {
            case PH_ZERO:
                single = true;
                clampSingle = false;
                nc = 1;
                break;
            case PH_MVP0:
            case PH_MVP1:
                single = true;
                nc = 1;
                fixedFlag = phase - PH_MVP0;
                singleX = (int16_t)(t.mvp[fixedFlag].x + 1) >> 2;
                singleY = (int16_t)(t.mvp[fixedFlag].y + 1) >> 2;
                break;
            case PH_PREV:
                single = true;
                nc = 1;
                singleX = t.prev2Nx2N.x >> 2;
                singleY = t.prev2Nx2N.y >> 2;
                break;
            case PH_MET_D4:
            case PH_STAR_SQUARE:
                nc = 4;
                pat = phase == PH_MET_D4 ? kPatD4 : kPatSq4;
                break;
            case PH_MET_H8:
                nc = 8;
                pat = kPatH8;
                break;
            case PH_STAR:
            case PH_REFINE:
                nc = 16 / stepP;
                pat = kPatD16;
                origin = start;
                pdist = dist;
                pstep = stepP;
                break;
            case PH_REFINE_SQUARE:
                nc = 4;
                pat = kPatSq4;
                origin = start;
                break;
            case PH_RASTER:
                isRaster = true;
                nc = min(16, rasterTotal - rasterBase);
                break;
            case PH_DIAMOND1:
                nc = 4;
                pat = kPatD1;
                break;
            default:
                break;
            }

            // ---- evaluate: this lane's candidates ci = l8 and l8 + 8 -------------------------------------------------
            long long myCost = 0x7fffffffffffffffLL;
            int myCi = 255, myMv = 0, myMvd = 0, myFlag = 0;
#pragma unroll 1
            for (int slot = 0; slot < 2; ++slot)
            {
                const int ci = l8 + 8 * slot;
                if (ci >= nc) break;
                int fx, fy;
                if (single)
                {
                    fx = singleX;
                    fy = singleY;
                }
                else if (isRaster)
                {
                    const int q = rasterBase + ci, row = q / rasterCols, col = q - row * rasterCols;
                    fx = (-raster + 20 * col) / 4;
                    fy = (-raster + 20 * row) / 4;
                }
                else
                {
                    const int8_t *p = sm.patterns + pat + 2 * ci * pstep;
                    fx = (int16_t)(((int16_t)origin + pdist * p[0]) / 4);
                    fy = (int16_t)(((origin >> 16) + pdist * p[1]) / 4);
                }
                if (!single || clampSingle)
                {
                    fx = min(max(fx, (int)t.limitMin.x), (int)t.limitMax.x);
                    fy = min(max(fy, (int)t.limitMin.y), (int)t.limitMax.y);
                }
                // SAD of the block at (fx, fy)
                int sad = 0;
                {
                    const Sample *r = ref + (intptr_t)fy * sr + fx;
                    const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(r) & 3) * 8;
                    const uint32_t *q = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(r) & ~uintptr_t(3));
                    const int srw = (sr * (int)sizeof(Sample)) >> 2;
                    // four rows at a time, every load issued before the first use; rows beyond h re-read row 0 and are not
                    // summed, the words beyond a 4-wide block are masked
#pragma unroll
                    for (int half = 0; half < 2; ++half)
                    {
                        if (half && h <= 4) break;
                        if (!k16)
                        {
                            uint32_t wv[4][3];
#pragma unroll
                            for (int yy = 0; yy < 4; ++yy)
                            {
                                const int y = half * 4 + yy;
                                const uint32_t *qr = q + (y < h ? y : 0) * srw;
                                wv[yy][0] = __ldg(qr);
                                wv[yy][1] = __ldg(qr + 1);
                                wv[yy][2] = __ldg(qr + 2);
                            }
#pragma unroll
                            for (int yy = 0; yy < 4; ++yy)
                            {
                                const int y = half * 4 + yy;
                                const uint32_t v0 = __funnelshift_r(wv[yy][0], wv[yy][1], sh);
                                const uint32_t v1 = w == 8 ? __funnelshift_r(wv[yy][1], wv[yy][2], sh) : 0u;
                                const int rowSad = __vsadu4(src[y][1], v1) + __vsadu4(src[y][0], v0);
                                sad += y < h ? rowSad : 0;
                            }
                        }
                        else
                        {
                            const uint32_t m = w == 8 ? ~0u : 0u;
                            uint32_t wv[4][5];
#pragma unroll
                            for (int yy = 0; yy < 4; ++yy)
                            {
                                const int y = half * 4 + yy;
                                const uint32_t *qr = q + (y < h ? y : 0) * srw;
#pragma unroll
                                for (int k = 0; k < 5; ++k) wv[yy][k] = __ldg(qr + k);
                            }
#pragma unroll
                            for (int yy = 0; yy < 4; ++yy)
                            {
                                const int y = half * 4 + yy;
                                const uint4 sv = *reinterpret_cast<const uint4 *>(sSrc[y]); // broadcast within the group
                                int rowSad = __vsadu2(sv.x, __funnelshift_r(wv[yy][0], wv[yy][1], sh));
                                rowSad = __vsadu2(sv.y, __funnelshift_r(wv[yy][1], wv[yy][2], sh)) + rowSad;
                                rowSad = __vsadu2(sv.z, __funnelshift_r(wv[yy][2], wv[yy][3], sh) & m) + rowSad;
                                rowSad = __vsadu2(sv.w, __funnelshift_r(wv[yy][3], wv[yy][4], sh) & m) + rowSad;
                                sad += y < h ? rowSad : 0;
                            }
                        }
                    }
                    if (k16) sad >>= 2; // havoc/sad.cpp:446-447: 16-bit samples
                }
                // the candidate (MvCandidate, Search.hpp:1262-1298; the predictor candidates :2131-2171 are tied to theirs)
                const int mvx = fx * 4, mvy = fy * 4;
                int dx = (int16_t)(mvx - t.mvp[0].x), dy = (int16_t)(mvy - t.mvp[0].y), flag = 0;
                long long cost = rateOfMvd(dx, dy) + t.rateMvpFlag[0];
                {
                    const int dx1 = (int16_t)(mvx - t.mvp[1].x), dy1 = (int16_t)(mvy - t.mvp[1].y);
                    const long long c1 = rateOfMvd(dx1, dy1) + t.rateMvpFlag[1];
                    if (fixedFlag == 1 || (fixedFlag < 0 && c1 < cost))
                    {
                        flag = 1;
                        dx = dx1;
                        dy = dy1;
                        cost = c1;
                    }
                }
                cost += (long long)lambda * sad;
                if (cost < myCost) // slot 1 (the later candidate) only wins when strictly cheaper
                {
                    myCost = cost;
                    myCi = ci;
                    myMv = (mvx & 0xffff) | (mvy << 16);
                    myMvd = (dx & 0xffff) | (dy << 16);
                    myFlag = flag;
                }
            }
            // ordered arg-min across the group
            long long cost = myCost;
            int who = myCi << 3 | l8; // candidate index, then the lane that holds it
#pragma unroll
            for (int o = 1; o < 8; o <<= 1)
            {
                const int lo = __shfl_xor_sync(0xffffffffu, (int)(cost & 0xffffffffLL), o);
                const int hi = __shfl_xor_sync(0xffffffffu, (int)(cost >> 32), o);
                const int ow = __shfl_xor_sync(0xffffffffu, who, o);
                const long long oc = ((long long)hi << 32) | (unsigned)lo;
                if (oc < cost || (oc == cost && ow < who))
                {
                    cost = oc;
                    who = ow;
                }
            }
            const int srcLane = base8 | (who & 7);
            const int wMv = __shfl_sync(0xffffffffu, myMv, srcLane), wMvd = __shfl_sync(0xffffffffu, myMvd, srcLane);
            const int wFlag = __shfl_sync(0xffffffffu, myFlag, srcLane);
            const bool active = phase != PH_DONE;
            const bool improved = active && cost < bestCost;
            if (improved)
            {
                bestCost = cost;
                bestMv = wMv;
                bestMvd = wMvd;
                bestFlag = wFlag;
            }
            if (active) nSad += nc;

            // ---- transition -----------------------------------------------------------------------------------------
            int next = phase;
            bool enterStar = false, afterStar = false, refineCheck = false, diamondCheck = false;
            switch (phase)
            {
            case PH_ZERO:
                next = PH_MVP0;
                break;
            case PH_MVP0:
                costMvdZero0 = cost;
                next = PH_MVP1;
                break;
            case PH_MVP1:
                costMvdZero1 = cost;
                next = t.usePrev2Nx2N ? PH_PREV : PH_STAR;
                break;
            case PH_PREV:
                next = PH_STAR;
                break;
            case PH_MET_D4:
                if (improved)
                    next = ret;
                else if (t.log2CbSize >= 5)
                    next = PH_MET_H8;
                else
                {
                    early = 1;
                    next = PH_DONE;
                }
                break;
            case PH_MET_H8:
                if (improved)
                    next = ret;
                else
                {
                    early = 1;
                    next = PH_DONE;
                }
                break;
            case PH_STAR:
                if (improved)
                {
                    distBest = dist;
                    counter = 0;
                }
                else
                    ++counter;
                dist <<= 1;
                if (dist <= window && counter < maxCounter)
                {
                    if (dist == 2 || dist == 8) stepP >>= 1;
                }
                else
                    afterStar = true;
                break;
            case PH_STAR_SQUARE:
                refineCheck = true; // distBest is 0 here
                break;
            case PH_RASTER:
                rasterBase += nc;
                if (rasterBase >= rasterTotal)
                {
                    distBest = 5;
                    refineCheck = true;
                }
                break;
            case PH_REFINE:
                if (improved) distBest = dist;
                dist <<= 1;
                if (dist <= window)
                {
                    if (dist == 2 || dist == 8) stepP >>= 1;
                }
                else if (distBest == 1)
                    next = PH_REFINE_SQUARE;
                else
                    refineCheck = true;
                break;
            case PH_REFINE_SQUARE:
                distBest = 0;
                diamondCheck = true;
                break;
            case PH_DIAMOND1:
                if (!improved) next = PH_DONE;
                break;
            default:
                break;
            }
            // the candidates up to the star search start MET when they improve on the best (:2125, :2168, :2196)
            if (phase <= PH_PREV)
            {
                if (improved && t.met)
                {
                    ret = next;
                    next = PH_MET_D4;
                }
            }
            if ((phase <= PH_PREV || phase == PH_MET_D4 || phase == PH_MET_H8) && next == PH_STAR) enterStar = true;
            if (enterStar)
            {
                start = bestMv;
                distBest = 0;
                counter = 0;
                stepP = 4;
                dist = 1;
            }
            if (afterStar)
            {
                if (distBest == 1)
                {
                    distBest = 0;
                    next = PH_STAR_SQUARE;
                }
                else if (distBest > 5)
                {
                    rasterBase = 0;
                    next = PH_RASTER;
                }
                else
                    refineCheck = true;
            }
            if (refineCheck)
            {
                if (distBest > 0)
                {
                    start = bestMv;
                    distBest = 0;
                    stepP = 4;
                    dist = 1;
                    next = PH_REFINE;
                }
                else
                    diamondCheck = true;
            }
            if (diamondCheck) next = t.smallSearchWindow ? PH_DONE : PH_DIAMOND1;
            phase = next;
        }

    }
}

} // namespace

// called by hvb_me_search_batch (hvb_me.cu) for 8-bit batches, before the warp-per-PU kernel takes the larger PUs
template <typename Sample>
static int launchMeSmall(hvb_context *ctx, const hvb_me_task *dTasks, int n, hvb_me_result *dOut)
{
    int perSm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, meSearchSmallKernel<Sample>, kWarps * 32, 0);
    // persistent: every group pulls PUs from the cursor until the batch is exhausted
    int blocks = (n + kWarps * kGroups - 1) / (kWarps * kGroups);
    const int cap = ctx->smCount * (perSm > 0 ? perSm : 1);
    if (blocks > cap) blocks = cap;
    cudaMemsetAsync(ctx->workCursors, 0, sizeof(int), ctx->stream);
    meSearchSmallKernel<Sample><<<blocks, kWarps * 32, 0, ctx->stream>>>(ctx->dPlanes, dTasks, n, dOut, ctx->workCursors);
    HVB_LAUNCH_CHECK(ctx, "meSearchSmallKernel");
    return HVB_OK;
}

int hvbLaunchMeSmall(hvb_context *ctx, const hvb_me_task *dTasks, int n, hvb_me_result *dOut)
{
    return ctx->bps == 1 ? launchMeSmall<uint8_t>(ctx, dTasks, n, dOut) : launchMeSmall<uint16_t>(ctx, dTasks, n, dOut);
}
