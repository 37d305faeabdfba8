// hvb_me.cu -- batched motion search WITH its control flow: the integer pattern search of the PUs larger than 8x8, one
// warp per prediction unit (smaller PUs: hvb_me_small.cu; the sub-pel refinement of every PU: hvb_me_subpel.cu), and
// the bi-directional refinement (searchMotionBi), whose fractional rounds keep the in-kernel sub-pel evaluation below.
//
// Reference semantics (bit-exact decisions):
//   fullPelMotionEstimation   turing/Search.hpp:2064-2336
//   considerPattern / LimitFullPelMv / MvCandidate   turing/Search.hpp:1254-1312, :1366-1495
//   subPelRefinement / patternSearch / costMv / costDistortionMv   :1965-2060, :2339-2357
//   rateOf                    turing/Measure.h:177-220
//   Cost / Lambda             turing/FixedPoint.h, Cost.h (Q16: int64 / int32)
//   interpolation / SATD      havoc/pred_inter.cpp:76-202, havoc/hadamard.cpp:58-98, turing/Measure.h:96-135
//
// In the reference every pattern step is a host round trip: 4 candidate vectors -> one
// havoc_sad_multiref call -> compare -> next origin.  A launch per call is hopeless (SURVEY.md
// section 7), so the data-dependent loop itself runs on the device.  A warp owns one PU and walks the
// reference's exact sequence of pattern calls; what it parallelises is the INSIDE of each call:
//
//   integer stage   the (up to 16, raster: 32) candidates of one considerPattern call are evaluated at
//                   once, 32/nCand lanes per candidate, then an ordered arg-min (cost, candidate index)
//                   reproduces the reference's sequential `consider` (strict <, first minimum wins);
//   sub-pel stage   the 9 half-pel (then 8 quarter-pel) candidates are evaluated at once: a job is one
//                   (candidate, 8x8 tile) -- 8 lanes, lane j owns column j: it streams the 15 rows of the
//                   8-tap horizontal filter through an 8-deep register window for the vertical filter,
//                   subtracts the source column and runs the Hadamard butterfly vertically in
//                   registers and horizontally with shuffles.  No prediction is ever stored.
//
// The source block lives in shared memory for the life of the search (it is compared against every
// candidate); candidate blocks stream from the L2/L1-resident reference window.  Serial depth per PU
// is the number of pattern calls (5-20), not the number of SADs (30-900).
// Algorithmic traffic per PU: w*h*B (source, once) + nSad*w*h*B + 17 ((w+7)(h+7) + w*h) B.
#include "hvb_internal.cuh"

namespace {

constexpr int kWarps = 8;
constexpr int kSrcWords = 64 * 64 / 2; // u16 worst case: 2 samples per word

// ---- sample-type helpers: 32-bit words of 4 (u8) or 2 (u16) samples -------------------------------
template <typename Sample>
struct Word;
template <>
struct Word<uint8_t>
{
    static constexpr int kLog2Spw = 2;
    static __device__ __forceinline__ uint32_t load(const uint8_t *p) { return hvbLoad4u8(p); }
    static __device__ __forceinline__ int sad(uint32_t a, uint32_t b, int acc) { return __vsadu4(a, b) + acc; }
};
template <>
struct Word<uint16_t>
{
    static constexpr int kLog2Spw = 1;
    static __device__ __forceinline__ uint32_t load(const uint16_t *p)
    {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
        const uint32_t lo = __ldg(q);
        if (!(a & 2)) return lo;
        return __funnelshift_r(lo, __ldg(q + 1), 16);
    }
    static __device__ __forceinline__ int sad(uint32_t a, uint32_t b, int acc) { return __vsadu2(a, b) + acc; }
};

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

__device__ __forceinline__ long long shflXor64(long long v, int m)
{
    const int lo = __shfl_xor_sync(0xffffffffu, (int)(v & 0xffffffffLL), m);
    const int hi = __shfl_xor_sync(0xffffffffu, (int)(v >> 32), m);
    return ((long long)hi << 32) | (unsigned)lo;
}

__device__ __constant__ int8_t kDiamond4[8] = {-4, 0, 0, 4, 4, 0, 0, -4};
__device__ __constant__ int8_t kHexagon8[16] = {0, -8, 8, -4, 8, 4, 0, 8, -8, 4, -8, -4, -8, 4, -8, -4};
__device__ __constant__ int8_t kDiamond16[32] = {0,  -4, 1,  -3, 2,  -2, 3,  -1, 4,  0, 3,  1,  2,  2,  1,  3,
                                                 0,  4,  -1, 3,  -2, 2,  -3, 1,  -4, 0, -3, -1, -2, -2, -1, -3};
__device__ __constant__ int8_t kSquare4[8] = {-4, -4, -4, 4, 4, 4, 4, -4};
__device__ __constant__ int8_t kDiamond1[8] = {0, -1, -1, 0, 0, 1, 1, 0};
__device__ __constant__ int8_t kLumaTaps[4][8] = {{0, 0, 0, 64, 0, 0, 0, 0},
                                                  {-1, 4, -10, 58, 17, -5, 1, 0},
                                                  {-1, 4, -11, 40, 40, -11, 4, -1},
                                                  {0, 1, -5, 17, 58, -10, 4, -1}};

template <typename Sample>
struct Search
{
    const hvb_me_task &t;
    const Sample *ref; // sample (x0, y0) of the reference plane
    int sr;
    const uint32_t *srcWords; // shared: the PU packed as words, row-major, wpr words per row
    const Sample *srcS;       // the same memory viewed as samples (row stride = w)
    int lane;
    int wpr, words, wprInv; // words per row, words in the block, ceil(65536 / wpr)
    Cand best;
    int nSad;

    // loadSource = false: the caller fills the shared block itself (the bi search compares against 2*src - predOther)
    __device__ Search(const hvb_me_task &task, const HvbPlane *planes, uint32_t *smemSrc, int lane_, bool loadSource = true)
        : t(task), lane(lane_)
    {
        const HvbPlane &sp = planes[t.src_pic * 3], &rp = planes[t.ref_pic * 3];
        sr = rp.stride;
        ref = reinterpret_cast<const Sample *>(rp.base) + (intptr_t)t.y0 * sr + t.x0;
        const Sample *src = reinterpret_cast<const Sample *>(sp.base) + (intptr_t)t.y0 * sp.stride + t.x0;
        wpr = t.w >> Word<Sample>::kLog2Spw;
        words = wpr * t.h;
        wprInv = (65536 + wpr - 1) / wpr;
        for (int i = lane; loadSource && i < words; i += 32)
        {
            const int y = (i * wprInv) >> 16, xw = i - y * wpr;
            smemSrc[i] = Word<Sample>::load(src + y * sp.stride + (xw << Word<Sample>::kLog2Spw));
        }
        srcWords = smemSrc;
        srcS = reinterpret_cast<const Sample *>(smemSrc);
        best.cost = 0x7fffffffffffffffLL;
        best.mv = best.mvd = hvb_mv{0, 0};
        best.mvpFlag = 0;
        nSad = 0;
        __syncwarp();
    }

    __device__ __forceinline__ void limit(int &x, int &y) const
    {
        x = min(max(x, (int)t.limitMin.x), (int)t.limitMax.x);
        y = min(max(y, (int)t.limitMin.y), (int)t.limitMax.y);
    }

    // SADs of `nc` (1..32) full-pel candidates at once.  Lane c < nc passes candidate c's displacement and
    // receives its SAD; the block's words are split over 32/nextpow2(nc) lanes per candidate.
    __device__ __noinline__ int sadMulti(int nc, int mvx, int mvy)
    {
        nSad += nc;
        const int log2p = nc <= 1 ? 0 : 32 - __clz(nc - 1); // ceil(log2(nc))
        const int lanesPer = 32 >> log2p;
        const int cand = lane >> (5 - log2p), sub = lane & (lanesPer - 1);
        const int cx = __shfl_sync(0xffffffffu, mvx, cand), cy = __shfl_sync(0xffffffffu, mvy, cand);
        int acc = 0;
        if (cand < nc)
        {
            const Sample *r = ref + (intptr_t)cy * sr + cx;
            if (wpr == 4 || wpr == 8 || wpr == 16 || wpr == 32)
            {
                // rows of 16, 32, 64 (128) bytes: a lane takes 16 bytes of a row at a time -- five aligned words of the
                // candidate (all requested before the first use), four funnel shifts, one 128-bit read of the source
                const int log2upr = wpr == 4 ? 0 : (wpr == 8 ? 1 : (wpr == 16 ? 2 : 3)), units = t.h << log2upr;
                const uintptr_t a0 = reinterpret_cast<uintptr_t>(r);
                const unsigned sh = (unsigned)(a0 & 3) * 8;
                const uint32_t *q0 = reinterpret_cast<const uint32_t *>(a0 & ~uintptr_t(3));
                const int srwB = (sr * (int)sizeof(Sample)) >> 2; // row stride in words
                for (int u = sub; u < units; u += lanesPer)
                {
                    const int y = u >> log2upr, xu = u & ((1 << log2upr) - 1);
                    const uint32_t *q = q0 + y * srwB + xu * 4;
                    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = __ldg(q + 3), w4 = __ldg(q + 4);
                    const uint4 s4 = *reinterpret_cast<const uint4 *>(srcWords + y * wpr + xu * 4);
                    acc = Word<Sample>::sad(s4.x, __funnelshift_r(w0, w1, sh), acc);
                    acc = Word<Sample>::sad(s4.y, __funnelshift_r(w1, w2, sh), acc);
                    acc = Word<Sample>::sad(s4.z, __funnelshift_r(w2, w3, sh), acc);
                    acc = Word<Sample>::sad(s4.w, __funnelshift_r(w3, w4, sh), acc);
                }
            }
            else if (wpr == 2)
            {
                // rows of 8 bytes: a lane takes a whole row (three aligned words, two funnel shifts)
                const uintptr_t a0 = reinterpret_cast<uintptr_t>(r);
                const unsigned sh = (unsigned)(a0 & 3) * 8;
                const uint32_t *q0 = reinterpret_cast<const uint32_t *>(a0 & ~uintptr_t(3));
                const int srwB = (sr * (int)sizeof(Sample)) >> 2;
                for (int y = sub; y < t.h; y += lanesPer)
                {
                    const uint32_t *q = q0 + y * srwB;
                    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
                    const uint2 s2 = *reinterpret_cast<const uint2 *>(srcWords + y * 2);
                    acc = Word<Sample>::sad(s2.x, __funnelshift_r(w0, w1, sh), acc);
                    acc = Word<Sample>::sad(s2.y, __funnelshift_r(w1, w2, sh), acc);
                }
            }
            else
                for (int i = sub; i < words; i += lanesPer)
                {
                    const int y = (i * wprInv) >> 16, xw = i - y * wpr;
                    acc = Word<Sample>::sad(srcWords[i], Word<Sample>::load(r + y * sr + (xw << Word<Sample>::kLog2Spw)), acc);
                }
        }
        for (int o = lanesPer >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        acc = __shfl_sync(0xffffffffu, acc, (lane << (5 - log2p)) & 31); // candidate `lane`'s group leader
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

    // The reference considers candidates one after the other with a strict `<`: the winner is the first
    // candidate of least cost, and it replaces `best` only if it is strictly cheaper.  Lane c holds candidate c.
    __device__ __noinline__ bool considerLanes(const Cand &mine, bool valid)
    {
        long long cost = valid ? mine.cost : 0x7fffffffffffffffLL;
        int who = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            const long long oc = shflXor64(cost, o);
            const int ow = __shfl_xor_sync(0xffffffffu, who, o);
            if (oc < cost || (oc == cost && ow < who))
            {
                cost = oc;
                who = ow;
            }
        }
        const bool improved = cost < best.cost;
        // every lane takes part in the broadcast; only the decision is conditional
        const int mvx = __shfl_sync(0xffffffffu, (int)mine.mv.x, who), mvy = __shfl_sync(0xffffffffu, (int)mine.mv.y, who);
        const int dx = __shfl_sync(0xffffffffu, (int)mine.mvd.x, who), dy = __shfl_sync(0xffffffffu, (int)mine.mvd.y, who);
        const int flag = __shfl_sync(0xffffffffu, mine.mvpFlag, who);
        if (improved)
        {
            best.cost = cost;
            best.mv = hvb_mv{(int16_t)mvx, (int16_t)mvy};
            best.mvd = hvb_mv{(int16_t)dx, (int16_t)dy};
            best.mvpFlag = flag;
        }
        return improved;
    }

    // one full-pel candidate given as a clamped full-pel vector: SAD with the whole warp, then consider
    __device__ bool considerOne(Cand c, int fx, int fy)
    {
        const int sad = sadMulti(1, fx, fy);
        c.cost += (long long)t.lambda * sad;
        if (c.cost < best.cost)
        {
            best = c;
            return true;
        }
        return false;
    }

    // StateMeFullPel::considerPattern (Search.hpp:1447-1482): pattern entries j = 0, step, 2*step, ... < n
    __device__ __noinline__ bool considerPattern(hvb_mv origin, const int8_t *pattern, int n, int step, int dist)
    {
        const int nc = n / step;
        int fx = 0, fy = 0;
        if (lane < nc)
        {
            const int8_t *p = pattern + 2 * lane * step;
            fx = (int16_t)((origin.x + dist * p[0]) / 4);
            fy = (int16_t)((origin.y + dist * p[1]) / 4);
            limit(fx, fy);
        }
        const int sad = sadMulti(nc, fx, fy);
        Cand c = makeCandidate(hvb_mv{(int16_t)(fx * 4), (int16_t)(fy * 4)});
        c.cost += (long long)t.lambda * sad;
        return considerLanes(c, lane < nc);
    }
};

template <typename Sample>
__device__ __noinline__ bool metTerminates(Search<Sample> &s)
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

    // zero vector, not clamped (:2103-2129)
    if (s.considerOne(s.makeCandidate(hvb_mv{0, 0}), 0, 0) && t.met && metTerminates(s)) return true;

    for (int flag = 0; flag < 2; ++flag) // the predictors rounded to full-pel (:2131-2171)
    {
        Cand c;
        c.mvpFlag = flag;
        int fx = (int16_t)(t.mvp[flag].x + 1) >> 2, fy = (int16_t)(t.mvp[flag].y + 1) >> 2;
        s.limit(fx, fy);
        c.mv = hvb_mv{(int16_t)(fx << 2), (int16_t)(fy << 2)};
        c.mvd.x = (int16_t)(c.mv.x - t.mvp[flag].x);
        c.mvd.y = (int16_t)(c.mv.y - t.mvp[flag].y);
        c.cost = rateOfMvd(c.mvd.x, c.mvd.y) + t.rateMvpFlag[flag];
        const int sad = s.sadMulti(1, c.mv.x >> 2, c.mv.y >> 2);
        c.cost += (long long)t.lambda * sad;
        costMvdZero[flag] = c.cost;
        bool better = false;
        if (c.cost < s.best.cost)
        {
            s.best = c;
            better = true;
        }
        if (better && t.met && metTerminates(s)) return true;
    }
    if (t.usePrev2Nx2N) // previous 2Nx2N integer vector (:2173-2198)
    {
        int fx = t.prev2Nx2N.x >> 2, fy = t.prev2Nx2N.y >> 2;
        s.limit(fx, fy);
        const hvb_mv mv{(int16_t)(fx << 2), (int16_t)(fy << 2)};
        if (s.considerOne(s.makeCandidate(mv), mv.x >> 2, mv.y >> 2) && t.met && metTerminates(s)) return true;
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
    if (distBest > 5) // raster (:2258-2273): rows of `line` patterns = a 5-sample grid of absolute displacements.
    {                 // None of its candidates depends on `best`, so 32 of them are evaluated per round.
        const int cols = 4 * ((2 * raster) / 80 + 1), rows = (2 * raster) / 20 + 1, total = rows * cols;
        for (int base = 0; base < total; base += 32)
        {
            const int q = base + s.lane, nc = min(32, total - base);
            int fx = 0, fy = 0;
            if (q < total)
            {
                const int row = q / cols, col = q - row * cols;
                fx = (-raster + 20 * col) / 4;
                fy = (-raster + 20 * row) / 4;
                s.limit(fx, fy);
            }
            const int sad = s.sadMulti(nc, fx, fy);
            Cand c = s.makeCandidate(hvb_mv{(int16_t)(fx * 4), (int16_t)(fy * 4)});
            c.cost += (long long)t.lambda * sad;
            s.considerLanes(c, q < total);
        }
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
        bool again;
        do
        {
            int fx = 0, fy = 0;
            if (s.lane < 4)
            {
                fx = (int16_t)(s.best.mv.x / 4 + kDiamond1[2 * s.lane]);
                fy = (int16_t)(s.best.mv.y / 4 + kDiamond1[2 * s.lane + 1]);
                s.limit(fx, fy);
            }
            const int sad = s.sadMulti(4, fx, fy);
            Cand c = s.makeCandidate(hvb_mv{(int16_t)(fx * 4), (int16_t)(fy * 4)});
            c.cost += (long long)t.lambda * sad;
            again = s.considerLanes(c, s.lane < 4);
        } while (again);
    }
    return false;
}

// The integer search of the PUs larger than 8x8, a warp each (the smaller ones: hvb_me_small.cu); hvb_me_subpel.cu then
// refines the whole batch in its own launch.
template <typename Sample>
__global__ void __launch_bounds__(kWarps * 32)
    meSearchKernel(const HvbPlane *__restrict__ planes, const hvb_me_task *__restrict__ tasks, int n, hvb_me_result *__restrict__ out,
                   int bitDepth)
{
    extern __shared__ __align__(16) uint32_t smemMe[];
    constexpr int kBlockWords = sizeof(Sample) == 1 ? 64 * 64 / 4 : kSrcWords;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *sSrc = smemMe + warp * kBlockWords;
    const int warpsTotal = gridDim.x * kWarps;
    for (int i = blockIdx.x * kWarps + warp; i < n; i += warpsTotal)
    {
        const hvb_me_task t = tasks[i];
        if (t.w <= 8 && t.h <= 8) continue; // PUs up to 8x8: hvb_me_small.cu, four per warp
        Search<Sample> s(t, planes, sSrc, lane);
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
        const hvb_mv mv = s.best.mv, mvd = s.best.mvd; // the sub-pel kernel starts from here
        r.mv = mv;
        r.mvd = mvd;
        r.nSad = s.nSad;
        if (lane == 0) out[i] = r;
        __syncwarp();
    }
}

// ---- bi-directional refinement: searchMotionBi (turing/Search.hpp:1498-1653) -------------------------
//
// One call refines list X's vector of a bi-predicted PU against the "ideal" block 2*src - predOther.  The warp
// builds that block in its shared source slot (8-tap prediction from the other list, SubtractBi at bit depth
// 6 + 2*sizeof(Sample)), then runs the reference's exhaustive (2r+1)^2 integer grid 32 candidates at a time and
// the two 3x3 fractional rounds 9 candidates at a time, each followed by the ordered arg-min.

// hvb_me_bi_task shares its first 56 bytes with hvb_me_task (pictures, block, predictors, rates, lambda, limits),
// so the Search<> machinery reads it through that type.
static_assert(offsetof(hvb_me_bi_task, mvp) == offsetof(hvb_me_task, mvp) &&
              offsetof(hvb_me_bi_task, rateMvpFlag) == offsetof(hvb_me_task, rateMvpFlag) &&
              offsetof(hvb_me_bi_task, lambda) == offsetof(hvb_me_task, lambda) &&
              offsetof(hvb_me_bi_task, limitMin) == offsetof(hvb_me_task, limitMin) &&
              offsetof(hvb_me_bi_task, limitMax) == offsetof(hvb_me_task, limitMax) &&
              sizeof(hvb_me_bi_task) == sizeof(hvb_me_task), "hvb_me_bi_task must overlay hvb_me_task");

// column q, rows r0..r0+3 of the 8-tap prediction at quarter-pel (mvx, mvy); `R` = sample (0,0) of the block in the
// reference plane.  Registers only: 11 horizontally filtered values feed 4 vertical filters.
template <typename Sample>
__device__ __forceinline__ void predColumn4(const Sample *R, int sr, int q, int r0, int mvx, int mvy, int bitDepth, int (&out)[4])
{
    const int shift1 = min(4, bitDepth - 8), shift3 = max(2, 14 - bitDepth);
    const int maxv = (1 << bitDepth) - 1;
    const int xf = mvx & 3, yf = mvy & 3;
    const Sample *p0 = R + (intptr_t)(r0 + (mvy >> 2) - 3) * sr + (q + (mvx >> 2) - 3);
    int mids[11];
#pragma unroll
    for (int r = 0; r < 11; ++r)
    {
        const Sample *p = p0 + r * sr;
        int mid = 0;
        if (xf)
        {
#pragma unroll
            for (int k = 0; k < 8; ++k) mid += kLumaTaps[xf][k] * (int)__ldg(p + k);
        }
        else
            mid = (int)__ldg(p + 3) << 6;
        mids[r] = mid >> shift1;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        int v = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) v += kLumaTaps[yf][k] * mids[i + k];
        out[i] = hvbClip3(0, maxv, (v + (1 << (5 + shift3))) >> (6 + shift3));
    }
}

template <typename Sample>
__global__ void __launch_bounds__(kWarps * 32)
    meBiSearchKernel(const HvbPlane *__restrict__ planes, const hvb_me_bi_task *__restrict__ tasks, int n,
                     hvb_me_bi_result *__restrict__ out, int bitDepth)
{
    extern __shared__ __align__(16) uint32_t smemMe[];
    constexpr int kBlockWords = sizeof(Sample) == 1 ? 64 * 64 / 4 : kSrcWords;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *sSrc = smemMe + warp * kBlockWords;
    const int warpsTotal = gridDim.x * kWarps;
    for (int i = blockIdx.x * kWarps + warp; i < n; i += warpsTotal)
    {
        const hvb_me_bi_task bt = tasks[i];
        const hvb_me_task &t = reinterpret_cast<const hvb_me_task &>(bt);
        Search<Sample> s(t, planes, sSrc, lane, false);

        // ---- ideal block (:1518-1548) ----
        {
            const HvbPlane &sp = planes[bt.src_pic * 3], &op = planes[bt.other_pic * 3];
            const Sample *src = reinterpret_cast<const Sample *>(sp.base) + (intptr_t)bt.y0 * sp.stride + bt.x0;
            int ox = bt.mvOther.x >> 2, oy = bt.mvOther.y >> 2;
            s.limit(ox, oy); // only the integer part is clamped; the fraction is kept
            const int omvx = (ox << 2) | (bt.mvOther.x & 3), omvy = (oy << 2) | (bt.mvOther.y & 3);
            const Sample *other = reinterpret_cast<const Sample *>(op.base) + (intptr_t)bt.y0 * op.stride + bt.x0;
            Sample *ideal = reinterpret_cast<Sample *>(sSrc);
            const int idealMax = (1 << (6 + 2 * (int)sizeof(Sample))) - 1;
            const int jobs = bt.w * (bt.h >> 2);
            for (int job = lane; job < jobs; job += 32)
            {
                const int strip = job / bt.w, q = job - strip * bt.w;
                int pred[4];
                predColumn4<Sample>(other, op.stride, q, strip * 4, omvx, omvy, bitDepth, pred);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    const int y = strip * 4 + k;
                    ideal[y * bt.w + q] = (Sample)hvbClip3(0, idealMax, 2 * (int)__ldg(src + (intptr_t)y * sp.stride + q) - pred[k]);
                }
            }
            __syncwarp();
        }

        // ---- exhaustive integer grid (:1551-1625) ----
        int sx = (int16_t)(bt.mvStart.x + 1) >> 2, sy = (int16_t)(bt.mvStart.y + 1) >> 2;
        s.limit(sx, sy);
        s.best.mv = hvb_mv{(int16_t)(sx << 2), (int16_t)(sy << 2)};
        const int range = bt.smallWindow ? 1 : 5, side = 2 * range + 1, total = side * side;
        for (int base = 0; base < total; base += 32)
        {
            const int q = min(base + lane, total - 1), nc = min(32, total - base);
            const int row = q / side, col = q - row * side, k = col & 3;
            int cx = sx + col - range, cy = sy + row - range; // the candidate's own vector
            s.limit(cx, cy);
            // its SAD comes from member k of a SAD4 group whose base was clamped BEFORE k was added (:1590-1608)
            int px = sx + (col - k) - range, py = sy + row - range;
            s.limit(px, py);
            px += k;
            s.limit(px, py);
            const int sad = s.sadMulti(nc, px, py);
            Cand c = s.makeCandidate(hvb_mv{(int16_t)(cx << 2), (int16_t)(cy << 2)});
            c.cost += (long long)bt.lambda * sad;
            s.considerLanes(c, base + lane < total);
        }
        hvb_me_bi_result r;
        r.mvInteger = s.best.mv;
        r.nSad = side * ((side + 3) / 4) * 4; // what the reference's SAD4 calls evaluate
        r.reserved = 0;

        // the fractional rounds (:1628-1650) run in hvb_me_subpel.cu's kernel (BI = true), from mvInteger
        r.mv = s.best.mv;
        r.mvd = s.best.mvd;
        r.mvpFlag = s.best.mvpFlag;
        r.cost = s.best.cost;
        if (lane == 0) out[i] = r;
        __syncwarp();
    }
}

} // namespace

int hvbLaunchMeSubpel(hvb_context *ctx, const hvb_me_task *dTasks, int n, hvb_me_result *dOut); // hvb_me_subpel.cu
int hvbLaunchMeSmall(hvb_context *ctx, const hvb_me_task *dTasks, int n, hvb_me_result *dOut);  // hvb_me_small.cu

extern "C" int hvb_me_search_batch(hvb_context *ctx, const hvb_me_task *tasks, int n, hvb_me_result *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(hvb_me_result) * n, mem, &st);
    if (rc) return rc;
    int blocks = (n + kWarps - 1) / kWarps;
    const int cap = ctx->smCount * 4;
    if (blocks > cap) blocks = cap;
    const auto *dT = static_cast<const hvb_me_task *>(st.dTasks);
    auto *dO = static_cast<hvb_me_result *>(st.dOut);
    if (ctx->bps == 1)
    {
        // integer search -- PUs up to 8x8 four per warp (hvb_me_small.cu), the larger ones a warp each -- then the
        // sub-pel refinement of the whole batch (hvb_me_subpel.cu), all on the same stream
        rc = hvbLaunchMeSmall(ctx, dT, n, dO);
        if (rc) return rc;
        const int smem = kWarps * (64 * 64 / 4) * 4;
        cudaFuncSetAttribute(meSearchKernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int perSm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, meSearchKernel<uint8_t>, kWarps * 32, smem);
        blocks = min((n + kWarps - 1) / kWarps, ctx->smCount * max(perSm, 1));
        meSearchKernel<uint8_t><<<blocks, kWarps * 32, smem, ctx->stream>>>(ctx->dPlanes, dT, n, dO, ctx->bitDepth);
        HVB_LAUNCH_CHECK(ctx, "meSearchKernel");
        rc = hvbLaunchMeSubpel(ctx, dT, n, dO);
        if (rc) return rc;
    }
    else
    {
        // 16-bit samples: the same three kernels (templates on the sample type)
        rc = hvbLaunchMeSmall(ctx, dT, n, dO);
        if (rc) return rc;
        const int smem = kWarps * kSrcWords * 4;
        cudaFuncSetAttribute(meSearchKernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int perSm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, meSearchKernel<uint16_t>, kWarps * 32, smem);
        blocks = min((n + kWarps - 1) / kWarps, ctx->smCount * max(perSm, 1));
        meSearchKernel<uint16_t><<<blocks, kWarps * 32, smem, ctx->stream>>>(ctx->dPlanes, dT, n, dO, ctx->bitDepth);
        HVB_LAUNCH_CHECK(ctx, "meSearchKernel");
        rc = hvbLaunchMeSubpel(ctx, dT, n, dO);
        if (rc) return rc;
    }
    return hvbStageOut(ctx, out, sizeof(hvb_me_result) * n, mem, st);
}

int hvbLaunchMeBiSubpel(hvb_context *ctx, const hvb_me_bi_task *dTasks, int n, hvb_me_bi_result *dOut); // hvb_me_subpel.cu

extern "C" int hvb_me_bi_search_batch(hvb_context *ctx, const hvb_me_bi_task *tasks, int n, hvb_me_bi_result *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(hvb_me_bi_result) * n, mem, &st);
    if (rc) return rc;
    int blocks = (n + kWarps - 1) / kWarps;
    const int cap = ctx->smCount * 4;
    if (blocks > cap) blocks = cap;
    const auto *dT = static_cast<const hvb_me_bi_task *>(st.dTasks);
    auto *dO = static_cast<hvb_me_bi_result *>(st.dOut);
    if (ctx->bps == 1)
    {
        const int smem = kWarps * (64 * 64 / 4) * 4;
        cudaFuncSetAttribute(meBiSearchKernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        meBiSearchKernel<uint8_t><<<blocks, kWarps * 32, smem, ctx->stream>>>(ctx->dPlanes, dT, n, dO, ctx->bitDepth);
    }
    else
    {
        const int smem = kWarps * kSrcWords * 4;
        cudaFuncSetAttribute(meBiSearchKernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        meBiSearchKernel<uint16_t><<<blocks, kWarps * 32, smem, ctx->stream>>>(ctx->dPlanes, dT, n, dO, ctx->bitDepth);
    }
    HVB_LAUNCH_CHECK(ctx, "meBiSearchKernel");
    rc = hvbLaunchMeBiSubpel(ctx, dT, n, dO);
    if (rc) return rc;
    return hvbStageOut(ctx, out, sizeof(hvb_me_bi_result) * n, mem, st);
}
