// hvb_rdoq.cuh -- rate-distortion optimised quantisation, one THREAD per transform block.
//
// Reference semantics (bit-exact): Rdoq::runQuantisation, turing/Rdoq.cpp:35-450, helpers :452-887,
// signDataHiding :889-1023, constructor arithmetic turing/Rdoq.h:170-188; fixed-point cost algebra
// turing/FixedPoint.h + Cost.h (Cost = Q16 in int64, Lambda = Q16 in int32); scan tables
// turing/ScanOrder.h:32-101; bit-cost table turing/Write.h:413-436 indexed by
// ContextModel::getState() ^ bin = (state >> 1) ^ bin (Rdoq.cpp:26-31, ContextModel.h:58-61).
//
// The level decision of a coefficient depends on the CABAC level-coding state left behind by the
// previous one (greater1/greater2 counters, Rice parameter, context set), so the walk over one block
// is inherently serial; parallelism has to come from running many blocks at
// once.  The first version of this file gave a warp to each block and let lane 0 walk the recurrence:
// the ncu capture showed 4 active threads per instruction and the kernel at 10x the cost of the
// transforms around it.  Now the work is split by what is parallel:
//
//   warp, cooperative (hvbRdoqPrepass)   everything before the reverse scan meets its first non-zero
//                                        rounding level: its position and the two "distortion if zero"
//                                        sums the reference accumulates on the way (integer sums: any order);
//   one thread per block (hvbRdoqThread) the recurrence, so a warp advances 32 blocks at a time.  A 4x4
//                                        group whose levels all round to zero -- the common case -- costs
//                                        two running sums; only groups with non-zero levels store
//                                        per-coefficient records (the reference's 40 KB of Rdoq members per
//                                        block shrink to 32 bytes per coefficient of a coded group).
#pragma once
#include "hvb_internal.cuh"

// per-coefficient record, indexed by scan position
struct HvbCoefRec
{
    long long rdCost;  // m_rdCostCoeff
    long long rateSig; // m_rateCostCoeffSig
    int rateUp, rateDown, sigDelta, deltaU;
};

// result of the cooperative pre-pass of one block
struct HvbRdoqMid
{
    int lastSp; // first position of the reverse scan whose rounding level is non-zero; -1: none; -2: block not RDOQ'd
    int reserved;
    long long totalDist0; // m_totalDistCoeff0
    long long tailDist0;  // m_rdCostTu when the reverse scan reaches lastSp
};

namespace hvb_rdoq {

// in global memory, read through L1: the index differs from thread to thread and a constant-bank read serialises per address
static __device__ const int32_t kEntropyBits[128] = {
    0x07b23, 0x085f9, 0x074a0, 0x08cbc, 0x06ee4, 0x09354, 0x067f4, 0x09c1b, 0x060b0, 0x0a62a, 0x05a9c, 0x0af5b, 0x0548d,
    0x0b955, 0x04f56, 0x0c2a9, 0x04a87, 0x0cbf7, 0x045d6, 0x0d5c3, 0x04144, 0x0e01b, 0x03d88, 0x0e937, 0x039e0, 0x0f2cd,
    0x03663, 0x0fc9e, 0x03347, 0x10600, 0x03050, 0x10f95, 0x02d4d, 0x11a02, 0x02ad3, 0x12333, 0x0286e, 0x12cad, 0x02604,
    0x136df, 0x02425, 0x13f48, 0x021f4, 0x149c4, 0x0203e, 0x1527b, 0x01e4d, 0x15d00, 0x01c99, 0x166de, 0x01b18, 0x17017,
    0x019a5, 0x17988, 0x01841, 0x18327, 0x016df, 0x18d50, 0x015d9, 0x19547, 0x0147c, 0x1a083, 0x0138e, 0x1a8a3, 0x01251,
    0x1b418, 0x01166, 0x1bd27, 0x01068, 0x1c77b, 0x00f7f, 0x1d18e, 0x00eda, 0x1d91a, 0x00e19, 0x1e254, 0x00d4f, 0x1ec9a,
    0x00c90, 0x1f6e0, 0x00c01, 0x1fef8, 0x00b5f, 0x208b1, 0x00ab6, 0x21362, 0x00a15, 0x21e46, 0x00988, 0x2285d, 0x00934,
    0x22ea8, 0x008a8, 0x239b2, 0x0081d, 0x24577, 0x007c9, 0x24ce6, 0x00763, 0x25663, 0x00710, 0x25e8f, 0x006a0, 0x26a26,
    0x00672, 0x26f23, 0x005e8, 0x27ef8, 0x005ba, 0x284b5, 0x0055e, 0x29057, 0x0050c, 0x29bab, 0x004c1, 0x2a674, 0x004a7,
    0x2aa5e, 0x0046f, 0x2b32f, 0x0041f, 0x2c0ad, 0x003e7, 0x2ca8d, 0x003ba, 0x2d323, 0x0010c, 0x3bfbb};

struct Engine;
__device__ __forceinline__ int bitsOf(const Engine &e, int bin, const uint8_t &state);

__device__ __forceinline__ long long shflXor64(long long v, int m)
{
    const int lo = __shfl_xor_sync(0xffffffffu, (int)(v & 0xffffffffLL), m);
    const int hi = __shfl_xor_sync(0xffffffffu, (int)(v >> 32), m);
    return ((long long)hi << 32) | (unsigned)lo;
}

// scan position -> (x, y) inside a (1 << log2)^2 grid, log2 <= 3 (ScanOrder.h:32-101).
// The up-right diagonal order is generated arithmetically: diagonal d holds min(d, n-1) - max(0, d-n+1) + 1 cells.
__host__ __device__ inline void scanXY(int log2, int scanIdx, int pos, int &x, int &y)
{
    const int n = 1 << log2;
    if (log2 == 0)
    {
        x = y = 0;
        return;
    }
    if (scanIdx == 1)
    {
        x = pos & (n - 1);
        y = pos >> log2;
        return;
    }
    if (scanIdx == 2)
    {
        x = pos >> log2;
        y = pos & (n - 1);
        return;
    }
    int d = 0, before = 0;
    for (;; ++d)
    {
        const int lo = d - (n - 1) > 0 ? d - (n - 1) : 0, hi = d < n - 1 ? d : n - 1;
        const int cells = hi - lo + 1;
        if (pos < before + cells)
        {
            x = lo + (pos - before);
            y = d - x;
            return;
        }
        before += cells;
    }
}

// raster position of scan position sp of a (1 << log2)^2 block: group order, then the 4x4 order (Rdoq.cpp:399-412)
__host__ __device__ inline int scanToRaster(int log2, int scanIdx, int sp)
{
    int gx, gy, x, y;
    scanXY(log2 - 2, scanIdx, sp >> 4, gx, gy);
    scanXY(2, scanIdx, sp & 15, x, y);
    return (((gy << 2) + y) << log2) + (gx << 2) + x;
}

// table of scanToRaster for log2 2..5 x scanIdx 0..2, 1024 entries each (filled by hvbRdoqInitTables)
static __device__ short gScanTable[4 * 3 * 1024];
__device__ __forceinline__ const short *scanTable(int log2, int scanIdx) { return gScanTable + ((log2 - 2) * 3 + scanIdx) * 1024; }

struct Engine
{
    const hvb_rdoq_ctx *cx;
    const int2 *bits; // per state byte of *cx: {bits of bin 0, bits of bin 1} (hvbLaunchRdoqBits), one load instead of two dependent ones
    const int *lastTab; // this block's (cIdx != 0, log2) slice of the last-position rate table: [0..9] x prefix, [10..19] y prefix
    HvbCoefRec *rec;
    const short *scan;
    const int16_t *src;
    int lambda, distScale, shdFactor;
    int iqScale, iqShift, iqOffset;
    int log2, cIdx, scanIdx;

    // Rdoq::Rdoq (Rdoq.h:170-188); FixedPoint<int32,16>::set(double) = int32(d * 65536 + 0.5)
    __device__ __forceinline__ void init(const hvb_rdoq_ctx *ctx, int iqScale_, int log2_, int cIdx_, int scanIdx_, int bitDepth)
    {
        cx = ctx;
        log2 = log2_;
        cIdx = cIdx_;
        scanIdx = scanIdx_;
        scan = scanTable(log2_, scanIdx_);
        const double lam = ctx->lambda;
        lambda = (int)(lam * 65536 + 0.5);
        shdFactor = (int)(iqScale_ * iqScale_ / lam / 16 + 0.5);
        const int transformShift = 15 - bitDepth - log2_;
        const int distShift = 15 - 2 * transformShift - 2 * (bitDepth - 8);
        distScale = (int)((double)(1 << distShift) * 65536 + 0.5);
        iqScale = iqScale_;
        iqShift = 20 - 14 - transformShift;
        iqOffset = 1 << (iqShift - 1);
    }
    __device__ __forceinline__ long long lam(int rate) const { return (long long)lambda * rate; }
    __device__ __forceinline__ long long dist(int err) const
    {
        const int sq = (int)((unsigned)err * (unsigned)err);
        return (long long)sq * distScale;
    }
};

// the bit cost of coding `bin` with the context whose state byte is `state` (a member of *e.cx)
__device__ __forceinline__ int bitsOf(const Engine &e, int bin, const uint8_t &state)
{
    const int2 v = e.bits[&state - reinterpret_cast<const uint8_t *>(e.cx)]; // (plain loads: the table may be in shared memory, tuFusedKernel)
    return bin ? v.y : v.x;
}

// both bins' bit costs of the three contexts a coefficient's decision reads, requested together (independent loads)
// instead of one by one inside the rate functions
struct CoefBits
{
    int2 sig, g1, g2;
};
__device__ __forceinline__ int2 bitsBoth(const Engine &e, const uint8_t &state)
{
    return e.bits[&state - reinterpret_cast<const uint8_t *>(e.cx)];
}

__device__ __forceinline__ int baseLevel(int g1Cnt, int g2Cnt) { return g1Cnt < 8 ? 2 + (g2Cnt < 1) : 1; }

// Rdoq.cpp:512-598
__device__ inline int sigCtxInc(int prevCsbf, int scanIdx, int xC, int yC, int log2, int cIdx)
{
    int inc;
    if (log2 == 2)
    {
        // {0,1,4,5, 2,3,4,5, 6,6,8,8, 7,7,8,8} packed 4 bits each, index (yC << 2) + xC
        const unsigned long long map = 0x8877886654325410ull;
        inc = (int)((map >> (4 * ((yC << 2) + xC))) & 15);
    }
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

// sigCtxInc for the 16 scan positions of one coefficient group without decoding positions.  Inside a group the increment
// is base(group) + f(neighbour pattern, xP, yP) with f in {0, 1, 2}; (xP, yP) of scan position k is fixed by the scan, so f
// is two bits per k of a word chosen by (scanIdx, neighbour pattern).  4x4 blocks: ctxIdxMap, four bits per k.  (Generated
// from sigCtxInc above and checked against it for every block size, plane class, scan, group and pattern.)
static __device__ const uint32_t kSigF[3][4] = {{0x556u, 0x1090926u, 0x10619au, 0xaaaaaaaau},
                                                {0x10516u, 0x55aau, 0x6060606u, 0xaaaaaaaau},
                                                {0x10516u, 0x6060606u, 0x55aau, 0xaaaaaaaau}};
static __device__ const unsigned long long kSigMap4[3] = {0x8885875467436120ull, 0x8877886654325410ull, 0x8855884476317620ull};

struct GroupSigCtx
{
    unsigned long long bits; // 2 (or 4) bits per scan position
    int base, dc, shift;     // dc: context of the block's DC coefficient (group 0, position 0), -1 elsewhere
    __device__ __forceinline__ GroupSigCtx(int prev, int scanIdx, int cg, int cgX, int cgY, int log2, int cIdx)
    {
        const int plane = cIdx == 0 ? 0 : 27;
        if (log2 == 2)
        {
            bits = __ldg(&kSigMap4[scanIdx]);
            base = plane;
            dc = -1;
            shift = 2; // 4 bits per position
        }
        else
        {
            bits = __ldg(&kSigF[scanIdx][prev]);
            base = cIdx == 0 ? (cgX + cgY > 0 ? 3 : 0) + (log2 == 3 ? (scanIdx == 0 ? 9 : 15) : 21) : 27 + (log2 == 3 ? 9 : 12);
            dc = cg == 0 ? plane : -1;
            shift = 1;
        }
    }
    __device__ __forceinline__ int at(int k) const
    {
        if (k == 0 && dc >= 0) return dc;
        return base + (int)((bits >> (k << shift)) & (shift == 2 ? 15 : 3));
    }
};

// Rdoq.cpp:619-673
__device__ inline long long levelRateCost(const Engine &e, int level, const CoefBits &cb, int rice, int g1Cnt, int g2Cnt)
{
    int rate = 32768;
    const int base = baseLevel(g1Cnt, g2Cnt);
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
            rate += cb.g1.y;
            if (g2Cnt < 1) rate += cb.g2.y;
        }
    }
    else if (level == 1)
        rate += cb.g1.x;
    else if (level == 2)
        rate += cb.g1.y + cb.g2.x;
    return e.lam(rate);
}

// Rdoq.cpp:805-870
__device__ inline int levelRate(const Engine &e, int level, const CoefBits &cb, int rice, int g1Cnt, int g2Cnt)
{
    int rate = 0;
    const int base = baseLevel(g1Cnt, g2Cnt);
    if (level >= base)
    {
        int symbol = level - base;
        // golombRiceRange {7,14,26,46,78} and golombRicePrefixLen {8,7,6,5,4}
        const int maxVlc = rice == 0 ? 7 : rice == 1 ? 14 : rice == 2 ? 26 : rice == 3 ? 46 : 78;
        const int maxPrefix = 8 - rice;
        if (symbol > maxVlc)
        {
            const int rest = symbol - maxVlc;
            int egs = 1;
            for (int m = 2; rest >= m; m <<= 1) egs += 2;
            rate += egs << 15;
            symbol = min(symbol, maxVlc + 1);
        }
        rate += (min(symbol >> (rice + 1), maxPrefix) + rice) << 15;
        if (g1Cnt < 8)
        {
            rate += cb.g1.y;
            if (g2Cnt < 1) rate += cb.g2.y;
        }
    }
    else if (level == 1)
        rate += cb.g1.x;
    else if (level == 2)
        rate += cb.g1.y + cb.g2.x;
    return rate;
}

// Rdoq.cpp:452-510
__device__ inline int adjustLevel(const Engine &e, int absCoeff, int q, const CoefBits &cb, int rice, int g1Cnt, int g2Cnt, bool isLast,
                                  long long &rdCost, long long &rateSig)
{
    long long sigCost = 0;
    int best = 0;
    if (!isLast && q < 3)
    {
        rateSig = e.lam(cb.sig.x);
        rdCost = e.dist(absCoeff) + rateSig;
        if (q == 0) return 0;
    }
    else
        rdCost = 0x7fffffffffffffffLL;
    if (!isLast) sigCost = e.lam(cb.sig.y);
    const int lowest = q > 1 ? q - 1 : 1;
    for (int level = q; level >= lowest; --level)
    {
        const int recon = hvbClip3(-32768, 32767, (hvbClip3(-32768, 32767, level) * e.iqScale + e.iqOffset) >> e.iqShift);
        const long long c = e.dist(absCoeff - recon) + levelRateCost(e, level, cb, rice, g1Cnt, g2Cnt) + sigCost;
        if (c < rdCost)
        {
            best = level;
            rdCost = c;
            rateSig = sigCost;
        }
    }
    return best;
}

// Rdoq.cpp:742-760
__device__ __forceinline__ int lastPrefixCtx(int binIdx, int cIdx, int log2)
{
    const int off = cIdx ? 15 : 3 * (log2 - 2) + ((log2 - 1) >> 2);
    const int sh = cIdx ? log2 - 2 : (log2 + 1) >> 2;
    return hvbClip3(0, 17, (binIdx >> sh) + off);
}

__device__ __forceinline__ int lastLen(int v) // binarisationLengthForPosition (Rdoq.cpp:704)
{
    return v < 4 ? v : (v < 6 ? 4 : (v < 8 ? 5 : (v < 12 ? 6 : (v < 16 ? 7 : (v < 24 ? 8 : 9)))));
}

// Rdoq.cpp:699-740: the rate of a last-significant-coefficient prefix of binarisation length `len` (0..9) in one direction,
// for one context snapshot.  It depends on the snapshot, the colour plane class and the block size only, so it is tabulated
// when the snapshots are uploaded (hvbLaunchRdoqBits) instead of being summed bin by bin -- up to 20 dependent lookups --
// for every candidate last position.
constexpr int kLastTabPerCtx = 2 * 4 * 20; // [cIdx != 0][log2 - 2][x: 0..9, y: 10..19]
__device__ inline int lastPrefixRate(const hvb_rdoq_ctx &cx, bool isY, int len, int cIdx, int log2)
{
    const uint8_t *prefix = isY ? cx.last_y_prefix : cx.last_x_prefix;
    int rate = 0;
    for (int i = 0; i < len; ++i) rate += kEntropyBits[(prefix[lastPrefixCtx(i, cIdx, log2)] >> 1) ^ 1];
    if (len < 9) rate += kEntropyBits[prefix[lastPrefixCtx(len, cIdx, log2)] >> 1];
    if (len > 3) rate += 32768 * ((len - 2) >> 1);
    return rate;
}

__device__ __forceinline__ long long lastPosCost(const Engine &e, int xC, int yC)
{
    return e.lam(e.lastTab[lastLen(xC)] + e.lastTab[10 + lastLen(yC)]);
}

// neighbours right / below of coefficient group (xS, yS) in the 64-bit csbf mask (Rdoq.cpp:601-617, :675-697)
__device__ __forceinline__ void cgNeighbours(unsigned long long csbf, int xS, int yS, int log2, int &right, int &below)
{
    const int wcg = 1 << (log2 - 2);
    right = xS < wcg - 1 ? (int)((csbf >> (yS * wcg + xS + 1)) & 1) : 0;
    below = yS < wcg - 1 ? (int)((csbf >> ((yS + 1) * wcg + xS)) & 1) : 0;
}

// Rdoq.cpp:889-1023 for ONE coefficient group whose final levels are in registers (lv[k], scan order; negMask bit k: the
// coefficient at scan position k is negative).  `lastCG` is the reference's state: 1 for the first group (from the top)
// that still holds a level, 0 afterwards.  The reference guards the whole step with "sum of levels >= 2"; a group is only
// ever touched when it holds two levels at least 4 scan positions apart, which implies that sum, so the guard is
// not needed here and the step can run while the group's levels are still in registers.
__device__ __forceinline__ void signDataHidingGroup(const Engine &e, int cg, const int (&lv)[16], unsigned nzMask, unsigned negMask, int lastCG,
                                                    int16_t *dst)
{
    const int lastNZ = 31 - __clz(nzMask), firstNZ = __ffs(nzMask) - 1;
    if (lastNZ - firstNZ < 4) return;
    int absSum = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) absSum += lv[k];
    const int signbit = (negMask >> firstNZ) & 1;
    if (signbit == (absSum & 1)) return;
    const short *sc = e.scan + (cg << 4);
    const int kStart = lastCG == 1 ? lastNZ : 15;
    int minCost = 0x7fffffff, minK = -1, finalChange = 0;
#pragma unroll
    for (int q4 = 3; q4 >= 0; --q4)
    {
        if (4 * q4 > kStart) continue;
        // {rateUp, rateDown, sigDelta, deltaU} of the quad's four coefficients, requested together
        int4 r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) r[j] = *reinterpret_cast<const int4 *>(&e.rec[(cg << 4) + 4 * q4 + j].rateUp);
#pragma unroll
        for (int j = 3; j >= 0; --j)
        {
            const int k = 4 * q4 + j;
            if (k > kStart) continue;
            const int level = lv[k];
            const int rateUp = r[j].x, rateDown = r[j].y, sigDelta = r[j].z, deltaU = r[j].w;
            int cost, change;
            if (level != 0)
            {
                const int up = e.shdFactor * (-deltaU) + rateUp;
                int down = e.shdFactor * deltaU + rateDown - (abs(level) == 1 ? ((1 << 15) + sigDelta) : 0);
                if (lastCG == 1 && lastNZ == k && abs(level) == 1) down -= 4 << 15;
                if (up < down)
                {
                    cost = up;
                    change = 1;
                }
                else
                {
                    change = -1;
                    cost = (k == firstNZ && abs(level) == 1) ? 0x7fffffff : down;
                }
            }
            else
            {
                cost = e.shdFactor * (-abs(deltaU)) + (1 << 15) + rateUp + sigDelta;
                change = 1;
                if (k < firstNZ && (int)((negMask >> k) & 1) != signbit) cost = 0x7fffffff;
            }
            if (cost < minCost)
            {
                minCost = cost;
                finalChange = change;
                minK = k;
            }
        }
    }
    if (minK >= 0)
    {
        const int minPos = sc[minK];
        const int cur = dst[minPos];
        if (cur == 32767 || cur == -32768) finalChange = -1;
        dst[minPos] = (int16_t)(((negMask >> minK) & 1) ? cur - finalChange : cur + finalChange);
    }
}

} // namespace hvb_rdoq

// Cooperative pre-pass on a warp (Rdoq.cpp:104-118, :165-169).  Also writes zero levels everywhere, so the serial
// stage only touches coded positions.  `dst` may be shared or global.
__device__ inline HvbRdoqMid hvbRdoqPrepass(int16_t *dst, const int16_t *src, const hvb_rdoq_ctx *ctx, int qScale, int qShift, int iqScale,
                                            int log2, int cIdx, int scanIdx, int bitDepth, int lane)
{
    using namespace hvb_rdoq;
    const int n = 1 << (2 * log2);
    Engine e;
    e.init(ctx, iqScale, log2, cIdx, scanIdx, bitDepth);
    int lastSp = -1;
    long long total = 0;
    for (int sp = lane; sp < n; sp += 32)
    {
        const int pos = e.scan[sp];
        const int a = abs((int)src[pos]);
        total += e.dist(a);
        if (((a * qScale + (1 << (qShift - 1))) >> qShift) > 0) lastSp = sp; // sp ascends: keeps this lane's maximum
        dst[pos] = 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        lastSp = max(lastSp, __shfl_xor_sync(0xffffffffu, lastSp, o));
        total += shflXor64(total, o);
    }
    long long tail = 0;
    for (int sp = lastSp + 1 + lane; sp < n; sp += 32) tail += e.dist(abs((int)src[e.scan[sp]]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tail += shflXor64(tail, o);
    HvbRdoqMid m;
    m.lastSp = lastSp;
    m.reserved = 0;
    m.totalDist0 = total;
    m.tailDist0 = tail;
    return m;
}

// The serial stages of Rdoq::runQuantisation for one block on one thread.  `dst` holds zeros on entry (pre-pass),
// `rec` has room for n records.  Returns the OR of the coded levels.
__device__ inline int hvbRdoqThread(int16_t *dst, const int16_t *src, const hvb_rdoq_ctx *ctx, const HvbRdoqMid &mid, int qScale, int qShift,
                                    int iqScale, int log2, int cIdx, int scanIdx, bool isIntra, bool sdh, int bitDepth, HvbCoefRec *rec,
                                    const int2 *bits, const int *lastTab)
{
    using namespace hvb_rdoq;
    const int lastSp = mid.lastSp;
    if (lastSp < 0) return 0; // every level rounds to zero (Rdoq.cpp:308-312)
    Engine e;
    e.init(ctx, iqScale, log2, cIdx, scanIdx, bitDepth);
    e.rec = rec;
    e.src = src;
    e.bits = bits;
    e.lastTab = lastTab + ((cIdx ? 4 : 0) + (log2 - 2)) * 20;
    const int log2Cg = log2 - 2, mask = (1 << log2) - 1;

    long long rdCostTu = mid.tailDist0;
    long long rateCostCgSig[64];
    unsigned long long csbf = 0;
    const int lastCg = lastSp >> 4;
    int ctxSet = (lastSp < 16 || cIdx != 0) ? 0 : 2, g1Idx = 1, g1Cnt = 0, g2Cnt = 0, rice = 0;
    const int g1Off = cIdx > 0 ? 16 : 0, g2Off = cIdx > 0 ? 4 : 0;
    for (int i = 0; i <= lastCg; ++i) rateCostCgSig[i] = 0;

    // ---- stage 1 (Rdoq.cpp:89-305), from the first coded position downwards
    for (int cg = lastCg; cg >= 0; --cg)
    {
        int cgX, cgY, right, below;
        scanXY(log2Cg, scanIdx, cg, cgX, cgY);
        const int cgPos = cgY * (1 << log2Cg) + cgX;
        cgNeighbours(csbf, cgX, cgY, log2, right, below);
        const int prev = right + (below << 1);
        const int cSig = (cIdx == 0 ? 0 : 2) + min(right + below, 1); // coded_sub_block_flag context (neighbours only)
        const short *sc = e.scan + (cg << 4);
        const GroupSigCtx sig(prev, scanIdx, cg, cgX, cgY, log2, cIdx);

        // A group whose levels all round to zero (never lastSp's group): adjustLevel takes its q == 0 exit for
        // every coefficient (Rdoq.cpp:466-476), the state does not move, and since an uncoded group is skipped
        // by stage 2 and by sign hiding, only the two sums are needed.
        // (The DC group is always treated as coded, so stage 2 reads its records: it takes the general path.)
        bool anyLevel = cg == 0;
        long long sumRd = 0, sumSig = 0;
        // four coefficients at a time: their loads are independent of each other (the early exit sits between the quads)
        for (int k0 = 0; k0 < 16 && !anyLevel; k0 += 4)
        {
            int pos4[4], a4[4], bits4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) pos4[j] = sc[k0 + j];
#pragma unroll
            for (int j = 0; j < 4; ++j) a4[j] = abs((int)src[pos4[j]]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                bits4[j] = bitsOf(e, 0, ctx->sig_coeff_flag[sig.at(k0 + j)]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                if ((cg * 16 + k0 + j <= lastSp) && ((a4[j] * qScale + (1 << (qShift - 1))) >> qShift) > 0) anyLevel = true;
                const long long rs = e.lam(bits4[j]);
                sumSig += rs;
                sumRd += e.dist(a4[j]) + rs;
            }
        }
        if (!anyLevel)
        {
            rdCostTu += sumRd;
            // group boundary of updateEntropyCodingEngine (Rdoq.cpp:791-802); cg > 0 here
            rice = 0;
            g1Cnt = 0;
            g2Cnt = 0;
            ctxSet = (cg == 1 || cIdx != 0) ? 0 : 2;
            if (g1Idx == 0) ctxSet++;
            g1Idx = 1;
            // uncoded group: pay for its flag, drop the significance costs (Rdoq.cpp:206-216)
            const long long zero = e.lam(bitsOf(e, 0, ctx->coded_sub_block_flag[cSig]));
            rdCostTu += zero - sumSig;
            rateCostCgSig[cg] = zero;
            continue;
        }

        int nzBeforePos0 = 0;
        long long cgDist0 = 0, cgRateSig = 0, cgRateSigPos0 = 0, cgRdCoeff = 0;
        bool cgCoded = false;
        // positions beyond lastSp (only in its own group) are accounted for by the pre-pass: their rate terms are zero.
        // The next coefficient is requested one iteration ahead of the chain of decisions that consumes it.
        const int kStart = cg == lastCg ? (lastSp & 15) : 15;
        int posNext = sc[kStart], aNext = abs((int)src[posNext]);
        for (int k = kStart; k >= 0; --k)
        {
            const int sp = cg * 16 + k;
            const int pos = posNext, a = aNext;
            if (k > 0)
            {
                posNext = sc[k - 1];
                aNext = abs((int)src[posNext]);
            }
            const int scaled = a * qScale;
            const int q = (scaled + (1 << (qShift - 1))) >> qShift;
            const int sigCtx = sig.at(k);
            const long long d0 = e.dist(a);
            const int g1Ctx = 4 * ctxSet + g1Idx + g1Off, g2Ctx = ctxSet + g2Off;
            const CoefBits cb{bitsBoth(e, ctx->sig_coeff_flag[sigCtx]), bitsBoth(e, ctx->greater1_flag[g1Ctx]), bitsBoth(e, ctx->greater2_flag[g2Ctx])};
            long long rdCost = 0, rateSig = 0;
            const int level = adjustLevel(e, a, q, cb, rice, g1Cnt, g2Cnt, sp == lastSp, rdCost, rateSig);
            rec[sp].rdCost = rdCost;
            rec[sp].rateSig = rateSig;
            HvbCoefRec &r = rec[sp]; // everything about a coefficient in one 32-byte record (one sector), indexed by scan position
            r.deltaU = (scaled - (level << qShift)) >> (qShift - 8);
            r.sigDelta = sp != lastSp ? cb.sig.y - cb.sig.x : 0;
            if (level > 0)
            {
                const int now = levelRate(e, level, cb, rice, g1Cnt, g2Cnt);
                r.rateUp = levelRate(e, level + 1, cb, rice, g1Cnt, g2Cnt) - now;
                r.rateDown = levelRate(e, level - 1, cb, rice, g1Cnt, g2Cnt) - now;
                dst[pos] = (int16_t)level;
            }
            else
            {
                r.rateUp = cb.g1.x;
                r.rateDown = 0;
            }
            rdCostTu += rdCost;

            // updateEntropyCodingEngine (Rdoq.cpp:762-803)
            if (level >= baseLevel(g1Cnt, g2Cnt) && level > 3 * (1 << rice)) rice = min(rice + 1, 4);
            if (level >= 1) g1Cnt++;
            if (level > 1)
            {
                g1Idx = 0;
                g2Cnt++;
            }
            else if (g1Idx < 3 && g1Idx > 0 && level)
                g1Idx++;
            if (k == 0 && sp > 0)
            {
                rice = 0;
                g1Cnt = 0;
                g2Cnt = 0;
                ctxSet = (sp == 16 || cIdx != 0) ? 0 : 2;
                if (g1Idx == 0) ctxSet++;
                g1Idx = 1;
            }

            cgRateSig += rateSig;
            if (k == 0) cgRateSigPos0 = rateSig;
            if (level)
            {
                cgCoded = true;
                cgRdCoeff += rdCost - rateSig;
                cgDist0 += d0;
                if (k != 0) nzBeforePos0++;
            }
        }
        if (cgCoded) csbf |= 1ull << cgPos;

        // coefficient-group zeroing (Rdoq.cpp:200-304)
        if (cg)
        {
            const long long zero = e.lam(bitsOf(e, 0, ctx->coded_sub_block_flag[cSig]));
            if (!cgCoded)
            {
                rdCostTu += zero - cgRateSig;
                rateCostCgSig[cg] = zero;
            }
            else if (cg < lastCg)
            {
                if (nzBeforePos0 == 0)
                {
                    rdCostTu -= cgRateSigPos0;
                    cgRateSig -= cgRateSigPos0;
                }
                const long long one = e.lam(bitsOf(e, 1, ctx->coded_sub_block_flag[cSig]));
                const long long allZero = rdCostTu + zero + cgDist0 - cgRdCoeff - cgRateSig;
                rdCostTu += one;
                rateCostCgSig[cg] = one;
                if (allZero < rdCostTu)
                {
                    // the group is dropped: it becomes invisible to stage 2 and to sign hiding, so its records die with it
                    csbf &= ~(1ull << cgPos);
                    rdCostTu = allZero;
                    rateCostCgSig[cg] = zero;
                    for (int k = 0; k < 16; ++k) dst[sc[k]] = 0;
                }
            }
        }
        else
            csbf |= 1ull << cgPos;
    }

    // ---- stage 2: last significant position (Rdoq.cpp:313-397)
    int lastIdx = 0;
    {
        const uint8_t &st = (!isIntra && cIdx == 0) ? ctx->rqt_root_cbf[0] : (cIdx == 0 ? ctx->cbf_luma[1] : ctx->cbf_cbcr[0]);
        long long best = mid.totalDist0 + e.lam(bitsOf(e, 0, st));
        rdCostTu += e.lam(bitsOf(e, 1, st));
        bool found = false;
        for (int cg = lastCg; cg >= 0 && !found; --cg)
        {
            int cgX, cgY;
            scanXY(log2Cg, scanIdx, cg, cgX, cgY);
            const int cgPos = cgY * (1 << log2Cg) + cgX;
            rdCostTu -= rateCostCgSig[cg];
            if (!((csbf >> cgPos) & 1)) continue;
            // four positions at a time: their levels and records are requested together, then consumed in order
            const int kStart = cg == lastCg ? (lastSp & 15) : 15;
#pragma unroll 1
            for (int q4 = kStart >> 2; q4 >= 0 && !found; --q4)
            {
                int p4[4], l4[4];
                longlong2 c4[4]; // {rdCost, rateSig}
#pragma unroll
                for (int j = 0; j < 4; ++j) p4[j] = e.scan[cg * 16 + 4 * q4 + j];
#pragma unroll
                for (int j = 0; j < 4; ++j) l4[j] = dst[p4[j]];
#pragma unroll
                for (int j = 0; j < 4; ++j) c4[j] = *reinterpret_cast<const longlong2 *>(&rec[cg * 16 + 4 * q4 + j].rdCost);
                int a4[4]; // |coefficient|: the "distortion if zero" of a level-1 coefficient that stage 2 walks past
#pragma unroll
                for (int j = 0; j < 4; ++j) a4[j] = abs((int)src[p4[j]]);
#pragma unroll
                for (int j = 3; j >= 0; --j)
                {
                    const int k = 4 * q4 + j, sp = cg * 16 + k;
                    if (k > kStart || found) continue;
                    const int pos = p4[j], level = l4[j];
                    if (level)
                    {
                        const int x = pos & mask, y = pos >> log2;
                        const long long lastCost = scanIdx == 2 ? lastPosCost(e, y, x) : lastPosCost(e, x, y);
                        const long long total = rdCostTu + lastCost - c4[j].y;
                        if (total < best)
                        {
                            lastIdx = sp + 1;
                            best = total;
                        }
                        if (level > 1)
                        {
                            found = true;
                            continue;
                        }
                        rdCostTu -= c4[j].x;
                        rdCostTu += e.dist(a4[j]);
                    }
                    else
                        rdCostTu -= c4[j].y;
                }
            }
        }
    }

    // signs back, uncoded tail to zero (Rdoq.cpp:414-431), and sign-data hiding (:889-1023) of each group while its final
    // levels are in registers.  Only coded groups can hold non-zero levels; from the top, as sign hiding walks them.
    int cbf = 0, lastCGstate = -1;
    for (int cg = lastCg; cg >= 0; --cg)
    {
        int cgX, cgY;
        scanXY(log2Cg, scanIdx, cg, cgX, cgY);
        if (!((csbf >> (cgY * (1 << log2Cg) + cgX)) & 1)) continue;
        const short *sc = e.scan + (cg << 4);
        int lv[16];
        unsigned nzMask = 0, negMask = 0;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
        {
            int p4[4], l4[4], s4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) p4[j] = sc[4 * q4 + j];
#pragma unroll
            for (int j = 0; j < 4; ++j) l4[j] = dst[p4[j]];
#pragma unroll
            for (int j = 0; j < 4; ++j) s4[j] = src[p4[j]];
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const int k = 4 * q4 + j, sp = cg * 16 + k;
                int level = l4[j];
                if (level)
                {
                    if (sp < lastIdx)
                    {
                        cbf |= level;
                        if (s4[j] < 0)
                        {
                            level = -level;
                            dst[p4[j]] = (int16_t)level;
                        }
                    }
                    else
                    {
                        level = 0;
                        dst[p4[j]] = 0;
                    }
                }
                lv[k] = level;
                nzMask |= (unsigned)(level != 0) << k;
                negMask |= (unsigned)(s4[j] < 0) << k;
            }
        }
        if (!sdh || !nzMask) continue; // an empty group changes nothing (and cannot be the first coded one)
        if (lastCGstate == -1) lastCGstate = 1;
        signDataHidingGroup(e, cg, lv, nzMask, negMask, lastCGstate, dst);
        if (lastCGstate == 1) lastCGstate = 0;
    }
    return cbf;
}
