// hvb_rdoq.cuh -- rate-distortion optimised quantisation of one transform block on one warp.
//
// Reference semantics (bit-exact): Rdoq::runQuantisation, turing/Rdoq.cpp:35-450, helpers :452-887,
// signDataHiding :889-1023, constructor arithmetic turing/Rdoq.h:170-188; fixed-point cost algebra
// turing/FixedPoint.h + Cost.h (Cost = Q16 in int64, Lambda = Q16 in int32); scan tables
// turing/ScanOrder.h:32-101; bit-cost table turing/Write.h:413-436 indexed by
// ContextModel::getState() ^ bin = (state >> 1) ^ bin (Rdoq.cpp:26-31, ContextModel.h:58-61).
//
// The level decision of a coefficient depends on the CABAC level-coding state left behind by the
// previous one (greater1/greater2 counters, Rice parameter, context set), so the walk over a block
// is inherently serial (SURVEY.md section 7 "hard parts"); parallelism comes from running many
// blocks at once.  The warp cooperates on what is parallel (scan table, sign restoration, zeroing)
// and lane 0 walks the recurrence.  Per-coefficient state that the reference keeps in the Rdoq
// object (40 KB per block) shrinks to two int64 and four int32 arrays in an L2-resident scratch
// slice; "distortion if zero" is recomputed from the coefficient instead of stored, and the
// reference's zero-initialised members are reproduced by construction (entries above the first
// non-zero level are never written and read as zero) instead of a 40 KB memset per block.
#pragma once
#include "hvb_internal.cuh"

struct HvbRdoqScratch
{
    long long rdCostCoeff[1024];
    long long rateCostSig[1024];
    int rateUp[1024], rateDown[1024], sigDelta[1024], deltaU[1024];
    short scan[1024];
};

__host__ __device__ inline size_t hvbRdoqScratchBytes() { return (sizeof(HvbRdoqScratch) + 255) & ~size_t(255); }

namespace hvb_rdoq {

__device__ __constant__ int32_t kEntropyBits[128] = {
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

__device__ __forceinline__ int bitsOf(int bin, uint8_t state) { return kEntropyBits[(state >> 1) ^ bin]; }

// scan position -> (x, y) inside a (1 << log2)^2 grid, log2 <= 3 (ScanOrder.h:32-101).
// The up-right diagonal order is generated arithmetically: diagonal d holds min(d, n-1) - max(0, d-n+1) + 1 cells.
__device__ __forceinline__ void scanXY(int log2, int scanIdx, int pos, int &x, int &y)
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

__device__ __forceinline__ long long shflXor64(long long v, int m)
{
    const int lo = __shfl_xor_sync(0xffffffffu, (int)(v & 0xffffffffLL), m);
    const int hi = __shfl_xor_sync(0xffffffffu, (int)(v >> 32), m);
    return ((long long)hi << 32) | (unsigned)lo;
}

struct Engine
{
    const hvb_rdoq_ctx *cx;
    HvbRdoqScratch *s;
    const int16_t *src;
    int lambda, distScale, shdFactor;
    int iqScale, iqShift, iqOffset;
    int log2, cIdx, scanIdx;

    __device__ __forceinline__ long long lam(int rate) const { return (long long)lambda * rate; }
    __device__ __forceinline__ long long dist(int err) const
    {
        const int sq = (int)((unsigned)err * (unsigned)err);
        return (long long)sq * distScale;
    }
    __device__ __forceinline__ long long dist0(int sp) const { return dist(abs((int)src[s->scan[sp]])); }
};

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

// Rdoq.cpp:619-673
__device__ inline long long levelRateCost(const Engine &e, int level, int g1Ctx, int g2Ctx, int rice, int g1Cnt, int g2Cnt)
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
            rate += bitsOf(1, e.cx->greater1_flag[g1Ctx]);
            if (g2Cnt < 1) rate += bitsOf(1, e.cx->greater2_flag[g2Ctx]);
        }
    }
    else if (level == 1)
        rate += bitsOf(0, e.cx->greater1_flag[g1Ctx]);
    else if (level == 2)
        rate += bitsOf(1, e.cx->greater1_flag[g1Ctx]) + bitsOf(0, e.cx->greater2_flag[g2Ctx]);
    return e.lam(rate);
}

// Rdoq.cpp:805-870
__device__ inline int levelRate(const Engine &e, int level, int g1Ctx, int g2Ctx, int rice, int g1Cnt, int g2Cnt)
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
            rate += bitsOf(1, e.cx->greater1_flag[g1Ctx]);
            if (g2Cnt < 1) rate += bitsOf(1, e.cx->greater2_flag[g2Ctx]);
        }
    }
    else if (level == 1)
        rate += bitsOf(0, e.cx->greater1_flag[g1Ctx]);
    else if (level == 2)
        rate += bitsOf(1, e.cx->greater1_flag[g1Ctx]) + bitsOf(0, e.cx->greater2_flag[g2Ctx]);
    return rate;
}

// Rdoq.cpp:452-510
__device__ inline int adjustLevel(Engine &e, int sp, int absCoeff, int q, int sigCtx, int g1Ctx, int g2Ctx, int rice, int g1Cnt,
                                  int g2Cnt, bool isLast, long long &rdCost, long long &rateSig)
{
    long long sigCost = 0;
    int best = 0;
    if (!isLast && q < 3)
    {
        rateSig = e.lam(bitsOf(0, e.cx->sig_coeff_flag[sigCtx]));
        rdCost = e.dist(absCoeff) + rateSig;
        if (q == 0) return 0;
    }
    else
        rdCost = 0x7fffffffffffffffLL;
    if (!isLast) sigCost = e.lam(bitsOf(1, e.cx->sig_coeff_flag[sigCtx]));
    const int lowest = q > 1 ? q - 1 : 1;
    for (int level = q; level >= lowest; --level)
    {
        const int recon = hvbClip3(-32768, 32767, (hvbClip3(-32768, 32767, level) * e.iqScale + e.iqOffset) >> e.iqShift);
        const long long c = e.dist(absCoeff - recon) + levelRateCost(e, level, g1Ctx, g2Ctx, rice, g1Cnt, g2Cnt) + sigCost;
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

// Rdoq.cpp:699-740
__device__ inline long long lastPosCost(const Engine &e, int xC, int yC)
{
    const int lx = lastLen(xC), ly = lastLen(yC);
    int rate = 0;
    for (int i = 0; i < lx; ++i) rate += bitsOf(1, e.cx->last_x_prefix[lastPrefixCtx(i, e.cIdx, e.log2)]);
    if (lx < 9) rate += bitsOf(0, e.cx->last_x_prefix[lastPrefixCtx(lx, e.cIdx, e.log2)]);
    for (int i = 0; i < ly; ++i) rate += bitsOf(1, e.cx->last_y_prefix[lastPrefixCtx(i, e.cIdx, e.log2)]);
    if (ly < 9) rate += bitsOf(0, e.cx->last_y_prefix[lastPrefixCtx(ly, e.cIdx, e.log2)]);
    if (lx > 3) rate += 32768 * ((lx - 2) >> 1);
    if (ly > 3) rate += 32768 * ((ly - 2) >> 1);
    return e.lam(rate);
}

// neighbours right (bit 0) / below (bit 1) of coefficient group (xS, yS) in the 64-bit csbf mask
__device__ __forceinline__ void cgNeighbours(unsigned long long csbf, int xS, int yS, int log2, int &right, int &below)
{
    const int wcg = 1 << (log2 - 2);
    right = xS < wcg - 1 ? (int)((csbf >> (yS * wcg + xS + 1)) & 1) : 0;
    below = yS < wcg - 1 ? (int)((csbf >> ((yS + 1) * wcg + xS)) & 1) : 0;
}

// Rdoq.cpp:889-1023
__device__ inline void signDataHiding(const Engine &e, int totalCg, int16_t *dst)
{
    const HvbRdoqScratch &s = *e.s;
    int lastCG = -1;
    for (int cg = totalCg - 1; cg >= 0; --cg)
    {
        const short *sc = s.scan + (cg << 4);
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
                int minCost = 0x7fffffff, minPos = -1, finalChange = 0;
                for (int k = (lastCG == 1 ? lastNZ : 15); k >= 0; --k)
                {
                    const int pos = sc[k];
                    const int level = dst[pos];
                    int cost, change;
                    if (level != 0)
                    {
                        const int up = e.shdFactor * (-s.deltaU[pos]) + s.rateUp[pos];
                        int down = e.shdFactor * s.deltaU[pos] + s.rateDown[pos] - (abs(level) == 1 ? ((1 << 15) + s.sigDelta[pos]) : 0);
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
                        cost = e.shdFactor * (-abs(s.deltaU[pos])) + (1 << 15) + s.rateUp[pos] + s.sigDelta[pos];
                        change = 1;
                        if (k < firstNZ && (e.src[pos] >= 0 ? 0 : 1) != signbit) cost = 0x7fffffff;
                    }
                    if (cost < minCost)
                    {
                        minCost = cost;
                        finalChange = change;
                        minPos = pos;
                    }
                }
                if (minPos >= 0)
                {
                    if (dst[minPos] == 32767 || dst[minPos] == -32768) finalChange = -1;
                    dst[minPos] = (int16_t)(e.src[minPos] >= 0 ? dst[minPos] + finalChange : dst[minPos] - finalChange);
                }
            }
        }
        if (lastCG == 1) lastCG = 0;
    }
}

} // namespace hvb_rdoq

// Runs on a full warp; dst/src are n*n int16 (shared or global); returns the OR of the coded levels on every lane.
//
// Structure (all lanes execute the same control flow; the level-coding state is kept identical on every lane):
//   pre-pass   parallel: first non-zero rounding level in reverse scan order (lastSp) and the two
//              "distortion if zero" sums the reference accumulates on the way there;
//   stage 1    per 4x4 coefficient group, from lastSp's group down: 16 lanes derive everything that does not
//              depend on the level-coding recurrence (position, |c|, rounding level, distortion, sig-flag
//              context and its bit costs).  A group whose levels all round to zero -- the common case -- is then
//              finished with two warp reductions; only groups containing non-zero levels walk their
//              coefficients serially (the recurrence of Rdoq.cpp:762-803);
//   stage 2    last-significant-position search, serial over the coded prefix;
//   finish     parallel sign restoration, optional sign-data hiding.
__device__ inline int hvbRdoqWarp(int16_t *dst, const int16_t *src, const hvb_rdoq_ctx *ctx, int qScale, int qShift, int iqScale,
                                  int log2, int cIdx, int scanIdx, bool isIntra, bool sdh, int bitDepth, HvbRdoqScratch *scratch,
                                  int lane)
{
    using namespace hvb_rdoq;
    const int n = 1 << (2 * log2), totalCg = n >> 4, log2Cg = log2 - 2;
    HvbRdoqScratch &s = *scratch;

    // scan table: coefficient-group order then the 4x4 order inside each group (Rdoq.cpp:399-412)
    for (int sp = lane; sp < n; sp += 32)
    {
        int gx, gy, x, y;
        scanXY(log2Cg, scanIdx, sp >> 4, gx, gy);
        scanXY(2, scanIdx, sp & 15, x, y);
        s.scan[sp] = (short)((((gy << 2) + y) << log2) + (gx << 2) + x);
    }
    __syncwarp();

    Engine e;
    e.cx = ctx;
    e.s = scratch;
    e.src = src;
    e.log2 = log2;
    e.cIdx = cIdx;
    e.scanIdx = scanIdx;
    {
        // Rdoq::Rdoq (Rdoq.h:170-188); FixedPoint<int32,16>::set(double) = int32(d * 65536 + 0.5)
        const double lambda = ctx->lambda;
        e.lambda = (int)(lambda * 65536 + 0.5);
        e.shdFactor = (int)(iqScale * iqScale / lambda / 16 + 0.5);
        const int transformShift = 15 - bitDepth - log2;
        const int distShift = 15 - 2 * transformShift - 2 * (bitDepth - 8);
        e.distScale = (int)((double)(1 << distShift) * 65536 + 0.5);
        e.iqScale = iqScale;
        e.iqShift = 20 - 14 - transformShift;
        e.iqOffset = 1 << (e.iqShift - 1);
    }

    // ---- pre-pass (Rdoq.cpp:104-118, :165-169: integer sums, any order gives the same value)
    int lastSp = -1;
    long long totalDist0 = 0;
    for (int sp = lane; sp < n; sp += 32)
    {
        const int a = abs((int)src[s.scan[sp]]);
        totalDist0 += e.dist(a);
        if (((a * qScale + (1 << (qShift - 1))) >> qShift) > 0) lastSp = sp; // sp ascends: keeps this lane's maximum
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        lastSp = max(lastSp, __shfl_xor_sync(0xffffffffu, lastSp, o));
        totalDist0 += shflXor64(totalDist0, o);
    }
    long long tailDist0 = 0; // what m_rdCostTu holds when the reverse scan reaches lastSp
    for (int sp = lastSp + 1 + lane; sp < n; sp += 32)
    {
        const int pos = s.scan[sp];
        tailDist0 += e.dist(abs((int)src[pos]));
        dst[pos] = 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tailDist0 += shflXor64(tailDist0, o);
    __syncwarp();
    if (lastSp < 0) return 0; // every level rounds to zero (Rdoq.cpp:308-312); dst is all zero

    long long rdCostTu = tailDist0;
    long long rateCostCgSig[64];
    unsigned long long csbf = 0;
    const int lastCg = lastSp >> 4;
    int ctxSet = (lastSp < 16 || cIdx != 0) ? 0 : 2, g1Idx = 1, g1Cnt = 0, g2Cnt = 0, rice = 0;
    const int g1Off = cIdx > 0 ? 16 : 0, g2Off = cIdx > 0 ? 4 : 0;
    for (int i = 0; i < totalCg; ++i) rateCostCgSig[i] = 0;
    const int kk = lane & 15; // lanes 16..31 mirror lanes 0..15

    // ---- stage 1 (Rdoq.cpp:89-305), from the first coded position downwards
    for (int cg = lastCg; cg >= 0; --cg)
    {
        int cgX, cgY, right, below;
        scanXY(log2Cg, scanIdx, cg, cgX, cgY);
        const int cgPos = cgY * (1 << log2Cg) + cgX;
        cgNeighbours(csbf, cgX, cgY, log2, right, below);
        const int prev = right + (below << 1);

        // recurrence-free part of coefficient kk of this group
        const int mySp = cg * 16 + kk, myPos = s.scan[mySp];
        const int myA = abs((int)src[myPos]);
        const int myScaled = myA * qScale;
        const int myQ = mySp <= lastSp ? (myScaled + (1 << (qShift - 1))) >> qShift : 0;
        const int mySc = sigCtxInc(prev, scanIdx, myPos & ((1 << log2) - 1), myPos >> log2, log2, cIdx);
        const int myBits0 = bitsOf(0, ctx->sig_coeff_flag[mySc]), myBits1 = bitsOf(1, ctx->sig_coeff_flag[mySc]);
        const long long myD0 = e.dist(myA);
        const unsigned nzMask = __ballot_sync(0xffffffffu, myQ > 0) & 0xffffu;

        const int cSig = (cIdx == 0 ? 0 : 2) + min(right + below, 1); // coded_sub_block_flag context (neighbours only)

        if (nzMask == 0)
        {
            // Every level of the group rounds to zero (so this is not lastSp's group and all 16 positions are
            // active): adjustLevel takes its q == 0 exit for each (Rdoq.cpp:466-476), the state does not move.
            const long long myRateSig = e.lam(myBits0), myRd = myD0 + myRateSig;
            if (lane < 16)
            {
                s.rdCostCoeff[mySp] = myRd;
                s.rateCostSig[mySp] = myRateSig;
                s.deltaU[myPos] = myScaled >> (qShift - 8);
                s.sigDelta[myPos] = myBits1 - myBits0;
                s.rateUp[myPos] = bitsOf(0, ctx->greater1_flag[4 * ctxSet + g1Idx + g1Off]);
                s.rateDown[myPos] = 0;
                dst[myPos] = 0;
            }
            long long sumRd = myRd, sumSig = myRateSig;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1)
            {
                sumRd += shflXor64(sumRd, o);
                sumSig += shflXor64(sumSig, o);
            }
            rdCostTu += sumRd;
            if (cg > 0)
            {
                // group boundary of updateEntropyCodingEngine (Rdoq.cpp:791-802)
                rice = 0;
                g1Cnt = 0;
                g2Cnt = 0;
                ctxSet = (cg == 1 || cIdx != 0) ? 0 : 2;
                if (g1Idx == 0) ctxSet++;
                g1Idx = 1;
                // uncoded group: pay the cost of its flag, drop the significance costs (Rdoq.cpp:206-216)
                const long long zero = e.lam(bitsOf(0, ctx->coded_sub_block_flag[cSig]));
                rdCostTu += zero - sumSig;
                rateCostCgSig[cg] = zero;
            }
            else
                csbf |= 1ull << cgPos; // the DC group always counts as coded (Rdoq.cpp:299-303)
            continue;
        }

        int nzBeforePos0 = 0;
        long long cgDist0 = 0, cgRateSig = 0, cgRateSigPos0 = 0, cgRdCoeff = 0;
        bool cgCoded = false;
        for (int k = 15; k >= 0; --k)
        {
            const int sp = cg * 16 + k;
            if (sp > lastSp) continue; // accounted for by the pre-pass (their rate terms are zero)
            const int pos = __shfl_sync(0xffffffffu, myPos, k), a = __shfl_sync(0xffffffffu, myA, k);
            const int q = __shfl_sync(0xffffffffu, myQ, k), sc = __shfl_sync(0xffffffffu, mySc, k);
            const int scaled = a * qScale;
            const long long d0 = e.dist(a);
            const int g1Ctx = 4 * ctxSet + g1Idx + g1Off, g2Ctx = ctxSet + g2Off;
            long long rdCost = 0, rateSig = 0;
            const int level = adjustLevel(e, sp, a, q, sc, g1Ctx, g2Ctx, rice, g1Cnt, g2Cnt, sp == lastSp, rdCost, rateSig);
            int up, down;
            if (level > 0)
            {
                const int now = levelRate(e, level, g1Ctx, g2Ctx, rice, g1Cnt, g2Cnt);
                up = levelRate(e, level + 1, g1Ctx, g2Ctx, rice, g1Cnt, g2Cnt) - now;
                down = levelRate(e, level - 1, g1Ctx, g2Ctx, rice, g1Cnt, g2Cnt) - now;
            }
            else
            {
                up = bitsOf(0, ctx->greater1_flag[g1Ctx]);
                down = 0;
            }
            if (lane == 0)
            {
                s.rdCostCoeff[sp] = rdCost;
                s.rateCostSig[sp] = rateSig;
                s.deltaU[pos] = (scaled - (level << qShift)) >> (qShift - 8);
                s.sigDelta[pos] = sp != lastSp ? bitsOf(1, ctx->sig_coeff_flag[sc]) - bitsOf(0, ctx->sig_coeff_flag[sc]) : 0;
                s.rateUp[pos] = up;
                s.rateDown[pos] = down;
                dst[pos] = (int16_t)level;
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
        __syncwarp();
        if (cgCoded) csbf |= 1ull << cgPos;

        // coefficient-group zeroing (Rdoq.cpp:200-304)
        if (cg)
        {
            const long long zero = e.lam(bitsOf(0, ctx->coded_sub_block_flag[cSig]));
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
                const long long one = e.lam(bitsOf(1, ctx->coded_sub_block_flag[cSig]));
                const long long allZero = rdCostTu + zero + cgDist0 - cgRdCoeff - cgRateSig;
                rdCostTu += one;
                rateCostCgSig[cg] = one;
                if (allZero < rdCostTu)
                {
                    csbf &= ~(1ull << cgPos);
                    rdCostTu = allZero;
                    rateCostCgSig[cg] = zero;
                    if (lane < 16 && dst[myPos])
                    {
                        dst[myPos] = 0;
                        s.rdCostCoeff[mySp] = myD0;
                        s.rateCostSig[mySp] = 0;
                    }
                    __syncwarp();
                }
            }
        }
        else
            csbf |= 1ull << cgPos;
    }

    // ---- stage 2: last significant position (Rdoq.cpp:313-397)
    int lastIdx = 0;
    {
        long long best;
        {
            const uint8_t st = (!isIntra && cIdx == 0) ? ctx->rqt_root_cbf[0] : (cIdx == 0 ? ctx->cbf_luma[1] : ctx->cbf_cbcr[0]);
            best = totalDist0 + e.lam(bitsOf(0, st));
            rdCostTu += e.lam(bitsOf(1, st));
        }
        bool found = false;
        for (int cg = lastCg; cg >= 0 && !found; --cg)
        {
            int cgX, cgY;
            scanXY(log2Cg, scanIdx, cg, cgX, cgY);
            const int cgPos = cgY * (1 << log2Cg) + cgX;
            rdCostTu -= rateCostCgSig[cg];
            if (!((csbf >> cgPos) & 1)) continue;
            for (int k = 15; k >= 0; --k)
            {
                const int sp = cg * 16 + k;
                if (sp > lastSp) continue;
                const int pos = s.scan[sp];
                const int level = dst[pos];
                if (level)
                {
                    const int x = pos & ((1 << log2) - 1), y = pos >> log2;
                    const long long lastCost = scanIdx == 2 ? lastPosCost(e, y, x) : lastPosCost(e, x, y);
                    const long long total = rdCostTu + lastCost - s.rateCostSig[sp];
                    if (total < best)
                    {
                        lastIdx = sp + 1;
                        best = total;
                    }
                    if (level > 1)
                    {
                        found = true;
                        break;
                    }
                    rdCostTu -= s.rdCostCoeff[sp];
                    rdCostTu += e.dist0(sp);
                }
                else
                    rdCostTu -= s.rateCostSig[sp];
            }
        }
    }
    __syncwarp();

    // signs back, uncoded tail to zero (Rdoq.cpp:414-431) -- data parallel
    int cbf = 0, absSum = 0;
    for (int sp = lane; sp <= lastSp; sp += 32)
    {
        const int pos = s.scan[sp];
        if (sp < lastIdx)
        {
            const int level = dst[pos];
            absSum += level;
            cbf |= level;
            dst[pos] = (int16_t)(src[pos] < 0 ? -level : level);
        }
        else
            dst[pos] = 0;
    }
    absSum = hvbWarpSum(absSum);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cbf |= __shfl_xor_sync(0xffffffffu, cbf, o);
    __syncwarp();
    if (sdh && absSum >= 2 && lane == 0) signDataHiding(e, totalCg, dst);
    __syncwarp();
    return cbf;
}
