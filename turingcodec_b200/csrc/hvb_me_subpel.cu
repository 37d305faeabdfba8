// hvb_me_subpel.cu -- the sub-pel refinement of the uni-directional motion search as its own kernel behind the integer
// search (hvb_me.cu): one launch refines every PU of the batch.  8-bit and 16-bit samples (template on the sample type;
// the text below describes the 8-bit arithmetic, the 16-bit differences are noted where they occur).
//
// Reference semantics (bit-exact decisions):
//   subPelRefinement / patternSearch / costMv / costDistortionMv   turing/Search.hpp:1965-2060, :2339-2357
//   HavocPredUni 8-tap    havoc/pred_inter.cpp:76-202
//   measureSatd, hadamard turing/Measure.h:96-135, havoc/hadamard.cpp:58-98
//   rateOf(mvd)           turing/Measure.h:177-220
//
// Work decomposition.  A PU is cut into UNITS: 8x8 blocks when both dimensions are multiples of 8 (the
// reference then measures SATD in 8x8 Hadamard tiles), otherwise 8x4 or 4x8 blocks of two 4x4 tiles.  A warp
// owns four consecutive PUs of the task array and walks the concatenation of their units four at a time, so a
// batch of 8x8 PUs fills the warp exactly like one 16x16 PU does.  Per group of four units and per round
// (half-pel: 9 candidates around the integer vector; quarter-pel: 8 around the half-pel winner):
//
//   H pass   64 (unit, support row) jobs, two per lane: the row's 20 bytes are loaded once as aligned words;
//            every output is two IDP.4A (u8 samples x s8 taps).  The half-pel round needs one filtered plane
//            (9 columns: the -1/2 and +1/2 candidates are the same plane one sample apart) plus the integer
//            samples; the quarter-pel round three planes.  Intermediates are the reference's 16-bit `mid`
//            values, stored column-major in shared memory.
//   V pass   (unit, plane, column) jobs: a lane loads its column (4 x LDS.64) and slides the 8-tap window down
//            it, 4 IDP.2A (s16 x s8) per output.  In the half-pel round each distinct (column, row) is computed
//            once (9x9 for the diagonal candidates) and the candidates read it at an offset.
//   SATD     the Hadamard transform of an 8x8 (4x4) tile is a product with the 64x64 (16x16) Kronecker matrix
//            H (x) H, so eight (unit, candidate) tiles at a time are one [H | -H] x [src ; pred] integer GEMM on
//            the tensor cores (IMMA m16n8k32, s8 x u8 -> s32) with fragments read from the 8-bit blocks in shared
//            memory; sum |.| is invariant to the row order of H, so the Sylvester matrix (-1)^popc(m & k) is
//            used and every A fragment register is one of four per-lane constants.
//
// Decisions (cost = rateOf(mvd) + lambda * SATD, candidates in the reference's pattern order, strict <) are taken
// by one lane per PU between the rounds.
#include "hvb_internal.cuh"
#include "hvb_unit.cuh"

namespace {

using namespace hvb_unit;

constexpr int kWarps = 8;
constexpr int kGroup = 4;                 // units per group = PUs per warp chunk
constexpr int kPlaneHalfwords = 9 * kColStride;
constexpr int kUnitMidBytes = 3 * kPlaneHalfwords * 2;      // 1080
constexpr int kPredRow = 12;              // half-pel round: samples per prediction row (9 columns + pad)
constexpr int kPredPlane = 112;           // 9 rows x 12, padded (samples)
constexpr int kUnitPredSamples = 576;     // half-pel: 4 planes x 112; quarter-pel: 8 (bi: 9) candidates x 8 rows x 8
constexpr int kUnitSrcSamples = 64;

template <typename Sample>
struct UnitDesc
{
    const Sample *ref; // sample (0,0) of the unit in the reference plane at the integer part of the round's centre
    int slot;           // PU of the chunk (0..3) the unit belongs to
    int uwuh;           // uw | uh << 8 | valid << 16 | x offset in the PU << 17 | y offset << 24 (padding units of the last group
                        // repeat the last unit and are not summed)
};

template <typename Sample, bool BI>
struct alignas(16) WarpSmem
{
    int16_t mids[kGroup][3 * kPlaneHalfwords];
    Sample preds[kGroup][BI ? kUnitPredSamples : 512]; // the uni search keeps four 256-thread blocks per SM at 8 bit
    Sample src[kGroup][kUnitSrcSamples];
    UnitDesc<Sample> unit[kGroup];
    int satd[kGroup][12];
    // per PU of the chunk
    int cx[kGroup], cy[kGroup];   // centre of the current round, quarter samples
    int units[kGroup];            // units of this PU in the current (round, tile mode) pass
    int geom[kGroup];             // uw | uh << 8 | unitsX << 16
    const Sample *refBase[kGroup]; // sample (x0, y0) of the reference plane
    const Sample *srcBase[kGroup];
    int stride[kGroup];           // reference stride; source stride in srcStride
    int srcStride[kGroup];
    // bi-prediction refinement only: the other list's picture at (x0, y0), its stride and its (clamped) vector
    const Sample *otherBase[BI ? kGroup : 1];
    int otherStride[BI ? kGroup : 1];
    int omvx[BI ? kGroup : 1], omvy[BI ? kGroup : 1];
};

// candidates in the reference's pattern order (Search.hpp:2346, :2352) as grid indices (dy + 1) * 3 + dx + 1
__device__ __constant__ int8_t kHalfOrder[9] = {4, 0, 1, 2, 3, 5, 6, 7, 8};
__device__ __constant__ int8_t kQuarterOrder[8] = {0, 1, 2, 3, 5, 6, 7, 8};

__device__ __forceinline__ long long rateOfMvd(int dx, int dy)
{
    const int rx = 32 - __clz(abs(dx)), ry = 32 - __clz(abs(dy));
    return (long long)(rx + ry + 1) << 17;
}

// ---- H pass ---------------------------------------------------------------------------------------------------
// HALF: plane 0 = integer samples << 6 (columns 0..uw-1), plane 1 = half-pel columns x - 1/2 (x = 0..uw), and the
// candidates that need no vertical filter: P00 (plane 0 of preds) and PB (plane 1 of preds).
// QUARTER: planes v = 0..2 at the horizontal quarter offsets hx + v - 1.
template <typename Sample, bool HALF, bool BI>
__device__ __forceinline__ void hPass(WarpSmem<Sample, BI> &s, const Depth &D, int lane)
{
#pragma unroll 1
    for (int j = 0; j < 2; ++j)
    {
        const int job = lane + 32 * j, u = job >> 4, r = job & 15;
        const UnitDesc<Sample> d = s.unit[u];
        const int uw = d.uwuh & 0xff, uh = (d.uwuh >> 8) & 0xff;
        if (r >= uh + 8) continue;
        const int slot = d.slot;
        const Sample *p = d.ref + (intptr_t)(r - 4) * s.stride[slot] - 4;
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
        const unsigned sh = (unsigned)(a & 3) * 8;
        int16_t *mid = s.mids[u] + r;
        const bool body = r >= 4 && r < 4 + uh;
        Sample *pb = s.preds[u] + kPredPlane + (r - 4) * kPredRow, *p00 = s.preds[u] + (r - 4) * kPredRow;
        if (sizeof(Sample) == 1)
        {
            // 8 bit: 20 bytes of the row, every output two IDP.4A (u8 samples x s8 taps); shift1 = 0
            uint32_t w[6], v[5];
#pragma unroll
            for (int i = 0; i < 6; ++i) w[i] = __ldg(q + i);
#pragma unroll
            for (int i = 0; i < 5; ++i) v[i] = __funnelshift_r(w[i], w[i + 1], sh);
            if (HALF)
            {
                const uint32_t t0 = kTapWords[2][0], t1 = kTapWords[2][1];
#pragma unroll
                for (int c = 0; c < 9; ++c)
                    if (c <= uw)
                    {
                        const int k = c >> 2, sft = (c & 3) * 8;
                        const uint32_t lo = sft ? __funnelshift_r(v[k], v[k + 1], sft) : v[k];
                        const uint32_t hi = sft ? __funnelshift_r(v[k + 1], v[k + 2 < 5 ? k + 2 : 4], sft) : v[k + 1];
                        const int m = dp4aUS(hi, t1, dp4aUS(lo, t0, 0));
                        mid[kPlaneHalfwords + c * kColStride] = (int16_t)m;
                        if (body) pb[c] = (Sample)D.outCopy(m);
                    }
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < uw)
                    {
                        const int smp = (v[(c + 4) >> 2] >> (((c + 4) & 3) * 8)) & 0xff;
                        mid[c * kColStride] = (int16_t)(smp << 6);
                        if (body) p00[c] = (Sample)smp;
                    }
            }
            else
            {
                const int hx = s.cx[slot] & 3;
#pragma unroll 1
                for (int pl = 0; pl < 3; ++pl)
                {
                    const int xq = hx + pl - 1, fx = xq & 3;
                    const uint32_t t0 = kTapWords[fx][0], t1 = kTapWords[fx][1];
                    // column c reads bytes c + (xq >> 2) + 1 .. + 8 of the row
                    uint32_t sv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) sv[i] = xq >= 0 ? __funnelshift_r(v[i], v[i + 1], 8) : v[i];
                    int16_t *mp = mid + pl * kPlaneHalfwords;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c < uw)
                        {
                            const int k = c >> 2, sft = (c & 3) * 8;
                            const uint32_t lo = sft ? __funnelshift_r(sv[k], sv[k + 1], sft) : sv[k];
                            const uint32_t hi = sft ? __funnelshift_r(sv[k + 1], sv[k + 2 < 4 ? k + 2 : 3], sft) : sv[k + 1];
                            mp[c * kColStride] = (int16_t)dp4aUS(hi, t1, dp4aUS(lo, t0, 0));
                        }
                }
            }
        }
        else
        {
            // 16 bit: 18 samples of the row as 9 words (two samples each), every output four IDP.2A (s16 samples x s8
            // taps), then >> shift1
            uint32_t w[10], v[9];
#pragma unroll
            for (int i = 0; i < 10; ++i) w[i] = __ldg(q + i);
#pragma unroll
            for (int i = 0; i < 9; ++i) v[i] = __funnelshift_r(w[i], w[i + 1], sh);
            if (HALF)
            {
                const uint32_t t0 = kTapWords[2][0], t1 = kTapWords[2][1];
#pragma unroll
                for (int c = 0; c < 9; ++c)
                    if (c <= uw)
                    {
                        const int k = c >> 1;
                        const bool odd = c & 1;
                        const uint32_t p0 = odd ? __funnelshift_r(v[k], v[k + 1], 16) : v[k];
                        const uint32_t p1 = odd ? __funnelshift_r(v[k + 1], v[k + 2], 16) : v[k + 1];
                        const uint32_t p2 = odd ? __funnelshift_r(v[k + 2], v[k + 3], 16) : v[k + 2];
                        const uint32_t p3 = odd ? __funnelshift_r(v[k + 3], v[k + 4 < 9 ? k + 4 : 8], 16) : v[k + 3];
                        const int m = tap8(p0, p1, p2, p3, t0, t1, 0) >> D.shift1;
                        mid[kPlaneHalfwords + c * kColStride] = (int16_t)m;
                        if (body) pb[c] = (Sample)D.outCopy(m);
                    }
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c < uw)
                    {
                        const int smp = (v[(c + 4) >> 1] >> (((c + 4) & 1) * 16)) & 0xffff;
                        mid[c * kColStride] = (int16_t)(smp << (6 - D.shift1));
                        if (body) p00[c] = (Sample)smp;
                    }
            }
            else
            {
                const int hx = s.cx[slot] & 3;
#pragma unroll 1
                for (int pl = 0; pl < 3; ++pl)
                {
                    const int xq = hx + pl - 1, fx = xq & 3;
                    const uint32_t t0 = kTapWords[fx][0], t1 = kTapWords[fx][1];
                    // column c reads samples c + (xq >> 2) + 1 .. + 8 of the row
                    uint32_t sv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) sv[i] = xq >= 0 ? __funnelshift_r(v[i], v[i + 1], 16) : v[i];
                    int16_t *mp = mid + pl * kPlaneHalfwords;
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (c < uw)
                        {
                            const int k = c >> 1;
                            const bool odd = c & 1;
                            const uint32_t p0 = odd ? __funnelshift_r(sv[k], sv[k + 1], 16) : sv[k];
                            const uint32_t p1 = odd ? __funnelshift_r(sv[k + 1], sv[k + 2], 16) : sv[k + 1];
                            const uint32_t p2 = odd ? __funnelshift_r(sv[k + 2], sv[k + 3], 16) : sv[k + 2];
                            const uint32_t p3 = odd ? __funnelshift_r(sv[k + 3], sv[k + 4 < 8 ? k + 4 : 7], 16) : sv[k + 3];
                            mp[c * kColStride] = (int16_t)(tap8(p0, p1, p2, p3, t0, t1, 0) >> D.shift1);
                        }
                }
            }
        }
    }
}

// ---- V pass, half-pel round: PC (plane 2 of preds) = vertical half-pel of the integer columns, (uh+1) x uw;
// PD (plane 3) = vertical half-pel of the half-pel columns, (uh+1) x (uw+1).  17 column jobs per unit.
template <typename Sample, bool BI>
__device__ __forceinline__ void vPassHalf(WarpSmem<Sample, BI> &s, const Depth &D, int lane)
{
    const uint32_t t0 = kTapWords[2][0], t1 = kTapWords[2][1];
#pragma unroll 1
    for (int it = 0; it < 3; ++it)
    {
        const int job = lane + 32 * it;
        if (job >= kGroup * 17) break;
        const int u = (job * 3856) >> 16, k = job - u * 17; // / 17
        const int uw = s.unit[u].uwuh & 0xff, uh = (s.unit[u].uwuh >> 8) & 0xff;
        const int pd = k >= 8, c = pd ? k - 8 : k;
        if (c >= uw + pd) continue;
        Column col;
        col.load(s.mids[u] + pd * kPlaneHalfwords + c * kColStride, 0);
        Sample *dst = s.preds[u] + (2 + pd) * kPredPlane + c;
        // rows 0..uh: output r is the half-sample position between picture rows r-1 and r
        dst[0 * kPredRow] = (Sample)col.out<0>(t0, t1, D);
        dst[1 * kPredRow] = (Sample)col.out<1>(t0, t1, D);
        dst[2 * kPredRow] = (Sample)col.out<2>(t0, t1, D);
        dst[3 * kPredRow] = (Sample)col.out<3>(t0, t1, D);
        dst[4 * kPredRow] = (Sample)col.out<4>(t0, t1, D);
        if (uh > 4)
        {
            dst[5 * kPredRow] = (Sample)col.out<5>(t0, t1, D);
            dst[6 * kPredRow] = (Sample)col.out<6>(t0, t1, D);
            dst[7 * kPredRow] = (Sample)col.out<7>(t0, t1, D);
            dst[8 * kPredRow] = (Sample)col.out<8>(t0, t1, D);
        }
    }
}

// ---- V pass, quarter-pel round: candidate q (grid index gi = q + (q >= 4)) = plane gi % 3 at the vertical quarter
// offset hy + gi / 3 - 1; (unit, candidate, column) jobs; preds[u][q][r][c], 8 bytes per row.
template <typename Sample, bool NINE>
__device__ __forceinline__ void vPassQuarter(WarpSmem<Sample, NINE> &s, const Depth &D, int lane)
{
#pragma unroll 1
    for (int it = 0; it < (NINE ? 9 : 8); ++it)
    {
        // uni: 8 candidates (the centre is not re-evaluated); bi: the whole 3x3 grid
        const int job = lane + 32 * it, c = job & 7, uq = job >> 3;
        const int u = NINE ? (uq * 7282) >> 16 : uq >> 3, q = NINE ? uq - 9 * u : uq & 7;
        const UnitDesc<Sample> d = s.unit[u];
        const int uw = d.uwuh & 0xff, uh = (d.uwuh >> 8) & 0xff;
        if (c >= uw) continue;
        const int gi = NINE ? q : q + (q >= 4), pl = gi % 3, yq = (s.cy[d.slot] & 3) + gi / 3 - 1;
        const uint32_t t0 = kTapWords[yq & 3][0], t1 = kTapWords[yq & 3][1];
        Column col;
        col.load(s.mids[u] + pl * kPlaneHalfwords + c * kColStride, (yq >> 2) + 1);
        Sample *dst = s.preds[u] + q * 64 + c;
        dst[0 * 8] = (Sample)col.out<0>(t0, t1, D);
        dst[1 * 8] = (Sample)col.out<1>(t0, t1, D);
        dst[2 * 8] = (Sample)col.out<2>(t0, t1, D);
        dst[3 * 8] = (Sample)col.out<3>(t0, t1, D);
        if (uh > 4)
        {
            dst[4 * 8] = (Sample)col.out<4>(t0, t1, D);
            dst[5 * 8] = (Sample)col.out<5>(t0, t1, D);
            dst[6 * 8] = (Sample)col.out<6>(t0, t1, D);
            dst[7 * 8] = (Sample)col.out<7>(t0, t1, D);
        }
    }
}

// where candidate gi of the half-pel round lives: pointer to its sample (0,0) rounded down to a word, and the
// sample offset (0 or 1) of its columns
template <typename Sample>
__device__ __forceinline__ const Sample *halfCand(const Sample *preds, int gi, int &off)
{
    const int dxi = gi % 3, dyi = gi / 3;
    const int plane = dxi == 1 ? (dyi == 1 ? 0 : 2) : (dyi == 1 ? 1 : 3);
    off = dxi == 2;
    return preds + plane * kPredPlane + (dyi == 2) * kPredRow;
}

// ---- SATD of every (unit, candidate) of the group on the tensor cores -----------------------------------------
template <typename Sample, bool HALF, bool T8, bool NINE>
__device__ __forceinline__ void satdPass(WarpSmem<Sample, NINE> &s, const HadamardA &A, int lane)
{
    constexpr int ncand = (HALF || NINE) ? 9 : 8;
    constexpr bool kDiv9 = ncand == 9;
    constexpr int ncols = kGroup * ncand * (T8 ? 1 : 2);
    constexpr bool k16 = sizeof(Sample) == 2;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
    for (int base = 0; base < ncols; base += 8)
    {
        const int col = min(base + g, ncols - 1);
        const int tile = T8 ? 0 : col & 1, uc = T8 ? col : col >> 1;
        const int u = kDiv9 ? (uc * 7282) >> 16 : uc >> 3; // / 9
        const int cand = uc - u * ncand;
        const int gi = kDiv9 ? cand : cand + (cand >= 4);
        const int uw = s.unit[u].uwuh & 0xff;
        int off = 0, prow = 8;
        const Sample *P;
        if (HALF)
        {
            P = halfCand(s.preds[u], gi, off);
            prow = kPredRow;
        }
        else
            P = s.preds[u] + cand * 64;
        const Sample *S = s.src[u];
        int s0, s1;
        if (T8)
        {
            // k = 32 ks + 4 t + j (+16): tile row 4 ks + (t >> 1) (+2), tile column 4 (t & 1) + j
            const int row = t >> 1, cx = (t & 1) * 4;
            Frag b[4][2];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
            {
                b[ks][0] = loadFrag(S + (ks * 4 + row) * 8 + cx, 0);
                b[ks][1] = loadFrag(S + (ks * 4 + row + 2) * 8 + cx, 0);
                b[ks + 2][0] = loadFrag(P + (ks * 4 + row) * prow + cx, off);
                b[ks + 2][1] = loadFrag(P + (ks * 4 + row + 2) * prow + cx, off);
            }
            s0 = s1 = 0;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
            {
                int acc[4] = {0, 0, 0, 0}, ach[4] = {0, 0, 0, 0};
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                {
                    const int n01 = (((mt >> 1) & ks) ^ (ks >> 1)) & 1; // bit 5 of m & k, and the prediction half of [H | -H]
                    const int n23 = n01 ^ (mt & 1);                     // bit 4 of m & k
                    imma16832(acc, A.e[n01], A.o[n01], A.e[n23], A.o[n23], b[ks][0].lo, b[ks][1].lo);
                    if (k16) imma16832(ach, A.e[n01], A.o[n01], A.e[n23], A.o[n23], b[ks][0].hi, b[ks][1].hi);
                }
                if (k16)
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[r] += ach[r] << 8;
                s0 = __sad(acc[0], 0, __sad(acc[2], 0, (unsigned)s0));
                s1 = __sad(acc[1], 0, __sad(acc[3], 0, (unsigned)s1));
            }
        }
        else
        {
            // two 4x4 tiles per unit: side by side in an 8x4 unit, stacked in a 4x8 unit; K = 16 + 16 is one k-step
            const int tx = uw == 8 ? tile * 4 : 0, ty = uw == 8 ? 0 : tile * 4;
            const Frag b0 = loadFrag(S + (ty + t) * 8 + tx, 0);
            const Frag b1 = loadFrag(P + (ty + t) * prow + tx, off);
            int acc[4] = {0, 0, 0, 0};
            imma16832(acc, A.e[0], A.o[0], A.e[1], A.o[1], b0.lo, b1.lo);
            if (k16)
            {
                int ach[4] = {0, 0, 0, 0};
                imma16832(ach, A.e[0], A.o[0], A.e[1], A.o[1], b0.hi, b1.hi);
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r] += ach[r] << 8;
            }
            s0 = __sad(acc[0], 0, __sad(acc[2], 0, 0u));
            s1 = __sad(acc[1], 0, __sad(acc[3], 0, 0u));
        }
        // The 8 lanes that share t hold partial sums of columns 2t (s0) and 2t+1 (s1).  First exchange so that even
        // g carries column 2t and odd g column 2t+1, then two more butterfly steps: 3 shuffles for both sums.
        int sum = (g & 1) ? s1 : s0;
        sum += __shfl_xor_sync(0xffffffffu, (g & 1) ? s0 : s1, 4);
        sum += __shfl_xor_sync(0xffffffffu, sum, 8);
        sum += __shfl_xor_sync(0xffffffffu, sum, 16);
        if (g < 2)
        {
            // havoc/hadamard.cpp:81-97: 4x4 (s + 1) >> 1, 8x8 (s + 2) >> 2, 16-bit samples >> 2 more; lane (g, t) reports
            // column base + 2t + g
            const int c2 = base + 2 * t + g;
            if (c2 < ncols)
            {
                const int uc2 = T8 ? c2 : c2 >> 1;
                const int u2 = kDiv9 ? (uc2 * 7282) >> 16 : uc2 >> 3;
                const int cand2 = uc2 - u2 * ncand;
                const int slot = s.unit[u2].slot;
                int v = (sum + (T8 ? 2 : 1)) >> (T8 ? 2 : 1);
                if (k16) v >>= 2;
                if ((s.unit[u2].uwuh >> 16) & 1) atomicAdd(&s.satd[slot][kDiv9 ? cand2 : cand2 + (cand2 >= 4)], v);
            }
        }
    }
}

// ---- bi-prediction refinement: the "source" is the ideal block 2 * src - pred(other list) (Search.hpp:1518-1548) -------
// Each unit's 8-tap prediction from the other list's picture at its (fractional) vector is rebuilt here -- one
// horizontal plane, one vertical pass, 1/8 of a quarter-pel round -- and SubtractBi'd into the unit's source slot.
template <typename Sample>
__device__ __forceinline__ void idealPass(WarpSmem<Sample, true> &s, const Depth &D, int lane)
{
    // H: (unit, support row) jobs, plane 0 of the unit's mids
#pragma unroll 1
    for (int j = 0; j < 2; ++j)
    {
        const int job = lane + 32 * j, u = job >> 4, r = job & 15;
        const UnitDesc<Sample> d = s.unit[u];
        const int uw = d.uwuh & 0xff, uh = (d.uwuh >> 8) & 0xff, slot = d.slot;
        if (r >= uh + 8) continue;
        const int ux = (d.uwuh >> 17) & 0x7f, uy = (d.uwuh >> 24) & 0x7f; // the unit's offset in the PU
        const int mvx = s.omvx[slot], mvy = s.omvy[slot], fx = mvx & 3;
        const Sample *p = s.otherBase[slot] + (intptr_t)(uy + (mvy >> 2) + r - 4) * s.otherStride[slot] + ux + (mvx >> 2) - 3;
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
        const unsigned sh = (unsigned)(a & 3) * 8;
        const uint32_t t0 = kTapWords[fx][0], t1 = kTapWords[fx][1];
        int16_t *mp = s.mids[u] + r;
        if (sizeof(Sample) == 1)
        {
            uint32_t w[5], v[4];
#pragma unroll
            for (int i = 0; i < 5; ++i) w[i] = __ldg(q + i);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = __funnelshift_r(w[i], w[i + 1], sh);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c < uw)
                {
                    const int k = c >> 2, sft = (c & 3) * 8;
                    const uint32_t lo = sft ? __funnelshift_r(v[k], v[k + 1], sft) : v[k];
                    const uint32_t hi = sft ? __funnelshift_r(v[k + 1], v[k + 2 < 4 ? k + 2 : 3], sft) : v[k + 1];
                    mp[c * kColStride] = (int16_t)dp4aUS(hi, t1, dp4aUS(lo, t0, 0));
                }
        }
        else
        {
            uint32_t w[9], v[8];
#pragma unroll
            for (int i = 0; i < 9; ++i) w[i] = __ldg(q + i);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __funnelshift_r(w[i], w[i + 1], sh);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (c < uw)
                {
                    const int k = c >> 1;
                    const bool odd = c & 1;
                    const uint32_t p0 = odd ? __funnelshift_r(v[k], v[k + 1], 16) : v[k];
                    const uint32_t p1 = odd ? __funnelshift_r(v[k + 1], v[k + 2], 16) : v[k + 1];
                    const uint32_t p2 = odd ? __funnelshift_r(v[k + 2], v[k + 3], 16) : v[k + 2];
                    const uint32_t p3 = odd ? __funnelshift_r(v[k + 3], v[k + 4 < 8 ? k + 4 : 7], 16) : v[k + 3];
                    mp[c * kColStride] = (int16_t)(tap8(p0, p1, p2, p3, t0, t1, 0) >> D.shift1);
                }
        }
    }
    __syncwarp();
    // V + SubtractBi: (unit, column) jobs; support row 0 is picture row -4, so output row r reads rows r + 1 .. r + 8
    {
        const int u = lane >> 3, c = lane & 7;
        const UnitDesc<Sample> d = s.unit[u];
        const int uw = d.uwuh & 0xff, uh = (d.uwuh >> 8) & 0xff;
        if (c < uw)
        {
            const int fy = s.omvy[d.slot] & 3;
            const uint32_t t0 = kTapWords[fy][0], t1 = kTapWords[fy][1];
            const int idealMax = (1 << (6 + 2 * (int)sizeof(Sample))) - 1; // SubtractBi's bit depth (Search.hpp:1541-1548)
            Column col;
            col.load(s.mids[u] + c * kColStride, 1);
            Sample *dst = s.src[u] + c;
            const int pred[8] = {col.template out<0>(t0, t1, D), col.template out<1>(t0, t1, D), col.template out<2>(t0, t1, D),
                                 col.template out<3>(t0, t1, D), col.template out<4>(t0, t1, D), col.template out<5>(t0, t1, D),
                                 col.template out<6>(t0, t1, D), col.template out<7>(t0, t1, D)};
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r < uh) dst[r * 8] = (Sample)min(max(2 * (int)dst[r * 8] - pred[r], 0), idealMax);
        }
    }
}

// one round over every unit of the chunk's PUs of one tile mode
template <typename Sample, bool HALF, bool T8, bool BI>
__device__ __forceinline__ void roundPass(WarpSmem<Sample, BI> &s, const HadamardA &A, const Depth &D, int lane)
{
    // units of the participating PUs, concatenated
    const int n0 = s.units[0], n1 = n0 + s.units[1], n2 = n1 + s.units[2], total = n2 + s.units[3];
#pragma unroll 1
    for (int base = 0; base < total; base += kGroup)
    {
        if (lane < kGroup)
        {
            const int id = base + lane;
            const int uid = min(id, total - 1);
            const int slot = (uid >= n0) + (uid >= n1) + (uid >= n2);
            const int local = uid - (slot == 0 ? 0 : slot == 1 ? n0 : slot == 2 ? n1 : n2);
            const int geom = s.geom[slot], uw = geom & 0xff, uh = (geom >> 8) & 0xff, unitsX = geom >> 16;
            const int uy = local / unitsX, ux = local - uy * unitsX;
            UnitDesc<Sample> d;
            d.ref = s.refBase[slot] + (intptr_t)(uy * uh + (s.cy[slot] >> 2)) * s.stride[slot] + ux * uw + (s.cx[slot] >> 2);
            d.slot = slot;
            d.uwuh = (geom & 0xffff) | (id < total) << 16 | (ux * uw) << 17 | (uy * uh) << 24;
            s.unit[lane] = d;
        }
        {
            // the source units: 4 x 8 rows of 8 samples
            const int u = lane >> 3, r = lane & 7;
            const int id = min(base + u, total - 1);
            const int slot = (id >= n0) + (id >= n1) + (id >= n2);
            const int local = id - (slot == 0 ? 0 : slot == 1 ? n0 : slot == 2 ? n1 : n2);
            const int geom = s.geom[slot], uw = geom & 0xff, uh = (geom >> 8) & 0xff, unitsX = geom >> 16;
            const int uy = local / unitsX, ux = local - uy * unitsX;
            if (r < uh)
            {
                // x0 is a multiple of 4 samples and the plane rows are 256-byte aligned: word loads (two words per 4 samples at 16 bit)
                const uint32_t *sp = reinterpret_cast<const uint32_t *>(s.srcBase[slot] + (intptr_t)(uy * uh + r) * s.srcStride[slot] + ux * uw);
                uint32_t *dp = reinterpret_cast<uint32_t *>(s.src[u] + r * 8);
                constexpr int kWordsPer4 = sizeof(Sample); // words per 4 samples
#pragma unroll
                for (int k = 0; k < 2 * kWordsPer4; ++k)
                    if (k < kWordsPer4 || uw == 8) dp[k] = __ldg(sp + k);
            }
        }
        __syncwarp();
        if constexpr (BI)
        {
            idealPass(s, D, lane);
            __syncwarp();
        }
        hPass<Sample, HALF, BI>(s, D, lane);
        __syncwarp();
        if (HALF)
            vPassHalf(s, D, lane);
        else
            vPassQuarter<Sample, BI>(s, D, lane);
        __syncwarp();
        satdPass<Sample, HALF, T8, BI>(s, A, lane);
        __syncwarp();
    }
}

// BI = false: uni-directional search (hvb_me_task / hvb_me_result).  BI = true: the fractional rounds of searchMotionBi
// (hvb_me_bi_task / hvb_me_bi_result, Search.hpp:1628-1650): the source is the ideal block, both rounds evaluate the
// whole 3x3 grid with the best cost reset before each, a candidate's rate is that of its cheaper predictor.
template <typename Sample, bool BI>
__global__ void __launch_bounds__(kWarps * 32)
    meSubpelKernel(const HvbPlane *__restrict__ planes, const void *__restrict__ tasksV, int n, void *__restrict__ outV, int bitDepth)
{
    extern __shared__ __align__(16) uint8_t smemSubpel[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpSmem<Sample, BI> &s = reinterpret_cast<WarpSmem<Sample, BI> *>(smemSubpel)[warp];
    const HadamardA A(lane);
    const Depth D(sizeof(Sample) == 1 ? 8 : bitDepth);
    // hvb_me_bi_task shares its first 56 bytes (pictures, block, predictors, rates, lambda, limits) with hvb_me_task
    const hvb_me_task *tasks = static_cast<const hvb_me_task *>(tasksV);
    const hvb_me_bi_task *biTasks = static_cast<const hvb_me_bi_task *>(tasksV);
    hvb_me_result *out = static_cast<hvb_me_result *>(outV);
    hvb_me_bi_result *biOut = static_cast<hvb_me_bi_result *>(outV);
    const int chunks = (n + kGroup - 1) / kGroup, warpsTotal = gridDim.x * kWarps;
    for (int chunk = blockIdx.x * kWarps + warp; chunk < chunks; chunk += warpsTotal)
    {
        // lane k < 4 owns PU 4 chunk + k for the bookkeeping
        const int i = chunk * kGroup + lane;
        const bool mine = lane < kGroup && i < n;
        int w = 8, h = 8, lambda = 0, halfPel = 0, quarterPel = 0, tiles8 = 1, mvpFlag = 0;
        hvb_mv mv{0, 0}, mvd{0, 0};
        long long bestCost = 0;
        if (mine)
        {
            const hvb_me_task &t = tasks[i];
            w = t.w;
            h = t.h;
            lambda = t.lambda;
            if constexpr (BI)
            {
                const hvb_me_bi_task &bt = biTasks[i];
                halfPel = bt.halfPel;
                quarterPel = bt.halfPel && bt.quarterPel;
                mv = biOut[i].mvInteger; // the integer grid left its winner here
                const HvbPlane &op = planes[bt.other_pic * 3];
                s.otherBase[lane] = reinterpret_cast<const Sample *>(op.base) + (intptr_t)bt.y0 * op.stride + bt.x0;
                s.otherStride[lane] = op.stride;
                // only the integer part of the other list's vector is clamped; the fraction is kept (:1518-1534)
                const int ox = min(max(bt.mvOther.x >> 2, (int)bt.limitMin.x), (int)bt.limitMax.x);
                const int oy = min(max(bt.mvOther.y >> 2, (int)bt.limitMin.y), (int)bt.limitMax.y);
                s.omvx[lane] = (ox << 2) | (bt.mvOther.x & 3);
                s.omvy[lane] = (oy << 2) | (bt.mvOther.y & 3);
            }
            else
            {
                halfPel = t.halfPel;
                quarterPel = t.halfPel && t.quarterPel;
                mv = out[i].mv; // the integer search left its winner here
                mvd = out[i].mvd;
            }
            tiles8 = ((w | h) & 7) == 0;
            const HvbPlane &sp = planes[t.src_pic * 3], &rp = planes[t.ref_pic * 3];
            s.refBase[lane] = reinterpret_cast<const Sample *>(rp.base) + (intptr_t)t.y0 * rp.stride + t.x0;
            s.srcBase[lane] = reinterpret_cast<const Sample *>(sp.base) + (intptr_t)t.y0 * sp.stride + t.x0;
            s.stride[lane] = rp.stride;
            s.srcStride[lane] = sp.stride;
            const int uw = tiles8 ? 8 : ((w & 7) ? 4 : 8), uh = tiles8 ? 8 : (uw == 8 ? 4 : 8);
            s.geom[lane] = uw | uh << 8 | (w / uw) << 16;
        }
        else if (lane < kGroup)
        {
            s.geom[lane] = 8 | 8 << 8 | 1 << 16;
            s.refBase[lane] = s.srcBase[lane] = nullptr;
            s.stride[lane] = s.srcStride[lane] = 0;
            if constexpr (BI)
            {
                s.otherBase[lane] = nullptr;
                s.otherStride[lane] = s.omvx[lane] = s.omvy[lane] = 0;
            }
        }
        const int nUnits = mine ? (w * h) >> (tiles8 ? 6 : 5) : 0;
#pragma unroll 1
        for (int round = 0; round < 2; ++round)
        {
            const bool takesPart = round == 0 ? halfPel : quarterPel;
            if (!__any_sync(0xffffffffu, takesPart)) break;
            if (lane < kGroup)
            {
                s.cx[lane] = mv.x;
                s.cy[lane] = mv.y;
#pragma unroll
                for (int c = 0; c < 9; ++c) s.satd[lane][c] = 0;
            }
#pragma unroll 1
            for (int mode = 0; mode < 2; ++mode) // 8x8-tile PUs, then 4x4-tile PUs
            {
                const bool now = takesPart && (mode == 0 ? tiles8 : !tiles8);
                if (!__any_sync(0xffffffffu, now)) continue;
                if (lane < kGroup) s.units[lane] = now ? nUnits : 0;
                __syncwarp();
                if (round == 0)
                {
                    if (mode == 0)
                        roundPass<Sample, true, true, BI>(s, A, D, lane);
                    else
                        roundPass<Sample, true, false, BI>(s, A, D, lane);
                }
                else
                {
                    if (mode == 0)
                        roundPass<Sample, false, true, BI>(s, A, D, lane);
                    else
                        roundPass<Sample, false, false, BI>(s, A, D, lane);
                }
            }
            __syncwarp();
            if (mine && takesPart)
            {
                const int step = round == 0 ? 2 : 1;
                if (BI)
                {
                    // 3x3 grid in raster order, best cost reset (Search.hpp:1628-1650); MvCandidate picks the cheaper predictor
                    const hvb_me_task &t = tasks[i];
                    const hvb_mv origin = mv;
                    bestCost = 0x7fffffffffffffffLL;
                    for (int gi = 0; gi < 9; ++gi)
                    {
                        const int cx = origin.x + (gi % 3 - 1) * step, cy = origin.y + (gi / 3 - 1) * step;
                        int dx = (int16_t)(cx - t.mvp[0].x), dy = (int16_t)(cy - t.mvp[0].y), flag = 0;
                        long long c = rateOfMvd(dx, dy) + t.rateMvpFlag[0];
                        const int dx1 = (int16_t)(cx - t.mvp[1].x), dy1 = (int16_t)(cy - t.mvp[1].y);
                        const long long c1 = rateOfMvd(dx1, dy1) + t.rateMvpFlag[1];
                        if (c1 < c)
                        {
                            c = c1;
                            dx = dx1;
                            dy = dy1;
                            flag = 1;
                        }
                        c += (long long)lambda * s.satd[lane][gi];
                        if (c < bestCost)
                        {
                            bestCost = c;
                            mv = hvb_mv{(int16_t)cx, (int16_t)cy};
                            mvd = hvb_mv{(int16_t)dx, (int16_t)dy};
                            mvpFlag = flag;
                        }
                    }
                }
                else
                {
                    // patternSearch with maxIterations = 1 (Search.hpp:2011-2060): costMv = rateOf(mvd) + lambda * SATD
                    const int ncand = round == 0 ? 9 : 8;
                    int best = -1;
                    for (int k = 0; k < ncand; ++k)
                    {
                        const int gi = round == 0 ? kHalfOrder[k] : kQuarterOrder[k];
                        const int dx = (gi % 3 - 1) * step, dy = (gi / 3 - 1) * step;
                        const long long c = rateOfMvd((int16_t)(mvd.x + dx), (int16_t)(mvd.y + dy)) + (long long)lambda * s.satd[lane][gi];
                        if (round == 0 && k == 0)
                            bestCost = c; // the origin (tryOrigin)
                        else if (c < bestCost)
                        {
                            best = gi;
                            bestCost = c;
                        }
                    }
                    if (best >= 0)
                    {
                        const int dx = (best % 3 - 1) * step, dy = (best / 3 - 1) * step;
                        mv.x = (int16_t)(mv.x + dx);
                        mv.y = (int16_t)(mv.y + dy);
                        mvd.x = (int16_t)(mvd.x + dx);
                        mvd.y = (int16_t)(mvd.y + dy);
                    }
                }
            }
            __syncwarp();
        }
        if (mine && halfPel)
        {
            if (BI)
            {
                biOut[i].mv = mv;
                biOut[i].mvd = mvd;
                biOut[i].mvpFlag = mvpFlag;
                biOut[i].cost = bestCost;
            }
            else
            {
                out[i].mv = mv;
                out[i].mvd = mvd;
                out[i].subpelCost = bestCost;
            }
        }
        __syncwarp();
    }
}

} // namespace

// called by hvb_me_search_batch / hvb_me_bi_search_batch (hvb_me.cu) after the integer stage, on the same stream
template <typename Sample, bool BI>
static int launchMeSubpel(hvb_context *ctx, const void *dTasks, int n, void *dOut)
{
    const int chunks = (n + kGroup - 1) / kGroup;
    int blocks = (chunks + kWarps - 1) / kWarps;
    const int smem = kWarps * (int)sizeof(WarpSmem<Sample, BI>);
    static_assert(sizeof(WarpSmem<Sample, BI>) % 16 == 0, "per-warp shared slices must stay 16-byte aligned");
    cudaFuncSetAttribute(meSubpelKernel<Sample, BI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int perSm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, meSubpelKernel<Sample, BI>, kWarps * 32, smem);
    const int cap = ctx->smCount * (perSm > 0 ? perSm : 1);
    if (blocks > cap) blocks = cap;
    meSubpelKernel<Sample, BI><<<blocks, kWarps * 32, smem, ctx->stream>>>(ctx->dPlanes, dTasks, n, dOut, ctx->bitDepth);
    HVB_LAUNCH_CHECK(ctx, "meSubpelKernel");
    return HVB_OK;
}

int hvbLaunchMeSubpel(hvb_context *ctx, const hvb_me_task *dTasks, int n, hvb_me_result *dOut)
{
    return ctx->bps == 1 ? launchMeSubpel<uint8_t, false>(ctx, dTasks, n, dOut) : launchMeSubpel<uint16_t, false>(ctx, dTasks, n, dOut);
}

int hvbLaunchMeBiSubpel(hvb_context *ctx, const hvb_me_bi_task *dTasks, int n, hvb_me_bi_result *dOut)
{
    return ctx->bps == 1 ? launchMeSubpel<uint8_t, true>(ctx, dTasks, n, dOut) : launchMeSubpel<uint16_t, true>(ctx, dTasks, n, dOut);
}
