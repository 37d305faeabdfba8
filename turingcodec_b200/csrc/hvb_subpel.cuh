// hvb_subpel.cuh -- sub-pel candidate evaluation for 8-bit pictures: interpolation planes shared between the
// candidates of a refinement round + Hadamard SATD on the integer tensor cores.
//
// Reference semantics (bit-exact):
//   costDistortionMv / patternSearch   turing/Search.hpp:1965-2060   (predict at a quarter-pel vector, measureSatd)
//   HavocPredUni 8-tap                 havoc/pred_inter.cpp:76-202   (mid = sum cx*s >> shift1; out = clip((sum cy*mid + 2^(5+shift3)) >> (6+shift3)))
//   measureSatd tiling, hadamard_satd  turing/Measure.h:96-135, havoc/hadamard.cpp:58-98
//
// The reference evaluates the 9 (then 8) candidates of a round one after the other, each with its own
// 8-tap horizontal + vertical pass and its own Hadamard.  Here a warp walks the PU in sub-blocks of at most
// 16x16 samples and, per sub-block,
//   H pass   one lane per row of the (h+8)-row support: the row is loaded once as aligned words and every
//            output column is two IDP.4A (u8 samples x s8 taps); intermediates (exactly the reference's `mid`
//            values, 16 bit) go to shared memory column-major.  Candidates that share a horizontal phase
//            share the plane: a half-pel round needs 2 planes (the +-2 columns are one plane shifted by a
//            sample), a quarter-pel round 3;
//   V pass   two vertically adjacent outputs per lane from 5 words of a column, 4 IDP.2A each (s16 x s8);
//            a half-pel round computes each distinct (column, row) once and stores it to every candidate it
//            belongs to;
//   SATD     sum |H (s - p) H^T| over a T x T tile is sum |(H (x) H) vec(s) - (H (x) H) vec(p)|: the Kronecker
//            Hadamard matrix is a 64x64 (16x16 for 4x4 tiles) +-1 matrix, so 8 tiles at a time are one
//            [H | -H] x [s ; p] integer GEMM -- IMMA m16n8k32 s8 x u8 -> s32 -- with the operand fragments read
//            straight from the 8-bit source and prediction blocks in shared memory.  The sum of absolute
//            values is invariant to the row order of H, so the natural-order (Sylvester) matrix is used,
//            whose entries are (-1)^popc(m & k): every A fragment register is one of four per-lane constants.
#pragma once
#include "hvb_internal.cuh"

namespace subpel {

constexpr int kMidStride = 26;                      // halfwords per plane column: 24 rows + 2 (13 words, odd)
constexpr int kMidCols = 17;
constexpr int kMidPlane = kMidCols * kMidStride;    // halfwords
constexpr int kPredStride = 256;                    // bytes per candidate prediction (16 x 16)
constexpr int kMaxCand = 9;
constexpr int kMidBytes = (3 * kMidPlane * 2 + 4 * 2 + 15) / 16 * 16; // + slack: the V pass may read one word past a column
constexpr int kScratchBytes = kMidBytes + kMaxCand * kPredStride;

// 8-tap luma filters (havoc/pred_inter.cpp:39-69) packed as s8x4 words, taps 0..3 and 4..7
__device__ __constant__ uint32_t kTapWords[4][2] = {{0x40000000u, 0x00000000u},
                                                    {0x3af604ffu, 0x0001fb11u},
                                                    {0x28f504ffu, 0xff04f528u},
                                                    {0x11fb0100u, 0xff04f63au}};

__device__ __forceinline__ int dp4aUS(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// Horizontal pass.  Lane r < nrows filters row r of the support.  `p0` points at the first byte the leftmost
// output column reads (sample x - 3 of column 0, support row 0); column c reads bytes c .. c+7 of its row.
// plane[c * kMidStride + r] = sum_k taps[fx][k] * p[c + k]   (shift1 = 0 at 8 bit)
__device__ __forceinline__ void hPass(const uint8_t *p0, int stride, int nrows, int ncols, int fx, int16_t *plane, int lane)
{
    if (lane < nrows)
    {
        const uint8_t *p = p0 + (intptr_t)lane * stride;
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
        const unsigned sh = (unsigned)(a & 3) * 8;
        uint32_t w[7], v[6];
#pragma unroll
        for (int i = 0; i < 7; ++i) w[i] = __ldg(q + i);
#pragma unroll
        for (int i = 0; i < 6; ++i) v[i] = __funnelshift_r(w[i], w[i + 1], sh);
        const uint32_t t0 = kTapWords[fx][0], t1 = kTapWords[fx][1];
        int16_t *dst = plane + lane;
#pragma unroll
        for (int c = 0; c < kMidCols; ++c)
            if (c < ncols)
            {
                const int k = c >> 2, s = (c & 3) * 8;
                const uint32_t lo = s ? __funnelshift_r(v[k], v[k + 1], s) : v[k];
                const uint32_t hi = s ? __funnelshift_r(v[k + 1], v[(k + 2) < 6 ? (k + 2) : 5], s) : v[k + 1];
                dst[c * kMidStride] = (int16_t)dp4aUS(hi, t1, dp4aUS(lo, t0, 0));
            }
    }
}

__device__ __forceinline__ int vFilter(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t t0, uint32_t t1)
{
    int acc = 1 << 11; // 2^(5 + shift3), shift3 = 6 at 8 bit
    acc = __dp2a_lo((int)p0, (int)t0, acc);
    acc = __dp2a_hi((int)p1, (int)t0, acc);
    acc = __dp2a_lo((int)p2, (int)t1, acc);
    acc = __dp2a_hi((int)p3, (int)t1, acc);
    return min(max(acc >> 12, 0), 255);
}

// Vertical pass over `ncols` x `nrows` outputs of one plane at the vertical quarter-pel offset yq (support row 0
// is picture row -4 of the sub-block, so output row r reads support rows r + (yq >> 2) + 1 .. + 8).
// Output (r, c) is stored to up to four candidate blocks (sbw bytes per row): dA at (r, c); dB at (r, c-1);
// dC at (r-1, c); dD at (r-1, c-1) -- the four half-pel candidates a (column, row) of a shifted plane serves.
// A negative index disables that destination.
__device__ __forceinline__ void vPass(const int16_t *plane, int ncols, int nrows, int yq, int sbw, int sbh, uint8_t *preds,
                                      int dA, int dB, int dC, int dD, int lane)
{
    const int fy = yq & 3, odd = ((yq >> 2) + 1) & 1; // first support row of output row 0 is 0 or 1
    const uint32_t t0 = kTapWords[fy][0], t1 = kTapWords[fy][1];
    const int rpairs = (nrows + 1) >> 1, jobs = ncols * rpairs;
    for (int job = lane; job < jobs; job += 32)
    {
        const int c = job / rpairs, rp = job - c * rpairs, r = 2 * rp;
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(plane + c * kMidStride) + rp;
        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3], w4 = wp[4];
        const uint32_t x0 = __funnelshift_r(w0, w1, 16), x1 = __funnelshift_r(w1, w2, 16), x2 = __funnelshift_r(w2, w3, 16),
                       x3 = __funnelshift_r(w3, w4, 16);
        int o0, o1;
        if (odd)
        {
            o0 = vFilter(x0, x1, x2, x3, t0, t1);
            o1 = vFilter(w1, w2, w3, w4, t0, t1);
        }
        else
        {
            o0 = vFilter(w0, w1, w2, w3, t0, t1);
            o1 = vFilter(x0, x1, x2, x3, t0, t1);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e)
        {
            const int rr = r + e, o = e ? o1 : o0;
            if (rr >= nrows) break;
            if (dA >= 0 && c < sbw && rr < sbh) preds[dA * kPredStride + rr * sbw + c] = (uint8_t)o;
            if (dB >= 0 && c >= 1 && rr < sbh) preds[dB * kPredStride + rr * sbw + c - 1] = (uint8_t)o;
            if (dC >= 0 && c < sbw && rr >= 1) preds[dC * kPredStride + (rr - 1) * sbw + c] = (uint8_t)o;
            if (dD >= 0 && c >= 1 && rr >= 1) preds[dD * kPredStride + (rr - 1) * sbw + c - 1] = (uint8_t)o;
        }
    }
}

__device__ __forceinline__ void imma16832(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// The four A-fragment constants of a lane.  Fragment register (m-tile mt, k-step ks, reg) holds
// A[m][k .. k+3] with m = 16 mt + (lane >> 2) + 8 (reg & 1) and k = 32 ks + 4 (lane & 3) + 16 (reg >> 1);
// the entry is (-1)^popc(m & k): its byte pattern depends on m & 3, its overall sign on the remaining bits,
// of which only two products involve the lane -- the rest are compile-time in the unrolled MMA loops.
struct HadamardA
{
    uint32_t e[2], o[2]; // [negated]: registers with reg & 1 == 0 / == 1
    __device__ __forceinline__ explicit HadamardA(int lane)
    {
        const int g = lane >> 2, t = lane & 3;
        // bytes j = 0..3: (-1)^popc((g & 3) & j)
        const uint32_t pat = (g & 2) ? ((g & 1) ? 0x01ffff01u : 0xffff0101u) : ((g & 1) ? 0xff01ff01u : 0x01010101u);
        const int q0 = (g >> 2) & t & 1, q1 = q0 ^ (t >> 1);
        e[0] = q0 ? pat ^ 0xfefefefeu : pat;
        e[1] = e[0] ^ 0xfefefefeu;
        o[0] = q1 ? pat ^ 0xfefefefeu : pat;
        o[1] = o[0] ^ 0xfefefefeu;
    }
};

// SATD of `ncand` candidate predictions (u8, `predStride` bytes apart, sbw bytes per row) against the source
// sub-block (`src`, srcStride bytes per row), T x T Hadamard tiles; adds each candidate's normalised tile sums
// to sSatd[cand].  8 (candidate, tile) columns per IMMA group.
template <int LOG2T>
__device__ __forceinline__ void satdMma(const uint8_t *src, int srcStride, const uint8_t *preds, int predStride, int sbw, int sbh,
                                        int ncand, int *sSatd, const HadamardA &A, int lane)
{
    const int tilesX = sbw >> LOG2T, tiles = tilesX * (sbh >> LOG2T), ncols = ncand * tiles;
    const int g = lane >> 2, t = lane & 3;
    for (int base = 0; base < ncols; base += 8)
    {
        const int col = min(base + g, ncols - 1);
        const int cand = col / tiles, tile = col - cand * tiles;
        const int ty = tile / tilesX, tx = tile - ty * tilesX;
        const uint8_t *S = src + ((ty * srcStride + tx) << LOG2T);
        const uint8_t *P = preds + cand * predStride + ((ty * sbw + tx) << LOG2T);
        int s0, s1;
        if (LOG2T == 3)
        {
            // k = 32 ks + 4 t + j (+16): tile row 4 ks + (t >> 1) (+2), tile column 4 (t & 1) + j
            const int row = t >> 1, cx = (t & 1) * 4;
            uint32_t b[4][2];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
            {
                b[ks][0] = *reinterpret_cast<const uint32_t *>(S + (ks * 4 + row) * srcStride + cx);
                b[ks][1] = *reinterpret_cast<const uint32_t *>(S + (ks * 4 + row + 2) * srcStride + cx);
                b[ks + 2][0] = *reinterpret_cast<const uint32_t *>(P + (ks * 4 + row) * sbw + cx);
                b[ks + 2][1] = *reinterpret_cast<const uint32_t *>(P + (ks * 4 + row + 2) * sbw + cx);
            }
            s0 = s1 = 0;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
            {
                int acc[4] = {0, 0, 0, 0};
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                {
                    // compile-time sign: bit 5 of m & k, and the prediction half of [H | -H]
                    const int n01 = (((mt >> 1) & ks) ^ (ks >> 1)) & 1;
                    const int n23 = n01 ^ (mt & 1); // bit 4 of m & k
                    imma16832(acc, A.e[n01], A.o[n01], A.e[n23], A.o[n23], b[ks][0], b[ks][1]);
                }
                s0 = __sad(acc[0], 0, __sad(acc[2], 0, (unsigned)s0));
                s1 = __sad(acc[1], 0, __sad(acc[3], 0, (unsigned)s1));
            }
        }
        else
        {
            // 4x4 tiles: K = 16 source + 16 prediction samples is exactly one k-step; lane t holds tile row t
            const uint32_t b0 = *reinterpret_cast<const uint32_t *>(S + t * srcStride);
            const uint32_t b1 = *reinterpret_cast<const uint32_t *>(P + t * sbw);
            int acc[4] = {0, 0, 0, 0};
            imma16832(acc, A.e[0], A.o[0], A.e[1], A.o[1], b0, b1);
            s0 = __sad(acc[0], 0, __sad(acc[2], 0, 0u));
            s1 = __sad(acc[1], 0, __sad(acc[3], 0, 0u));
        }
        // columns 2t and 2t+1 of this group: sum over the 8 lanes that share t
#pragma unroll
        for (int o = 4; o < 32; o <<= 1)
        {
            s0 += __shfl_xor_sync(0xffffffffu, s0, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        }
        if (g == 0)
        {
            // havoc/hadamard.cpp:81-97: 4x4 (s + 1) >> 1, 8x8 (s + 2) >> 2
            const int c0 = base + 2 * t;
            if (c0 < ncols) atomicAdd(&sSatd[c0 / tiles], (s0 + (1 << (LOG2T - 2))) >> (LOG2T - 1));
            if (c0 + 1 < ncols) atomicAdd(&sSatd[(c0 + 1) / tiles], (s1 + (1 << (LOG2T - 2))) >> (LOG2T - 1));
        }
    }
}

// One refinement round: SATD of the prediction at centre + step * (dx, dy), (dx, dy) in {-1,0,1}^2, for the
// candidates named by candOf[(dy + 1) * 3 + dx + 1] (a 9-entry table in constant memory; >= 0: index into
// sSatd; < 0: not evaluated).
//   src      the PU's source block in shared memory, w bytes per row
//   ref      sample (0, 0) of the PU in the reference plane; (cx, cy) the centre in quarter samples
//   step     2: half-pel round around an integer centre (cx, cy multiples of 4); 1: quarter-pel round
// sSatd[0 .. ncand) must be zero on entry (and visible to the warp).
__device__ __noinline__ void evalRound(const uint8_t *src, int w, int h, const uint8_t *ref, int refStride, int cx, int cy, int step,
                                       const int8_t *candOf, int ncand, uint8_t *scratch, int *sSatd, int lane)
{
    int16_t *mids = reinterpret_cast<int16_t *>(scratch);
    uint8_t *preds = scratch + kMidBytes;
    const HadamardA A(lane);
    const bool tiles8 = ((w | h) & 7) == 0;
    for (int by = 0; by < h; by += 16)
        for (int bx = 0; bx < w; bx += 16)
        {
            const int sbw = min(16, w - bx), sbh = min(16, h - by);
            // support row 0 = picture row -4 of the sub-block at the centre's integer position
            const uint8_t *R = ref + (intptr_t)(by + (cy >> 2) - 4) * refStride + bx + (cx >> 2);
            if (step == 2)
            {
                // plane 0: integer columns; plane 1: the half-pel columns x - 1/2, x = 0 .. sbw
                hPass(R - 3, refStride, sbh + 8, sbw, 0, mids, lane);
                hPass(R - 4, refStride, sbh + 8, sbw + 1, 2, mids + kMidPlane, lane);
                __syncwarp();
                vPass(mids, sbw, sbh, 0, sbw, sbh, preds, candOf[4], -1, -1, -1, lane);
                vPass(mids, sbw, sbh + 1, -2, sbw, sbh, preds, candOf[1], -1, candOf[7], -1, lane);
                vPass(mids + kMidPlane, sbw + 1, sbh, 0, sbw, sbh, preds, candOf[3], candOf[5], -1, -1, lane);
                vPass(mids + kMidPlane, sbw + 1, sbh + 1, -2, sbw, sbh, preds, candOf[0], candOf[2], candOf[6], candOf[8], lane);
            }
            else
            {
                const int hx = cx & 3, hy = cy & 3;
#pragma unroll 1
                for (int v = 0; v < 3; ++v)
                {
                    const int xq = hx + v - 1;
                    hPass(R + (xq >> 2) - 3, refStride, sbh + 8, sbw, xq & 3, mids + v * kMidPlane, lane);
                }
                __syncwarp();
#pragma unroll 1
                for (int i = 0; i < 9; ++i)
                    if (candOf[i] >= 0) vPass(mids + (i % 3) * kMidPlane, sbw, sbh, hy + i / 3 - 1, sbw, sbh, preds, candOf[i], -1, -1, -1, lane);
            }
            __syncwarp();
            const uint8_t *S = src + by * w + bx;
            if (tiles8)
                satdMma<3>(S, w, preds, kPredStride, sbw, sbh, ncand, sSatd, A, lane);
            else
                satdMma<2>(S, w, preds, kPredStride, sbw, sbh, ncand, sSatd, A, lane);
            __syncwarp();
        }
}

} // namespace subpel
