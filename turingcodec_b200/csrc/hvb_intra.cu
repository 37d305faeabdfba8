// hvb_intra.cu -- batched HEVC intra prediction and the 35-mode SATD sweep.
//
// Reference semantics (bit-exact):
//   havoc::intra::Function      havoc/pred_intra.cpp:20282-20401 (planar / DC / angular, edge filters)
//   neighbour addressing        havoc/pred_intra.cpp:43-51: p(x,y) = neighbours[x - y - 1]
//   filterFlag                  turing/Dsp.h:57-70 (which modes use filtered reference samples)
//   reference-sample filter     turing/IntraReferenceSamples.h:373-419 ([1 2 1] and strong bi-linear)
//   35-mode sweep + SATD        turing/Reconstruct.cpp:630-701 (PredictIntraLumaBlock), Search.hpp:39-267
//
// The reference materialises every prediction into a stack buffer and then runs SATD on it, one mode
// at a time.  Here a predicted sample is a pure function of (mode, x, y) and the 4n+1 neighbours: the sweep
// generates the 35 predictions of a tile row-wise into shared memory and takes their SATDs on the integer
// tensor cores (see "the sweep" below).  Per partition that is (4n+1)B + n*n*B bytes in and 35 * 4 bytes out.
#include "hvb_internal.cuh"

namespace {

constexpr int kWarps = 4;
constexpr int kNbMax = 4 * 32 + 1;

__device__ __constant__ int8_t kAngle[35] = {0,   0,   32,  26,  21,  17, 13, 9,  5,  2,  0,  -2, -5, -9, -13, -17, -21, -26,
                                             -32, -26, -21, -17, -13, -9, -5, -2, 0,  2,  5,  9,  13, 17, 21,  26,  32};
__device__ __constant__ int16_t kInvAngle[35] = {0,    0,    0,    0,    0,    0,    0,    0,    0,     0,     0,    -4096,
                                                 -1638, -910, -630, -482, -390, -315, -256, -315, -390,  -482,  -630, -910,
                                                 -1638, -4096, 0,   0,    0,    0,    0,    0,    0,     0,     0};

// Reference samples of one partition: A[c + 1 + x] = p(x,-1), A[c] = p(-1,-1), A[c - 1 - y] = p(-1,y).
struct Neighbours
{
    const int16_t *A;
    int c; // = 2n
    __device__ __forceinline__ int top(int x) const { return A[c + 1 + x]; }
    __device__ __forceinline__ int left(int y) const { return A[c - 1 - y]; }
    __device__ __forceinline__ int corner() const { return A[c]; }
};

// Projected 1-D reference of the angular modes (pred_intra.cpp:20330-20344 / :20368-20382),
// evaluated on the fly: index i >= 0 walks the main side, i < 0 the inverse-angle projection.
__device__ __forceinline__ int angularRef(const Neighbours &nb, int i, bool vertical, int inv)
{
    const int s = i >= 0 ? i : -((i * inv + 128) >> 8);
    return vertical ? nb.A[nb.c + s] : nb.A[nb.c - s];
}

// `angle` / `inv` are kAngle[mode] / kInvAngle[mode], fetched once per job by the caller: with a different mode on
// every lane the constant-bank read serialises, and the ncu capture had it at a third of all stall samples when it
// sat inside the per-sample code.
__device__ __forceinline__ int intraSample(const Neighbours &nb, int mode, int angle, int inv, int x, int y, int log2n, int dcVal, bool edge,
                                           int maxv)
{
    const int n = 1 << log2n;
    if (mode == 0)
        return ((n - 1 - x) * nb.left(y) + (x + 1) * nb.top(n) + (n - 1 - y) * nb.top(x) + (y + 1) * nb.left(n) + n) >> (log2n + 1);
    if (mode == 1)
    {
        if (edge)
        {
            if (x == 0 && y == 0) return (nb.left(0) + 2 * dcVal + nb.top(0) + 2) >> 2;
            if (y == 0) return (nb.top(x) + 3 * dcVal + 2) >> 2;
            if (x == 0) return (nb.left(y) + 3 * dcVal + 2) >> 2;
        }
        return dcVal;
    }
    const bool vertical = mode >= 18;
    if (edge && mode == 26 && x == 0) return hvbClip3(0, maxv, nb.top(0) + ((nb.left(y) - nb.corner()) >> 1));
    if (edge && mode == 10 && y == 0) return hvbClip3(0, maxv, nb.left(0) + ((nb.top(x) - nb.corner()) >> 1));
    const int major = vertical ? y : x, minor = vertical ? x : y;
    const int t = (major + 1) * angle;
    const int idx = t >> 5, fact = t & 31;
    const int r0 = angularRef(nb, minor + idx + 1, vertical, inv);
    if (!fact) return r0;
    const int r1 = angularRef(nb, minor + idx + 2, vertical, inv);
    return ((32 - fact) * r0 + fact * r1 + 16) >> 5;
}

__device__ __forceinline__ int dcValue(const Neighbours &nb, int log2n, int lane)
{
    const int n = 1 << log2n;
    int acc = 0;
    for (int i = lane; i < n; i += 32) acc += nb.top(i) + nb.left(i);
    return (hvbWarpSum(acc) + n) >> (log2n + 1);
}

// turing/Dsp.h:57-70 as the rule it tabulates: filter when the mode is further from pure
// horizontal/vertical than the size-dependent threshold (8:7, 16:1, 32:0); planar always for n >= 8.
__device__ __forceinline__ bool filterFlag(int cIdx, int mode, int n)
{
    if (cIdx != 0 || mode == 1 || n == 4) return false;
    if (mode == 0) return true;
    const int dist = min(abs(mode - 26), abs(mode - 10));
    const int thres = n == 8 ? 7 : (n == 16 ? 1 : 0);
    return dist > thres;
}

// turing/IntraReferenceSamples.h:373-419
__device__ void filterNeighbours(int16_t *F, const int16_t *U, int n, int bitDepth, bool strongEnabled, int lane)
{
    const int c = 2 * n;
    const Neighbours p{U, c};
    bool strong = false;
    if (strongEnabled && n == 32)
        strong = abs(p.corner() + p.top(63) - 2 * p.top(31)) < (1 << (bitDepth - 5)) &&
                 abs(p.corner() + p.left(63) - 2 * p.left(31)) < (1 << (bitDepth - 5));
    for (int i = lane; i <= 4 * n; i += 32)
    {
        int v;
        if (i == 0 || i == 4 * n)
            v = U[i];
        else if (strong)
        {
            if (i == c)
                v = U[c];
            else if (i > c) // top row, x = i - c - 1 in 0..62
            {
                const int x = i - c - 1;
                v = ((63 - x) * p.corner() + (x + 1) * p.top(63) + 32) >> 6;
            }
            else
            {
                const int y = c - 1 - i;
                v = ((63 - y) * p.corner() + (y + 1) * p.left(63) + 32) >> 6;
            }
        }
        else
            v = (U[i - 1] + 2 * U[i] + U[i + 1] + 2) >> 2;
        F[i] = (int16_t)v;
    }
}

template <typename Sample>
__global__ void __launch_bounds__(kWarps * 32)
    intraPredKernel(const HvbPlane *__restrict__ planes, const Sample *__restrict__ pool, const hvb_intra_task *__restrict__ tasks,
                    int n, int bitDepth)
{
    __shared__ int16_t sNb[kWarps][kNbMax + 3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warpsTotal = gridDim.x * kWarps;
    const int maxv = (1 << bitDepth) - 1;
    for (int i = blockIdx.x * kWarps + warp; i < n; i += warpsTotal)
    {
        const hvb_intra_task t = tasks[i];
        const int log2n = t.log2n, nn = 1 << log2n;
        for (int k = lane; k <= 4 * nn; k += 32) sNb[warp][k] = (int16_t)pool[t.nb - 2 * nn + k];
        __syncwarp();
        const Neighbours nb{sNb[warp], 2 * nn};
        const int dc = t.mode == 1 ? dcValue(nb, log2n, lane) : 0;
        int sd;
        Sample *dst = hvbBlockPtrW<Sample>(planes, t.dst, sd);
        for (int j = lane; j < nn * nn; j += 32)
        {
            const int y = j >> log2n, x = j & (nn - 1);
            dst[y * sd + x] = (Sample)intraSample(nb, t.mode, kAngle[t.mode], kInvAngle[t.mode], x, y, log2n, dc, t.edge_flag != 0, maxv);
        }
        __syncwarp();
    }
}

// ---- the sweep: predictions to shared memory row by row, SATD on the integer tensor cores (8- and 16-bit samples) ----
//
// The sum of absolute Hadamard coefficients of a tile equals that of the transposed tile, and a horizontal mode's
// prediction is the transpose of the vertical mode 36 - m evaluated on mirrored neighbours (left <-> top).  So
// modes 2..17 are swept on the transposed source tile and all 33 angular modes run the same "row of the
// vertical-mode formula" code: the row's two interpolation weights are constant and its samples slide along
// the (projected) reference.  A lane owns one (mode, tile row); the T predicted bytes go to shared memory where
// the (mode, tile) blocks are the B fragments of a [H | -H] x [src ; pred] IMMA, 8 modes at a time
// (hvb_me_subpel.cu has the same construction).
template <typename Sample>
struct SweepSmem
{
    int16_t front[8]; // mode 2's last row computes index -1 of U for a weight-0 sample; keep it inside the struct
    int16_t U[kNbMax + 7], F[kNbMax + 7];
    Sample pred[35][64];
    Sample src[2][64]; // [0] the tile, [1] the transposed source's tile
    int sum[36];
};

struct HadamardA8
{
    uint32_t e[2], o[2];
    __device__ __forceinline__ explicit HadamardA8(int lane)
    {
        const int g = lane >> 2, t = lane & 3;
        const uint32_t pat = (g & 2) ? ((g & 1) ? 0x01ffff01u : 0xffff0101u) : ((g & 1) ? 0xff01ff01u : 0x01010101u);
        const int q0 = (g >> 2) & t & 1, q1 = q0 ^ (t >> 1);
        e[0] = q0 ? pat ^ 0xfefefefeu : pat;
        e[1] = e[0] ^ 0xfefefefeu;
        o[0] = q1 ? pat ^ 0xfefefefeu : pat;
        o[1] = o[0] ^ 0xfefefefeu;
    }
};

__device__ __forceinline__ void imma16832(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// The 8 lanes that share (lane & 3) hold partial sums of accumulator columns 2t (s0) and 2t+1 (s1): exchange so that
// even g carries column 2t and odd g column 2t+1, then two more butterfly steps.  Lane (g < 2, t) ends with column 2t + g.
__device__ __forceinline__ int columnSums(int s0, int s1, int g)
{
    int sum = (g & 1) ? s1 : s0;
    sum += __shfl_xor_sync(0xffffffffu, (g & 1) ? s0 : s1, 4);
    sum += __shfl_xor_sync(0xffffffffu, sum, 8);
    sum += __shfl_xor_sync(0xffffffffu, sum, 16);
    return sum;
}

// B-fragment register(s) for 4 consecutive samples at p (aligned to 4 samples): 8 bit one word; 16 bit the low bytes and
// the high bytes as two words, the product runs on both planes and the sums are recombined (hvb_me_subpel.cu)
struct SweepFrag
{
    uint32_t lo, hi;
};
__device__ __forceinline__ SweepFrag sweepFrag(const uint8_t *p) { return SweepFrag{*reinterpret_cast<const uint32_t *>(p), 0u}; }
__device__ __forceinline__ SweepFrag sweepFrag(const uint16_t *p)
{
    const uint2 w = *reinterpret_cast<const uint2 *>(p);
    return SweepFrag{__byte_perm(w.x, w.y, 0x6420), __byte_perm(w.x, w.y, 0x7531)};
}
__device__ __forceinline__ void storeRow4(uint8_t *p, const int (&o)[4]) { *reinterpret_cast<uint32_t *>(p) = o[0] | o[1] << 8 | o[2] << 16 | o[3] << 24; }
__device__ __forceinline__ void storeRow4(uint16_t *p, const int (&o)[4]) { *reinterpret_cast<uint2 *>(p) = make_uint2(o[0] | o[1] << 16, o[2] | o[3] << 16); }
__device__ __forceinline__ void storeRow8(uint8_t *p, const int (&o)[8])
{
    *reinterpret_cast<uint2 *>(p) = make_uint2(o[0] | o[1] << 8 | o[2] << 16 | o[3] << 24, o[4] | o[5] << 8 | o[6] << 16 | o[7] << 24);
}
__device__ __forceinline__ void storeRow8(uint16_t *p, const int (&o)[8])
{
    *reinterpret_cast<uint4 *>(p) = make_uint4(o[0] | o[1] << 16, o[2] | o[3] << 16, o[4] | o[5] << 16, o[6] | o[7] << 16);
}

// T samples of row yy, columns x0 .. x0+T-1, of mode `mode` in its own frame (transposed for modes 2..17)
template <int T, typename Sample>
__device__ __forceinline__ void sweepRow(const SweepSmem<Sample> &s, int mode, int cIdx, int log2n, int dc, bool edge, int maxv, int x0, int yy,
                                         int (&out)[T])
{
    const int n = 1 << log2n, c = 2 * n;
    const int16_t *A = filterFlag(cIdx, mode, n) ? s.F : s.U;
    if (mode >= 2)
    {
        const bool vertical = mode >= 18;
        const int angle = kAngle[mode], inv = kInvAngle[mode];
        const int t = (yy + 1) * angle, idx = t >> 5, fact = t & 31;
        const int i0 = x0 + idx + 1;
        if (i0 >= 0)
        {
            // the whole row reads the main side: consecutive reference samples, no projection
            const int16_t *R = A + c + (vertical ? i0 : -i0);
            const int dir = vertical ? 1 : -1;
            int prev = R[0];
#pragma unroll
            for (int x = 0; x < T; ++x)
            {
                // when fact == 0 the reference does not read the second sample; reading it is harmless (the arrays are
                // padded) and (32 r0 + 16) >> 5 == r0
                const int next = R[dir * (x + 1)];
                out[x] = ((32 - fact) * prev + fact * next + 16) >> 5;
                prev = next;
            }
        }
        else
        {
            int prev;
            {
                const int sft = -((i0 * inv + 128) >> 8);
                prev = vertical ? A[c + sft] : A[c - sft];
            }
#pragma unroll
            for (int x = 0; x < T; ++x)
            {
                const int i = i0 + x + 1, sft = i >= 0 ? i : -((i * inv + 128) >> 8);
                const int next = vertical ? A[c + sft] : A[c - sft];
                out[x] = ((32 - fact) * prev + fact * next + 16) >> 5;
                prev = next;
            }
        }
        if (edge && (mode == 26 || mode == 10) && x0 == 0)
        {
            // pred_intra.cpp:20354-20358 / :20392-20396; for mode 10 in the mirrored frame the roles of top and left swap
            const int top0 = vertical ? A[c + 1] : A[c - 1], side = vertical ? A[c - 1 - yy] : A[c + 1 + yy];
            out[0] = hvbClip3(0, maxv, top0 + ((side - A[c]) >> 1));
        }
    }
    else if (mode == 1)
    {
#pragma unroll
        for (int x = 0; x < T; ++x)
        {
            const int xx = x0 + x;
            int v = dc;
            if (edge)
            {
                if (xx == 0 && yy == 0)
                    v = (A[c - 1] + 2 * dc + A[c + 1] + 2) >> 2;
                else if (yy == 0)
                    v = (A[c + 1 + xx] + 3 * dc + 2) >> 2;
                else if (xx == 0)
                    v = (A[c - 1 - yy] + 3 * dc + 2) >> 2;
            }
            out[x] = v;
        }
    }
    else
    {
        const int left = A[c - 1 - yy], topN = A[c + 1 + n], leftN = A[c - 1 - n];
#pragma unroll
        for (int x = 0; x < T; ++x)
        {
            const int xx = x0 + x;
            out[x] = ((n - 1 - xx) * left + (xx + 1) * topN + (n - 1 - yy) * A[c + 1 + xx] + (yy + 1) * leftN + n) >> (log2n + 1);
        }
    }
}

template <typename Sample>
__global__ void __launch_bounds__(kWarps * 32)
    intraSweepKernel8(const HvbPlane *__restrict__ planes, const Sample *__restrict__ pool, const hvb_intra_sweep_task *__restrict__ tasks,
                      int n, int32_t *__restrict__ out, int bitDepth)
{
    __shared__ __align__(16) SweepSmem<Sample> sAll[kWarps];
    constexpr bool k16 = sizeof(Sample) == 2;
    const int maxv = (1 << bitDepth) - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    SweepSmem<Sample> &s = sAll[warp];
    const HadamardA8 A(lane);
    const int g = lane >> 2, tq = lane & 3;
    const int warpsTotal = gridDim.x * kWarps;
    for (int i = blockIdx.x * kWarps + warp; i < n; i += warpsTotal)
    {
        const hvb_intra_sweep_task t = tasks[i];
        const int log2n = t.log2n, nn = 1 << log2n;
        for (int k = lane; k <= 4 * nn; k += 32)
        {
            s.U[k] = (int16_t)pool[t.nb_unfiltered - 2 * nn + k];
            if (t.nb_filtered >= 0) s.F[k] = (int16_t)pool[t.nb_filtered - 2 * nn + k];
        }
        if (lane < 6) // padding read (never used) by the fact == 0 rows
            s.U[4 * nn + 1 + lane] = s.F[4 * nn + 1 + lane] = 0;
        for (int k = lane; k < 36; k += 32) s.sum[k] = 0;
        __syncwarp();
        if (t.nb_filtered < 0) filterNeighbours(s.F, s.U, nn, bitDepth, t.strong_intra_smoothing != 0, lane);
        __syncwarp();
        const Neighbours nbU{s.U, 2 * nn};
        const int dc = dcValue(nbU, log2n, lane);
        const bool edge = t.cIdx == 0 && log2n < 5;
        int ss;
        const Sample *src = hvbBlockPtr<Sample>(planes, t.src, ss);

        if (log2n == 2)
        {
            // one 4x4 tile: 35 x 4 row jobs, then 5 groups of 8 modes, one IMMA each (K = 16 source + 16 prediction)
            if (lane < 16)
            {
                const int r = lane >> 2, cc = lane & 3;
                const Sample v = src[r * ss + cc];
                s.src[0][r * 4 + cc] = v;
                s.src[1][cc * 4 + r] = v;
            }
            for (int job = lane; job < 35 * 4; job += 32)
            {
                const int mode = job >> 2, r = job & 3;
                int o[4];
                sweepRow<4>(s, mode, t.cIdx, 2, dc, edge, maxv, 0, r, o);
                storeRow4(&s.pred[mode][r * 4], o);
            }
            __syncwarp();
#pragma unroll 1
            for (int base = 0; base < 40; base += 8)
            {
                const int mode = min(base + g, 34);
                const Sample *S = s.src[mode >= 2 && mode < 18];
                const SweepFrag b0 = sweepFrag(S + tq * 4), b1 = sweepFrag(&s.pred[mode][tq * 4]);
                int acc[4] = {0, 0, 0, 0};
                imma16832(acc, A.e[0], A.o[0], A.e[1], A.o[1], b0.lo, b1.lo);
                if (k16)
                {
                    int ach[4] = {0, 0, 0, 0};
                    imma16832(ach, A.e[0], A.o[0], A.e[1], A.o[1], b0.hi, b1.hi);
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[r] += ach[r] << 8;
                }
                int s0 = __sad(acc[0], 0, __sad(acc[2], 0, 0u)), s1 = __sad(acc[1], 0, __sad(acc[3], 0, 0u));
                const int sum = columnSums(s0, s1, g);
                if (g < 2 && base + 2 * tq + g < 35) s.sum[base + 2 * tq + g] = ((sum + 1) >> 1) >> (k16 ? 2 : 0);
            }
        }
        else
        {
            const int tilesPerRow = nn >> 3;
#pragma unroll 1
            for (int ty = 0; ty < tilesPerRow; ++ty)
#pragma unroll 1
                for (int tx = 0; tx < tilesPerRow; ++tx)
                {
                    {
                        // the tile (two words per row) and the tile of the transposed source at the same tile coordinates
                        const int r = lane >> 2, cc = (lane & 3) * 2;
                        const Sample *p = src + (ty * 8 + r) * ss + tx * 8 + cc;
                        s.src[0][r * 8 + cc] = p[0];
                        s.src[0][r * 8 + cc + 1] = p[1];
                        const Sample *q = src + (tx * 8 + cc) * ss + ty * 8 + r; // srcT(x = tx*8+cc, y = ty*8+r) = src(x = ty*8+r, y = tx*8+cc)
                        s.src[1][r * 8 + cc] = q[0];
                        s.src[1][r * 8 + cc + 1] = q[ss];
                    }
                    for (int job = lane; job < 35 * 8; job += 32)
                    {
                        const int mode = job >> 3, r = job & 7;
                        int o[8];
                        sweepRow<8>(s, mode, t.cIdx, log2n, dc, edge, maxv, tx * 8, ty * 8 + r, o);
                        storeRow8(&s.pred[mode][r * 8], o);
                    }
                    __syncwarp();
#pragma unroll 1
                    for (int base = 0; base < 40; base += 8)
                    {
                        const int mode = min(base + g, 34);
                        const Sample *S = s.src[mode >= 2 && mode < 18], *P = s.pred[mode];
                        const int row = tq >> 1, cx = (tq & 1) * 4;
                        SweepFrag b[4][2];
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                        {
                            b[ks][0] = sweepFrag(S + (ks * 4 + row) * 8 + cx);
                            b[ks][1] = sweepFrag(S + (ks * 4 + row + 2) * 8 + cx);
                            b[ks + 2][0] = sweepFrag(P + (ks * 4 + row) * 8 + cx);
                            b[ks + 2][1] = sweepFrag(P + (ks * 4 + row + 2) * 8 + cx);
                        }
                        int s0 = 0, s1 = 0;
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt)
                        {
                            int acc[4] = {0, 0, 0, 0}, ach[4] = {0, 0, 0, 0};
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                            {
                                const int n01 = (((mt >> 1) & ks) ^ (ks >> 1)) & 1, n23 = n01 ^ (mt & 1);
                                imma16832(acc, A.e[n01], A.o[n01], A.e[n23], A.o[n23], b[ks][0].lo, b[ks][1].lo);
                                if (k16) imma16832(ach, A.e[n01], A.o[n01], A.e[n23], A.o[n23], b[ks][0].hi, b[ks][1].hi);
                            }
                            if (k16)
#pragma unroll
                                for (int r = 0; r < 4; ++r) acc[r] += ach[r] << 8;
                            s0 = __sad(acc[0], 0, __sad(acc[2], 0, (unsigned)s0));
                            s1 = __sad(acc[1], 0, __sad(acc[3], 0, (unsigned)s1));
                        }
                        const int sum = columnSums(s0, s1, g);
                        if (g < 2 && base + 2 * tq + g < 35) s.sum[base + 2 * tq + g] += ((sum + 2) >> 2) >> (k16 ? 2 : 0);
                    }
                    __syncwarp();
                }
        }
        __syncwarp();
        for (int k = lane; k < 35; k += 32) out[i * 35 + k] = s.sum[k];
        __syncwarp();
    }
}

int gridWarps(hvb_context *ctx, int n)
{
    const int blocks = (n + kWarps - 1) / kWarps;
    const int cap = ctx->smCount * 8;
    return blocks < cap ? blocks : cap;
}

} // namespace

extern "C" int hvb_intra_pred_batch(hvb_context *ctx, const hvb_intra_task *tasks, int n, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || tasks) && ctx->samplePool);
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, nullptr, 0, mem, &st);
    if (rc) return rc;
    const auto *dT = static_cast<const hvb_intra_task *>(st.dTasks);
    if (ctx->bps == 1)
        intraPredKernel<uint8_t><<<gridWarps(ctx, n), kWarps * 32, 0, ctx->stream>>>(
            ctx->dPlanes, static_cast<const uint8_t *>(ctx->samplePool), dT, n, ctx->bitDepth);
    else
        intraPredKernel<uint16_t><<<gridWarps(ctx, n), kWarps * 32, 0, ctx->stream>>>(
            ctx->dPlanes, static_cast<const uint16_t *>(ctx->samplePool), dT, n, ctx->bitDepth);
    HVB_LAUNCH_CHECK(ctx, "intraPredKernel");
    return hvbStageOut(ctx, nullptr, 0, mem, st);
}

extern "C" int hvb_intra_satd35_batch(hvb_context *ctx, const hvb_intra_sweep_task *tasks, int n, int32_t *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)) && ctx->samplePool);
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(int32_t) * 35 * n, mem, &st);
    if (rc) return rc;
    const auto *dT = static_cast<const hvb_intra_sweep_task *>(st.dTasks);
    auto *dO = static_cast<int32_t *>(st.dOut);
    int perSm = 1;
    if (ctx->bps == 1)
    {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, intraSweepKernel8<uint8_t>, kWarps * 32, 0);
        const int blocks = min((n + kWarps - 1) / kWarps, ctx->smCount * max(perSm, 1));
        intraSweepKernel8<uint8_t><<<blocks, kWarps * 32, 0, ctx->stream>>>(ctx->dPlanes, static_cast<const uint8_t *>(ctx->samplePool), dT, n, dO,
                                                                           ctx->bitDepth);
    }
    else
    {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, intraSweepKernel8<uint16_t>, kWarps * 32, 0);
        const int blocks = min((n + kWarps - 1) / kWarps, ctx->smCount * max(perSm, 1));
        intraSweepKernel8<uint16_t><<<blocks, kWarps * 32, 0, ctx->stream>>>(ctx->dPlanes, static_cast<const uint16_t *>(ctx->samplePool), dT, n,
                                                                            dO, ctx->bitDepth);
    }
    HVB_LAUNCH_CHECK(ctx, "intraSweepKernel");
    return hvbStageOut(ctx, out, sizeof(int32_t) * 35 * n, mem, st);
}
