// hvb_pu_cost.cu -- the distortion half of measurePuCost: predictInter (uni / bi, luma 8-tap, chroma 4-tap) followed by
// the Hadamard SATD of Y, Cb and Cr against the source picture, one launch per batch of PU candidates.
//
// Reference semantics (bit-exact):
//   measurePuCost                 turing/Search.hpp:1668-1682
//   predictUni / predictBi        turing/Dsp.h:769-864 (clipMvLumaComponent :723-731, chroma origin = luma origin >> 1,
//                                 chroma phase = mv & 7)
//   HavocPredUni / HavocPredBi    havoc/pred_inter.cpp:76-202, :1207-1252
//   measureSatd                   turing/Measure.h:96-135, :156-160 (a chroma block that is not a multiple of 4: 0)
//   havoc_hadamard_satd           havoc/hadamard.cpp:58-98
//
// Mapping.  A warp pulls PUs from a device-side cursor and walks each PU in STRIPS: 8 rows of the block when the
// component is measured in 8x8 Hadamard tiles (both sides multiples of 8), else 4 rows.  Luma is one pass over the
// PU's strips; Cb and Cr are a second pass in which the two chroma blocks lie side by side (columns 0..wc-1 and
// wc..2wc-1), so that a small PU's chroma still fills lanes.  Per strip and per reference list:
//   H pass   (support row, 4 columns) jobs: the row's bytes are loaded once as aligned words, every output is two
//            IDP.4A (8 bit; chroma one) or four IDP.2A (16 bit); intermediates are the reference's 16-bit `mid`
//            values, stored column-major in shared memory.
//   V pass   a lane owns a column: its 16 support rows are four 64-bit shared loads, the window slides down them
//            with 4 (chroma 2) IDP.2A per output.  For a bi-predicted PU the first list's 14-bit values stay in the
//            lane's registers until the second list's arrive.
//   SATD     eight tiles at a time as one [H | -H] x [src ; pred] product on the integer tensor cores (hvb_unit.cuh,
//            derivation in hvb_me_subpel.cu), from the 8- or 16-bit strips of source and prediction in shared memory.
// A warp needs 3.5 KB (8 bit) of shared memory instead of the 17 KB of the first-generation kernel (a whole 64x64
// block per warp), which is what lets 48 warps per SM hide the latency of the reference-picture reads.
#include "hvb_internal.cuh"
#include "hvb_unit.cuh"

namespace {

using namespace hvb_unit;

constexpr int kWarps = 8;
constexpr int kMaxW = 64; // columns of a strip: the luma PU width, or Cb | Cr side by side
constexpr int kGrab = 4;  // PUs a warp takes from the cursor at a time

// 4-tap chroma filters (havoc/pred_inter.cpp:52-69) packed as s8x4 words
__device__ __constant__ uint32_t kChromaWords[8] = {0x00004000u, 0xfe0a3afeu, 0xfe1036fcu, 0xfc1c2efau,
                                                    0xfc2424fcu, 0xfa2e1cfcu, 0xfc3610feu, 0xfe3a0afeu};

template <typename Sample>
struct alignas(16) WarpSmem
{
    static constexpr int kRow = sizeof(Sample) == 1 ? 64 : 72; // samples per strip row (16 bit: padded against bank conflicts)
    int16_t mids[kMaxW * kColStride];                         // [column][support row]
    Sample pred[8 * kRow];
    Sample src[8 * kRow];
    int16_t first[kMaxW * 8]; // [column][row]: the first list's values of a bi-predicted strip
};

__device__ __forceinline__ int clipMvLumaComponent(int component, int nPbSize, int pictureSize)
{
    if (component + nPbSize + 4 < 0) return -nPbSize - 4;
    if (component > pictureSize + 2) return pictureSize + 2;
    return component;
}

// ---- H pass: support rows 0..S+TAPS-2 of a strip, columns of `ncomp` blocks of width wc side by side ----------------
// base[c] points at (strip row 0, column 0) of block c in its reference plane.
template <typename Sample, int TAPS, int S>
__device__ __forceinline__ void hPass(int16_t *mids, const Sample *base0, const Sample *base1, int stride, int wc, int ncomp, int xFrac,
                                      const Depth &D, int lane)
{
    constexpr int R = S + TAPS - 1, M = TAPS / 2 - 1;
    const int quads = (wc + 3) >> 2, jobs = ncomp * quads * R;
    const uint32_t t0 = TAPS == 8 ? kTapWords[xFrac][0] : kChromaWords[xFrac], t1 = TAPS == 8 ? kTapWords[xFrac][1] : 0u;
#pragma unroll 1
    for (int job = lane; job < jobs; job += 32)
    {
        const int qk = job / R, r = job - qk * R;
        const int comp = qk >= quads, k = comp ? qk - quads : qk;
        const Sample *p = (comp ? base1 : base0) + (intptr_t)(r - M) * stride + 4 * k - M;
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
        const unsigned sh = (unsigned)(a & 3) * 8;
        int m[4];
        if (sizeof(Sample) == 1)
        {
            // 8 bit: 11 (chroma 7) bytes of the row; u8 samples x s8 taps; shift1 = 0
            constexpr int NW = TAPS == 8 ? 4 : 3;
            uint32_t w[NW], v[NW - 1];
#pragma unroll
            for (int i = 0; i < NW; ++i) w[i] = __ldg(q + i);
#pragma unroll
            for (int i = 0; i < NW - 1; ++i) v[i] = __funnelshift_r(w[i], w[i + 1], sh);
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const uint32_t lo = j ? __funnelshift_r(v[0], v[1], 8 * j) : v[0];
                if constexpr (TAPS == 8)
                {
                    const uint32_t hi = j ? __funnelshift_r(v[1], v[NW - 2], 8 * j) : v[1];
                    m[j] = dp4aUS(hi, t1, dp4aUS(lo, t0, 0));
                }
                else
                    m[j] = dp4aUS(lo, t0, 0);
            }
        }
        else
        {
            // 16 bit: 11 (chroma 7) samples of the row as words of two; s16 samples x s8 taps, then >> shift1
            constexpr int NV = TAPS == 8 ? 6 : 4;
            uint32_t w[NV + 1], v[NV];
#pragma unroll
            for (int i = 0; i < NV + 1; ++i) w[i] = __ldg(q + i);
#pragma unroll
            for (int i = 0; i < NV; ++i) v[i] = __funnelshift_r(w[i], w[i + 1], sh);
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
                const int kk = j >> 1;
                const bool odd = j & 1;
                const uint32_t p0 = odd ? __funnelshift_r(v[kk], v[kk + 1], 16) : v[kk];
                const uint32_t p1 = odd ? __funnelshift_r(v[kk + 1], v[kk + 2], 16) : v[kk + 1];
                int sum;
                if constexpr (TAPS == 8)
                {
                    const uint32_t p2 = odd ? __funnelshift_r(v[kk + 2], v[kk + 3], 16) : v[kk + 2];
                    const uint32_t p3 = odd ? __funnelshift_r(v[kk + 3], v[kk + 4 < NV ? kk + 4 : NV - 1], 16) : v[kk + 3];
                    sum = tap8(p0, p1, p2, p3, t0, t1, 0);
                }
                else
                    sum = __dp2a_hi((int)p1, (int)t0, __dp2a_lo((int)p0, (int)t0, 0));
                m[j] = sum >> D.shift1;
            }
        }
        int16_t *col = mids + (comp * wc + 4 * k) * kColStride + r;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (4 * k + j < wc) col[j * kColStride] = (int16_t)m[j];
    }
}

// second-stage sum of output row ROW of a column (support rows ROW .. ROW + TAPS - 1), without rounding or shift
template <int TAPS, int ROW>
__device__ __forceinline__ int vSum(const Column &c, uint32_t t0, uint32_t t1)
{
    constexpr int k = ROW >> 1;
    if constexpr (TAPS == 8)
    {
        if (ROW & 1) return tap8(c.x[k], c.x[k + 1], c.x[k + 2], c.x[k + 3 < 7 ? k + 3 : 6], t0, t1, 0);
        return tap8(c.w[k], c.w[k + 1], c.w[k + 2], c.w[k + 3 < 8 ? k + 3 : 7], t0, t1, 0);
    }
    if (ROW & 1) return __dp2a_hi((int)c.x[k + 1], (int)t0, __dp2a_lo((int)c.x[k], (int)t0, 0));
    return __dp2a_hi((int)c.w[k + 1], (int)t0, __dp2a_lo((int)c.w[k], (int)t0, 0));
}

template <int TAPS, int S>
__device__ __forceinline__ void vSums(const int16_t *col, int yFrac, int (&sum)[S])
{
    const uint32_t t0 = TAPS == 8 ? kTapWords[yFrac][0] : kChromaWords[yFrac], t1 = TAPS == 8 ? kTapWords[yFrac][1] : 0u;
    Column c;
    c.load(col, 0);
    sum[0] = vSum<TAPS, 0>(c, t0, t1);
    sum[1] = vSum<TAPS, 1>(c, t0, t1);
    sum[2] = vSum<TAPS, 2>(c, t0, t1);
    sum[3] = vSum<TAPS, 3>(c, t0, t1);
    if constexpr (S == 8)
    {
        sum[4] = vSum<TAPS, 4>(c, t0, t1);
        sum[5] = vSum<TAPS, 5>(c, t0, t1);
        sum[6] = vSum<TAPS, 6>(c, t0, t1);
        sum[7] = vSum<TAPS, 7>(c, t0, t1);
    }
}

// ---- SATD of the strip's tiles (8x8 when S == 8, else 4x4) on the tensor cores; lane (g < 2, t) of each group of
// eight tiles receives the value of tile base + 2t + g and adds it to acc[tile's block] ---------------------------------
template <typename Sample, int S>
__device__ __forceinline__ void satdStrip(const WarpSmem<Sample> &s, int tiles, int wc, const HadamardA &A, int lane, int (&acc)[2])
{
    constexpr int kRow = WarpSmem<Sample>::kRow;
    constexpr bool k16 = sizeof(Sample) == 2, T8 = S == 8;
    constexpr int TW = T8 ? 8 : 4;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
    for (int base = 0; base < tiles; base += 8)
    {
        const int tile = min(base + g, tiles - 1);
        const Sample *Sp = s.src + tile * TW, *Pp = s.pred + tile * TW;
        int s0, s1;
        if (T8)
        {
            // k = 32 ks + 4 t + j (+16): tile row 4 ks + (t >> 1) (+2), tile column 4 (t & 1) + j
            const int row = t >> 1, cx = (t & 1) * 4;
            Frag b[4][2];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
            {
                b[ks][0] = loadFrag(Sp + (ks * 4 + row) * kRow + cx, 0);
                b[ks][1] = loadFrag(Sp + (ks * 4 + row + 2) * kRow + cx, 0);
                b[ks + 2][0] = loadFrag(Pp + (ks * 4 + row) * kRow + cx, 0);
                b[ks + 2][1] = loadFrag(Pp + (ks * 4 + row + 2) * kRow + cx, 0);
            }
            s0 = s1 = 0;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
            {
                int c[4] = {0, 0, 0, 0}, ch[4] = {0, 0, 0, 0};
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                {
                    const int n01 = (((mt >> 1) & ks) ^ (ks >> 1)) & 1; // bit 5 of m & k, and the prediction half of [H | -H]
                    const int n23 = n01 ^ (mt & 1);                     // bit 4 of m & k
                    imma16832(c, A.e[n01], A.o[n01], A.e[n23], A.o[n23], b[ks][0].lo, b[ks][1].lo);
                    if (k16) imma16832(ch, A.e[n01], A.o[n01], A.e[n23], A.o[n23], b[ks][0].hi, b[ks][1].hi);
                }
                if (k16)
#pragma unroll
                    for (int r = 0; r < 4; ++r) c[r] += ch[r] << 8;
                s0 = __sad(c[0], 0, __sad(c[2], 0, (unsigned)s0));
                s1 = __sad(c[1], 0, __sad(c[3], 0, (unsigned)s1));
            }
        }
        else
        {
            // a 4x4 tile: K = 16 + 16 is one k-step, lane (g, t) supplies row t of tile g from both blocks
            const Frag b0 = loadFrag(Sp + t * kRow, 0);
            const Frag b1 = loadFrag(Pp + t * kRow, 0);
            int c[4] = {0, 0, 0, 0};
            imma16832(c, A.e[0], A.o[0], A.e[1], A.o[1], b0.lo, b1.lo);
            if (k16)
            {
                int ch[4] = {0, 0, 0, 0};
                imma16832(ch, A.e[0], A.o[0], A.e[1], A.o[1], b0.hi, b1.hi);
#pragma unroll
                for (int r = 0; r < 4; ++r) c[r] += ch[r] << 8;
            }
            s0 = __sad(c[0], 0, __sad(c[2], 0, 0u));
            s1 = __sad(c[1], 0, __sad(c[3], 0, 0u));
        }
        // the 8 lanes that share t hold partial sums of tiles 2t (s0) and 2t+1 (s1) of the group
        int sum = (g & 1) ? s1 : s0;
        sum += __shfl_xor_sync(0xffffffffu, (g & 1) ? s0 : s1, 4);
        sum += __shfl_xor_sync(0xffffffffu, sum, 8);
        sum += __shfl_xor_sync(0xffffffffu, sum, 16);
        const int mineTile = base + 2 * t + g;
        if (g < 2 && mineTile < tiles)
        {
            // havoc/hadamard.cpp:81-97: 4x4 (s + 1) >> 1, 8x8 (s + 2) >> 2, 16-bit samples >> 2 more
            int v = (sum + (T8 ? 2 : 1)) >> (T8 ? 2 : 1);
            if (k16) v >>= 2;
            if (mineTile * TW >= wc)
                acc[1] += v;
            else
                acc[0] += v;
        }
    }
}

// ---- one pass over a PU: luma (TAPS = 8, one block) or chroma (TAPS = 4, Cb | Cr), strips of S rows ------------------
template <typename Sample, int TAPS, int S>
__device__ __forceinline__ void puPass(WarpSmem<Sample> &s, const HvbPlane *__restrict__ planes, const hvb_pu_cost_task &t,
                                       int bx0, int bx1, int by0, int by1, bool satd, const Depth &D, const HadamardA &A, int lane,
                                       int (&acc)[2])
{
    constexpr int kRow = WarpSmem<Sample>::kRow;
    constexpr bool C = TAPS == 4;
    constexpr int sh = C ? 1 : 0, ncomp = C ? 2 : 1, c0 = C ? 1 : 0, fracMask = C ? 7 : 3;
    const int wc = t.w >> sh, hc = t.h >> sh, W = wc * ncomp;
    const bool bi = t.ref_pic[0] >= 0 && t.ref_pic[1] >= 0;
    const int x0 = t.x0 >> sh, y0 = t.y0 >> sh;

#pragma unroll 1
    for (int ys = 0; ys < hc; ys += S)
    {
        if (satd)
        {
            // the source strip: S rows of W samples in chunks of four (x0 is a multiple of 4 luma samples)
            const int chunks = W >> 2, perBlock = wc >> 2;
            for (int i = lane; i < chunks * S; i += 32)
            {
                const int r = i / chunks, k = i - r * chunks;
                const int comp = k >= perBlock, kk = comp ? k - perBlock : k;
                const HvbPlane &sp = planes[t.src_pic * 3 + c0 + comp];
                const Sample *p = reinterpret_cast<const Sample *>(sp.base) + (intptr_t)(y0 + ys + r) * sp.stride + x0 + 4 * kk;
                uint32_t *d = reinterpret_cast<uint32_t *>(s.src + r * kRow + 4 * k);
                if (sizeof(Sample) == 1)
                    d[0] = hvbLoad4u8(reinterpret_cast<const uint8_t *>(p));
                else
                {
                    const uint32_t *q = reinterpret_cast<const uint32_t *>(p);
                    d[0] = __ldg(q);
                    d[1] = __ldg(q + 1);
                }
            }
        }

        bool second = false;
#pragma unroll 1
        for (int l = 0; l < 2; ++l)
        {
            const int refPic = l ? t.ref_pic[1] : t.ref_pic[0];
            if (refPic < 0) continue;
            const int mvx = l ? t.mvx[1] : t.mvx[0], mvy = l ? t.mvy[1] : t.mvy[0];
            const HvbPlane &rp = planes[refPic * 3 + c0];
            const intptr_t off = (intptr_t)(((l ? by1 : by0) >> sh) + ys) * rp.stride + ((l ? bx1 : bx0) >> sh);
            const Sample *base0 = reinterpret_cast<const Sample *>(rp.base) + off;
            const Sample *base1 = C ? reinterpret_cast<const Sample *>(planes[refPic * 3 + 2].base) + off : base0;
            hPass<Sample, TAPS, S>(s.mids, base0, base1, rp.stride, wc, ncomp, mvx & fracMask, D, lane);
            __syncwarp();
            const int yFrac = mvy & fracMask;
#pragma unroll 1
            for (int c = lane; c < W; c += 32)
            {
                int sum[S];
                vSums<TAPS, S>(s.mids + c * kColStride, yFrac, sum);
                uint32_t *keep = reinterpret_cast<uint32_t *>(s.first + c * 8); // the lane's own column: no synchronisation
                if (bi && !second)
                {
                    // HavocPredBi's 14-bit intermediates of the first list (int16 in the reference too)
#pragma unroll
                    for (int r = 0; r < S; r += 2) keep[r >> 1] = ((uint32_t)(sum[r] >> 6) & 0xffffu) | ((uint32_t)(sum[r + 1] >> 6) << 16);
                }
                else
                {
                    Sample *dst = s.pred + c;
#pragma unroll
                    for (int r = 0; r < S; ++r)
                    {
                        if (!bi)
                            dst[r * kRow] = (Sample)D.out(sum[r]);
                        else
                        {
                            const uint32_t pair = keep[r >> 1];
                            const int f = (r & 1) ? (int)pair >> 16 : (int)(int16_t)(pair & 0xffffu);
                            dst[r * kRow] = (Sample)__vimin_s32_relu((f + (sum[r] >> 6) + (1 << D.shift3)) >> (D.shift3 + 1), D.maxv);
                        }
                    }
                }
            }
            __syncwarp();
            second = true;
        }

        if (t.dst_pic >= 0)
        {
            // what predictInter leaves in the reconstructed picture
            for (int i = lane; i < W * S; i += 32)
            {
                const int r = i / W, c = i - r * W;
                const int comp = c >= wc, x = comp ? c - wc : c;
                if (ys + r < hc)
                {
                    const HvbPlane &dp = planes[t.dst_pic * 3 + c0 + comp];
                    reinterpret_cast<Sample *>(dp.base)[(intptr_t)(y0 + ys + r) * dp.stride + x0 + x] = s.pred[r * kRow + c];
                }
            }
        }
        if (satd) satdStrip<Sample, S>(s, W / (S == 8 ? 8 : 4), wc, A, lane, acc);
        __syncwarp();
    }
}

template <typename Sample>
__global__ void __launch_bounds__(kWarps * 32, 4)
    puCostKernel(const HvbPlane *__restrict__ planes, const hvb_pu_cost_task *__restrict__ tasks, int n, int32_t *__restrict__ out,
                 int bitDepth, int *__restrict__ cursor)
{
    extern __shared__ __align__(16) uint8_t smemPuCost[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpSmem<Sample> &s = reinterpret_cast<WarpSmem<Sample> *>(smemPuCost)[warp];
    const HadamardA A(lane);
    const Depth D(sizeof(Sample) == 1 ? 8 : bitDepth);
    // a warp takes kGrab consecutive PUs per visit to the cursor (the atomic and its broadcast were 10 % of the stall samples
    // at one PU per visit, profiles/r01g_hot_lines.txt; consecutive PUs of a task list are of similar size)
    int i = 0, end = 0;
    for (;;)
    {
        if (i == end)
        {
            if (lane == 0) i = atomicAdd(cursor, kGrab);
            i = __shfl_sync(0xffffffffu, i, 0);
            if (i >= n) break;
            end = min(i + kGrab, n);
        }
        const hvb_pu_cost_task t = tasks[i];
        int bx0 = 0, bx1 = 0, by0 = 0, by1 = 0;
        if (t.ref_pic[0] >= 0)
        {
            const HvbPlane &rp = planes[t.ref_pic[0] * 3];
            bx0 = clipMvLumaComponent(t.x0 + (t.mvx[0] >> 2), t.w, rp.width);
            by0 = clipMvLumaComponent(t.y0 + (t.mvy[0] >> 2), t.h, rp.height);
        }
        if (t.ref_pic[1] >= 0)
        {
            const HvbPlane &rp = planes[t.ref_pic[1] * 3];
            bx1 = clipMvLumaComponent(t.x0 + (t.mvx[1] >> 2), t.w, rp.width);
            by1 = clipMvLumaComponent(t.y0 + (t.mvy[1] >> 2), t.h, rp.height);
        }
        int y[2] = {0, 0}, c[2] = {0, 0};
        // measureSatd picks the tile from the alignment of (w | h) of the component (turing/Measure.h:96-135)
        if ((t.w | t.h) & 7)
            puPass<Sample, 8, 4>(s, planes, t, bx0, bx1, by0, by1, true, D, A, lane, y);
        else
            puPass<Sample, 8, 8>(s, planes, t, bx0, bx1, by0, by1, true, D, A, lane, y);
        const int wh = (t.w | t.h) >> 1;
        const bool chromaSatd = (wh & 3) == 0;
        if (chromaSatd || t.dst_pic >= 0)
        {
            if (wh & 7)
                puPass<Sample, 4, 4>(s, planes, t, bx0, bx1, by0, by1, chromaSatd, D, A, lane, c);
            else
                puPass<Sample, 4, 8>(s, planes, t, bx0, bx1, by0, by1, true, D, A, lane, c);
        }
        const int sy = hvbWarpSum(y[0]), scb = hvbWarpSum(c[0]), scr = hvbWarpSum(c[1]);
        if (lane == 0)
        {
            out[3 * i] = sy;
            out[3 * i + 1] = scb;
            out[3 * i + 2] = scr;
        }
        ++i;
    }
}

template <typename Sample>
int launch(hvb_context *ctx, const hvb_pu_cost_task *dTasks, int n, int32_t *dOut)
{
    const int smem = kWarps * (int)sizeof(WarpSmem<Sample>);
    static_assert(sizeof(WarpSmem<Sample>) % 16 == 0, "per-warp shared slices must stay 16-byte aligned");
    cudaFuncSetAttribute(puCostKernel<Sample>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int perSm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, puCostKernel<Sample>, kWarps * 32, smem);
    int blocks = (n + kWarps - 1) / kWarps;
    const int cap = ctx->smCount * (perSm > 0 ? perSm : 1);
    if (blocks > cap) blocks = cap;
    int *cursor = ctx->workCursors + 1;
    cudaMemsetAsync(cursor, 0, sizeof(int), ctx->stream);
    puCostKernel<Sample><<<blocks, kWarps * 32, smem, ctx->stream>>>(ctx->dPlanes, dTasks, n, dOut, ctx->bitDepth, cursor);
    return 0;
}

} // namespace

extern "C" int hvb_pu_cost_batch(hvb_context *ctx, const hvb_pu_cost_task *tasks, int n, int32_t *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(int32_t) * 3 * n, mem, &st);
    if (rc) return rc;
    const auto *dT = static_cast<const hvb_pu_cost_task *>(st.dTasks);
    auto *dO = static_cast<int32_t *>(st.dOut);
    if (ctx->bps == 1)
        launch<uint8_t>(ctx, dT, n, dO);
    else
        launch<uint16_t>(ctx, dT, n, dO);
    HVB_LAUNCH_CHECK(ctx, "puCostKernel");
    return hvbStageOut(ctx, out, sizeof(int32_t) * 3 * n, mem, st);
}
