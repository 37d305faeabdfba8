// hvb_metrics.cu -- batched SAD / SAD4 / SSD / Hadamard-SATD over device-resident pictures.
//
// Reference semantics (bit-exact):
//   havoc_sad           havoc/sad.cpp:432-449     (u16: >> 2)
//   havoc_sad_multiref  havoc/sad.cpp:513-542     (4 references, u16: each >> 2)
//   havoc_ssd           havoc/ssd.cpp:28-43       (uint32 wrap, u16: >> 4)
//   havoc_hadamard_satd havoc/hadamard.cpp:58-98  (2x2 / 4x4 (s+1)>>1 / 8x8 (s+2)>>2, u16: >> 2)
//   measureSatd         turing/Measure.h:96-135   (tiling of a w x h block)
//
// Mapping: one warp per candidate.  These kernels are pure streaming reductions (every sample is
// read once per candidate), so the bound is memory: 2*w*h*B algorithmic bytes per candidate
// (5*w*h*B for SAD4).  Blocks are arbitrarily aligned (motion-search candidates): every 16-byte chunk
// of a row is one aligned 128-bit load, or two and a word funnel.  Byte / halfword lanes are reduced
// with the SIMD-in-word video instructions (VABSDIFF4 / VABSDIFF2 / dp4a), then a 5-step shuffle tree.
#include "hvb_internal.cuh"
#include "hvb_satd.cuh"
#include "hvb_unit.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;

// 16 bytes at an arbitrary byte address: two aligned 128-bit loads and a word funnel.  The alignment is a property of
// the block (one warp per candidate), so the branches are warp-uniform.  May touch up to 31 bytes past the last byte
// wanted: picture rows carry that slack (hvb_picture_create).
// bytes sh .. sh + 15 (sh = 1..15) of the 32 bytes (lo, hi)
__device__ __forceinline__ uint4 funnel16(const uint4 &lo, const uint4 &hi, unsigned sh)
{
    const unsigned b = (sh & 3) * 8;
    uint32_t w0, w1, w2, w3, w4;
    switch (sh >> 2)
    {
    case 0: w0 = lo.x, w1 = lo.y, w2 = lo.z, w3 = lo.w, w4 = hi.x; break;
    case 1: w0 = lo.y, w1 = lo.z, w2 = lo.w, w3 = hi.x, w4 = hi.y; break;
    case 2: w0 = lo.z, w1 = lo.w, w2 = hi.x, w3 = hi.y, w4 = hi.z; break;
    default: w0 = lo.w, w1 = hi.x, w2 = hi.y, w3 = hi.z, w4 = hi.w; break;
    }
    return make_uint4(__funnelshift_r(w0, w1, b), __funnelshift_r(w1, w2, b), __funnelshift_r(w2, w3, b), __funnelshift_r(w3, w4, b));
}

__device__ __forceinline__ uint4 load16(const uint8_t *p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint4 *q = reinterpret_cast<const uint4 *>(a & ~uintptr_t(15));
    const unsigned sh = (unsigned)(a & 15);
    const uint4 lo = __ldg(q);
    if (sh == 0) return lo;
    return funnel16(lo, __ldg(q + 1), sh);
}

// keep the first `valid` (1..16) bytes of a 16-byte chunk
__device__ __forceinline__ uint4 maskChunk(uint4 v, int valid)
{
    if (valid >= 16) return v;
    const auto m = [valid](int k) -> uint32_t {
        const int b = valid - 4 * k;
        return b >= 4 ? 0xffffffffu : (b <= 0 ? 0u : (1u << (8 * b)) - 1u);
    };
    return make_uint4(v.x & m(0), v.y & m(1), v.z & m(2), v.w & m(3));
}

template <typename Sample>
__device__ __forceinline__ unsigned sadWords(uint4 a, uint4 b)
{
    if (sizeof(Sample) == 1) return __vsadu4(a.x, b.x) + __vsadu4(a.y, b.y) + __vsadu4(a.z, b.z) + __vsadu4(a.w, b.w);
    return __vsadu2(a.x, b.x) + __vsadu2(a.y, b.y) + __vsadu2(a.z, b.z) + __vsadu2(a.w, b.w);
}

template <typename Sample>
__device__ __forceinline__ unsigned ssdWords(uint4 a, uint4 b, unsigned acc)
{
    const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        if (sizeof(Sample) == 1)
        {
            const uint32_t d = __vabsdiffu4(wa[k], wb[k]);
            acc = __dp4a(d, d, acc);
        }
        else
        {
            const uint32_t d = __vabsdiffu2(wa[k], wb[k]);
            const uint32_t l = d & 0xffffu, h = d >> 16;
            acc += l * l + h * h;
        }
    }
    return acc;
}

// Streaming walk of a block pair as 16-byte chunks, `consume(chunk of a, chunk of b)` per chunk.  Rows of 4 or 8 whole
// chunks (64- and 128-byte rows; narrower blocks measured slower here than in the general loop -- B200 call 11: 32x32 8-bit
// 51 % / 31 % of the HBM peak co-located / as a search candidate against 60 % / 49 %) with the first operand on the 16-byte
// grid -- a source block always is -- take one of two lean loops in which the warp's 32 lanes cover 32 / cpr rows per turn,
// a lane keeps its column, and the loads of several turns are requested before the first is consumed (two loads in
// flight per lane do not cover the HBM latency at full bandwidth, and address arithmetic between the loads costs more than
// the loads):
//   second operand on the grid too (co-located blocks)        three turns in flight, six 128-bit loads per lane;
//   second operand anywhere (a motion-search candidate)       a chunk is the five aligned words that contain it and four
//                                                             funnel shifts; two turns in flight.
// Everything else (partial chunks, AMP widths, a source off the grid) returns false and takes the caller's general loop.
template <typename F>
__device__ __forceinline__ bool alignedChunks(const uint8_t *pa, intptr_t pitchA, const uint8_t *pb, intptr_t pitchB, int wb, int h, int lane, F consume)
{
    const int cpr = wb >> 4;
    if (((reinterpret_cast<uintptr_t>(pa) | (uintptr_t)pitchA | (uintptr_t)pitchB | (uintptr_t)wb) & 15) || (cpr != 4 && cpr != 8))
        return false;
    const int shift = __ffs(cpr) - 1, rowsPerTurn = 32 >> shift;
    const int x = (lane & (cpr - 1)) << 4, y0 = lane >> shift;
    const uint8_t *qa = pa + (intptr_t)y0 * pitchA + x, *qb = pb + (intptr_t)y0 * pitchB + x;
    const intptr_t stepA = rowsPerTurn * pitchA, stepB = rowsPerTurn * pitchB;
    const unsigned off = (unsigned)(reinterpret_cast<uintptr_t>(pb) & 15);
    if (off)
    {
        // the five aligned words that contain the chunk, funnel-shifted by the byte part of the offset: no word selection
        // (a switch over the word part costs the compiler four live variants of every chunk)
        const unsigned bits = (off & 3) * 8;
        qb -= off & 3;
#pragma unroll 1
        for (int y = y0; y < h; y += 2 * rowsPerTurn)
        {
            uint4 va[2];
            uint32_t w[2][5];
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (y + k * rowsPerTurn < h)
                {
                    va[k] = __ldg(reinterpret_cast<const uint4 *>(qa + k * stepA));
                    const uint32_t *q = reinterpret_cast<const uint32_t *>(qb + k * stepB);
#pragma unroll
                    for (int i = 0; i < 5; ++i) w[k][i] = __ldg(q + i);
                }
#pragma unroll
            for (int k = 0; k < 2; ++k)
                if (y + k * rowsPerTurn < h)
                    consume(va[k], make_uint4(__funnelshift_r(w[k][0], w[k][1], bits), __funnelshift_r(w[k][1], w[k][2], bits),
                                              __funnelshift_r(w[k][2], w[k][3], bits), __funnelshift_r(w[k][3], w[k][4], bits)));
            qa += 2 * stepA;
            qb += 2 * stepB;
        }
        return true;
    }
#pragma unroll 1
    for (int y = y0; y < h; y += 3 * rowsPerTurn)
    {
        uint4 va[3], vb[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (y + k * rowsPerTurn < h)
            {
                va[k] = __ldg(reinterpret_cast<const uint4 *>(qa + k * stepA));
                vb[k] = __ldg(reinterpret_cast<const uint4 *>(qb + k * stepB));
            }
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (y + k * rowsPerTurn < h) consume(va[k], vb[k]);
        qa += 3 * stepA;
        qb += 3 * stepB;
    }
    return true;
}

// Streaming form shared by SAD, SAD4 and SSD: a block is h rows of ceil(w * B / 16) 16-byte chunks, the chunks go
// round the lanes, every operand chunk is one or two 128-bit loads whatever the block's alignment (co-located blocks,
// motion-search candidates, 8- and 16-bit samples alike); the excess bytes of a row's last chunk are masked in both
// operands, so they contribute |0 - 0|.
template <typename Sample>
__device__ __forceinline__ int sadBlock(const Sample *a, int sa, const Sample *b, int sb, int w, int h, int lane)
{
    const int B = (int)sizeof(Sample), wb = w * B, cpr = (wb + 15) >> 4, total = cpr * h;
    const uint8_t *pa = reinterpret_cast<const uint8_t *>(a), *pb = reinterpret_cast<const uint8_t *>(b);
    unsigned acc = 0;
    if (alignedChunks(pa, (intptr_t)sa * B, pb, (intptr_t)sb * B, wb, h, lane, [&](const uint4 &va, const uint4 &vb) { acc += sadWords<Sample>(va, vb); }))
        return (int)acc;
#pragma unroll 2
    for (int i = lane; i < total; i += 32)
    {
        const int y = i / cpr, x = (i - y * cpr) << 4;
        const uint4 va = maskChunk(load16(pa + (intptr_t)y * sa * B + x), wb - x);
        const uint4 vb = maskChunk(load16(pb + (intptr_t)y * sb * B + x), wb - x);
        acc += sadWords<Sample>(va, vb);
    }
    return (int)acc;
}

template <typename Sample>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 3)
    sadKernel(const HvbPlane *__restrict__ planes, const hvb_metric_task *__restrict__ tasks, int n, int32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int warpsTotal = gridDim.x * kWarpsPerBlock;
    int t = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (t >= n) return;
    hvb_metric_task task = tasks[t];
    for (;;)
    {
        // the next block's task is requested while this block streams: one memory latency per block instead of two
        const int tn = t + warpsTotal;
        hvb_metric_task next = task;
        if (tn < n) next = tasks[tn];
        int sa, sb;
        const Sample *a = hvbBlockPtr<Sample>(planes, task.a, sa);
        const Sample *b = hvbBlockPtr<Sample>(planes, task.b, sb);
        int acc = hvbWarpSum(sadBlock<Sample>(a, sa, b, sb, task.w, task.h, lane));
        if (sizeof(Sample) == 2) acc >>= 2;
        if (lane == 0) out[t] = acc;
        if (tn >= n) break;
        task = next;
        t = tn;
    }
}

template <typename Sample>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    sad4Kernel(const HvbPlane *__restrict__ planes, const hvb_sad4_task *__restrict__ tasks, int n, int32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int warpsTotal = gridDim.x * kWarpsPerBlock;
    const int B = (int)sizeof(Sample);
    for (int t = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); t < n; t += warpsTotal)
    {
        const hvb_sad4_task task = tasks[t];
        int ss;
        const uint8_t *src = reinterpret_cast<const uint8_t *>(hvbBlockPtr<Sample>(planes, task.src, ss));
        const HvbPlane &rp = planes[task.ref_pic * 3 + task.ref_cIdx];
        const uint8_t *rbase = reinterpret_cast<const uint8_t *>(rp.base);
        const int sr = rp.stride;
        const int wb = task.w * B, h = task.h, cpr = (wb + 15) >> 4, total = cpr * h;
        unsigned acc[4] = {0, 0, 0, 0};
        const uint8_t *ref[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) ref[k] = rbase + ((intptr_t)task.ry[k] * sr + task.rx[k]) * B;
        // the source chunk is loaded once and compared with the four references (havoc/sad.cpp:513-542)
        for (int i = lane; i < total; i += 32)
        {
            const int y = i / cpr, x = (i - y * cpr) << 4;
            const uint4 vs = maskChunk(load16(src + (intptr_t)y * ss * B + x), wb - x);
            uint4 vr[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) vr[k] = load16(ref[k] + (intptr_t)y * sr * B + x);
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[k] += sadWords<Sample>(vs, maskChunk(vr[k], wb - x));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            int v = hvbWarpSum((int)acc[k]);
            if (sizeof(Sample) == 2) v >>= 2;
            if (lane == 0) out[t * 4 + k] = v;
        }
    }
}

template <typename Sample>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 3)
    ssdKernel(const HvbPlane *__restrict__ planes, const hvb_metric_task *__restrict__ tasks, int n, uint32_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int warpsTotal = gridDim.x * kWarpsPerBlock;
    const int B = (int)sizeof(Sample);
    int t = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (t >= n) return;
    hvb_metric_task task = tasks[t];
    for (;;)
    {
        const int tn = t + warpsTotal; // (the next block's task is requested while this block streams)
        hvb_metric_task next = task;
        if (tn < n) next = tasks[tn];
        int sa, sb;
        const uint8_t *pa = reinterpret_cast<const uint8_t *>(hvbBlockPtr<Sample>(planes, task.a, sa));
        const uint8_t *pb = reinterpret_cast<const uint8_t *>(hvbBlockPtr<Sample>(planes, task.b, sb));
        const int wb = task.w * B, h = task.h, cpr = (wb + 15) >> 4, total = cpr * h;
        unsigned acc = 0; // modulo 2^32, like the reference's uint32_t accumulator (havoc/ssd.cpp:28-43)
        if (!alignedChunks(pa, (intptr_t)sa * B, pb, (intptr_t)sb * B, wb, h, lane, [&](const uint4 &va, const uint4 &vb) { acc = ssdWords<Sample>(va, vb, acc); }))
        {
#pragma unroll 2
            for (int i = lane; i < total; i += 32)
            {
                const int y = i / cpr, x = (i - y * cpr) << 4;
                const uint4 va = maskChunk(load16(pa + (intptr_t)y * sa * B + x), wb - x);
                const uint4 vb = maskChunk(load16(pb + (intptr_t)y * sb * B + x), wb - x);
                acc = ssdWords<Sample>(va, vb, acc);
            }
        }
        acc = hvbWarpSumU(acc);
        if (sizeof(Sample) == 2) acc >>= 4;
        if (lane == 0) out[t] = acc;
        if (tn >= n) break;
        task = next;
        t = tn;
    }
}

} // namespace

namespace {

using hvb_unit::imma16832;

// ---- 8-bit SATD of 8x8-tiled blocks on the integer tensor cores ---------------------------------------------------
// sum |H d H^T| over a tile = sum |(H (x) H) vec(a) - (H (x) H) vec(b)|: eight tiles at a time are one
// [H | -H] x [a ; b] product (IMMA m16n8k32, s8 x u8 -> s32, M = 64, K = 64 + 64), the B fragments loaded straight
// from the two pictures (a lane reads 4 bytes of a tile row; the eight tiles of a group are neighbours in the block,
// so a load instruction covers whole 32-byte sectors), the A fragments four per-lane constants (-1)^popc(m & k).
// See hvb_me_subpel.cu for the derivation; ~7 instructions per tile instead of ~16 for the register butterfly.
// Here the K index is laid out so that a lane's two B registers of a k-step are one whole tile row: position bits
// [row | col] = [ks, t1, t0 | reg, j1, j0], i.e. lane (g, t) loads rows t and t + 4 of tile g with one 64-bit load each.
// A register (m-tile mt, k-step ks, reg r) is then x[mt & 1][r], negated when ((mt >> 1) & ks) ^ (ks >> 1) is odd.
struct HadamardFrag
{
    uint32_t x[2][4];
    __device__ __forceinline__ explicit HadamardFrag(int lane)
    {
        const int g = lane >> 2, t = lane & 3, t0 = t & 1, t1 = t >> 1, g2 = g >> 2;
        const uint32_t pat = (g & 2) ? ((g & 1) ? 0x01ffff01u : 0xffff0101u) : ((g & 1) ? 0xff01ff01u : 0x01010101u);
#pragma unroll
        for (int m0 = 0; m0 < 2; ++m0)
#pragma unroll
            for (int r = 0; r < 4; ++r) x[m0][r] = (((r & 1) & t0) ^ ((r >> 1) & g2) ^ (m0 & t1)) ? pat ^ 0xfefefefeu : pat;
    }
};

// the k-steps of m-tile mt in the order the streaming kernels run them: phase 0 = the two whose products enter negated
// (((mt >> 1) & ks) ^ (ks >> 1) odd: ks in {2, 3} for mt < 2, {1, 2} otherwise), phase 1 = the other two
__device__ __forceinline__ constexpr int satdStep(int mt, int phase, int j)
{
    return mt < 2 ? (phase == 0 ? 2 + j : j) : (phase == 0 ? 1 + j : 3 * j);
}

// 8 bytes of a tile row: one 64-bit load when the row is 8-byte aligned, else three aligned words and two funnel shifts
__device__ __forceinline__ uint2 loadRow8(const uint8_t *p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    if ((a & 7) == 0) return __ldg(reinterpret_cast<const uint2 *>(p));
    const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
    const unsigned sh = (unsigned)(a & 3) * 8;
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1);
    if (sh == 0) return make_uint2(w0, w1);
    const uint32_t w2 = __ldg(q + 2);
    return make_uint2(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh));
}

// ---- the streaming form: tile rows staged through shared memory by asynchronous copies ---------------------------
// A warp treats the tile groups of all its tasks as ONE stream and keeps STAGES - 1 groups (1 KB each: rows t and
// t + 4 of tile g from both pictures, 8 bytes per copy, a lane's four copies land in slots only that lane reads, so no
// warp synchronisation is needed -- cp.async.wait_group orders a lane's own copies) in flight ahead of the group whose
// products run.  Bytes in flight no longer cost registers, and the pipeline does not drain at block boundaries (a
// 32x32 block is two groups).  Rows that are not 8-byte aligned take the register path (loadRow8) into the same slots.
// Only the issuing side walks the task array; what the consuming side needs of a group in flight -- its task and its
// place in the block -- travels in a register queue as deep as the pipeline.
constexpr int kSmallTiles = 8; // blocks of up to this many 8x8 tiles share tile groups with their neighbours in the task list

struct SatdCursor
{
    int t, base, tiles, tilesX, sa, sb;
    unsigned recip; // ceil(2^32 / tilesX): umulhi(tile, recip) = tile / tilesX for tile < 2^16, tilesX > 1
    const uint8_t *a, *b;
    // blocks 1, 2, 4 or 8 tiles wide made of whole groups: a group is 8 / tilesX tile rows, so a lane's rows of the next group
    // are its rows of this one plus a constant -- no tile arithmetic after a block's first group
    bool fast;
    int dS, dP;          // bytes from a group to the next
    const uint8_t *S, *P; // this lane's row t of its tile in the NEXT group to be requested (valid once base > 0)
};

__device__ __forceinline__ void satdCursorFast(SatdCursor &c)
{
    c.fast = c.tilesX <= 8 && !(c.tilesX & (c.tilesX - 1)) && !(c.tiles & 7);
    c.dS = (64 / c.tilesX) * c.sa;
    c.dP = (64 / c.tilesX) * c.sb;
}

// position the cursor on the first task at or after c.t (stepping by `step`) that this kernel owns
__device__ __forceinline__ void satdCursorOpen(SatdCursor &c, const HvbPlane *__restrict__ planes, const hvb_metric_task *__restrict__ tasks,
                                               int n, int step, int *__restrict__ leftover)
{
    while (c.t < n)
    {
        const hvb_metric_task task = tasks[c.t];
        if (!((task.w | task.h) & 7) && (task.w >> 3) * (task.h >> 3) <= kSmallTiles)
            leftover[2] = 1; // a block for satdMmaSmallKernel
        else if (!((task.w | task.h) & 7))
        {
            c.a = hvbBlockPtr<uint8_t>(planes, task.a, c.sa);
            c.b = hvbBlockPtr<uint8_t>(planes, task.b, c.sb);
            c.tilesX = task.w >> 3;
            c.recip = c.tilesX > 1 ? 0xffffffffu / (unsigned)c.tilesX + 1u : 0u;
            c.tiles = c.tilesX * (task.h >> 3);
            c.base = 0;
            satdCursorFast(c);
            return;
        }
        else
            leftover[0] = 1; // a block for satdKernel (every lane stores the same value)
        c.t += step;
    }
}

__device__ __forceinline__ void cpAsync8(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpAsync4(uint32_t dst, const void *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int PENDING>
__device__ __forceinline__ void cpAsyncWait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
}

// 8-bit blocks whose sides are multiples of 8 (the other blocks of the batch belong to satdKernel)
template <int STAGES>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
    satdMmaKernel(const HvbPlane *__restrict__ planes, const hvb_metric_task *__restrict__ tasks, int n, int32_t *__restrict__ out,
                  int *__restrict__ leftover)
{
    // per warp and stage: [part 0..3][lane] 8-byte row slots (1 KB), then [part 2..3][lane] the third word of a prediction row
    // that is not 8-byte aligned (256 B)
    extern __shared__ __align__(16) uint2 satdStage[];
    constexpr int kStageWords = 128 + 32; // in uint2
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int step = gridDim.x * kWarpsPerBlock;
    const HadamardFrag A(lane);
    uint2 *mine = satdStage + warp * STAGES * kStageWords + lane;
    const uint32_t mineAddr = (uint32_t)__cvta_generic_to_shared(mine);
    const uint32_t *mineThird = reinterpret_cast<const uint32_t *>(satdStage + warp * STAGES * kStageWords + 128) + lane;
    const uint32_t thirdAddr = (uint32_t)__cvta_generic_to_shared(mineThird);

    SatdCursor issue;
    issue.t = blockIdx.x * kWarpsPerBlock + warp;
    satdCursorOpen(issue, planes, tasks, n, step, leftover);

    // requests one group into `slot`; (qt, qb) = its task (>= n: none left) and the tiles its block has left, this group
    // included; qs = the bit shift that realigns this lane's prediction rows when they are read back.
    // The source block of a search sits on a PU boundary, the prediction anywhere: an 8-byte aligned row is one 8-byte
    // asynchronous copy; any other prediction row is copied as the three aligned words that contain it (still asynchronous:
    // nothing waits at issue) and funnel-shifted into place by the consuming side.
    auto issueGroup = [&](int slot, int &qt, int &qb, unsigned &qs) {
        qt = issue.t;
        qb = 0;
        qs = 0;
        if (issue.t < n)
        {
            qb = issue.tiles - issue.base;
            const uint8_t *S, *P;
            if (issue.fast && issue.base > 0)
                S = issue.S, P = issue.P;
            else
            {
                const int tile = min(issue.base + g, issue.tiles - 1);
                const int ty = issue.tilesX == 1 ? tile : issue.tiles < 65536 ? (int)__umulhi((unsigned)tile, issue.recip) : tile / issue.tilesX;
                const int tx = tile - ty * issue.tilesX;
                S = issue.a + (intptr_t)(ty * 8 + t) * issue.sa + tx * 8;
                P = issue.b + (intptr_t)(ty * 8 + t) * issue.sb + tx * 8;
            }
            issue.S = S + issue.dS;
            issue.P = P + issue.dP;
            const uint32_t dst = mineAddr + slot * (kStageWords * 8);
            if (!((reinterpret_cast<uintptr_t>(S) | (uintptr_t)issue.sa) & 7))
            {
                cpAsync8(dst, S);
                cpAsync8(dst + 256, S + 4 * issue.sa);
            }
            else
            {
                uint2 *d = mine + slot * kStageWords;
                d[0] = loadRow8(S);
                d[32] = loadRow8(S + 4 * issue.sa);
            }
            if (!((reinterpret_cast<uintptr_t>(P) | (uintptr_t)issue.sb) & 7))
            {
                cpAsync8(dst + 512, P);
                cpAsync8(dst + 768, P + 4 * issue.sb);
            }
            else
            {
                // (a stride that is no multiple of 4 cannot occur: plane pitches are multiples of 256 bytes)
                const uintptr_t a = reinterpret_cast<uintptr_t>(P);
                const uint8_t *q = reinterpret_cast<const uint8_t *>(a & ~uintptr_t(3));
                qs = (unsigned)(a & 3) * 8;
                const uint32_t third = thirdAddr + slot * (kStageWords * 8);
                cpAsync4(dst + 512, q);
                cpAsync4(dst + 516, q + 4);
                cpAsync4(third, q + 8);
                q += 4 * (intptr_t)issue.sb;
                cpAsync4(dst + 768, q);
                cpAsync4(dst + 772, q + 4);
                cpAsync4(third + 128, q + 8);
            }
            issue.base += 8;
            if (issue.base >= issue.tiles)
            {
                issue.t += step;
                satdCursorOpen(issue, planes, tasks, n, step, leftover);
            }
        }
        cpAsyncCommit();
    };

    int qt[STAGES - 1], qb[STAGES - 1]; // the groups in flight, oldest first
    unsigned qs[STAGES - 1];
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issueGroup(s, qt[s], qb[s], qs[s]);
    int slot = 0, total = 0;
    while (qt[0] < n)
    {
        int nt, nb;
        unsigned ns;
        issueGroup(slot == 0 ? STAGES - 1 : slot - 1, nt, nb, ns);
        cpAsyncWait<STAGES - 1>();
        const uint2 *src = mine + slot * kStageWords;
        const uint2 r0 = src[0], r1 = src[32];
        uint2 r2 = src[64], r3 = src[96];
        if (qs[0])
        {
            // prediction rows that arrived as three aligned words: bytes (shift / 8) .. (shift / 8) + 7 of them
            const uint32_t *third = mineThird + slot * (kStageWords * 2);
            const uint32_t w2 = third[0], w3 = third[32];
            r2 = make_uint2(__funnelshift_r(r2.x, r2.y, qs[0]), __funnelshift_r(r2.y, w2, qs[0]));
            r3 = make_uint2(__funnelshift_r(r3.x, r3.y, qs[0]), __funnelshift_r(r3.y, w3, qs[0]));
        }
        const uint32_t bx[4] = {r0.x, r1.x, r2.x, r3.x}, by[4] = {r0.y, r1.y, r2.y, r3.y};
        // the four m-tiles' products advance together, k-step by k-step: four independent accumulator chains, so that an
        // IMMA is followed by three that do not wait for it (the ncu capture of the mt-outer order showed `wait` on the
        // dependent chains as the top stall, profiles/r01g_summary.txt)
        // The A registers of (m-tile mt, k-step ks) are x[mt & 1], negated when ((mt >> 1) & ks) ^ (ks >> 1) is odd.  Instead
        // of negating four registers per product, the k-steps whose products enter negated accumulate in one register set and
        // the others in a second one, and the coefficient's absolute value |P - N| is one absolute-difference-and-add.  Two
        // m-tiles at a time: four independent accumulator chains (an IMMA is followed by three that do not wait for it) in
        // the sixteen registers one set of all four m-tiles would take.
        int s0 = 0, s1 = 0;
#pragma unroll
        for (int pair = 0; pair < 2; ++pair)
        {
            int accN[2][4] = {}, accP[2][4] = {};
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int m = 0; m < 2; ++m)
                {
                    const int mt = 2 * pair + m, kn = satdStep(mt, 0, j), kp = satdStep(mt, 1, j);
                    imma16832(accN[m], A.x[m][0], A.x[m][1], A.x[m][2], A.x[m][3], bx[kn], by[kn]);
                    imma16832(accP[m], A.x[m][0], A.x[m][1], A.x[m][2], A.x[m][3], bx[kp], by[kp]);
                }
#pragma unroll
            for (int m = 0; m < 2; ++m)
            {
                s0 = __sad(accP[m][0], accN[m][0], __sad(accP[m][2], accN[m][2], (unsigned)s0));
                s1 = __sad(accP[m][1], accN[m][1], __sad(accP[m][3], accN[m][3], (unsigned)s1));
            }
        }
        // column 2t + (g & 1) of the group, summed over the 8 lanes that share t
        int sum = (g & 1) ? s1 : s0;
        sum += __shfl_xor_sync(0xffffffffu, (g & 1) ? s0 : s1, 4);
        sum += __shfl_xor_sync(0xffffffffu, sum, 8);
        sum += __shfl_xor_sync(0xffffffffu, sum, 16);
        if (g < 2 && 2 * t + g < qb[0]) total += (sum + 2) >> 2; // havoc/hadamard.cpp:319-323
        if (qb[0] <= 8)
        {
            total = hvbWarpSum(total);
            if (lane == 0) out[qt[0]] = total;
            total = 0;
        }
#pragma unroll
        for (int s = 0; s + 1 < STAGES - 1; ++s)
        {
            qt[s] = qt[s + 1];
            qb[s] = qb[s + 1];
            qs[s] = qs[s + 1];
        }
        qt[STAGES - 2] = nt;
        qb[STAGES - 2] = nb;
        qs[STAGES - 2] = ns;
        slot = slot == STAGES - 1 ? 0 : slot + 1;
    }
    cpAsyncWait<0>();
}

// ---- 16-bit samples: the same stream, the products on the samples' byte planes ---------------------------------------
// A tile row is 16 bytes.  Lane (g, t) requests rows t and t + 4 of tile g of both pictures: one 16-byte asynchronous copy
// where the row is 16-byte aligned (the source block of a search always is), otherwise the five aligned words that contain
// it, realigned by the consuming side.  The rows are split into their low-byte and high-byte planes (two byte permutes per
// 8 bytes); the high plane's products run first, the accumulators are scaled by 256 and the low plane's products continue
// in the same registers: sum_k H[m][k] (256 hi_k + lo_k), exact in 32 bits, so the absolute value is taken of the
// reference's coefficient.  Blocks of up to kSmallTiles tiles stay with satdKernel (two per warp).
__device__ __forceinline__ void cpAsync16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__device__ __forceinline__ void satdCursorOpen16(SatdCursor &c, const HvbPlane *__restrict__ planes, const hvb_metric_task *__restrict__ tasks,
                                                 int n, int step, int *__restrict__ leftover)
{
    while (c.t < n)
    {
        const hvb_metric_task task = tasks[c.t];
        if (!((task.w | task.h) & 7) && (task.w >> 3) * (task.h >> 3) > kSmallTiles)
        {
            int sa, sb;
            c.a = reinterpret_cast<const uint8_t *>(hvbBlockPtr<uint16_t>(planes, task.a, sa));
            c.b = reinterpret_cast<const uint8_t *>(hvbBlockPtr<uint16_t>(planes, task.b, sb));
            c.sa = 2 * sa, c.sb = 2 * sb; // in bytes
            c.tilesX = task.w >> 3;
            c.recip = c.tilesX > 1 ? 0xffffffffu / (unsigned)c.tilesX + 1u : 0u;
            c.tiles = c.tilesX * (task.h >> 3);
            c.base = 0;
            satdCursorFast(c);
            return;
        }
        leftover[0] = 1; // a block for satdKernel (every lane stores the same value)
        c.t += step;
    }
}

// 16 bytes of a 16-bit tile row at any 2-byte aligned address, synchronously (source rows off the 16-byte grid: not the
// case for the blocks of a search)
__device__ __forceinline__ uint4 loadRow16(const uint8_t *p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    if ((a & 15) == 0) return __ldg(reinterpret_cast<const uint4 *>(p));
    const uint32_t *q = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
    const unsigned sh = (unsigned)(a & 3) * 8;
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = __ldg(q + 3);
    if (sh == 0) return make_uint4(w0, w1, w2, w3);
    const uint32_t w4 = __ldg(q + 4);
    return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}

template <int STAGES>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 3)
    satdMma16Kernel(const HvbPlane *__restrict__ planes, const hvb_metric_task *__restrict__ tasks, int n, int32_t *__restrict__ out,
                    int *__restrict__ leftover)
{
    // per warp and stage: [part 0..3][lane] 16-byte row slots (2 KB), then [part 2..3][lane] the fifth word of a prediction
    // row that is not 16-byte aligned (256 B)
    extern __shared__ __align__(16) uint4 satdStage16[];
    constexpr int kStageQuads = 128 + 16; // in uint4
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int step = gridDim.x * kWarpsPerBlock;
    const HadamardFrag A(lane);
    uint4 *mine = satdStage16 + warp * STAGES * kStageQuads + lane;
    const uint32_t mineAddr = (uint32_t)__cvta_generic_to_shared(mine);
    const uint32_t *mineFifth = reinterpret_cast<const uint32_t *>(satdStage16 + warp * STAGES * kStageQuads + 128) + lane;
    const uint32_t fifthAddr = (uint32_t)__cvta_generic_to_shared(mineFifth);

    SatdCursor issue;
    issue.t = blockIdx.x * kWarpsPerBlock + warp;
    satdCursorOpen16(issue, planes, tasks, n, step, leftover);

    auto issueGroup = [&](int slot, int &qt, int &qb, unsigned &qs) {
        qt = issue.t;
        qb = 0;
        qs = 0;
        if (issue.t < n)
        {
            qb = issue.tiles - issue.base;
            const uint8_t *S, *P;
            if (issue.fast && issue.base > 0)
                S = issue.S, P = issue.P;
            else
            {
                const int tile = min(issue.base + g, issue.tiles - 1);
                const int ty = issue.tilesX == 1 ? tile : issue.tiles < 65536 ? (int)__umulhi((unsigned)tile, issue.recip) : tile / issue.tilesX;
                const int tx = tile - ty * issue.tilesX;
                S = issue.a + (intptr_t)(ty * 8 + t) * issue.sa + tx * 16;
                P = issue.b + (intptr_t)(ty * 8 + t) * issue.sb + tx * 16;
            }
            issue.S = S + issue.dS;
            issue.P = P + issue.dP;
            const uint32_t dst = mineAddr + slot * (kStageQuads * 16);
            if (!((reinterpret_cast<uintptr_t>(S) | (uintptr_t)issue.sa) & 15))
            {
                cpAsync16(dst, S);
                cpAsync16(dst + 512, S + 4 * issue.sa);
            }
            else
            {
                uint4 *d = mine + slot * kStageQuads;
                d[0] = loadRow16(S);
                d[32] = loadRow16(S + 4 * issue.sa);
            }
            if (!((reinterpret_cast<uintptr_t>(P) | (uintptr_t)issue.sb) & 15))
            {
                cpAsync16(dst + 1024, P);
                cpAsync16(dst + 1536, P + 4 * issue.sb);
            }
            else
            {
                const uintptr_t a = reinterpret_cast<uintptr_t>(P);
                const uint8_t *q = reinterpret_cast<const uint8_t *>(a & ~uintptr_t(3));
                qs = (unsigned)(a & 3) * 8; // 0 or 16
                const uint32_t fifth = fifthAddr + slot * (kStageQuads * 16);
#pragma unroll
                for (int r = 0; r < 2; ++r)
                {
                    const uint32_t d = dst + 1024 + 512 * r;
#pragma unroll
                    for (int i = 0; i < 4; ++i) cpAsync4(d + 4 * i, q + 4 * i);
                    cpAsync4(fifth + 128 * r, q + 16);
                    q += 4 * (intptr_t)issue.sb;
                }
            }
            issue.base += 8;
            if (issue.base >= issue.tiles)
            {
                issue.t += step;
                satdCursorOpen16(issue, planes, tasks, n, step, leftover);
            }
        }
        cpAsyncCommit();
    };

    int qt[STAGES - 1], qb[STAGES - 1];
    unsigned qs[STAGES - 1];
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issueGroup(s, qt[s], qb[s], qs[s]);
    int slot = 0, total = 0;
    while (qt[0] < n)
    {
        int nt, nb;
        unsigned ns;
        issueGroup(slot == 0 ? STAGES - 1 : slot - 1, nt, nb, ns);
        cpAsyncWait<STAGES - 1>();
        const uint4 *src = mine + slot * kStageQuads;
        uint4 r[4] = {src[0], src[32], src[64], src[96]};
        if (qs[0])
        {
            const uint32_t *fifth = mineFifth + slot * (kStageQuads * 4);
            const uint32_t w4[2] = {fifth[0], fifth[32]};
#pragma unroll
            for (int i = 0; i < 2; ++i)
            {
                uint4 &v = r[2 + i];
                v = make_uint4(__funnelshift_r(v.x, v.y, qs[0]), __funnelshift_r(v.y, v.z, qs[0]), __funnelshift_r(v.z, v.w, qs[0]),
                               __funnelshift_r(v.w, w4[i], qs[0]));
            }
        }
        // Byte planes of the four rows (high bytes, low bytes), and the order of the products: with N / P the k-steps of an
        // m-tile that enter negated / as they are (satdStep), the accumulators go through
        //   N_hi, change sign, P_hi              = P_hi - N_hi
        //   times -256, N_lo, change sign        = 256 (P_hi - N_hi) - N_lo
        //   P_lo                                 = sum_k H[m][k] (256 hi_k + lo_k), exact in 32 bits
        // so no A register is ever negated.
        uint32_t bxh[4], byh[4], bxl[4], byl[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            bxh[i] = __byte_perm(r[i].x, r[i].y, 0x7531u);
            byh[i] = __byte_perm(r[i].z, r[i].w, 0x7531u);
            bxl[i] = __byte_perm(r[i].x, r[i].y, 0x6420u);
            byl[i] = __byte_perm(r[i].z, r[i].w, 0x6420u);
        }
        int acc[4][4] = {};
#pragma unroll
        for (int step = 0; step < 4; ++step) // N_hi, P_hi, N_lo, P_lo
        {
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
                {
                    const int ks = satdStep(mt, step & 1, j);
                    imma16832(acc[mt], A.x[mt & 1][0], A.x[mt & 1][1], A.x[mt & 1][2], A.x[mt & 1][3], step < 2 ? bxh[ks] : bxl[ks],
                              step < 2 ? byh[ks] : byl[ks]);
                }
            if (step < 3)
            {
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[mt][i] = step == 1 ? acc[mt][i] * -256 : -acc[mt][i];
            }
        }
        int s0 = 0, s1 = 0;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt)
        {
            s0 = __sad(acc[mt][0], 0, __sad(acc[mt][2], 0, (unsigned)s0));
            s1 = __sad(acc[mt][1], 0, __sad(acc[mt][3], 0, (unsigned)s1));
        }
        int sum = (g & 1) ? s1 : s0;
        sum += __shfl_xor_sync(0xffffffffu, (g & 1) ? s0 : s1, 4);
        sum += __shfl_xor_sync(0xffffffffu, sum, 8);
        sum += __shfl_xor_sync(0xffffffffu, sum, 16);
        if (g < 2 && 2 * t + g < qb[0]) total += ((sum + 2) >> 2) >> 2; // havoc/hadamard.cpp:319-323, 16-bit samples: >> 2 per tile
        if (qb[0] <= 8)
        {
            total = hvbWarpSum(total);
            if (lane == 0) out[qt[0]] = total;
            total = 0;
        }
#pragma unroll
        for (int s = 0; s + 1 < STAGES - 1; ++s)
        {
            qt[s] = qt[s + 1];
            qb[s] = qb[s + 1];
            qs[s] = qs[s + 1];
        }
        qt[STAGES - 2] = nt;
        qb[STAGES - 2] = nb;
        qs[STAGES - 2] = ns;
        slot = slot == STAGES - 1 ? 0 : slot + 1;
    }
    cpAsyncWait<0>();
}

// ---- small blocks: tile groups filled across consecutive tasks --------------------------------------------------------
// A block of one to eight tiles (8x8 .. 32x16 / 64x8) would leave most of a group's eight tile slots empty, and a warp per
// block most of the machine idle (an 8x8 batch ran at 3 % of the HBM bandwidth that way).  Here a warp takes 32 consecutive
// tasks, lays their tiles end to end (a prefix sum over the lanes) and walks that sequence eight tiles at a time, so a
// group holds the tiles of up to eight blocks.  Per warp in shared memory: the 32 blocks' addresses, a tile -> block map
// and the blocks' sums (integer atomics: order-independent).  Blocks of more than kSmallTiles tiles belong to the streaming
// kernel above, which notes in leftover[2] whether this kernel has anything to do.
struct SmallDesc
{
    const uint8_t *a, *b;
    int sa, sb, tilesX, first; // first: position of the block's tile 0 in the warp's tile sequence
};

__global__ void __launch_bounds__(kWarpsPerBlock * 32, 4)
    satdMmaSmallKernel(const HvbPlane *__restrict__ planes, const hvb_metric_task *__restrict__ tasks, int n, int32_t *__restrict__ out,
                       const int *__restrict__ leftover)
{
    if (leftover[2] == 0) return;
    __shared__ SmallDesc sDesc[kWarpsPerBlock][32];
    __shared__ uint8_t sMap[kWarpsPerBlock][32 * kSmallTiles];
    __shared__ int sSum[kWarpsPerBlock][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const HadamardFrag A(lane);
    const int chunks = (n + 31) >> 5;
    for (int chunk = blockIdx.x * kWarpsPerBlock + warp; chunk < chunks; chunk += gridDim.x * kWarpsPerBlock)
    {
        const int i = chunk * 32 + lane;
        int tiles = 0;
        SmallDesc d = {nullptr, nullptr, 0, 0, 1, 0};
        if (i < n)
        {
            const hvb_metric_task task = tasks[i];
            const int count = (task.w >> 3) * (task.h >> 3);
            if (!((task.w | task.h) & 7) && count <= kSmallTiles)
            {
                tiles = count;
                d.a = hvbBlockPtr<uint8_t>(planes, task.a, d.sa);
                d.b = hvbBlockPtr<uint8_t>(planes, task.b, d.sb);
                d.tilesX = task.w >> 3;
            }
        }
        int incl = tiles;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue; // nothing of this kernel's in the chunk (the same for every lane)
        d.first = incl - tiles;
        sDesc[warp][lane] = d;
        sSum[warp][lane] = 0;
        for (int k = 0; k < tiles; ++k) sMap[warp][d.first + k] = (uint8_t)lane;
        __syncwarp();
        for (int base = 0; base < total; base += 8)
        {
            const int f = min(base + g, total - 1);
            const SmallDesc &e = sDesc[warp][sMap[warp][f]];
            const int tile = f - e.first, ty = tile / e.tilesX, tx = tile - ty * e.tilesX;
            const uint8_t *S = e.a + (intptr_t)(ty * 8 + t) * e.sa + tx * 8, *P = e.b + (intptr_t)(ty * 8 + t) * e.sb + tx * 8;
            const uint2 r0 = loadRow8(S), r1 = loadRow8(S + 4 * e.sa), r2 = loadRow8(P), r3 = loadRow8(P + 4 * e.sb);
            const uint32_t bx[4] = {r0.x, r1.x, r2.x, r3.x}, by[4] = {r0.y, r1.y, r2.y, r3.y};
            int acc[4][4] = {};
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
                {
                    const uint32_t neg = ((((mt >> 1) & ks) ^ (ks >> 1)) & 1) ? 0xfefefefeu : 0u;
                    imma16832(acc[mt], A.x[mt & 1][0] ^ neg, A.x[mt & 1][1] ^ neg, A.x[mt & 1][2] ^ neg, A.x[mt & 1][3] ^ neg, bx[ks], by[ks]);
                }
            int s0 = 0, s1 = 0;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
            {
                s0 = __sad(acc[mt][0], 0, __sad(acc[mt][2], 0, (unsigned)s0));
                s1 = __sad(acc[mt][1], 0, __sad(acc[mt][3], 0, (unsigned)s1));
            }
            int sum = (g & 1) ? s1 : s0;
            sum += __shfl_xor_sync(0xffffffffu, (g & 1) ? s0 : s1, 4);
            sum += __shfl_xor_sync(0xffffffffu, sum, 8);
            sum += __shfl_xor_sync(0xffffffffu, sum, 16);
            const int mine = base + 2 * t + g; // lane (g < 2, t) holds the tile at this position of the sequence
            if (g < 2 && mine < total) atomicAdd(&sSum[warp][sMap[warp][mine]], (sum + 2) >> 2); // havoc/hadamard.cpp:319-323
        }
        __syncwarp();
        if (tiles) out[i] = sSum[warp][lane];
        __syncwarp(); // the tables are rewritten for the next chunk
    }
}

// one register-resident Hadamard tile per lane: 16-bit samples, and the 4x4 / 2x2 tiled blocks of 8-bit batches
template <typename Sample>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    satdKernel(const HvbPlane *__restrict__ planes, const hvb_metric_task *__restrict__ tasks, int n, int32_t *__restrict__ out,
               const int *__restrict__ leftover)
{
    // 8-bit batches: satdMmaKernel ran first on this stream and noted whether it left any block to this kernel
    if (leftover && *leftover == 0) return;
    const int lane = threadIdx.x & 31;
    const int warpsTotal = gridDim.x * kWarpsPerBlock;
    // a warp takes two consecutive tasks at a time: a half-warp each when both blocks have at most 16 tiles (a 32x32 block of
    // 8x8 tiles would leave half the lanes idle), otherwise the whole warp one block after the other
    for (int t2 = 2 * (blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)); t2 < n; t2 += 2 * warpsTotal)
    {
        const bool pair = t2 + 1 < n;
        const hvb_metric_task task0 = tasks[t2], task1 = pair ? tasks[t2 + 1] : task0;
        const bool split = pair && hvbSatdTiles(task0.w, task0.h) <= 16 && hvbSatdTiles(task1.w, task1.h) <= 16;
        for (int k = 0; k < (split || !pair ? 1 : 2); ++k)
        {
            const int half = lane >> 4;
            const int index = split ? t2 + half : t2 + k;
            const hvb_metric_task task = (split ? half : k) ? task1 : task0;
            // blocks tiled 8x8 belong to the tensor-core kernels: all of them with 8-bit samples (satdMmaKernel, satdMmaSmallKernel),
            // those of more than kSmallTiles tiles with 16-bit samples (satdMma16Kernel)
            const bool skip = ((task.w | task.h) & 7) == 0 && (sizeof(Sample) == 1 || (task.w >> 3) * (task.h >> 3) > kSmallTiles);
            int acc = 0;
            if (!skip)
            {
                int sa, sb;
                const Sample *a = hvbBlockPtr<Sample>(planes, task.a, sa);
                const Sample *b = hvbBlockPtr<Sample>(planes, task.b, sb);
                acc = hvbMeasureSatdLanes<Sample, Sample>(a, sa, b, sb, task.w, task.h, split ? (lane & 15) : lane, split ? 16 : 32,
                                                          sizeof(Sample) == 2 ? 2 : 0);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (!split) acc += __shfl_xor_sync(0xffffffffu, acc, 16);
            if ((lane & (split ? 15 : 31)) == 0 && !skip) out[index] = acc;
        }
    }
}

template <typename Task>
int gridFor(hvb_context *ctx, int n)
{
    const int blocks = (n + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int cap = ctx->smCount * 8; // 8 resident 256-thread CTAs per SM
    return blocks < cap ? blocks : cap;
}

} // namespace

#define HVB_DISPATCH_SAMPLE(ctx, kernel, grid, ...)                                                      \
    do                                                                                                   \
    {                                                                                                    \
        if ((ctx)->bps == 1)                                                                             \
            kernel<uint8_t><<<(grid), kWarpsPerBlock * 32, 0, (ctx)->stream>>>(__VA_ARGS__);             \
        else                                                                                             \
            kernel<uint16_t><<<(grid), kWarpsPerBlock * 32, 0, (ctx)->stream>>>(__VA_ARGS__);           \
    } while (0)

extern "C" int hvb_sad_batch(hvb_context *ctx, const hvb_metric_task *tasks, int n, int32_t *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(int32_t) * n, mem, &st);
    if (rc) return rc;
#ifdef __CUDACC__
    if (ctx->useTma) // blocks staged by the TMA (hvb_metrics_tma.cu); an experiment until it wins everywhere: HVB_TMA=1
    {
        rc = hvbLaunchSadTma(ctx, st.dTasks, n, static_cast<int32_t *>(st.dOut), 1);
        if (rc) return rc;
        return hvbStageOut(ctx, out, sizeof(int32_t) * n, mem, st);
    }
#endif
    HVB_DISPATCH_SAMPLE(ctx, sadKernel, gridFor<hvb_metric_task>(ctx, n), ctx->dPlanes,
                        static_cast<const hvb_metric_task *>(st.dTasks), n, static_cast<int32_t *>(st.dOut));
    HVB_LAUNCH_CHECK(ctx, "sadKernel");
    return hvbStageOut(ctx, out, sizeof(int32_t) * n, mem, st);
}

extern "C" int hvb_sad4_batch(hvb_context *ctx, const hvb_sad4_task *tasks, int n, int32_t *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(int32_t) * 4 * n, mem, &st);
    if (rc) return rc;
#ifdef __CUDACC__
    if (ctx->useTma)
    {
        rc = hvbLaunchSadTma(ctx, st.dTasks, n, static_cast<int32_t *>(st.dOut), 4);
        if (rc) return rc;
        return hvbStageOut(ctx, out, sizeof(int32_t) * 4 * n, mem, st);
    }
#endif
    HVB_DISPATCH_SAMPLE(ctx, sad4Kernel, gridFor<hvb_sad4_task>(ctx, n), ctx->dPlanes,
                        static_cast<const hvb_sad4_task *>(st.dTasks), n, static_cast<int32_t *>(st.dOut));
    HVB_LAUNCH_CHECK(ctx, "sad4Kernel");
    return hvbStageOut(ctx, out, sizeof(int32_t) * 4 * n, mem, st);
}

extern "C" int hvb_ssd_batch(hvb_context *ctx, const hvb_metric_task *tasks, int n, uint32_t *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(uint32_t) * n, mem, &st);
    if (rc) return rc;
    HVB_DISPATCH_SAMPLE(ctx, ssdKernel, gridFor<hvb_metric_task>(ctx, n), ctx->dPlanes,
                        static_cast<const hvb_metric_task *>(st.dTasks), n, static_cast<uint32_t *>(st.dOut));
    HVB_LAUNCH_CHECK(ctx, "ssdKernel");
    return hvbStageOut(ctx, out, sizeof(uint32_t) * n, mem, st);
}

extern "C" int hvb_satd_batch(hvb_context *ctx, const hvb_metric_task *tasks, int n, int32_t *out, hvb_mem mem)
{
    HVB_CHECK_ARGS(ctx, n >= 0 && (n == 0 || (tasks && out)));
    if (!n) return HVB_OK;
    cudaSetDevice(ctx->device);
    HvbStaged st;
    int rc = hvbStageIn(ctx, tasks, sizeof(*tasks) * n, out, sizeof(int32_t) * n, mem, &st);
    if (rc) return rc;
    int *leftover = nullptr;
    if (ctx->bps == 1)
    {
        const auto *dT = static_cast<const hvb_metric_task *>(st.dTasks);
        auto *dO = static_cast<int32_t *>(st.dOut);
        leftover = ctx->workCursors + 2; // [0]: blocks for satdKernel, [2]: blocks for satdMmaSmallKernel
        cudaMemsetAsync(leftover, 0, 3 * sizeof(int), ctx->stream);
        // two stages measured best (64.5 % of HBM peak at 64x64 against 64.1 % with three and 62.3 % with four: the
        // products, not the copies, bound the kernel, and a deeper queue costs registers and shared memory)
        constexpr int kStages = 2;
        const int smem = kWarpsPerBlock * kStages * 1280; // per warp and stage: 1 KB of row slots + 256 B of third words
        int perSm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, satdMmaKernel<kStages>, kWarpsPerBlock * 32, smem);
        const int blocks = min((n + kWarpsPerBlock - 1) / kWarpsPerBlock, ctx->smCount * max(perSm, 1));
        satdMmaKernel<kStages><<<blocks, kWarpsPerBlock * 32, smem, ctx->stream>>>(ctx->dPlanes, dT, n, dO, leftover);
        HVB_LAUNCH_CHECK(ctx, "satdMmaKernel");
        const int smallBlocks = min((n + 32 * kWarpsPerBlock - 1) / (32 * kWarpsPerBlock), ctx->smCount * 4);
        satdMmaSmallKernel<<<smallBlocks, kWarpsPerBlock * 32, 0, ctx->stream>>>(ctx->dPlanes, dT, n, dO, leftover);
        HVB_LAUNCH_CHECK(ctx, "satdMmaSmallKernel");
    }
    else
    {
        const auto *dT = static_cast<const hvb_metric_task *>(st.dTasks);
        auto *dO = static_cast<int32_t *>(st.dOut);
        leftover = ctx->workCursors + 2;
        cudaMemsetAsync(leftover, 0, 3 * sizeof(int), ctx->stream);
        constexpr int kStages = 2;
        const int smem = kWarpsPerBlock * kStages * 2304; // per warp and stage: 2 KB of row slots + 256 B of fifth words
        cudaFuncSetAttribute(satdMma16Kernel<kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        int perSm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, satdMma16Kernel<kStages>, kWarpsPerBlock * 32, smem);
        const int blocks = min((n + kWarpsPerBlock - 1) / kWarpsPerBlock, ctx->smCount * max(perSm, 1));
        satdMma16Kernel<kStages><<<blocks, kWarpsPerBlock * 32, smem, ctx->stream>>>(ctx->dPlanes, dT, n, dO, leftover);
        HVB_LAUNCH_CHECK(ctx, "satdMma16Kernel");
    }
    HVB_DISPATCH_SAMPLE(ctx, satdKernel, gridFor<hvb_metric_task>(ctx, n), ctx->dPlanes,
                        static_cast<const hvb_metric_task *>(st.dTasks), n, static_cast<int32_t *>(st.dOut), leftover);
    HVB_LAUNCH_CHECK(ctx, "satdKernel");
    return hvbStageOut(ctx, out, sizeof(int32_t) * n, mem, st);
}
