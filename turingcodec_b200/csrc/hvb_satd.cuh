// hvb_satd.cuh -- register-resident Hadamard SATD tiles shared by the metric, fused
// interpolation+SATD and intra-sweep kernels.
//   havoc_hadamard_satd  havoc/hadamard.cpp:58-98
//   measureSatd          turing/Measure.h:96-135
#pragma once
#include "hvb_internal.cuh"

// One N x N Hadamard tile computed entirely in one thread's registers (N = 2, 4, 8).
// Shared with the fused interpolation+SATD and intra-sweep kernels.
template <typename SampleA, typename SampleB, int LOG2N>
__device__ __forceinline__ int hvbSatdTile(const SampleA *a, int sa, const SampleB *b, int sb, int postShift)
{
    constexpr int N = 1 << LOG2N;
    int m[N][N];
    if (N == 8 && sizeof(SampleA) == 1 && sizeof(SampleB) == 1 &&
        !((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | (uintptr_t)sa | (uintptr_t)sb) & 7))
    {
        // both 8-bit rows 8-byte aligned: one 64-bit load per row per operand instead of eight byte loads
#pragma unroll
        for (int y = 0; y < N; ++y)
        {
            const uint2 va = *reinterpret_cast<const uint2 *>(reinterpret_cast<const uint8_t *>(a) + y * sa);
            const uint2 vb = *reinterpret_cast<const uint2 *>(reinterpret_cast<const uint8_t *>(b) + y * sb);
#pragma unroll
            for (int x = 0; x < 4; ++x)
            {
                m[y][x] = (int)((va.x >> (8 * x)) & 0xff) - (int)((vb.x >> (8 * x)) & 0xff);
                m[y][x + 4] = (int)((va.y >> (8 * x)) & 0xff) - (int)((vb.y >> (8 * x)) & 0xff);
            }
        }
    }
    else if (N == 8 && sizeof(SampleA) == 2 && sizeof(SampleB) == 2 &&
             !((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | (uintptr_t)(2 * sa) | (uintptr_t)(2 * sb)) & 3))
    {
        // 16-bit rows on a 4-byte boundary: four 32-bit loads per row per operand (one 128-bit load when the row is 16-byte
        // aligned) instead of eight 16-bit loads
#pragma unroll
        for (int y = 0; y < N; ++y)
        {
            const uint32_t *pa = reinterpret_cast<const uint32_t *>(a + y * sa), *pb = reinterpret_cast<const uint32_t *>(b + y * sb);
            // (plain loads: the prediction operand of the fused interpolation + SATD kernel lives in shared memory)
            uint4 va, vb;
            if (!(reinterpret_cast<uintptr_t>(pa) & 15))
                va = *reinterpret_cast<const uint4 *>(pa);
            else
                va = make_uint4(pa[0], pa[1], pa[2], pa[3]);
            if (!(reinterpret_cast<uintptr_t>(pb) & 15))
                vb = *reinterpret_cast<const uint4 *>(pb);
            else
                vb = make_uint4(pb[0], pb[1], pb[2], pb[3]);
            const uint32_t wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int x = 0; x < 4; ++x)
            {
                m[y][2 * x] = (int)(wa[x] & 0xffff) - (int)(wb[x] & 0xffff);
                m[y][2 * x + 1] = (int)(wa[x] >> 16) - (int)(wb[x] >> 16);
            }
        }
    }
    else
    {
#pragma unroll
        for (int y = 0; y < N; ++y)
#pragma unroll
            for (int x = 0; x < N; ++x) m[y][x] = (int)a[y * sa + x] - (int)b[y * sb + x];
    }

        // rows
#pragma unroll
    for (int y = 0; y < N; ++y)
#pragma unroll
        for (int half = N / 2; half >= 1; half >>= 1)
#pragma unroll
            for (int base = 0; base < N; base += 2 * half)
#pragma unroll
                for (int j = 0; j < half; ++j)
                {
                    const int p = m[y][base + j], q = m[y][base + j + half];
                    m[y][base + j] = p + q;
                    m[y][base + j + half] = p - q;
                }
        // columns
#pragma unroll
    for (int x = 0; x < N; ++x)
#pragma unroll
        for (int half = N / 2; half >= 1; half >>= 1)
#pragma unroll
            for (int base = 0; base < N; base += 2 * half)
#pragma unroll
                for (int j = 0; j < half; ++j)
                {
                    const int p = m[base + j][x], q = m[base + j + half][x];
                    m[base + j][x] = p + q;
                    m[base + j + half][x] = p - q;
                }
    int acc = N / 4;
#pragma unroll
    for (int y = 0; y < N; ++y)
#pragma unroll
        for (int x = 0; x < N; ++x) acc += abs(m[y][x]);
    // havoc/hadamard.cpp:93-96: normalise, then the 16-bit sample paths shift right by 2 per tile
    return (acc >> (LOG2N - 1)) >> postShift;
}

// the number of Hadamard tiles measureSatd cuts a w x h block into (turing/Measure.h:96-135)
__device__ __forceinline__ int hvbSatdTiles(int w, int h)
{
    const int log2 = ((w | h) & 3) ? 1 : (((w | h) & 7) ? 2 : 3);
    return (w >> log2) * (h >> log2);
}

template <typename SampleA, typename SampleB>
__device__ __forceinline__ int hvbMeasureSatdLanes(const SampleA *a, int sa, const SampleB *b, int sb, int w, int h, int lane,
                                                   int lanes, int postShift)
{
    // turing/Measure.h:96-135: tile size from the alignment of (w | h)
    int acc = 0;
    if ((w | h) & 3)
    {
        const int tw = w >> 1, tiles = tw * (h >> 1);
        for (int t = lane; t < tiles; t += lanes)
        {
            const int ty = t / tw, tx = t - ty * tw;
            acc += hvbSatdTile<SampleA, SampleB, 1>(a + 2 * ty * sa + 2 * tx, sa, b + 2 * ty * sb + 2 * tx, sb, postShift);
        }
    }
    else if ((w | h) & 7)
    {
        const int tw = w >> 2, tiles = tw * (h >> 2);
        for (int t = lane; t < tiles; t += lanes)
        {
            const int ty = t / tw, tx = t - ty * tw;
            acc += hvbSatdTile<SampleA, SampleB, 2>(a + 4 * ty * sa + 4 * tx, sa, b + 4 * ty * sb + 4 * tx, sb, postShift);
        }
    }
    else
    {
        const int tw = w >> 3, tiles = tw * (h >> 3);
        for (int t = lane; t < tiles; t += lanes)
        {
            const int ty = t / tw, tx = t - ty * tw;
            acc += hvbSatdTile<SampleA, SampleB, 3>(a + 8 * ty * sa + 8 * tx, sa, b + 8 * ty * sb + 8 * tx, sb, postShift);
        }
    }
    return acc;
}


