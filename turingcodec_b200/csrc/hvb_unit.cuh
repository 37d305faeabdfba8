// hvb_unit.cuh -- arithmetic shared by the kernels that work on 8x8 / 8x4 / 4x8 units of a prediction block: the
// sub-pel stage of the motion search (hvb_me_subpel.cu) and the PU cost (hvb_pu_cost.cu).
//   filter taps, shift schedule   havoc/pred_inter.cpp:39-69, :76-110
//   Hadamard SATD as a product    havoc/hadamard.cpp:58-98 (see hvb_me_subpel.cu for the derivation)
#pragma once
#include "hvb_internal.cuh"

namespace hvb_unit {

constexpr int kColStride = 20; // halfwords per mid column: 16 support rows + 4 (40 bytes: 8-byte aligned, odd/2 banks)

// 8-tap luma filters (havoc/pred_inter.cpp:39-69) packed as s8x4 words, taps 0..3 and 4..7
static __device__ __constant__ uint32_t kTapWords[4][2] = {{0x40000000u, 0x00000000u},
                                                    {0x3af604ffu, 0x0001fb11u},
                                                    {0x28f504ffu, 0xff04f528u},
                                                    {0x11fb0100u, 0xff04f63au}};
__device__ __forceinline__ int dp4aUS(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// 8 taps over four (s16, s16) pairs
__device__ __forceinline__ int tap8(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t t0, uint32_t t1, int acc)
{
    acc = __dp2a_lo((int)p0, (int)t0, acc);
    acc = __dp2a_hi((int)p1, (int)t0, acc);
    acc = __dp2a_lo((int)p2, (int)t1, acc);
    acc = __dp2a_hi((int)p3, (int)t1, acc);
    return acc;
}

// arithmetic of one bit depth (havoc/pred_inter.cpp:76-110): shift1 = min(4, bd - 8), shift3 = max(2, 14 - bd)
struct Depth
{
    int shift1, shift3, maxv;
    __device__ __forceinline__ explicit Depth(int bd) : shift1(min(4, bd - 8)), shift3(max(2, 14 - bd)), maxv((1 << bd) - 1) {}
    // second-stage output from the 8-tap sum of mids
    __device__ __forceinline__ int out(int sum) const { return __vimin_s32_relu((sum + (1 << (5 + shift3))) >> (6 + shift3), maxv); }
    // the same for a zero vertical phase (taps {0,0,0,64,..}) applied to one mid
    __device__ __forceinline__ int outCopy(int mid) const { return __vimin_s32_relu((mid + (1 << (shift3 - 1))) >> shift3, maxv); }
};

__device__ __forceinline__ int vFilter(uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, uint32_t t0, uint32_t t1, const Depth &D)
{
    return D.out(tap8(p0, p1, p2, p3, t0, t1, 0));
}

__device__ __forceinline__ void imma16832(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// A-fragment register (m-tile mt, k-step ks, reg) holds A[m][k..k+3], m = 16 mt + (lane >> 2) + 8 (reg & 1),
// k = 32 ks + 4 (lane & 3) + 16 (reg >> 1), entries (-1)^popc(m & k): the byte pattern depends on m & 3, the sign on
// the other bits, of which two products involve the lane; the rest is compile-time in the unrolled loops.
struct HadamardA
{
    uint32_t e[2], o[2]; // [negated], registers with reg & 1 == 0 / 1
    __device__ __forceinline__ explicit HadamardA(int lane)
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

// B-fragment register(s) for 4 consecutive samples at `p + off` (p aligned to 4 samples, off = 0 or 1).
// 8 bit: one word.  16 bit: the samples' low bytes and high bytes as two words (the products run on both planes).
struct Frag
{
    uint32_t lo, hi;
};
__device__ __forceinline__ Frag loadFrag(const uint8_t *p, int off)
{
    const uint32_t *q = reinterpret_cast<const uint32_t *>(p);
    const uint32_t lo = q[0];
    return Frag{off ? __funnelshift_r(lo, q[1], 8) : lo, 0u};
}
__device__ __forceinline__ Frag loadFrag(const uint16_t *p, int off)
{
    const uint32_t *q = reinterpret_cast<const uint32_t *>(p);
    uint32_t x = q[0], y = q[1];
    if (off)
    {
        const uint32_t z = q[2];
        x = __funnelshift_r(x, y, 16);
        y = __funnelshift_r(y, z, 16);
    }
    return Frag{__byte_perm(x, y, 0x6420), __byte_perm(x, y, 0x7531)};
}

// the 8 words of a mid column (support rows 0..15) and the 8-tap window slid down it.  Output row r reads support
// rows r + j0 .. r + j0 + 7 (j0 = 0 or 1).
struct Column
{
    uint32_t w[8], x[7];
    __device__ __forceinline__ void load(const int16_t *col, int j0)
    {
        const uint2 *q = reinterpret_cast<const uint2 *>(col);
#pragma unroll
        for (int i = 0; i < 4; ++i)
        {
            const uint2 t = q[i];
            w[2 * i] = t.x;
            w[2 * i + 1] = t.y;
        }
        if (j0)
        {
#pragma unroll
            for (int i = 0; i < 7; ++i) w[i] = __funnelshift_r(w[i], w[i + 1], 16);
            w[7] >>= 16;
        }
#pragma unroll
        for (int i = 0; i < 7; ++i) x[i] = __funnelshift_r(w[i], w[i + 1], 16);
    }
    template <int R>
    __device__ __forceinline__ int out(uint32_t t0, uint32_t t1, const Depth &D) const
    {
        constexpr int k = R >> 1;
        if (R & 1) return vFilter(x[k], x[k + 1], x[k + 2], x[k + 3 < 7 ? k + 3 : 6], t0, t1, D);
        return vFilter(w[k], w[k + 1], w[k + 2], w[k + 3 < 8 ? k + 3 : 7], t0, t1, D);
    }
};

} // namespace hvb_unit
