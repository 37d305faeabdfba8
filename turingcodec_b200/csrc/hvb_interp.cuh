// hvb_interp.cuh -- HEVC fractional-sample interpolation passes shared by the prediction kernels
// (hvb_pred.cu) and the sub-pel stage of the motion search (hvb_me.cu).
//   filter taps        havoc/pred_inter.cpp:39-69
//   separable passes   havoc/pred_inter.cpp:76-110, :145-202 (shift1 / shift2 / shift3 schedule)
#pragma once
#include "hvb_internal.cuh"

namespace hvb_interp {

constexpr int kMidElems = (64 + 7) * 64; // horizontal-pass tile: (h + taps - 1) rows of w int16
constexpr int kPredElems = 64 * 64;
constexpr int kSmemPerWarp = (kMidElems + kPredElems) * 2; // bytes

static __device__ __constant__ int8_t kLuma[4][8] = {{0, 0, 0, 64, 0, 0, 0, 0},
                                              {-1, 4, -10, 58, 17, -5, 1, 0},
                                              {-1, 4, -11, 40, 40, -11, 4, -1},
                                              {0, 1, -5, 17, 58, -10, 4, -1}};
static __device__ __constant__ int8_t kChroma[8][4] = {{0, 64, 0, 0},   {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4},
                                                {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2}};

template <int TAPS>
__device__ __forceinline__ int coef(int frac, int k)
{
    return TAPS == 8 ? kLuma[frac][k] : kChroma[frac][k];
}

// Horizontal pass of one reference block into `mid` ((h + TAPS - 1) rows of w int16, row stride w).
template <typename Sample, int TAPS>
__device__ __forceinline__ void passH(int16_t *mid, const Sample *ref, int sr, int w, int h, int xFrac, int shift1, int lane)
{
    constexpr int M = TAPS / 2 - 1;
    int c[TAPS];
#pragma unroll
    for (int k = 0; k < TAPS; ++k) c[k] = coef<TAPS>(xFrac, k);
    const int rows = h + TAPS - 1, total = rows * w;
    const Sample *origin = ref - M * sr - M;
    for (int i = lane; i < total; i += 32)
    {
        const int y = i / w, x = i - y * w;
        const Sample *p = origin + y * sr + x;
        int acc = 0;
#pragma unroll
        for (int k = 0; k < TAPS; ++k) acc += c[k] * (int)__ldg(p + k);
        mid[i] = (int16_t)(acc >> shift1);
    }
}

// Vertical pass value at (x, y) without rounding/shift.
template <int TAPS>
__device__ __forceinline__ int passV(const int16_t *mid, int w, int x, int y, const int (&c)[TAPS])
{
    int acc = 0;
#pragma unroll
    for (int k = 0; k < TAPS; ++k) acc += c[k] * (int)mid[(y + k) * w + x];
    return acc;
}

} // namespace hvb_interp
