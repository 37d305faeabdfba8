/*
 * oracle/oracle_havoc.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * CPU restatement of the havoc primitive library's C reference path.  Written from the
 * algorithm, not transliterated: transforms are stated as the matrix products the
 * reference's partial butterflies factorise (integer sums are exact in int32, so the two
 * are bit-identical), the DCT matrix is generated from its 31 distinct magnitudes, and the
 * Hadamard SATD is a separable +/- butterfly.  Pinned against oracle/_ref in
 * tests/test_oracle_pin.py.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

static inline int sample_at(const void *p, intptr_t i, int bps)
{
    return bps == 1 ? ((const uint8_t *)p)[i] : ((const uint16_t *)p)[i];
}

static inline void sample_put(void *p, intptr_t i, int v, int bps)
{
    if (bps == 1) ((uint8_t *)p)[i] = (uint8_t)v;
    else ((uint16_t *)p)[i] = (uint16_t)v;
}

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

/* ------------------------------------------------------------------------------------ */
/* SAD / SSD                                                                            */
/* ------------------------------------------------------------------------------------ */

/* havoc/sad.cpp:432-449 */
int orc_sad(const void *src, intptr_t ss, const void *ref, intptr_t sr, int w, int h, int bps)
{
    int acc = 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            acc += abs(sample_at(src, x + y * ss, bps) - sample_at(ref, x + y * sr, bps));
    return bps == 2 ? acc >> 2 : acc;
}

/* havoc/sad.cpp:513-542 */
void orc_sad_multiref4(const void *src, intptr_t ss, const void *const ref[4], intptr_t sr,
                       int sad[4], int w, int h, int bps)
{
    for (int way = 0; way < 4; ++way)
        sad[way] = orc_sad(src, ss, ref[way], sr, w, h, bps);
}

/* havoc/ssd.cpp:28-43 -- accumulates in uint32 (wraps), 16-bit samples scale down by 16 */
uint32_t orc_ssd(const void *a, intptr_t sa, const void *b, intptr_t sb, int w, int h, int bps)
{
    uint32_t acc = 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            int d = sample_at(a, x + y * sa, bps) - sample_at(b, x + y * sb, bps);
            acc += (uint32_t)(d * d);
        }
    return bps == 2 ? acc >> 4 : acc;
}

/* havoc/diff.cpp:29-38 */
int orc_ssd_linear(const uint8_t *a, const uint8_t *b, int n)
{
    int acc = 0;
    for (int i = 0; i < n; ++i)
    {
        int d = a[i] - b[i];
        acc += d * d;
    }
    return acc;
}

/* ------------------------------------------------------------------------------------ */
/* Hadamard SATD                                                                        */
/* ------------------------------------------------------------------------------------ */

/* in-place length-n (n = 2,4,8) Hadamard butterfly over v[0], v[step], ... */
static void hadamard_1d(int *v, int n, int step)
{
    for (int half = n / 2; half >= 1; half /= 2)
        for (int base = 0; base < n; base += 2 * half)
            for (int j = 0; j < half; ++j)
            {
                int a = v[(base + j) * step], b = v[(base + j + half) * step];
                v[(base + j) * step] = a + b;
                v[(base + j + half) * step] = a - b;
            }
}

/* havoc/hadamard.cpp:58-98: the order of the transformed coefficients does not affect the
 * sum of magnitudes, so any Hadamard ordering gives the reference's value. */
int orc_hadamard_satd(const void *a, intptr_t sa, const void *b, intptr_t sb, int log2n, int bps)
{
    const int n = 1 << log2n;
    int m[8 * 8];
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x)
            m[y * 8 + x] = sample_at(a, x + y * sa, bps) - sample_at(b, x + y * sb, bps);
    for (int y = 0; y < n; ++y) hadamard_1d(&m[y * 8], n, 1);
    for (int x = 0; x < n; ++x) hadamard_1d(&m[x], n, 8);
    int acc = n / 4; /* rounding offset: 0, 1, 2 */
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x)
            acc += abs(m[y * 8 + x]);
    acc /= n / 2; /* 1, 2, 4 */
    return bps == 2 ? acc >> 2 : acc;
}

/* turing/Measure.h:96-135 */
int32_t orc_measure_satd(const void *a, intptr_t sa, const void *b, intptr_t sb, int w, int h, int bps)
{
    int log2n = ((w | h) & 3) ? 1 : (((w | h) & 7) ? 2 : 3);
    int n = 1 << log2n;
    int32_t acc = 0;
    for (int y = 0; y < h; y += n)
        for (int x = 0; x < w; x += n)
            acc += orc_hadamard_satd((const char *)a + (x + y * sa) * bps, sa,
                                     (const char *)b + (x + y * sb) * bps, sb, log2n, bps);
    return acc;
}

/* ------------------------------------------------------------------------------------ */
/* Inter prediction                                                                     */
/* ------------------------------------------------------------------------------------ */

/* havoc/pred_inter.cpp:39-69: HEVC luma (8-tap, quarter-pel) and chroma (4-tap, eighth-pel) filters */
int orc_pred_coefficient(int taps, int frac, int k)
{
    static const int8_t luma[4][8] = {
        {0, 0, 0, 64, 0, 0, 0, 0},
        {-1, 4, -10, 58, 17, -5, 1, 0},
        {-1, 4, -11, 40, 40, -11, 4, -1},
        {0, 1, -5, 17, 58, -10, 4, -1}};
    static const int8_t chroma[8][4] = {
        {0, 64, 0, 0}, {-2, 58, 10, -2}, {-4, 54, 16, -2}, {-6, 46, 28, -4},
        {-4, 36, 36, -4}, {-4, 28, 46, -6}, {-2, 16, 54, -4}, {-2, 10, 58, -2}};
    return taps == 8 ? luma[frac][k] : chroma[frac][k];
}

/* One separable pass (havoc/pred_inter.cpp:76-110 havoc_pred_uni_generic).  Reads integers via
 * `get`, taps run along `tap_stride`; rounding term is (round << shift) >> 1. */
typedef int (*get_fn)(const void *, intptr_t);
static int get_u8(const void *p, intptr_t i) { return ((const uint8_t *)p)[i]; }
static int get_u16(const void *p, intptr_t i) { return ((const uint16_t *)p)[i]; }
static int get_i32(const void *p, intptr_t i) { return ((const int *)p)[i]; }

static void filter_pass(int *dst, intptr_t sd, const void *src, get_fn get, intptr_t ss,
                        int w, int h, intptr_t tap_stride, int taps, int frac, int shift,
                        int round, int clipBits)
{
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            int acc = (round << shift) >> 1;
            for (int k = 0; k < taps; ++k)
                acc += orc_pred_coefficient(taps, frac, k) *
                       get(src, x + y * ss + (k - taps / 2 + 1) * tap_stride);
            acc >>= shift;
            if (clipBits) acc = clip3(0, (1 << clipBits) - 1, acc);
            dst[x + y * sd] = acc;
        }
}

/* havoc/pred_inter.cpp:113-202 with the C_OPT dispatch of havocPopulatePredUni (:941-962) */
void orc_pred_uni(void *dst, intptr_t sd, const void *ref, intptr_t sr, int w, int h,
                  int xFrac, int yFrac, int bitDepth, int taps, int bps)
{
    get_fn get = bps == 1 ? get_u8 : get_u16;
    int *out = (int *)malloc(sizeof(int) * 64 * 64);
    if (!xFrac && !yFrac)
    {
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x)
                out[x + y * 64] = get(ref, x + y * sr);
    }
    else if (xFrac && !yFrac)
        filter_pass(out, 64, ref, get, sr, w, h, 1, taps, xFrac, 6, 1, bitDepth);
    else if (!xFrac && yFrac)
        filter_pass(out, 64, ref, get, sr, w, h, sr, taps, yFrac, 6, 1, bitDepth);
    else
    {
        int shift1 = bitDepth - 8 > 4 ? 4 : bitDepth - 8;
        int shift3 = 14 - bitDepth < 2 ? 2 : 14 - bitDepth;
        int margin = taps / 2 - 1;
        int *mid = (int *)malloc(sizeof(int) * 64 * (64 + 7));
        filter_pass(mid, 64, (const char *)ref - margin * sr * bps, get, sr, w, h + taps - 1, 1, taps,
                    xFrac, shift1, 0, 0);
        filter_pass(out, 64, mid + margin * 64, get_i32, 64, w, h, 64, taps, yFrac, 6 + shift3, 1,
                    bitDepth);
        free(mid);
    }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            sample_put(dst, x + y * sd, out[x + y * 64], bps);
    free(out);
}

/* havoc/pred_inter.cpp:1222-1252: both references go through H (shift1) then V (shift 6) at
 * 14-bit precision, then mean with rounding and clip. */
void orc_pred_bi(void *dst, intptr_t sd, const void *ref0, const void *ref1, intptr_t sr, int w,
                 int h, int xFrac0, int yFrac0, int xFrac1, int yFrac1, int bitDepth, int taps, int bps)
{
    get_fn get = bps == 1 ? get_u8 : get_u16;
    int shift1 = bitDepth - 8 > 4 ? 4 : bitDepth - 8;
    int shift3 = 14 - bitDepth < 2 ? 2 : 14 - bitDepth;
    int margin = taps / 2 - 1;
    int *mid = (int *)malloc(sizeof(int) * 64 * (64 + 7));
    int *p[2];
    const void *refs[2] = {ref0, ref1};
    int xf[2] = {xFrac0, xFrac1}, yf[2] = {yFrac0, yFrac1};
    for (int i = 0; i < 2; ++i)
    {
        p[i] = (int *)malloc(sizeof(int) * 64 * 64);
        filter_pass(mid, 64, (const char *)refs[i] - margin * sr * bps, get, sr, w, h + taps - 1, 1,
                    taps, xf[i], shift1, 0, 0);
        filter_pass(p[i], 64, mid + margin * 64, get_i32, 64, w, h, 64, taps, yf[i], 6, 0, 0);
    }
    const int shift = shift3 + 1;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            int v = (p[0][x + y * 64] + p[1][x + y * 64] + (1 << (shift - 1))) >> shift;
            sample_put(dst, x + y * sd, clip3(0, (1 << bitDepth) - 1, v), bps);
        }
    free(p[0]);
    free(p[1]);
    free(mid);
}

/* havoc/pred_inter.cpp:2063-2080 */
void orc_subtract_bi(void *dst, intptr_t sd, const void *pred, intptr_t sp, const void *src,
                     intptr_t ss, int w, int h, int bitDepth, int bps)
{
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            int v = 2 * sample_at(src, x + y * ss, bps) - sample_at(pred, x + y * sp, bps);
            sample_put(dst, x + y * sd, clip3(0, (1 << bitDepth) - 1, v), bps);
        }
}

/* ------------------------------------------------------------------------------------ */
/* Intra prediction                                                                     */
/* ------------------------------------------------------------------------------------ */

static const int8_t intra_angle[35] = {0, 0, 32, 26, 21, 17, 13, 9, 5, 2, 0, -2, -5, -9, -13, -17, -21,
                                       -26, -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21,
                                       26, 32};
static const int16_t intra_inv_angle[26] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, -4096, -1638, -910, -630,
                                            -482, -390, -315, -256, -315, -390, -482, -630, -910,
                                            -1638, -4096};

/* havoc/pred_intra.cpp:20282-20401.  nb(x,y) = neighbours[x - y - 1] (:43-51). */
void orc_pred_intra(void *dst, intptr_t sd, const void *neighbours, int mode, int log2n,
                    int bitDepth, int edge_flag, int bps)
{
    const int n = 1 << log2n;
#define NB(x, y) sample_at(neighbours, (x) - (y)-1, bps)
    if (mode == 0)
    {
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x)
                sample_put(dst, x + y * sd,
                           ((n - 1 - x) * NB(-1, y) + (x + 1) * NB(n, -1) + (n - 1 - y) * NB(x, -1) +
                            (y + 1) * NB(-1, n) + n) >> (log2n + 1),
                           bps);
        return;
    }
    if (mode == 1)
    {
        int dc = n;
        for (int i = 0; i < n; ++i) dc += NB(i, -1) + NB(-1, i);
        dc >>= log2n + 1;
        for (int y = 0; y < n; ++y)
            for (int x = 0; x < n; ++x) sample_put(dst, x + y * sd, dc, bps);
        if (edge_flag)
        {
            sample_put(dst, 0, (NB(-1, 0) + 2 * dc + NB(0, -1) + 2) >> 2, bps);
            for (int x = 1; x < n; ++x) sample_put(dst, x, (NB(x, -1) + 3 * dc + 2) >> 2, bps);
            for (int y = 1; y < n; ++y) sample_put(dst, y * sd, (NB(-1, y) + 3 * dc + 2) >> 2, bps);
        }
        return;
    }

    /* angular: build the 1-D reference array ref[-n .. 2n], then interpolate along it.
     * vertical family (mode >= 18) walks the row above, horizontal family the left column;
     * the two are transposes of each other. */
    const int vertical = mode >= 18;
    const int angle = intra_angle[mode];
    int refbuf[32 + 65];
    int *ref = refbuf + 32;
    for (int i = 0; i <= n; ++i) ref[i] = vertical ? NB(-1 + i, -1) : NB(-1, -1 + i);
    if (angle < 0)
    {
        const int inv = intra_inv_angle[mode];
        const int last = (n * angle) >> 5;
        if (last < -1)
            for (int i = -1; i >= last; --i)
            {
                int j = -1 + ((i * inv + 128) >> 8);
                ref[i] = vertical ? NB(-1, j) : NB(j, -1);
            }
    }
    else
        for (int i = n + 1; i <= 2 * n; ++i) ref[i] = vertical ? NB(-1 + i, -1) : NB(-1, -1 + i);

    for (int major = 0; major < n; ++major) /* y for the vertical family, x for the horizontal */
    {
        const int idx = ((major + 1) * angle) >> 5;
        const int fact = ((major + 1) * angle) & 31;
        for (int minor = 0; minor < n; ++minor)
        {
            int v = fact ? ((32 - fact) * ref[minor + idx + 1] + fact * ref[minor + idx + 2] + 16) >> 5
                         : ref[minor + idx + 1];
            if (vertical) sample_put(dst, minor + major * sd, v, bps);
            else sample_put(dst, major + minor * sd, v, bps);
        }
    }
    if (edge_flag && mode == 26)
        for (int y = 0; y < n; ++y)
            sample_put(dst, y * sd,
                       clip3(0, (1 << bitDepth) - 1, NB(0, -1) + ((NB(-1, y) - NB(-1, -1)) >> 1)), bps);
    if (edge_flag && mode == 10)
        for (int x = 0; x < n; ++x)
            sample_put(dst, x,
                       clip3(0, (1 << bitDepth) - 1, NB(-1, 0) + ((NB(x, -1) - NB(-1, -1)) >> 1)), bps);
#undef NB
}

/* ------------------------------------------------------------------------------------ */
/* Transforms                                                                           */
/* ------------------------------------------------------------------------------------ */

/* HEVC core transform matrix entry M_N[k][i] (the tables in havoc/transform.cpp:76-330 /
 * :3087-3352 are this matrix).  The 32-point matrix has 31 distinct magnitudes
 * mag[j] ~ 64*sqrt(2)*cos(j*pi/64); smaller sizes take every (32/N)-th row. */
static int dct_coeff(int log2n, int k, int i)
{
    static const int8_t mag[33] = {64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                                   61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9, 4, 0};
    if (k == 0) return 64;
    /* cos(pi*k*(2i+1)/(2N)) = cos(pi*(k*32/N)*(2i+1)/64): angle in units of pi/64, mod 2pi */
    int m = ((k << (5 - log2n)) * (2 * i + 1)) & 127;
    if (m > 64) m = 128 - m;
    return m > 32 ? -mag[64 - m] : mag[m];
}

static int dst_coeff(int k, int i)
{
    static const int8_t m[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};
    return m[k][i];
}

static int tr_coeff(int trType, int log2n, int k, int i)
{
    return trType ? dst_coeff(k, i) : dct_coeff(log2n, k, i);
}

/* havoc/transform.cpp:3087-3397.  pass: out[k*n + j] = (int16)((sum_i M[k][i]*in[j*stride+i] + add) >> shift).
 * The narrowing to int16 wraps (shiftRight :3071-3084 stores into a short before clamping). */
void orc_transform_fwd(int16_t *coeffs, const int16_t *src, intptr_t stride, int trType, int log2n,
                       int bitDepth)
{
    const int n = 1 << log2n;
    int16_t tmp[32 * 32];
    const int shift1 = log2n - 1 + bitDepth - 8, shift2 = log2n + 6;
    for (int pass = 0; pass < 2; ++pass)
    {
        const int16_t *in = pass ? tmp : src;
        const intptr_t is = pass ? n : stride;
        int16_t *out = pass ? coeffs : tmp;
        const int shift = pass ? shift2 : shift1;
        const int add = 1 << (shift - 1);
        for (int j = 0; j < n; ++j)
            for (int k = 0; k < n; ++k)
            {
                int acc = add;
                for (int i = 0; i < n; ++i) acc += tr_coeff(trType, log2n, k, i) * in[j * is + i];
                out[k * n + j] = (int16_t)(acc >> shift);
            }
    }
}

/* havoc/transform.cpp:50-401.  pass: out[j*n + k] = clip16((sum_i M[i][k]*in[i*n + j] + add) >> shift),
 * shifts 7 then 20 - bitDepth. */
void orc_inverse_transform(int16_t *res, const int16_t *coeffs, int trType, int log2n, int bitDepth)
{
    const int n = 1 << log2n;
    int16_t tmp[32 * 32];
    for (int pass = 0; pass < 2; ++pass)
    {
        const int16_t *in = pass ? tmp : coeffs;
        int16_t *out = pass ? res : tmp;
        const int shift = pass ? 20 - bitDepth : 7;
        const int add = 1 << (shift - 1);
        for (int j = 0; j < n; ++j)
            for (int k = 0; k < n; ++k)
            {
                int acc = add;
                for (int i = 0; i < n; ++i) acc += tr_coeff(trType, log2n, i, k) * in[i * n + j];
                out[j * n + k] = (int16_t)clip3(-32768, 32767, acc >> shift);
            }
    }
}

/* havoc/transform.cpp:366-375 + transform.h:104-114 */
void orc_inverse_transform_add(void *dst, intptr_t sd, const void *pred, intptr_t sp,
                               const int16_t *coeffs, int trType, int log2n, int bitDepth, int bps)
{
    const int n = 1 << log2n;
    int16_t res[32 * 32];
    orc_inverse_transform(res, coeffs, trType, log2n, bitDepth);
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x)
            sample_put(dst, x + y * sd,
                       clip3(0, (1 << bitDepth) - 1, sample_at(pred, x + y * sp, bps) + res[x + y * n]), bps);
}

/* ------------------------------------------------------------------------------------ */
/* Quantisation                                                                         */
/* ------------------------------------------------------------------------------------ */

/* havoc/quantize.cpp:278-304 */
int orc_quantize(int16_t *dst, const int16_t *src, int scale, int shift, int offset, int n)
{
    int cbf = 0;
    const int off = offset << (shift - 16);
    for (int i = 0; i < n; ++i)
    {
        int v = src[i];
        int mag = ((v < 0 ? -v : v) * scale + off) >> shift;
        int q = clip3(-32768, 32767, v < 0 ? -mag : mag);
        cbf |= q;
        dst[i] = (int16_t)q;
    }
    return cbf;
}

/* havoc/quantize.cpp:37-46 */
void orc_quantize_inverse(int16_t *dst, const int16_t *src, int scale, int shift, int n)
{
    for (int i = 0; i < n; ++i)
        dst[i] = (int16_t)clip3(-32768, 32767, (src[i] * scale + (1 << (shift - 1))) >> shift);
}

/* havoc/quantize.cpp:538-548 */
void orc_quantize_reconstruct(uint8_t *rec, intptr_t sr, const uint8_t *pred, intptr_t sp,
                              const int16_t *res, int n)
{
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x)
            rec[x + y * sr] = (uint8_t)clip3(0, 255, pred[x + y * sp] + res[x + y * n]);
}
