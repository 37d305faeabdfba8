/*
 * oracle/oracle_preanalysis.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * CPU restatement of the intra-complexity measure of the reference's pre-analysis:
 *   EstimateIntraComplexity::computeSatd8x8     turing/EstimateIntraComplexity.h:55-157
 *   EstimateIntraComplexity::preAnalysis        turing/EstimateIntraComplexity.h:159-176
 * The measure is the sum of the absolute 8x8 Hadamard coefficients of the SOURCE samples themselves without the DC term,
 * (s + 2) >> 2, and >> 2 again for 16-bit samples.  Written as a separable transform with generic butterflies (sum |.| does
 * not depend on the coefficient order the reference's hand-unrolled stages produce).  Pinned against the reference function
 * by tests/test_oracle_pin_preanalysis.py through oracle/ref_shim_preanalysis.cpp.
 *
 * Further down, also pinned there:
 *   AdaptiveQuantisation::preAnalysis                       turing/AdaptiveQuantisation.h:172-246
 *   ShotChangeDetection: luma histogram, getLikelihood      turing/SCDetection.h:40-178, :237-262
 */
#include "oracle.h"
#include <stdlib.h>

static void hadamard8(int *v, int step)
{
    for (int half = 4; half >= 1; half >>= 1)
        for (int base = 0; base < 8; base += 2 * half)
            for (int j = 0; j < half; ++j)
            {
                const int a = v[(base + j) * step], b = v[(base + j + half) * step];
                v[(base + j) * step] = a + b;
                v[(base + j + half) * step] = a - b;
            }
}

int orc_intra_complexity_8x8(const void *p, intptr_t stride, int bps)
{
    int m[64], total = 0;
    for (int y = 0; y < 8; ++y)
        for (int x = 0; x < 8; ++x) m[8 * y + x] = bps == 1 ? ((const uint8_t *)p)[y * stride + x] : ((const uint16_t *)p)[y * stride + x];
    for (int y = 0; y < 8; ++y) hadamard8(m + 8 * y, 1);
    for (int x = 0; x < 8; ++x) hadamard8(m + x, 8);
    for (int i = 1; i < 64; ++i) total += abs(m[i]); /* m[0] is the DC term (the sum of the samples) */
    total = (total + 2) >> 2;
    return bps == 2 ? total >> 2 : total;
}

/* preAnalysis: every whole 8x8 block of the plane, raster order; returns the sum (m_satdSum) */
int orc_intra_complexity(const void *plane, intptr_t stride, int width, int height, int bps, int32_t *out)
{
    int sum = 0, k = 0;
    for (int by = 0; by < (height >> 3); ++by)
        for (int bx = 0; bx < (width >> 3); ++bx)
        {
            out[k] = orc_intra_complexity_8x8((const uint8_t *)plane + ((intptr_t)(8 * by) * stride + 8 * bx) * bps, stride, bps);
            sum += out[k++];
        }
    return sum;
}


/* ---- adaptive quantisation: the activity of the units of one layer ------------------------------------------------- */

static uint64_t sample_at(const void *plane, intptr_t stride, int x, int y, int bps)
{
    return bps == 1 ? ((const uint8_t *)plane)[y * stride + x] : ((const uint16_t *)plane)[y * stride + x];
}

/* turing/AdaptiveQuantisation.h:183-241 for the layer whose units are `unit` samples square: out[] = the smallest of the four
 * quadrants' integer variances (activity = 1.0 + out), raster order of units; returns (int)(mean activity), the layer's
 * average activity (:244-245).  Quadrant 0 "sums" the squares by assignment and quadrant 1 adds the samples where the
 * squares were meant: both kept, they are what the encoder's QP offsets are computed from. */
int orc_aq_layer(const void *plane, intptr_t stride, int width, int height, int bps, int unit, int64_t *out)
{
    int k = 0;
    double total = 0.0;
    for (int row = 0; row < height; row += unit)
        for (int col = 0; col < width; col += unit, ++k)
        {
            const int uh = height - row < unit ? height - row : unit, uw = width - col < unit ? width - col : unit;
            const int hh = uh >> 1, hw = uw >> 1, num = (uh * uw) >> 2;
            uint64_t sum[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0};
            for (int r = 0; r < uh; ++r)
                for (int c = 0; c < uw; ++c)
                {
                    const uint64_t v = sample_at(plane, stride, col + c, row + r, bps);
                    const int b = (r >= hh ? 2 : 0) + (c >= hw ? 1 : 0);
                    sum[b] += v;
                    if (b == 0) sq[0] = v * v; /* the value left by the last sample of the quadrant */
                    else if (b == 1) sq[1] += v;
                    else sq[b] += v * v;
                }
            int64_t best = 0;
            if (num)
                for (int b = 0; b < 4; ++b)
                {
                    const int64_t avg = (int64_t)(sum[b] / (uint64_t)num);
                    const int64_t var = (int64_t)(sq[b] / (uint64_t)num) - avg * avg;
                    if (b == 0 || var < best) best = var;
                }
            out[k] = best;
            total += 1.0 + (double)best;
        }
    return k ? (int)(total / k) : 0;
}

/* ---- shot-change detection ------------------------------------------------------------------------------------------ */

static int scd_byte(const void *plane, intptr_t stride, int x, int y, int bps)
{
    return bps == 1 ? ((const uint8_t *)plane)[y * stride + x] : (uint8_t)(((const uint16_t *)plane)[y * stride + x] >> 2);
}

/* turing/SCDetection.h:237-262: the 64-bin histogram of the picture's 8-bit luma values >> 2 */
void orc_scd_histogram(const void *plane, intptr_t stride, int width, int height, int bps, int32_t *hist)
{
    for (int i = 0; i < 64; ++i) hist[i] = 0;
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) ++hist[scd_byte(plane, stride, x, y, bps) >> 2];
}

/* turing/SCDetection.h:71-147: avg and var of the blocks of the grid with the given margin (1: previous picture, 2: current),
 * out[2 k], out[2 k + 1]; returns the number of blocks, or -1 where the reference's `h * height` addressing would leave the
 * picture.  The packed byte vector of the reference is addressed through (index / width, index % width). */
int orc_scd_block_stats(const void *plane, intptr_t stride, int width, int height, int bps, int margin, double *out)
{
    const int bw = width >> 3, bh = height >> 3;
    int k = 0;
    if (bw <= 0 || bh <= 0) return 0;
    for (int j = margin * bh; j < height - margin * bh; j += bh)
        for (int i = margin * bw; i < width - margin * bw; i += bw, ++k)
        {
            const long origin = (long)j * width + i;
            if (origin + (long)(bh - 1) * height + bw - 1 >= (long)width * height) return -1;
            int64_t sum = 0;
            for (int h = 0; h < bh; ++h)
                for (int w = 0; w < bw; ++w)
                {
                    const long f = origin + (long)h * height + w;
                    sum += scd_byte(plane, stride, (int)(f % width), (int)(f / width), bps);
                }
            const double avg = (double)sum / (bh * bw);
            volatile double var = 0.0; /* volatile: every addition rounded to double, in this order */
            for (int h = 0; h < bh; ++h)
                for (int w = 0; w < bw; ++w)
                {
                    const long f = origin + (long)h * height + w;
                    const double elem = (double)scd_byte(plane, stride, (int)(f % width), (int)(f / width), bps);
                    var += (elem - avg) * (elem - avg);
                }
            out[2 * k] = avg;
            out[2 * k + 1] = var / (bh * bw);
        }
    return k;
}

/* turing/SCDetection.h:40-47, :149-178: the likelihood ratio of a picture pair from the block statistics (prev: the 6 x 6
 * grid of margin 1, cur: the 4 x 4 grid of margin 2; the reference hard-codes these extents) */
double orc_scd_likelihood(const double *prev, const double *cur)
{
    double total = 0.0;
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i)
        {
            double best = 10000000.0;
            for (int s = j; s < j + 3; ++s)
                for (int k = i; k < i + 3; ++k)
                {
                    const double avg1 = prev[2 * (k + s * 6)], var1 = prev[2 * (k + s * 6) + 1];
                    const double avg2 = cur[2 * (i + j * 4)], var2 = cur[2 * (i + j * 4) + 1];
                    double t = (avg2 - avg1) / 2.0;
                    t = t * t;
                    const double tv = (var1 + var2) / 2.0;
                    t = (t + tv) * (t + tv);
                    const double l = t / (var1 * var2);
                    if (l < best) best = l;
                }
            total += best;
        }
    return total / 16;
}
