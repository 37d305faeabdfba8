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
