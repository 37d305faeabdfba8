/*
 * oracle/oracle_codeddata.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * CPU restatement of CodedData::storeResidual (turing/CodedData.h:457-517; SubBlock / Residual word layout :117-270; the
 * scans of turing/ScanOrder.h:32-101, :152-187): the record a transform block's quantised levels become in the encoder's
 * coded-data stream.  Layout, in uint16 words:
 *   [0]            transform-skip word (left 0 here: storeResidual does not write it)
 *   [1] or [1..4]  coded_sub_block flags, bit (i & 15) of word 1 + (i >> 4) for sub-block i in scan order (4 words at 32x32)
 *   then, for each SIGNIFICANT 4x4 sub-block from the last in scan order to the first:
 *       significance mask, greater-than-1 mask, sign mask (bit 15 - n for scan position n), and the magnitudes > 1 from
 *       scan position 15 down to 0.
 * Written from that description (two passes: size, then fill); pinned against the reference function by
 * tests/test_oracle_pin_codeddata.py through oracle/ref_shim_codeddata.cpp.
 */
#include "oracle.h"
#include <stdlib.h>

/* i-th position of the scan `scanIdx` (0 up-right diagonal, 1 horizontal, 2 vertical) of a size x size grid */
static void scan_pos(int size, int scanIdx, int i, int *x, int *y)
{
    if (scanIdx == 1)
    {
        *x = i % size, *y = i / size;
        return;
    }
    if (scanIdx == 2)
    {
        *x = i / size, *y = i % size;
        return;
    }
    for (int d = 0;; ++d)
    {
        /* anti-diagonal d: cells x + y = d inside the grid, walked from the bottom-left end upwards */
        const int yTop = d < size ? d : size - 1, count = d < size ? d + 1 : 2 * size - 1 - d;
        if (i < count)
        {
            *y = yTop - i, *x = d - *y;
            return;
        }
        i -= count;
    }
}

int orc_coded_residual(const int16_t *levels, int log2n, int scanIdx, uint16_t *out)
{
    const int n = 1 << log2n, grid = n >> 2, subBlocks = grid * grid, header = 1 + (log2n == 5 ? 4 : 1);
    int inner[16];
    for (int k = 0; k < 16; ++k)
    {
        int x, y;
        scan_pos(4, scanIdx, k, &x, &y);
        inner[k] = y * n + x;
    }
    int any = 0;
    for (int i = 0; i < n * n; ++i) any |= levels[i] != 0;
    if (!any) return 0;
    for (int i = 0; i < header; ++i) out[i] = 0;
    uint16_t *p = out + header;
    for (int i = subBlocks - 1; i >= 0; --i)
    {
        int sx, sy;
        scan_pos(grid, scanIdx, i, &sx, &sy);
        const int16_t *block = levels + (sy * 4) * n + sx * 4;
        uint16_t sig = 0, greater1 = 0, sign = 0, magnitudes[16];
        int count = 0;
        for (int k = 15; k >= 0; --k)
        {
            const int v = block[inner[k]];
            if (!v) continue;
            sig |= 1u << (15 - k);
            if (v < 0) sign |= 1u << (15 - k);
            if (abs(v) > 1)
            {
                greater1 |= 1u << (15 - k);
                magnitudes[count++] = (uint16_t)abs(v);
            }
        }
        if (!sig) continue;
        out[1 + (i >> 4)] |= 1u << (i & 15);
        *p++ = sig;
        *p++ = greater1;
        *p++ = sign;
        for (int k = 0; k < count; ++k) *p++ = magnitudes[k];
    }
    return (int)(p - out);
}
