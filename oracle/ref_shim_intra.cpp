/*
 * oracle/ref_shim_intra.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" face over two header-only pieces of the UNMODIFIED reference encoder that the intra sweep relies on:
 * filterFlag (turing/Dsp.h:57-70) and IntraReferenceSamples<Sample>::filter (turing/IntraReferenceSamples.h:373-419).
 * The neighbour array convention of this repository (4n+1 samples, corner at index 2n, left(y) at 2n-1-y,
 * top(x) at 2n+1+x) is mapped onto the reference's p(x,-1) / p(-1,y) accessors; no algorithm lives here.
 */
#include "turing/IntraReferenceSamples.h"
#include "turing/Dsp.h"
#include <cstdint>

namespace {

template <typename Sample>
void run(Sample *out, const Sample *in, int n, int bitDepth, int strong)
{
    IntraReferenceSamples<Sample> p, f;
    const int c = 2 * n;
    p(-1, -1) = in[c];
    for (int k = 0; k < 2 * n; ++k)
    {
        p(-1, k) = in[c - 1 - k];
        p(k, -1) = in[c + 1 + k];
    }
    f.filter(p, strong, bitDepth, n);
    out[c] = f(-1, -1);
    for (int k = 0; k < 2 * n; ++k)
    {
        out[c - 1 - k] = f(-1, k);
        out[c + 1 + k] = f(k, -1);
    }
}

} // namespace

extern "C" int ref_intra_filter_flag(int cIdx, int mode, int nTbS)
{
    return filterFlag(cIdx, mode, nTbS) ? 1 : 0;
}

extern "C" void ref_intra_filter_neighbours(void *out, const void *in, int n, int bitDepth, int strong, int bps)
{
    if (bps == 1) run(static_cast<uint8_t *>(out), static_cast<const uint8_t *>(in), n, bitDepth, strong);
    else run(static_cast<uint16_t *>(out), static_cast<const uint16_t *>(in), n, bitDepth, strong);
}
