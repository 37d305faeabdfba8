/*
 * oracle/ref_shim_saostats.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" face over the UNMODIFIED reference encoder's SAO statistics -- the static members of class EncSao
 * (turing/EncSao.h:151-284 edge_offset_stats_class0..3, :111-149 band_offset_luma_stats), called exactly as
 * saoRdEstimateLuma / saoRdEstimateChroma call them per CTU and plane (:328-478, :581-742) -- so that
 * oracle_loopfilter.c's orc_sao_stats (and through it the device statistics pass) can be pinned against them.
 * EncSao.h is included the way the encoder includes it (through Write.h); no statistics logic lives here.
 */
#include "turing/Write.h"
#include <cstdint>

namespace {

template <typename Sample>
int run(const void *org, intptr_t strideOrg, const void *rec, intptr_t strideRec, int w, int h, int shift, int64_t *out)
{
    const Sample *o = static_cast<const Sample *>(org), *r = static_cast<const Sample *>(rec);
    /* out: [class 0..3][E[5], count[5]], then band E[32], band count[32] */
    EncSao::edge_offset_stats_class0<Sample>(o, strideOrg, r, strideRec, out + 0, out + 5, h, w);
    EncSao::edge_offset_stats_class1<Sample>(o, strideOrg, r, strideRec, out + 10, out + 15, h, w);
    EncSao::edge_offset_stats_class2<Sample>(o, strideOrg, r, strideRec, out + 20, out + 25, h, w);
    EncSao::edge_offset_stats_class3<Sample>(o, strideOrg, r, strideRec, out + 30, out + 35, h, w);
    for (int b = 0; b < 64; ++b) out[40 + b] = 0; /* as :472-476 */
    return EncSao::band_offset_luma_stats<Sample>(o, strideOrg, r, strideRec, out + 40, out + 72, h, w, shift);
}

} // namespace

/* returns the reference's startBand for the block (EncSao.h:137-148) */
extern "C" int ref_sao_stats(const void *org, intptr_t strideOrg, const void *rec, intptr_t strideRec, int w, int h, int shift, int bps,
                             int64_t out[104])
{
    return bps == 1 ? run<uint8_t>(org, strideOrg, rec, strideRec, w, h, shift, out) : run<uint16_t>(org, strideOrg, rec, strideRec, w, h, shift, out);
}
