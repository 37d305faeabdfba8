/*
 * oracle/ref_shim_preanalysis.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" face over the UNMODIFIED reference's intra-complexity measure, EstimateIntraComplexity::computeSatd8x8
 * (turing/EstimateIntraComplexity.h:55-157): the AC Hadamard energy of an 8x8 block of source samples, which preAnalysis
 * (:159-176) evaluates for every 8x8 luma block of an intra picture for the rate control.  No arithmetic lives here.
 */
#include "turing/Picture.h"
#include "turing/EstimateIntraComplexity.h"
#include <cstdint>

extern "C" int ref_intra_complexity_8x8(const void *p, intptr_t stride, int bps)
{
    return bps == 1 ? EstimateIntraComplexity::computeSatd8x8<uint8_t>(static_cast<const uint8_t *>(p), stride)
                    : EstimateIntraComplexity::computeSatd8x8<uint16_t>(static_cast<const uint16_t *>(p), stride);
}
