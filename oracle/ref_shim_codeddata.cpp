/*
 * oracle/ref_shim_codeddata.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" face over the UNMODIFIED reference's CodedData::storeResidual (turing/CodedData.h:457-517), the function that
 * serialises a transform block's quantised levels into the encoder's coded-data stream (coded-sub-block flags, then per
 * significant 4x4 sub-block in reverse scan order: significance / greater-than-1 / sign masks and the magnitudes above 1)
 * for the CABAC writer and the rate estimator to read.  The coding-unit and transform-tree words it also touches (cbf bits)
 * are dummies here; no serialisation logic lives in this file.
 */
#include "turing/Picture.h" /* Raster, which CodedData.h uses without including it */
#include "turing/CodedData.h"
#include <cstdint>
#include <cstring>

/* returns the number of uint16 words the record occupies (0 for an all-zero block: storeResidual is only called with cbf) */
extern "C" int ref_coded_residual(int16_t *coefficients, int log2TrafoSize, int scanIdx, uint16_t *out, int capacityWords)
{
    bool any = false;
    for (int i = 0; i < (1 << (2 * log2TrafoSize)); ++i) any |= coefficients[i] != 0;
    if (!any) return 0;
    std::memset(out, 0, sizeof(uint16_t) * capacityWords);
    CodedData::Type cuWords[4] = {0, 0, 0, 0}, ttWords[4] = {0, 0, 0, 0};
    CodedData::CodingUnit cu;
    cu.p = cuWords;
    CodedData::TransformTree tt;
    tt.p = ttWords;
    CodedData::Residual residual;
    residual.p = out;
    CodedData::storeResidual(cu, residual, coefficients, log2TrafoSize, scanIdx, true, tt, 0);
    return (int)(residual.p - out);
}
