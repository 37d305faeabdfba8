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

/*
 * Adaptive quantisation and shot-change detection: the UNMODIFIED reference classes driven on a
 * caller's plane.  AdaptiveQuantisation::preAnalysis (turing/AdaptiveQuantisation.h:172-246) fills its layers from a
 * PictureWrap; ShotChangeDetection::getLikelihood (turing/SCDetection.h:71-178) works on packed byte vectors, and the
 * histogram lives inline in processSeq (:237-262), so the shim packs the plane exactly as processSeq does (:239-246,
 * 16-bit samples >> 2 into unsigned char) and lets the reference's own members do every computation that has one.
 */
#include "turing/AdaptiveQuantisation.h"
#include "turing/SCDetection.h"
#include <memory>
#include <vector>

template <typename Sample>
static std::shared_ptr<PictureWrapper> wrapPlane(const void *plane, intptr_t stride, int width, int height)
{
    auto picture = std::make_shared<PictureWrap<Sample>>(width, height, 1, 0, 0, 32);
    auto &luma = (*picture)[0];
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) luma(x, y) = static_cast<const Sample *>(plane)[y * stride + x];
    return picture;
}

/* activities of layer `depth` (unit = maxCuSize >> depth) as the reference stores them (doubles), and its average activity */
extern "C" int ref_aq_layer(const void *plane, intptr_t stride, int width, int height, int bps, int maxCuSize, int depth, int aqDepth, double *activity,
                            double *average)
{
    AdaptiveQuantisation aq(aqDepth, 6, height, width, maxCuSize);
    if (bps == 1)
        aq.preAnalysis<uint8_t>(wrapPlane<uint8_t>(plane, stride, width, height));
    else
        aq.preAnalysis<uint16_t>(wrapPlane<uint16_t>(plane, stride, width, height));
    AdaptiveQuantisationLayer &layer = aq.getAqLayerArray()[depth];
    const int n = layer.getPicHeightInPartUnits() * layer.getPicWidthInPartUnits();
    for (int i = 0; i < n; ++i) activity[i] = layer.getUnitArray()[i].getActivity();
    *average = layer.getAverageActivity();
    return n;
}

/* the QP offset the encoder derives from the layers (turing/AdaptiveQuantisation.h:147-169) */
extern "C" int ref_aq_offset(const void *plane, intptr_t stride, int width, int height, int bps, int maxCuSize, int aqDepth, int aqRange, int row, int col,
                             int depth)
{
    AdaptiveQuantisation aq(aqDepth, aqRange, height, width, maxCuSize);
    if (bps == 1)
        aq.preAnalysis<uint8_t>(wrapPlane<uint8_t>(plane, stride, width, height));
    else
        aq.preAnalysis<uint16_t>(wrapPlane<uint16_t>(plane, stride, width, height));
    return aq.getAqOffset(row, col, depth);
}

static ShotChangeDetection::FramePtr packPlane(const void *plane, intptr_t stride, int width, int height, int bps)
{
    auto frame = std::make_shared<std::vector<unsigned char>>((size_t)width * height);
    int z = 0;
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x)
            (*frame)[z++] = bps == 1 ? static_cast<const uint8_t *>(plane)[y * stride + x] : (static_cast<const uint16_t *>(plane)[y * stride + x] >> 2);
    return frame;
}

extern "C" double ref_scd_likelihood(const void *prev, const void *cur, intptr_t stride, int width, int height, int bps)
{
    ShotChangeDetection scd;
    return scd.getLikelihood(packPlane(prev, stride, width, height, bps), packPlane(cur, stride, width, height, bps), width, height);
}

/* the histogram loop of processSeq (:248-252) on the packed frame: `++cur[v >> SHIFT_DOWN]` with the reference's constant */
extern "C" void ref_scd_histogram(const void *plane, intptr_t stride, int width, int height, int bps, int *hist)
{
    ShotChangeDetection scd;
    auto frame = packPlane(plane, stride, width, height, bps);
    for (auto &v : *frame)
    {
        unsigned char s = v >> SHIFT_DOWN;
        ++(scd.cur)[s];
    }
    for (size_t i = 0; i < scd.cur.size(); ++i) hist[i] = scd.cur[i];
}
