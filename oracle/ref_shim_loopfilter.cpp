/*
 * oracle/ref_shim_loopfilter.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives the UNMODIFIED reference deblocking filter -- LoopFilter::Picture::deblock<edgeType> with its
 * LumaBlockEdge / ChromaBlockEdge workers (turing/LoopFilter.h:165-423, :739-777) -- in isolation, the way
 * TaskDeblock::run calls it per CTU (turing/TaskDeblock.cpp:104-127), so that oracle_loopfilter.c (and through it a
 * device deblocking pass) can be pinned against it.
 *
 * deblock() is a template over the encoder's handler type H; it is instantiated here with a stand-in that answers the
 * h[Tag()] questions it asks (bit depths, CTB geometry, chroma format, the PPS chroma QP offsets) from plain integers.
 * The per-8x8 Block records (QP, filter-disable bit, the four 2-bit boundary strengths) and the per-CTU slice offsets
 * are the reference's own structs, filled from the caller's arrays; no filter logic lives here.
 */
#include "turing/LoopFilter.h"
#include <cstdint>
#include <cstring>

namespace {

struct Stand
{
    int bitDepthY, bitDepthC, picWidthInCtbs, picHeightInCtbs, ctbLog2, cbOffset, crOffset;
    int operator[](BitDepthY) { return bitDepthY; }
    int operator[](BitDepthC) { return bitDepthC; }
    int operator[](PicOrderCntVal) { return 0; }
    int operator[](PicWidthInCtbsY) { return picWidthInCtbs; }
    int operator[](PicHeightInCtbsY) { return picHeightInCtbs; }
    int operator[](PicSizeInCtbsY) { return picWidthInCtbs * picHeightInCtbs; }
    int operator[](CtbLog2SizeY) { return ctbLog2; }
    int operator[](SubWidthC) { return 2; }
    int operator[](SubHeightC) { return 2; }
    int operator[](pps_cb_qp_offset) { return cbOffset; }
    int operator[](pps_cr_qp_offset) { return crOffset; }
};

template <typename Sample>
void run(void *const planes[3], const intptr_t strides[3], Stand &h, const uint8_t *blockData, const int8_t *ctuOffsets, int edgeType,
         int xBegin, int yBegin, int xEnd, int yEnd)
{
    LoopFilter::Picture lf(h);
    /* blockData: (data, packedBs) byte pairs in the reference's own grid: stride = (picture width in 8x8 blocks) + 1 */
    for (size_t i = 0; i < lf.blocks.size(); ++i)
    {
        lf.blocks[i].data = (int8_t)blockData[2 * i];
        lf.blocks[i].packedBs = blockData[2 * i + 1];
    }
    for (size_t i = 0; i < lf.ctus.size(); ++i)
    {
        std::memset(&lf.ctus[i], 0, sizeof(lf.ctus[i]));
        lf.ctus[i].tc_offset_div2 = ctuOffsets[2 * i];
        lf.ctus[i].beta_offset_div2 = ctuOffsets[2 * i + 1];
    }
    Raster<Sample> l(static_cast<Sample *>(planes[0]), strides[0]), cb(static_cast<Sample *>(planes[1]), strides[1]),
        cr(static_cast<Sample *>(planes[2]), strides[2]);
    if (edgeType == EDGE_VER)
        lf.deblock<EDGE_VER>(h, l, cb, cr, xBegin, yBegin, xEnd, yEnd);
    else
        lf.deblock<EDGE_HOR>(h, l, cb, cr, xBegin, yBegin, xEnd, yEnd);
}

} // namespace

/* grid geometry the caller must match: blocks per row, rows (turing/LoopFilter.h:436-443) */
extern "C" void ref_deblock_grid(int picWidthInCtbs, int picHeightInCtbs, int ctbLog2, int *stride, int *rows)
{
    *stride = (picWidthInCtbs << ctbLog2 >> 3) + 1;
    *rows = (picHeightInCtbs << ctbLog2 >> 3) + 1;
}

extern "C" void ref_deblock(void *const planes[3], const intptr_t strides[3], int bps, int bitDepthY, int bitDepthC, int picWidthInCtbs,
                            int picHeightInCtbs, int ctbLog2, int cbOffset, int crOffset, const uint8_t *blockData,
                            const int8_t *ctuOffsets, int edgeType, int xBegin, int yBegin, int xEnd, int yEnd)
{
    Stand h{bitDepthY, bitDepthC, picWidthInCtbs, picHeightInCtbs, ctbLog2, cbOffset, crOffset};
    if (bps == 1)
        run<uint8_t>(planes, strides, h, blockData, ctuOffsets, edgeType, xBegin, yBegin, xEnd, yEnd);
    else
        run<uint16_t>(planes, strides, h, blockData, ctuOffsets, edgeType, xBegin, yBegin, xEnd, yEnd);
}
