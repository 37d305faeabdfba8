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
 *
 * ref_sao does the same for sample adaptive offset: LoopFilter::Picture::filterBlockSao (turing/LoopFilter.h:885-1017,
 * with sao_filter_edge / sao_filter_band of turing/sao.cpp and restoreUnfilteredRegions :849-877) per CTU and component,
 * in the order of applySaoCTU (:794-811), from one picture into another.
 */
#include "turing/LoopFilter.h"
#include <cstdint>
#include <cstring>

namespace {

struct Stand
{
    int bitDepthY, bitDepthC, picWidthInCtbs, picHeightInCtbs, ctbLog2, cbOffset, crOffset;
    int picWidth = 0, picHeight = 0;
    HavocTablePredUni<uint8_t> *pred8 = nullptr;
    HavocTablePredUni<uint16_t> *pred16 = nullptr;
    operator HavocTablePredUni<uint8_t> *() { return pred8; }
    operator HavocTablePredUni<uint16_t> *() { return pred16; }
    int operator[](pic_width_in_luma_samples) { return picWidth; }
    int operator[](pic_height_in_luma_samples) { return picHeight; }
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

/* One CTU's SAO parameters and neighbourhood as LoopFilter::Ctu holds them (turing/LoopFilter.h:92-163, :476-534);
 * layout shared with oracle.h's orc_sao_ctu / include/hvb.h's hvb_sao_ctu. */
struct SaoCtuRecord
{
    int16_t left, top, right, bottom;
    uint8_t topLeft, topRight, bottomLeft, bottomRight;
    struct
    {
        int8_t typeIdx, classOrBand;
        int16_t offset[4];
    } plane[3];
};

namespace {

template <typename Sample>
void runSao(void *const dst[3], void *const src[3], const intptr_t strides[3], Stand &h, const uint8_t *blockData, const SaoCtuRecord *rec,
            int lumaFlag, int chromaFlag)
{
    LoopFilter::Picture lf(h);
    for (size_t i = 0; i < lf.blocks.size(); ++i)
    {
        lf.blocks[i].data = (int8_t)blockData[2 * i];
        lf.blocks[i].packedBs = blockData[2 * i + 1];
    }
    for (size_t i = 0; i < lf.ctus.size(); ++i)
    {
        LoopFilter::Ctu &c = lf.ctus[i];
        std::memset(&c, 0, sizeof(c));
        c.left = rec[i].left, c.top = rec[i].top, c.right = rec[i].right, c.bottom = rec[i].bottom;
        c.topLeft = rec[i].topLeft, c.topRight = rec[i].topRight, c.bottomLeft = rec[i].bottomLeft, c.bottomRight = rec[i].bottomRight;
        for (int k = 0; k < 3; ++k)
        {
            c.planes[k].SaoTypeIdx = rec[i].plane[k].typeIdx;
            c.planes[k].u.eoClass = rec[i].plane[k].classOrBand; /* a union with saoLeftClass */
            for (int j = 0; j < 4; ++j) c.planes[k].SaoOffsetVal[j + 1] = rec[i].plane[k].offset[j];
        }
    }
    const int n = 1 << h.ctbLog2;
    for (int ry = 0; ry < h.picHeightInCtbs; ++ry)
        for (int rx = 0; rx < h.picWidthInCtbs; ++rx)
            for (int c = 0; c < 3; ++c)
            {
                if (!(c ? chromaFlag : lumaFlag)) continue;
                Raster<Sample> d(static_cast<Sample *>(dst[c]), strides[c]), s(static_cast<Sample *>(src[c]), strides[c]);
                lf.filterBlockSao<Sample>(h, d, s, rx, ry, c ? n / 2 : n, c ? n / 2 : n, c);
            }
}

} // namespace

extern "C" void ref_sao(void *const dst[3], void *const src[3], const intptr_t strides[3], int bps, int bitDepthY, int bitDepthC,
                        int picWidth, int picHeight, int ctbLog2, const uint8_t *blockData, const SaoCtuRecord *ctus, int lumaFlag,
                        int chromaFlag)
{
    const int n = 1 << ctbLog2;
    Stand h{bitDepthY, bitDepthC, (picWidth + n - 1) >> ctbLog2, (picHeight + n - 1) >> ctbLog2, ctbLog2, 0, 0};
    h.picWidth = picWidth;
    h.picHeight = picHeight;
    /* the SaoTypeIdx == 0 path copies through havoc's own pred_uni copy kernel (turing/LoopFilter.h:1009-1015) */
    havoc_code code = havoc_new_code((havoc_instruction_set)(HAVOC_C_REF | HAVOC_C_OPT), 2000000);
    HavocTablePredUni<uint8_t> t8;
    HavocTablePredUni<uint16_t> t16;
    havocPopulatePredUni<uint8_t>(&t8, code);
    havocPopulatePredUni<uint16_t>(&t16, code);
    h.pred8 = &t8;
    h.pred16 = &t16;
    if (bps == 1)
        runSao<uint8_t>(dst, src, strides, h, blockData, ctus, lumaFlag, chromaFlag);
    else
        runSao<uint16_t>(dst, src, strides, h, blockData, ctus, lumaFlag, chromaFlag);
    havoc_delete_code(code);
}
