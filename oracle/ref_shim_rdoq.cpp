/*
 * oracle/ref_shim_rdoq.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * extern "C" face over the UNMODIFIED reference RDOQ: it fills a reference `Contexts` object
 * (turing/Cabac.h:411-436) with the caller's context-state bytes, constructs the reference's
 * `Rdoq` engine exactly as ReconstructInterBlock does (turing/Reconstruct.cpp:806-808) and calls
 * Rdoq::runQuantisation (turing/Rdoq.cpp:35).  Rdoq.cpp and ScanOrder.cpp are compiled from where
 * they lie under /root/reference/turing by oracle/Makefile.  No algorithm lives here.
 */
#include "turing/Rdoq.h"
#include <cstdint>

extern "C" {

struct ref_rdoq_ctx /* same bytes as orc_rdoq_ctx / hvb_rdoq_ctx */
{
    uint8_t sig_coeff_flag[44];
    uint8_t greater1_flag[24];
    uint8_t greater2_flag[6];
    uint8_t coded_sub_block_flag[4];
    uint8_t last_x_prefix[18];
    uint8_t last_y_prefix[18];
    uint8_t cbf_luma[2];
    uint8_t cbf_cbcr[5];
    uint8_t rqt_root_cbf[1];
    uint8_t reserved[6];
    double lambda;
};

int ref_rdoq(int16_t *dst, const int16_t *src, const ref_rdoq_ctx *c, int quantiserScale, int quantiserShift,
             int invQuantScale, int log2n, int cIdx, int scanIdx, int isIntra, int sdh, int bitDepth)
{
    Contexts contexts;
    for (int i = 0; i < 44; ++i) contexts.get<sig_coeff_flag>(i).state = c->sig_coeff_flag[i];
    for (int i = 0; i < 24; ++i) contexts.get<coeff_abs_level_greater1_flag>(i).state = c->greater1_flag[i];
    for (int i = 0; i < 6; ++i) contexts.get<coeff_abs_level_greater2_flag>(i).state = c->greater2_flag[i];
    for (int i = 0; i < 4; ++i) contexts.get<coded_sub_block_flag>(i).state = c->coded_sub_block_flag[i];
    for (int i = 0; i < 18; ++i) contexts.get<last_sig_coeff_x_prefix>(i).state = c->last_x_prefix[i];
    for (int i = 0; i < 18; ++i) contexts.get<last_sig_coeff_y_prefix>(i).state = c->last_y_prefix[i];
    for (int i = 0; i < 2; ++i) contexts.get<cbf_luma>(i).state = c->cbf_luma[i];
    for (int i = 0; i < 4; ++i) contexts.get<cbf_cX>(i).state = c->cbf_cbcr[i];
    contexts.get<rqt_root_cbf>(0).state = c->rqt_root_cbf[0];

    Rdoq engine(c->lambda, &contexts, quantiserScale, invQuantScale, log2n, bitDepth);
    residual_coding rc(0, 0, log2n, cIdx);
    return engine.runQuantisation(dst, src, quantiserScale, quantiserShift, 1 << (2 * log2n), rc, scanIdx,
                                  isIntra != 0, sdh != 0);
}

} // extern "C"
