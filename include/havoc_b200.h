/*
 * havoc_b200.h -- the LITERAL drop-in boundary: the function-table types and populate entry points that
 * the reference encoder's `StateFunctionTables` (turing/StateFunctionTables.h:37-102) links against,
 * re-declared so that csrc/havoc_b200.cpp can define them on top of the batched C-ABI (hvb.h).
 *
 * Building the reference's turing/*.cpp against its own havoc headers and linking the objects with
 * libhavoc_b200.so instead of libhavoc.a gives an encoder whose every pixel primitive runs on the B200
 * (one tiny batch per call -- correct and slow; the fast path is the batched ABI).
 *
 * ABI contract: type names, member order and array extents below mirror the reference headers, because
 * the reference's inline getters index straight into these structs:
 *   havoc/sad.h:28-115        havoc_table_sad, havoc_table_sad_multiref
 *   havoc/ssd.h:36-48         havoc_table_ssd
 *   havoc/hadamard.h:35-46    havoc_table_hadamard_satd
 *   havoc/pred_inter.h:39-98  HavocTablePredUni, HavocTablePredBi, havoc::TableSubtractBi
 *   havoc/pred_intra.h:34-55  havoc::intra::Table
 *   havoc/transform.h:36-140  havoc::table_inverse_transform(_add), havoc::table_transform
 *   havoc/quantize.h:44-95    havoc_table_quantize_inverse / _quantize / _quantize_reconstruct
 *   havoc/havoc.h:107-149     havoc_instruction_set, havoc_code
 * A populate leaves a slot null where the reference would (callers do not check, havoc/README.md).
 */
#ifndef HAVOC_B200_H
#define HAVOC_B200_H

#include <stdint.h>
#include <stddef.h>
#include <stdio.h>

extern "C" {

typedef enum
{
    HAVOC_NONE = 0,
    HAVOC_C_REF = 1 << 0,
    HAVOC_C_OPT = 1 << 1,
    HAVOC_SSE2 = 1 << 2,
    HAVOC_SSE3 = 1 << 3,
    HAVOC_SSSE3 = 1 << 4,
    HAVOC_SSE41 = 1 << 5,
    HAVOC_SSE42 = 1 << 6,
    HAVOC_LZCNT = 1 << 7,
    HAVOC_POPCNT = 1 << 8,
    HAVOC_AVX = 1 << 9,
    HAVOC_AVX2 = 1 << 10
} havoc_instruction_set;

typedef struct
{
    void *implementation; /* here: the per-encoder B200 state instead of a JIT buffer */
} havoc_code;

havoc_instruction_set havoc_instruction_set_support();
void havoc_print_instruction_set_support(FILE *f, havoc_instruction_set mask);
havoc_code havoc_new_code(havoc_instruction_set mask, int size);
void havoc_delete_code(havoc_code);
int havoc_main(int argc, const char *argv[]);

typedef void havoc_quantize_inverse(int16_t *dst, const int16_t *src, int scale, int shift, int n);
typedef struct { havoc_quantize_inverse *p[2]; } havoc_table_quantize_inverse;
void havoc_populate_quantize_inverse(havoc_table_quantize_inverse *table, havoc_code code);

typedef int havoc_quantize(int16_t *dst, const int16_t *src, int scale, int shift, int offset, int n);
typedef struct { havoc_quantize *p; } havoc_table_quantize;
void havoc_populate_quantize(havoc_table_quantize *table, havoc_code code);

typedef void havoc_quantize_reconstruct(uint8_t *rec, intptr_t stride_rec, const uint8_t *pred, intptr_t stride_pred, const int16_t *res, int n);
typedef struct { havoc_quantize_reconstruct *p[4]; } havoc_table_quantize_reconstruct;
void havoc_populate_quantize_reconstruct(havoc_table_quantize_reconstruct *table, havoc_code code);

typedef int havoc_ssd_linear(const uint8_t *src0, const uint8_t *src1, int size);
havoc_ssd_linear *havoc_get_ssd_linear(int size, havoc_code code);

} /* extern "C" */

/* ---- SAD ---------------------------------------------------------------------------------- */
template <typename Sample>
using havoc_sad = int(const Sample *src, intptr_t stride_src, const Sample *ref, intptr_t stride_ref, uint32_t rect);

/* the 23 HEVC PU shapes in the reference's order, then the generic slot */
#define HAVOC_B200_PU_SIZES(X) \
    X(64, 64) X(64, 48) X(64, 32) X(64, 16) X(48, 64) X(32, 64) X(32, 32) X(32, 24) X(32, 16) X(32, 8) X(24, 32) X(16, 64) \
    X(16, 32) X(16, 16) X(16, 12) X(16, 8) X(16, 4) X(12, 16) X(8, 32) X(8, 16) X(8, 8) X(8, 4) X(4, 8)

template <typename Sample>
struct havoc_table_sad
{
#define X(w, h) havoc_sad<Sample> *sad##w##x##h;
    HAVOC_B200_PU_SIZES(X)
#undef X
    havoc_sad<Sample> *sadGeneric;
};
template <typename Sample>
void havoc_populate_sad(havoc_table_sad<Sample> *table, havoc_code code);

template <typename Sample>
using havoc_sad_multiref = void(const Sample *src, intptr_t stride_src, const Sample *ref[], intptr_t stride_ref, int sad[], uint32_t rect);
template <typename Sample>
struct havoc_table_sad_multiref
{
    havoc_sad_multiref<Sample> *lookup[16][16];
    havoc_sad_multiref<Sample> *sadGeneric_4;
};
template <typename Sample>
void havoc_populate_sad_multiref(havoc_table_sad_multiref<Sample> *table, havoc_code code);

/* ---- SSD / SATD ---------------------------------------------------------------------------- */
template <typename Sample>
using havoc_ssd = uint32_t(Sample const *srcA, intptr_t stride_srcA, Sample const *srcB, intptr_t stride_srcB, int w, int h);
template <typename Sample>
struct havoc_table_ssd { havoc_ssd<Sample> *ssd[5]; };
template <typename Sample>
void havoc_populate_ssd(havoc_table_ssd<Sample> *table, havoc_code code);

template <typename Sample>
using havoc_hadamard_satd = int(Sample const *srcA, intptr_t stride_srcA, Sample const *srcB, intptr_t stride_srcB);
template <typename Sample>
struct havoc_table_hadamard_satd { havoc_hadamard_satd<Sample> *satd[3]; };
template <typename Sample>
void havoc_populate_hadamard_satd(havoc_table_hadamard_satd<Sample> *table, havoc_code code);

/* ---- inter prediction ------------------------------------------------------------------------ */
template <typename Sample>
using HavocPredUni = void(Sample *dst, intptr_t stride_dst, Sample const *ref, intptr_t stride_ref, int nPbW, int nPbH, int xFrac, int yFrac, int bitDepth);
template <typename Sample>
struct HavocTablePredUni { HavocPredUni<Sample> *p[3][2][17][2][2]; };
template <typename Sample>
void havocPopulatePredUni(HavocTablePredUni<Sample> *table, havoc_code code);

template <typename Sample>
using HavocPredBi = void(Sample *dst0, intptr_t stride_dst, const Sample *ref0, const Sample *ref1, intptr_t stride_ref, int nPbW, int nPbH,
                         int xFrac0, int yFrac0, int xFrac1, int yFrac1, int bitDepth);
template <typename Sample>
struct HavocTablePredBi { HavocPredBi<Sample> *p[3][2][9][2]; };
template <typename Sample>
void havocPopulatePredBi(HavocTablePredBi<Sample> *table, havoc_code code);

namespace havoc {

template <typename Sample>
using SubtractBi = void(Sample *dst0, intptr_t stride_dst, const Sample *ref0, intptr_t stride_ref, const Sample *src, intptr_t stride_src,
                        int nPbW, int nPbH, int bitDepth);
template <typename Sample>
struct TableSubtractBi
{
    SubtractBi<Sample> *p;
    SubtractBi<Sample> *&get() { return this->p; }
};
template <typename Sample>
void populateSubtractBi(TableSubtractBi<Sample> *table, havoc_code code, int bitDepth = 0);

/* ---- intra prediction ------------------------------------------------------------------------ */
namespace intra {
template <typename Sample>
using Function = void(Sample *dst, intptr_t dstStride, Sample const *neighbours, int predModeIntra);
template <typename Sample>
struct Table
{
    Function<Sample> *entries[3 * sizeof(Sample) - 2][4][38];
    void populate(havoc_code code);
};
} // namespace intra

/* ---- transforms -------------------------------------------------------------------------------- */
using inverse_transform = void(int16_t dst[], int16_t const coeffs[], int bitDepth);
struct table_inverse_transform
{
    inverse_transform *sine;
    inverse_transform *cosine[4];
};
void populate_inverse_transform(table_inverse_transform *table, havoc_code code, int encoder);

template <typename Sample>
using inverse_transform_add = void(Sample *dst, intptr_t stride_dst, Sample const *pred, intptr_t stride_pred, int16_t const coeffs[], int bitDepth);
template <typename Sample>
struct table_inverse_transform_add
{
    inverse_transform_add<Sample> *sine;
    inverse_transform_add<Sample> *cosine[4];
};
template <typename Sample>
void populate_inverse_transform_add(table_inverse_transform_add<Sample> *table, havoc_code code, int encoder);

typedef void Transform(int16_t *coeffs, const int16_t *src, intptr_t src_stride);
template <int bitDepth>
struct table_transform
{
    Transform *dst;
    Transform *dct[4];
};
template <int bitDepth>
void populate_transform(table_transform<bitDepth> *table, havoc_code code);

} // namespace havoc

#endif
