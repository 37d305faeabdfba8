/*
 * oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * A thin extern "C" face over the UNMODIFIED reference havoc library, so that Python tests
 * (ctypes) and the bench's CPU-baseline leg can call the reference's own function tables.
 * It contains no algorithm: it builds the reference's `StateFunctionTables`
 * (/root/reference/turing/StateFunctionTables.h:37-102, included from where it lies) exactly as
 * the encoder does and forwards each call to the table entry the encoder would fetch.
 * Built by oracle/Makefile into oracle/_ref/libhavoc_ref.so together with the reference's
 * havoc/*.cpp objects.  `use_asm` = 0 gives the HAVOC_C_REF|HAVOC_C_OPT tables (the `--asm 0`
 * identity oracle, turing/StateEncode.h:753-759); 1 gives the xbyak-JIT tables the CPU
 * supports (the speed baseline).
 */
#include "turing/StateFunctionTables.h"
#include "havoc/diff.h"
#include <cstdint>
#include <cstring>

namespace {

struct Ref
{
    StateFunctionTables tables;
    Ref(havoc_instruction_set mask) : tables(true, mask) {}
};

template <class T, class U> T *as(U *u) { return static_cast<T *>(u); }

} // namespace

extern "C" {

void *ref_create(int use_asm)
{
    havoc_instruction_set mask = use_asm ? havoc_instruction_set_support()
                                         : (havoc_instruction_set)(HAVOC_C_REF | HAVOC_C_OPT);
    return new Ref(mask);
}

void ref_destroy(void *h) { delete static_cast<Ref *>(h); }

unsigned ref_isa(void *h) { return (unsigned)static_cast<Ref *>(h)->tables.instruction_set_support; }

int ref_sad(void *h, const void *src, intptr_t ss, const void *ref, intptr_t sr, int w, int hgt, int bps)
{
    auto &t = static_cast<Ref *>(h)->tables;
    if (bps == 1)
        return (*havoc_get_sad<uint8_t>(&t, w, hgt))((const uint8_t *)src, ss, (const uint8_t *)ref, sr, HAVOC_RECT(w, hgt));
    return (*havoc_get_sad<uint16_t>(&t, w, hgt))((const uint16_t *)src, ss, (const uint16_t *)ref, sr, HAVOC_RECT(w, hgt));
}

void ref_sad_multiref4(void *h, const void *src, intptr_t ss, const void *const ref[4], intptr_t sr,
                       int sad[4], int w, int hgt, int bps)
{
    auto &t = static_cast<Ref *>(h)->tables;
    if (bps == 1)
        (*havoc_get_sad_multiref<uint8_t>(&t, 4, w, hgt))((const uint8_t *)src, ss, (const uint8_t **)ref, sr, sad, HAVOC_RECT(w, hgt));
    else
        (*havoc_get_sad_multiref<uint16_t>(&t, 4, w, hgt))((const uint16_t *)src, ss, (const uint16_t **)ref, sr, sad, HAVOC_RECT(w, hgt));
}

uint32_t ref_ssd(void *h, const void *a, intptr_t sa, const void *b, intptr_t sb, int log2n, int bps)
{
    auto &t = static_cast<Ref *>(h)->tables;
    const int n = 1 << log2n;
    if (bps == 1)
        return (*havoc_get_ssd<uint8_t>(&t, log2n))((const uint8_t *)a, sa, (const uint8_t *)b, sb, n, n);
    return (*havoc_get_ssd<uint16_t>(&t, log2n))((const uint16_t *)a, sa, (const uint16_t *)b, sb, n, n);
}

int ref_ssd_linear(void *h, const uint8_t *a, const uint8_t *b, int n)
{
    return havoc_get_ssd_linear(n, static_cast<Ref *>(h)->tables.code)(a, b, n);
}

int ref_hadamard_satd(void *h, const void *a, intptr_t sa, const void *b, intptr_t sb, int log2n, int bps)
{
    auto &t = static_cast<Ref *>(h)->tables;
    if (bps == 1)
        return (*havoc_get_hadamard_satd<uint8_t>(&t, log2n))((const uint8_t *)a, sa, (const uint8_t *)b, sb);
    return (*havoc_get_hadamard_satd<uint16_t>(&t, log2n))((const uint16_t *)a, sa, (const uint16_t *)b, sb);
}

/* returns 0 when the table slot is empty (populate leaves unsupported slots null) */
int ref_pred_uni(void *h, void *dst, intptr_t sd, const void *ref, intptr_t sr, int w, int hgt,
                 int xFrac, int yFrac, int bitDepth, int taps, int bps)
{
    auto &t = static_cast<Ref *>(h)->tables;
    if (bps == 1)
    {
        auto *f = *havocGetPredUni<uint8_t>(&t, taps, w, hgt, xFrac, yFrac, bitDepth);
        if (!f) return 0;
        f((uint8_t *)dst, sd, (const uint8_t *)ref, sr, w, hgt, xFrac, yFrac, bitDepth);
    }
    else
    {
        auto *f = *havocGetPredUni<uint16_t>(&t, taps, w, hgt, xFrac, yFrac, bitDepth);
        if (!f) return 0;
        f((uint16_t *)dst, sd, (const uint16_t *)ref, sr, w, hgt, xFrac, yFrac, bitDepth);
    }
    return 1;
}

int ref_pred_bi(void *h, void *dst, intptr_t sd, const void *ref0, const void *ref1, intptr_t sr, int w,
                int hgt, int xf0, int yf0, int xf1, int yf1, int bitDepth, int taps, int bps)
{
    auto &t = static_cast<Ref *>(h)->tables;
    if (bps == 1)
    {
        auto *f = *havocGetPredBi<uint8_t>(&t, taps, w, hgt, xf0, yf0, xf1, yf1, bitDepth);
        if (!f) return 0;
        f((uint8_t *)dst, sd, (const uint8_t *)ref0, (const uint8_t *)ref1, sr, w, hgt, xf0, yf0, xf1, yf1, bitDepth);
    }
    else
    {
        auto *f = *havocGetPredBi<uint16_t>(&t, taps, w, hgt, xf0, yf0, xf1, yf1, bitDepth);
        if (!f) return 0;
        f((uint16_t *)dst, sd, (const uint16_t *)ref0, (const uint16_t *)ref1, sr, w, hgt, xf0, yf0, xf1, yf1, bitDepth);
    }
    return 1;
}

void ref_subtract_bi(void *h, void *dst, intptr_t sd, const void *pred, intptr_t sp, const void *src,
                     intptr_t ss, int w, int hgt, int bitDepth, int bps)
{
    auto &t = static_cast<Ref *>(h)->tables;
    if (bps == 1)
        as<havoc::TableSubtractBi<uint8_t>>(&t)->get()((uint8_t *)dst, sd, (const uint8_t *)pred, sp, (const uint8_t *)src, ss, w, hgt, bitDepth);
    else
        as<havoc::TableSubtractBi<uint16_t>>(&t)->get()((uint16_t *)dst, sd, (const uint16_t *)pred, sp, (const uint16_t *)src, ss, w, hgt, bitDepth);
}

/* cIdx selects the edge-filter aliasing exactly as turing does (pred_intra.h:41-49) */
int ref_pred_intra(void *h, void *dst, intptr_t sd, const void *neighbours, int mode, int log2n,
                   int bitDepth, int cIdx, int bps)
{
    auto &t = static_cast<Ref *>(h)->tables;
    if (bps == 1)
    {
        auto *f = as<havoc::intra::Table<uint8_t>>(&t)->lookup(cIdx, bitDepth, log2n, mode);
        if (!f) return 0;
        f((uint8_t *)dst, sd, (const uint8_t *)neighbours, mode);
    }
    else
    {
        auto *f = as<havoc::intra::Table<uint16_t>>(&t)->lookup(cIdx, bitDepth, log2n, mode);
        if (!f) return 0;
        f((uint16_t *)dst, sd, (const uint16_t *)neighbours, mode);
    }
    return 1;
}

void ref_transform_fwd(void *h, int16_t *coeffs, const int16_t *src, intptr_t stride, int trType,
                       int log2n, int bitDepth)
{
    auto &t = static_cast<Ref *>(h)->tables;
    if (bitDepth == 8)
        (*havoc::get_transform<8>(&t, trType, log2n))(coeffs, src, stride);
    else
        (*havoc::get_transform<10>(&t, trType, log2n))(coeffs, src, stride);
}

void ref_inverse_transform_add(void *h, void *dst, intptr_t sd, const void *pred, intptr_t sp,
                               const int16_t *coeffs, int trType, int log2n, int bitDepth, int bps)
{
    auto &t = static_cast<Ref *>(h)->tables;
    if (bps == 1)
        (*havoc::get_inverse_transform_add<uint8_t>(&t, trType, log2n))((uint8_t *)dst, sd, (const uint8_t *)pred, sp, coeffs, bitDepth);
    else
        (*havoc::get_inverse_transform_add<uint16_t>(&t, trType, log2n))((uint16_t *)dst, sd, (const uint16_t *)pred, sp, coeffs, bitDepth);
}

void ref_inverse_transform(void *h, int16_t *res, const int16_t *coeffs, int trType, int log2n, int bitDepth)
{
    auto &t = static_cast<Ref *>(h)->tables;
    (*havoc::get_inverse_transform(&t, trType, log2n))(res, coeffs, bitDepth);
}

int ref_quantize(void *h, int16_t *dst, const int16_t *src, int scale, int shift, int offset, int n)
{
    auto &t = static_cast<Ref *>(h)->tables;
    return (*havoc_get_quantize(&t))(dst, src, scale, shift, offset, n);
}

void ref_quantize_inverse(void *h, int16_t *dst, const int16_t *src, int scale, int shift, int n)
{
    auto &t = static_cast<Ref *>(h)->tables;
    (*havoc_get_quantize_inverse(&t, scale, shift))(dst, src, scale, shift, n);
}

void ref_quantize_reconstruct(void *h, uint8_t *rec, intptr_t sr, const uint8_t *pred, intptr_t sp,
                              const int16_t *res, int log2n)
{
    auto &t = static_cast<Ref *>(h)->tables;
    (*havoc_get_quantize_reconstruct(&t, log2n))(rec, sr, pred, sp, res, 1 << log2n);
}

} // extern "C"
