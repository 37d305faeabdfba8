// havoc_b200.cpp -- the literal drop-in: the reference's havoc populate entry points, defined on top of
// the batched C-ABI.  Every table slot gets a host function with the reference's signature
// (include/havoc_b200.h) that stages its operands into small device pictures, issues a batch of ONE
// through hvb_*, and copies the result back.  It exists to prove the boundary: the unmodified encoder
// objects link against it and every pixel primitive then runs on the B200 (the test-side `encoder` make
// target, tests/test_gpu_dropin.py).  It is not the fast path -- that is the batched ABI.
//
// Table filling mirrors turing/StateFunctionTables.h:63-92 and the populate bodies it calls
// (havoc/sad.cpp:494-504, :1006-1017; ssd.cpp:155-170; hadamard.cpp:747-760; pred_inter.cpp:913-1016,
// :1846-1883, :2082-2094; pred_intra.cpp:20403-20434; transform.cpp:2861-2962, :5262-5272;
// quantize.cpp:164-190, :425-448, :686-700).
#include "../../include/havoc_b200.h"
#include "../../include/hvb.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace {

// One staging area per (host thread, sample size, bit depth): the reference calls the tables concurrently
// from its pool threads (turing/ThreadPool.cpp:87-103) on disjoint data.
struct Stage
{
    hvb_context *ctx = nullptr;
    int pic[4] = {-1, -1, -1, -1}; // 160x160 luma staging pictures
    int bps = 0, bitDepth = 0;
};

constexpr int kStageSize = 160, kStagePad = 16, kOrigin = 16; // blocks are staged at (kOrigin, kOrigin)

[[noreturn]] void die(const char *what, hvb_context *ctx)
{
    std::fprintf(stderr, "havoc_b200: %s failed: %s\n", what, ctx ? hvb_last_error(ctx) : "no context");
    std::abort(); // the reference's primitives cannot fail; there is no CPU fallback to fall back to
}

Stage &stage(int bps, int bitDepth)
{
    static thread_local Stage stages[2][3];
    Stage &s = stages[bps - 1][bitDepth - 8];
    if (!s.ctx)
    {
        const char *dev = std::getenv("HVB_DEVICE");
        if (hvb_create(dev ? std::atoi(dev) : 0, bps, bitDepth, &s.ctx) != HVB_OK) die("hvb_create (is a B200 visible?)", nullptr);
        for (int &p : s.pic)
            if (hvb_picture_create(s.ctx, kStageSize, kStageSize, kStagePad, &p) != HVB_OK) die("hvb_picture_create", s.ctx);
        s.bps = bps;
        s.bitDepth = bitDepth;
    }
    return s;
}

#define CHECK(call, s)                                  \
    do                                                  \
    {                                                   \
        if ((call) != HVB_OK) die(#call, (s).ctx);      \
    } while (0)

template <typename Sample>
void put(Stage &s, int pic, int cIdx, const Sample *p, intptr_t stride, int x, int y, int w, int h)
{
    CHECK(hvb_picture_upload_rect(s.ctx, s.pic[pic], cIdx, p, stride, x, y, w, h), s);
}

template <typename Sample>
void get(Stage &s, int pic, int cIdx, Sample *p, intptr_t stride, int x, int y, int w, int h)
{
    CHECK(hvb_picture_download_rect(s.ctx, s.pic[pic], cIdx, p, stride, x, y, w, h), s);
}

hvb_block blk(Stage &s, int pic, int cIdx, int x, int y) { return hvb_block{(int16_t)s.pic[pic], (int16_t)cIdx, (int16_t)x, (int16_t)y}; }

template <typename Sample>
constexpr int defaultDepth() { return sizeof(Sample) == 1 ? 8 : 10; }

// ---- metrics ---------------------------------------------------------------------------------------
template <typename Sample>
hvb_metric_task stageMetric(Stage &s, const Sample *a, intptr_t sa, const Sample *b, intptr_t sb, int w, int h)
{
    put(s, 0, 0, a, sa, kOrigin, kOrigin, w, h);
    put(s, 1, 0, b, sb, kOrigin, kOrigin, w, h);
    hvb_metric_task t{};
    t.a = blk(s, 0, 0, kOrigin, kOrigin);
    t.b = blk(s, 1, 0, kOrigin, kOrigin);
    t.w = (int16_t)w;
    t.h = (int16_t)h;
    return t;
}

template <typename Sample>
int sadShim(const Sample *src, intptr_t ss, const Sample *ref, intptr_t sr, uint32_t rect)
{
    Stage &s = stage(sizeof(Sample), defaultDepth<Sample>());
    const hvb_metric_task t = stageMetric(s, src, ss, ref, sr, (int)(rect >> 8), (int)(rect & 0xff));
    int32_t out = 0;
    CHECK(hvb_sad_batch(s.ctx, &t, 1, &out, HVB_HOST), s);
    return out;
}

template <typename Sample>
void sad4Shim(const Sample *src, intptr_t ss, const Sample *ref[], intptr_t sr, int sad[], uint32_t rect)
{
    Stage &s = stage(sizeof(Sample), defaultDepth<Sample>());
    const int w = (int)(rect >> 8), h = (int)(rect & 0xff);
    put(s, 0, 0, src, ss, kOrigin, kOrigin, w, h);
    hvb_metric_task t[4]{};
    for (int k = 0; k < 4; ++k)
    {
        put(s, 1, 0, ref[k], sr, kOrigin, kOrigin, w, h);
        // one candidate at a time through the same staging block keeps the shim trivially correct
        t[k].a = blk(s, 0, 0, kOrigin, kOrigin);
        t[k].b = blk(s, 1, 0, kOrigin, kOrigin);
        t[k].w = (int16_t)w;
        t[k].h = (int16_t)h;
        int32_t out = 0;
        CHECK(hvb_sad_batch(s.ctx, &t[k], 1, &out, HVB_HOST), s);
        sad[k] = out;
    }
}

template <typename Sample>
uint32_t ssdShim(const Sample *a, intptr_t sa, const Sample *b, intptr_t sb, int w, int h)
{
    Stage &s = stage(sizeof(Sample), defaultDepth<Sample>());
    const hvb_metric_task t = stageMetric(s, a, sa, b, sb, w, h);
    uint32_t out = 0;
    CHECK(hvb_ssd_batch(s.ctx, &t, 1, &out, HVB_HOST), s);
    return out;
}

template <typename Sample, int LOG2N>
int satdShim(const Sample *a, intptr_t sa, const Sample *b, intptr_t sb)
{
    Stage &s = stage(sizeof(Sample), defaultDepth<Sample>());
    const hvb_metric_task t = stageMetric(s, a, sa, b, sb, 1 << LOG2N, 1 << LOG2N);
    int32_t out = 0;
    CHECK(hvb_satd_batch(s.ctx, &t, 1, &out, HVB_HOST), s);
    return out;
}

int ssdLinearShim(const uint8_t *a, const uint8_t *b, int n)
{
    // havoc/diff.cpp:29-38 over n samples: rows of 64 through the block SSD (8-bit: no post-shift, sums < 2^31 for n <= 2^15)
    Stage &s = stage(1, 8);
    int total = 0;
    for (int done = 0; done < n;)
    {
        const int w = n - done >= 64 ? 64 : n - done, rows = w == 64 ? (n - done) / 64 > 64 ? 64 : (n - done) / 64 : 1;
        put(s, 0, 0, a + done, 64, kOrigin, kOrigin, w, rows);
        put(s, 1, 0, b + done, 64, kOrigin, kOrigin, w, rows);
        hvb_metric_task t{};
        t.a = blk(s, 0, 0, kOrigin, kOrigin);
        t.b = blk(s, 1, 0, kOrigin, kOrigin);
        t.w = (int16_t)w;
        t.h = (int16_t)rows;
        uint32_t out = 0;
        CHECK(hvb_ssd_batch(s.ctx, &t, 1, &out, HVB_HOST), s);
        total += (int)out;
        done += w * rows;
    }
    return total;
}

// ---- inter prediction ---------------------------------------------------------------------------------
template <typename Sample, int TAPS>
void predUniShim(Sample *dst, intptr_t sd, const Sample *ref, intptr_t sr, int w, int h, int xFrac, int yFrac, int bitDepth)
{
    Stage &s = stage(sizeof(Sample), bitDepth);
    constexpr int M = TAPS / 2 - 1, cIdx = TAPS == 8 ? 0 : 1;
    put(s, 1, cIdx, ref - M * sr - M, sr, kOrigin - M, kOrigin - M, w + TAPS - 1, h + TAPS - 1);
    hvb_pred_task t{};
    t.dst = blk(s, 0, cIdx, kOrigin, kOrigin);
    t.ref_pic[0] = (int16_t)s.pic[1];
    t.ref_pic[1] = -1;
    t.x = t.y = kOrigin;
    t.w = (int16_t)w;
    t.h = (int16_t)h;
    t.mvx[0] = (int16_t)xFrac;
    t.mvy[0] = (int16_t)yFrac;
    CHECK(hvb_pred_batch(s.ctx, &t, 1, HVB_HOST), s);
    get(s, 0, cIdx, dst, sd, kOrigin, kOrigin, w, h);
}

template <typename Sample, int TAPS>
void predBiShim(Sample *dst, intptr_t sd, const Sample *ref0, const Sample *ref1, intptr_t sr, int w, int h, int xf0, int yf0, int xf1, int yf1,
                int bitDepth)
{
    Stage &s = stage(sizeof(Sample), bitDepth);
    constexpr int M = TAPS / 2 - 1, cIdx = TAPS == 8 ? 0 : 1;
    put(s, 1, cIdx, ref0 - M * sr - M, sr, kOrigin - M, kOrigin - M, w + TAPS - 1, h + TAPS - 1);
    put(s, 2, cIdx, ref1 - M * sr - M, sr, kOrigin - M, kOrigin - M, w + TAPS - 1, h + TAPS - 1);
    hvb_pred_task t{};
    t.dst = blk(s, 0, cIdx, kOrigin, kOrigin);
    t.ref_pic[0] = (int16_t)s.pic[1];
    t.ref_pic[1] = (int16_t)s.pic[2];
    t.x = t.y = kOrigin;
    t.w = (int16_t)w;
    t.h = (int16_t)h;
    t.mvx[0] = (int16_t)xf0;
    t.mvy[0] = (int16_t)yf0;
    t.mvx[1] = (int16_t)xf1;
    t.mvy[1] = (int16_t)yf1;
    CHECK(hvb_pred_batch(s.ctx, &t, 1, HVB_HOST), s);
    get(s, 0, cIdx, dst, sd, kOrigin, kOrigin, w, h);
}

template <typename Sample>
void subtractBiShim(Sample *dst, intptr_t sd, const Sample *pred, intptr_t sp, const Sample *src, intptr_t ss, int w, int h, int bitDepth)
{
    Stage &s = stage(sizeof(Sample), bitDepth);
    put(s, 1, 0, pred, sp, kOrigin, kOrigin, w, h);
    put(s, 2, 0, src, ss, kOrigin, kOrigin, w, h);
    hvb_subtract_bi_task t{};
    t.dst = blk(s, 0, 0, kOrigin, kOrigin);
    t.pred = blk(s, 1, 0, kOrigin, kOrigin);
    t.src = blk(s, 2, 0, kOrigin, kOrigin);
    t.w = (int16_t)w;
    t.h = (int16_t)h;
    CHECK(hvb_subtract_bi_batch(s.ctx, &t, 1, HVB_HOST), s);
    get(s, 0, 0, dst, sd, kOrigin, kOrigin, w, h);
}

// ---- intra prediction -----------------------------------------------------------------------------------
template <typename Sample, int BITDEPTH, int LOG2N, bool EDGE>
void intraShim(Sample *dst, intptr_t sd, const Sample *neighbours, int mode)
{
    Stage &s = stage(sizeof(Sample), BITDEPTH);
    constexpr int n = 1 << LOG2N;
    // neighbours points one past p(-1,-1); the array spans [-(2n+1), 2n) around it
    CHECK(hvb_pool_upload(s.ctx, neighbours - 1 - 2 * n, 4 * n + 1, 0), s);
    hvb_intra_task t{};
    t.dst = blk(s, 0, 0, kOrigin, kOrigin);
    t.nb = 2 * n;
    t.log2n = LOG2N;
    t.mode = (int8_t)(mode >= 35 ? (mode == 35 ? 1 : (mode == 36 ? 10 : 26)) : mode);
    t.edge_flag = EDGE;
    CHECK(hvb_intra_pred_batch(s.ctx, &t, 1, HVB_HOST), s);
    get(s, 0, 0, dst, sd, kOrigin, kOrigin, n, n);
}

// ---- transforms / quantisation (coefficient pool offsets: src at 0, dst at 4096) --------------------------
constexpr int kPoolDst = 4096;

template <int BITDEPTH, int LOG2N, int TRTYPE>
void transformShim(int16_t *coeffs, const int16_t *src, intptr_t stride)
{
    Stage &s = stage(BITDEPTH == 8 ? 1 : 2, BITDEPTH);
    constexpr int n = 1 << LOG2N;
    int16_t packed[32 * 32];
    for (int y = 0; y < n; ++y) std::memcpy(packed + y * n, src + y * stride, n * sizeof(int16_t));
    CHECK(hvb_coeff_upload(s.ctx, packed, n * n, 0), s);
    CHECK(hvb_coeff_upload(s.ctx, packed, n * n, kPoolDst), s); // sizes the pool
    const hvb_transform_task t{0, kPoolDst, n, LOG2N, TRTYPE, 0};
    CHECK(hvb_transform_fwd_batch(s.ctx, &t, 1, HVB_HOST), s);
    CHECK(hvb_coeff_download(s.ctx, coeffs, n * n, kPoolDst), s);
}

template <int LOG2N, int TRTYPE>
void inverseTransformShim(int16_t dst[], int16_t const coeffs[], int bitDepth)
{
    Stage &s = stage(bitDepth == 8 ? 1 : 2, bitDepth);
    constexpr int n = 1 << LOG2N;
    CHECK(hvb_coeff_upload(s.ctx, coeffs, n * n, 0), s);
    CHECK(hvb_coeff_upload(s.ctx, coeffs, n * n, kPoolDst), s);
    const hvb_transform_task t{0, kPoolDst, n, LOG2N, TRTYPE, 0};
    CHECK(hvb_transform_inv_batch(s.ctx, &t, 1, HVB_HOST), s);
    CHECK(hvb_coeff_download(s.ctx, dst, n * n, kPoolDst), s);
}

template <typename Sample, int LOG2N, int TRTYPE>
void inverseTransformAddShim(Sample *dst, intptr_t sd, Sample const *pred, intptr_t sp, int16_t const coeffs[], int bitDepth)
{
    Stage &s = stage(sizeof(Sample), bitDepth);
    constexpr int n = 1 << LOG2N;
    CHECK(hvb_coeff_upload(s.ctx, coeffs, n * n, 0), s);
    put(s, 1, 0, pred, sp, kOrigin, kOrigin, n, n);
    hvb_ita_task t{};
    t.dst = blk(s, 0, 0, kOrigin, kOrigin);
    t.pred = blk(s, 1, 0, kOrigin, kOrigin);
    t.coeffs = 0;
    t.log2n = LOG2N;
    t.trType = TRTYPE;
    CHECK(hvb_inverse_transform_add_batch(s.ctx, &t, 1, HVB_HOST), s);
    get(s, 0, 0, dst, sd, kOrigin, kOrigin, n, n);
}

int quantizeShim(int16_t *dst, const int16_t *src, int scale, int shift, int offset, int n)
{
    Stage &s = stage(1, 8);
    CHECK(hvb_coeff_upload(s.ctx, src, n, 0), s);
    CHECK(hvb_coeff_upload(s.ctx, src, n, kPoolDst), s);
    const hvb_quant_task t{0, kPoolDst, n, scale, shift, offset};
    int32_t cbf = 0;
    CHECK(hvb_quantize_batch(s.ctx, &t, 1, &cbf, HVB_HOST), s);
    CHECK(hvb_coeff_download(s.ctx, dst, n, kPoolDst), s);
    return cbf;
}

void quantizeInverseShim(int16_t *dst, const int16_t *src, int scale, int shift, int n)
{
    Stage &s = stage(1, 8);
    CHECK(hvb_coeff_upload(s.ctx, src, n, 0), s);
    CHECK(hvb_coeff_upload(s.ctx, src, n, kPoolDst), s);
    const hvb_quant_task t{0, kPoolDst, n, scale, shift, 0};
    CHECK(hvb_quantize_inverse_batch(s.ctx, &t, 1, HVB_HOST), s);
    CHECK(hvb_coeff_download(s.ctx, dst, n, kPoolDst), s);
}

} // namespace

// ==========================================================================================================
// populate entry points
// ==========================================================================================================

extern "C" {

havoc_instruction_set havoc_instruction_set_support()
{
    // every "instruction set" the caller may ask for is served by the same device code
    return (havoc_instruction_set)0x7ff;
}

void havoc_print_instruction_set_support(FILE *f, havoc_instruction_set)
{
    std::fprintf(f ? f : stdout, "havoc_b200: pixel primitives run on an NVIDIA B200 (sm_100a) through libhvb\n");
}

havoc_code havoc_new_code(havoc_instruction_set, int)
{
    havoc_code c;
    c.implementation = nullptr; // device state is per host thread (Stage), created on first use
    return c;
}

void havoc_delete_code(havoc_code) {}

int havoc_main(int, const char *[])
{
    std::printf("havoc_b200: the self-test lives in tests/ (pytest -m gpu); nothing to run here\n");
    return 0;
}

void havoc_populate_quantize_inverse(havoc_table_quantize_inverse *table, havoc_code)
{
    table->p[0] = table->p[1] = quantizeInverseShim;
}

void havoc_populate_quantize(havoc_table_quantize *table, havoc_code) { table->p = quantizeShim; }

void havoc_populate_quantize_reconstruct(havoc_table_quantize_reconstruct *table, havoc_code)
{
    for (auto &p : table->p) p = nullptr; // never called by turing/ 
}

havoc_ssd_linear *havoc_get_ssd_linear(int, havoc_code) { return ssdLinearShim; }

} // extern "C"

template <typename Sample>
void havoc_populate_sad(havoc_table_sad<Sample> *table, havoc_code)
{
#define X(w, h) table->sad##w##x##h = sadShim<Sample>;
    HAVOC_B200_PU_SIZES(X)
#undef X
    table->sadGeneric = sadShim<Sample>;
}
template void havoc_populate_sad<uint8_t>(havoc_table_sad<uint8_t> *, havoc_code);
template void havoc_populate_sad<uint16_t>(havoc_table_sad<uint16_t> *, havoc_code);

template <typename Sample>
void havoc_populate_sad_multiref(havoc_table_sad_multiref<Sample> *table, havoc_code)
{
    for (auto &row : table->lookup)
        for (auto &p : row) p = sad4Shim<Sample>;
    table->sadGeneric_4 = sad4Shim<Sample>;
}
template void havoc_populate_sad_multiref<uint8_t>(havoc_table_sad_multiref<uint8_t> *, havoc_code);
template void havoc_populate_sad_multiref<uint16_t>(havoc_table_sad_multiref<uint16_t> *, havoc_code);

template <typename Sample>
void havoc_populate_ssd(havoc_table_ssd<Sample> *table, havoc_code)
{
    for (auto &p : table->ssd) p = ssdShim<Sample>;
}
template void havoc_populate_ssd<uint8_t>(havoc_table_ssd<uint8_t> *, havoc_code);
template void havoc_populate_ssd<uint16_t>(havoc_table_ssd<uint16_t> *, havoc_code);

template <typename Sample>
void havoc_populate_hadamard_satd(havoc_table_hadamard_satd<Sample> *table, havoc_code)
{
    table->satd[0] = satdShim<Sample, 1>;
    table->satd[1] = satdShim<Sample, 2>;
    table->satd[2] = satdShim<Sample, 3>;
}
template void havoc_populate_hadamard_satd<uint8_t>(havoc_table_hadamard_satd<uint8_t> *, havoc_code);
template void havoc_populate_hadamard_satd<uint16_t>(havoc_table_hadamard_satd<uint16_t> *, havoc_code);

template <typename Sample>
void havocPopulatePredUni(HavocTablePredUni<Sample> *table, havoc_code)
{
    std::memset(table, 0, sizeof(*table));
    const int maxBitDepth = 6 + 2 * (int)sizeof(Sample);
    for (int bd = 8; bd <= maxBitDepth; ++bd)
        for (int i = 0; i < 17; ++i)
            for (int xf = 0; xf < 2; ++xf)
                for (int yf = 0; yf < 2; ++yf)
                {
                    table->p[bd - 8][0][i][xf][yf] = predUniShim<Sample, 4>;
                    if (i <= 8) table->p[bd - 8][1][i][xf][yf] = predUniShim<Sample, 8>;
                }
}
template void havocPopulatePredUni<uint8_t>(HavocTablePredUni<uint8_t> *, havoc_code);
template void havocPopulatePredUni<uint16_t>(HavocTablePredUni<uint16_t> *, havoc_code);

template <typename Sample>
void havocPopulatePredBi(HavocTablePredBi<Sample> *table, havoc_code)
{
    std::memset(table, 0, sizeof(*table));
    const int maxBitDepth = 6 + 2 * (int)sizeof(Sample);
    for (int bd = 8; bd <= maxBitDepth; ++bd)
        for (int i = 0; i < 9; ++i)
            for (int frac = 0; frac < 2; ++frac)
            {
                table->p[bd - 8][0][i][frac] = predBiShim<Sample, 4>;
                if (i <= 4) table->p[bd - 8][1][i][frac] = predBiShim<Sample, 8>;
            }
}
template void havocPopulatePredBi<uint8_t>(HavocTablePredBi<uint8_t> *, havoc_code);
template void havocPopulatePredBi<uint16_t>(HavocTablePredBi<uint16_t> *, havoc_code);

namespace havoc {

template <typename Sample>
void populateSubtractBi(TableSubtractBi<Sample> *table, havoc_code, int)
{
    table->get() = subtractBiShim<Sample>;
}
template void populateSubtractBi<uint8_t>(TableSubtractBi<uint8_t> *, havoc_code, int);
template void populateSubtractBi<uint16_t>(TableSubtractBi<uint16_t> *, havoc_code, int);

namespace intra {

template <typename Sample, int BITDEPTH, int LOG2N>
void fillIntra(Function<Sample> *(&row)[38])
{
    // 0..34 unfiltered-edge variants; 35/36/37 are DC / 10 / 26 with the luma edge filter (pred_intra.h:41-49)
    for (int m = 0; m < 35; ++m) row[m] = intraShim<Sample, BITDEPTH, LOG2N, false>;
    for (int m = 35; m < 38; ++m) row[m] = LOG2N < 5 ? intraShim<Sample, BITDEPTH, LOG2N, true> : nullptr;
}

template <typename Sample, int BITDEPTH>
void fillIntraDepth(Table<Sample> &t)
{
    // the reference indexes uint16_t tables by 10 - bitDepth and uint8_t tables by 0 (pred_intra.h:51-52)
    constexpr int bd = sizeof(Sample) == 2 ? 10 - BITDEPTH : 0;
    fillIntra<Sample, BITDEPTH, 2>(t.entries[bd][0]);
    fillIntra<Sample, BITDEPTH, 3>(t.entries[bd][1]);
    fillIntra<Sample, BITDEPTH, 4>(t.entries[bd][2]);
    fillIntra<Sample, BITDEPTH, 5>(t.entries[bd][3]);
}

template <>
void Table<uint8_t>::populate(havoc_code)
{
    std::memset(this->entries, 0, sizeof(this->entries));
    fillIntraDepth<uint8_t, 8>(*this);
}

template <>
void Table<uint16_t>::populate(havoc_code)
{
    std::memset(this->entries, 0, sizeof(this->entries));
    fillIntraDepth<uint16_t, 8>(*this);
    fillIntraDepth<uint16_t, 9>(*this);
    fillIntraDepth<uint16_t, 10>(*this);
}

} // namespace intra

void populate_inverse_transform(table_inverse_transform *table, havoc_code, int)
{
    table->sine = inverseTransformShim<2, 1>;
    table->cosine[0] = inverseTransformShim<2, 0>;
    table->cosine[1] = inverseTransformShim<3, 0>;
    table->cosine[2] = inverseTransformShim<4, 0>;
    table->cosine[3] = inverseTransformShim<5, 0>;
}

template <typename Sample>
void populate_inverse_transform_add(table_inverse_transform_add<Sample> *table, havoc_code, int)
{
    table->sine = inverseTransformAddShim<Sample, 2, 1>;
    table->cosine[0] = inverseTransformAddShim<Sample, 2, 0>;
    table->cosine[1] = inverseTransformAddShim<Sample, 3, 0>;
    table->cosine[2] = inverseTransformAddShim<Sample, 4, 0>;
    table->cosine[3] = inverseTransformAddShim<Sample, 5, 0>;
}
template void populate_inverse_transform_add<uint8_t>(table_inverse_transform_add<uint8_t> *, havoc_code, int);
template void populate_inverse_transform_add<uint16_t>(table_inverse_transform_add<uint16_t> *, havoc_code, int);

template <int bitDepth>
void populate_transform(table_transform<bitDepth> *table, havoc_code)
{
    table->dst = transformShim<bitDepth, 2, 1>;
    table->dct[0] = transformShim<bitDepth, 2, 0>;
    table->dct[1] = transformShim<bitDepth, 3, 0>;
    table->dct[2] = transformShim<bitDepth, 4, 0>;
    table->dct[3] = transformShim<bitDepth, 5, 0>;
}
template void populate_transform<8>(table_transform<8> *, havoc_code);
template void populate_transform<10>(table_transform<10> *, havoc_code);

} // namespace havoc
