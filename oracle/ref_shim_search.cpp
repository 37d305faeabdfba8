/*
 * oracle/ref_shim_search.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives the UNMODIFIED reference motion search -- fullPelMotionEstimation (turing/Search.hpp:2064-2336) and
 * subPelRefinement (:2339-2357), chained the way searchMotionUni does (:1315-1352) -- in isolation, so that
 * oracle_search.c (and through it the CUDA search) can be pinned against it.
 *
 * Both functions are templates over the encoder's handler type H.  Search.hpp is included exactly as the
 * reference's own Search.cpp includes it, and the templates are instantiated with a small stand-in handler that
 * answers the questions they ask of H (pointer conversions, h[Tag()] values) from objects of the REFERENCE's
 * own types: Picture<Sample>, Contexts, Speed, Mvp::Predictors, StateEncodePicture, the havoc function tables.
 * No search logic lives here; the stand-in only holds state.  Linked against the reference encoder's objects
 * (oracle/Makefile `searchref`) so every non-inline symbol resolves to reference code.
 */
#include "turing/Search.hpp"
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>

namespace {

struct SubstreamFields /* the two members fullPelMotionEstimation touches (turing/StateEncode.h:470-472) */
{
    MotionVector mvPreviousInteger2Nx2N[2];
    Cost costMvdZero[2][2];
};

struct DecodedPictureStandIn
{
    std::shared_ptr<StateReconstructedPictureBase> reconstructedPicture;
};

struct RefPicEntry
{
    DecodedPictureStandIn *dp;
};

struct RefPicListStandIn
{
    RefPicEntry entry;
    RefPicEntry &operator[](int) { return entry; }
};

struct CodedPuWords /* storage CodedData::PredictionUnit points into (turing/CodedData.h:41-95) */
{
    CodedData::Type words[8];
};

template <typename Sample, bool haveLzcnt>
struct Stand
{
    struct Mode { static const bool value = haveLzcnt; };

    Profiler::Timers timers;
    Mvp::Predictors predictors;
    Speed speed;
    prediction_unit pu{0, 0, 0, 0};
    coding_quadtree cqt{0, 0, 0, 0};
    StateEncodePicture picture;
    Contexts contexts;
    StateEncode *encode; /* zero-filled storage: only met / useRateControl / concurrentFrames are read */
    SubstreamFields substream;
    DecodedPictureStandIn decoded[2];
    RefPicListStandIn refPicList[2];
    StateCodedData codedData;
    CodedPuWords codedWords;
    havoc::TableSubtractBi<Sample> tableSubtractBi;
    havoc_table_sad<Sample> tableSad;
    havoc_table_sad_multiref<Sample> tableSad4;
    HavocTablePredUni<Sample> tablePred;
    havoc_table_hadamard_satd<Sample> tableSatd;
    int ctbSize, width, height, x_ctb, y_ctb, partMode, bitDepth;

    Stand()
    {
        void *raw = std::calloc(1, sizeof(StateEncode));
        encode = static_cast<StateEncode *>(raw);
        refPicList[0].entry.dp = &decoded[0];
        refPicList[1].entry.dp = &decoded[1];
        std::memset(&codedWords, 0, sizeof(codedWords));
        codedData.codedPu.p = codedWords.words;
    }
    ~Stand() { std::free(encode); }

    operator Profiler::Timers *() { return &timers; }
    operator StateCodedData *() { return &codedData; }
    operator StateEncodeSubstream<Sample> *() { return nullptr; } /* declared but unused by searchMotionBi */
    operator StatePicture *() { return &picture; }                /* dpbIndexPlus1 lookup of PuData::setRefIdx */
    operator havoc::TableSubtractBi<Sample> *() { return &tableSubtractBi; }
    operator Mvp::Predictors *() { return &predictors; }
    operator StateEncode *() { return encode; }
    operator Speed *() { return &speed; }
    operator prediction_unit *() { return &pu; }
    operator coding_quadtree *() { return &cqt; }
    operator StateEncodePicture *() { return &picture; }
    operator Contexts *() { return &contexts; }
    operator havoc_table_sad<Sample> *() { return &tableSad; }
    operator havoc_table_sad_multiref<Sample> *() { return &tableSad4; }
    operator HavocTablePredUni<Sample> *() { return &tablePred; }
    operator havoc_table_hadamard_satd<Sample> *() { return &tableSatd; }

    SubstreamFields &operator[](Concrete<StateSubstream>) { return substream; }
    int operator[](CtbSizeY) { return ctbSize; }
    int operator[](pic_width_in_luma_samples) { return width; }
    int operator[](pic_height_in_luma_samples) { return height; }
    int operator[](xCtb) { return x_ctb; }
    int operator[](yCtb) { return y_ctb; }
    int operator[](PartMode) { return partMode; }
    int operator[](BitDepthY) { return bitDepth; }
    int operator[](PicOrderCntVal) { return 0; }
    int operator[](CtbAddrInRs) { return 0; }
    RefPicListStandIn &operator[](RefPicList list) { return refPicList[list.x]; }
    /* what setPuDataMvpPredFlags (turing/Mvp.h:772-800) reads: answered from the coded-data words, like the encoder */
    int operator[](ref_idx_l0) { return 0; }
    int operator[](ref_idx_l1) { return 0; }
    int operator[](mvp_l0_flag) { return codedData.codedPu.word0().metadata[0].mvp_lX_flag; }
    int operator[](mvp_l1_flag) { return codedData.codedPu.word0().metadata[1].mvp_lX_flag; }
    MotionVector operator[](Mvd e) { return codedData.codedPu.mvd(e.refList); }
};

} // namespace

template <typename Sample, bool lz> struct SampleType<Stand<Sample, lz>> { typedef Sample Type; };

extern "C" {

struct ref_search_task
{
    int x0, y0, w, h;          /* prediction_unit */
    int cqtX0, cqtY0, log2CbSize, cqtDepth;
    int partMode;              /* 0 = PART_2Nx2N ... (turing/Global.h PartModeType) */
    int16_t mvp[2][2];
    int mvpFlagState;          /* ContextModel::state of mvp_lX_flag ctx 0 */
    double reciprocalSqrtLambda;
    int speed;                 /* Speed::Type: 0 slow, 1 medium, 2 fast */
    int met;
    int concurrentFrames;
    int xCtb, yCtb;
    int16_t prev2Nx2N[2];
    int bitDepth;
    int refList;
};

struct ref_search_result
{
    int16_t mv[2], mvd[2], mvInteger[2], prev2Nx2NAfter[2];
    int64_t cost;
    int mvpFlag;
    int reserved;
    int64_t costMvdZero[2];
    int64_t rateMvpFlag[2];
    int32_t lambda;
    int32_t reserved2;
};

struct ref_search_pictures
{
    const void *src, *ref; /* luma sample (0,0) of planes padded by `pad` on every side */
    intptr_t strideSrc, strideRef; /* in samples */
    int width, height, pad, bps;
    int isa;               /* havoc_instruction_set mask for the tables (C_REF|C_OPT = the --asm 0 path) */
    int ctbSize;
    int lzcnt;             /* which rateOf<> instantiation (Search.cpp vs SearchLzcnt.cpp) */
};

} // extern "C"

namespace {

template <typename Sample>
std::shared_ptr<PictureWrap<Sample>> makePicture(const void *origin, intptr_t stride, int width, int height, int pad)
{
    /* allocated the way the encoder allocates its pictures (turing/StatePictures.h:154-156) */
    std::shared_ptr<PictureWrap<Sample>> p(new PictureWrap<Sample>(width, height, 1, pad, pad, 32));
    auto &plane = (*p)[0];
    const Sample *s = static_cast<const Sample *>(origin);
    for (int y = -pad; y < height + pad; ++y)
        std::memcpy(&plane(-pad, y), s + y * stride - pad, sizeof(Sample) * (width + 2 * pad));
    return p;
}

template <typename Sample, bool lz>
void run(const ref_search_pictures &pics, const ref_search_task *tasks, ref_search_result *results, int count)
{
    typedef Stand<Sample, lz> H;
    std::unique_ptr<H> hp(new H);
    H &h = *hp;

    auto source = makePicture<Sample>(pics.src, pics.strideSrc, pics.width, pics.height, pics.pad);
    auto recon = std::make_shared<StateReconstructedPicture<Sample>>();
    {
        auto refPic = makePicture<Sample>(pics.ref, pics.strideRef, pics.width, pics.height, pics.pad);
        recon->picture = std::shared_ptr<Picture<Sample>>(refPic, static_cast<Picture<Sample> *>(refPic.get()));
    }
    h.decoded[0].reconstructedPicture = recon;
    h.decoded[1].reconstructedPicture = recon;
    auto docket = std::make_shared<InputQueue::Docket>();
    docket->picture = source;
    docket->segmentPoc = 0;
    h.picture.docket = docket;

    havoc_code code = havoc_new_code(havoc_instruction_set(pics.isa), 12000000) /* the encoder's size, StateFunctionTables.h:67 */;
    havoc_populate_sad(&h.tableSad, code);
    havoc_populate_sad_multiref(&h.tableSad4, code);
    havocPopulatePredUni(&h.tablePred, code);
    havoc_populate_hadamard_satd(&h.tableSatd, code);

    h.ctbSize = pics.ctbSize;
    h.width = pics.width;
    h.height = pics.height;

    for (int i = 0; i < count; ++i)
    {
        const ref_search_task &t = tasks[i];
        ref_search_result &r = results[i];
        std::memset(&r, 0, sizeof(r));

        h.pu = prediction_unit{t.x0, t.y0, t.w, t.h};
        h.cqt = coding_quadtree{t.cqtX0, t.cqtY0, t.log2CbSize, t.cqtDepth};
        h.partMode = t.partMode;
        h.bitDepth = t.bitDepth;
        h.x_ctb = t.xCtb;
        h.y_ctb = t.yCtb;
        h.speed = Speed(static_cast<Speed::Type>(t.speed));
        h.encode->met = t.met != 0;
        h.encode->useRateControl = false;
        h.encode->concurrentFrames = t.concurrentFrames;
        h.picture.reciprocalSqrtLambda = t.reciprocalSqrtLambda;
        h.contexts.template get<mvp_lX_flag>(0).state = static_cast<uint8_t>(t.mvpFlagState);
        for (int k = 0; k < 2; ++k)
        {
            h.predictors.mvp[0][t.refList][k][0] = t.mvp[k][0];
            h.predictors.mvp[0][t.refList][k][1] = t.mvp[k][1];
        }
        h.substream.mvPreviousInteger2Nx2N[t.refList][0] = t.prev2Nx2N[0];
        h.substream.mvPreviousInteger2Nx2N[t.refList][1] = t.prev2Nx2N[1];
        h.substream.costMvdZero[t.refList][0] = Cost();
        h.substream.costMvdZero[t.refList][1] = Cost();

        /* --- the body of searchMotionUni (Search.hpp:1331-1348) --- */
        mvd_coding const mvdc{t.x0, t.y0, t.refList};
        MvCandidate best;
        fullPelMotionEstimation(mvdc, h, best);
        MotionVector mvd = best.mvd;
        MotionVector mv = best.mv;
        r.mvInteger[0] = best.mv[0];
        r.mvInteger[1] = best.mv[1];
        if (h.speed.doHalfPelRefinement())
        {
            auto const input = (*source)(t.x0, t.y0, 0);
            auto const reference = (*recon->picture)(t.x0, t.y0, 0);
            subPelRefinement(mv, mvd, mvdc, h, input, reference);
        }

        r.mv[0] = mv[0];
        r.mv[1] = mv[1];
        r.mvd[0] = mvd[0];
        r.mvd[1] = mvd[1];
        r.prev2Nx2NAfter[0] = h.substream.mvPreviousInteger2Nx2N[t.refList][0];
        r.prev2Nx2NAfter[1] = h.substream.mvPreviousInteger2Nx2N[t.refList][1];
        r.cost = best.cost.value;
        r.mvpFlag = best.mvpFlag;
        r.costMvdZero[0] = h.substream.costMvdZero[t.refList][0].value;
        r.costMvdZero[1] = h.substream.costMvdZero[t.refList][1].value;
        EstimateRateBin<mvp_lX_flag> bin(h, 0);
        r.rateMvpFlag[0] = bin.rate(0).value;
        r.rateMvpFlag[1] = bin.rate(1).value;
        Lambda lambda;
        lambda.set(t.reciprocalSqrtLambda);
        r.lambda = lambda.value;
    }

    havoc_delete_code(code);
}

} // namespace

extern "C" int ref_search_batch(const ref_search_pictures *pics, const ref_search_task *tasks,
                                ref_search_result *results, int count)
{
    if (pics->bps == 1)
    {
        if (pics->lzcnt) run<uint8_t, true>(*pics, tasks, results, count);
        else run<uint8_t, false>(*pics, tasks, results, count);
    }
    else
    {
        if (pics->lzcnt) run<uint16_t, true>(*pics, tasks, results, count);
        else run<uint16_t, false>(*pics, tasks, results, count);
    }
    return 0;
}


/* ---- searchMotionBi (turing/Search.hpp:1498-1653) ------------------------------------------------------- */

extern "C" {

struct ref_bi_task
{
    int x0, y0, w, h;
    int16_t mvp[2][2][2];      /* [refList][mvp flag][x,y] */
    int16_t mvd[2][2];         /* coded mvd of each list before the call (the uni-directional results) */
    int mvpFlag[2];
    int mvpFlagState;
    double reciprocalSqrtLambda;
    int speed, concurrentFrames, xCtb, yCtb, bitDepth;
    int chain;                 /* 0: searchMotionBi(L0) only (the mvd_l1_zero_flag case); 1: L0 then L1 (Search.hpp:1817-1822) */
};

struct ref_bi_result
{
    int16_t mvd[2][2];         /* coded mvd after the call(s) */
    int mvpFlag[2];
    int64_t rateMvpFlag[2];
    int32_t lambdaHalf;        /* Lambda::set(reciprocalSqrtLambda * 0.5).value */
    int32_t reserved;
};

} // extern "C"

namespace {

template <typename Sample, bool lz>
void runBi(const ref_search_pictures &pics, const void *ref1, intptr_t strideRef1, const ref_bi_task *tasks,
           ref_bi_result *results, int count)
{
    typedef Stand<Sample, lz> H;
    std::unique_ptr<H> hp(new H);
    H &h = *hp;

    auto source = makePicture<Sample>(pics.src, pics.strideSrc, pics.width, pics.height, pics.pad);
    const void *planes[2] = {pics.ref, ref1};
    const intptr_t strides[2] = {pics.strideRef, strideRef1};
    for (int list = 0; list < 2; ++list)
    {
        auto recon = std::make_shared<StateReconstructedPicture<Sample>>();
        auto refPic = makePicture<Sample>(planes[list], strides[list], pics.width, pics.height, pics.pad);
        recon->picture = std::shared_ptr<Picture<Sample>>(refPic, static_cast<Picture<Sample> *>(refPic.get()));
        h.decoded[list].reconstructedPicture = recon;
    }
    auto docket = std::make_shared<InputQueue::Docket>();
    docket->picture = source;
    docket->segmentPoc = 0;
    h.picture.docket = docket;

    havoc_code code = havoc_new_code(havoc_instruction_set(pics.isa), 12000000);
    havoc_populate_sad(&h.tableSad, code);
    havoc_populate_sad_multiref(&h.tableSad4, code);
    havocPopulatePredUni(&h.tablePred, code);
    havoc_populate_hadamard_satd(&h.tableSatd, code);
    havoc::populateSubtractBi(&h.tableSubtractBi, code);

    h.ctbSize = pics.ctbSize;
    h.width = pics.width;
    h.height = pics.height;

    for (int i = 0; i < count; ++i)
    {
        const ref_bi_task &t = tasks[i];
        ref_bi_result &r = results[i];
        std::memset(&r, 0, sizeof(r));

        h.pu = prediction_unit{t.x0, t.y0, t.w, t.h};
        h.cqt = coding_quadtree{t.x0, t.y0, 6, 0};
        h.partMode = 0;
        h.bitDepth = t.bitDepth;
        h.x_ctb = t.xCtb;
        h.y_ctb = t.yCtb;
        h.speed = Speed(static_cast<Speed::Type>(t.speed));
        h.encode->useRateControl = false;
        h.encode->concurrentFrames = t.concurrentFrames;
        h.picture.reciprocalSqrtLambda = t.reciprocalSqrtLambda;
        h.contexts.template get<mvp_lX_flag>(0).state = static_cast<uint8_t>(t.mvpFlagState);

        /* coded data of a bi-predicted PU as Search<prediction_unit>::searchBi leaves it (Search.hpp:1795-1803) */
        auto &codedPu = h.codedData.codedPu;
        codedPu.init();
        for (int list = 0; list < 2; ++list)
        {
            codedPu.word0().metadata[list].predFlag = 1;
            codedPu.word0().metadata[list].ref_idx_lX = 0;
            codedPu.word0().metadata[list].mvp_lX_flag = t.mvpFlag[list];
        }
        for (int list = 0; list < 2; ++list)
        {
            codedPu.mvd(list) = MotionVector{t.mvd[list][0], t.mvd[list][1]};
            for (int k = 0; k < 2; ++k)
                h.predictors.mvp[0][list][k] = MotionVector{t.mvp[list][k][0], t.mvp[list][k][1]};
        }

        searchMotionBi(h, 0);
        if (t.chain) searchMotionBi(h, 1);

        for (int list = 0; list < 2; ++list)
        {
            r.mvd[list][0] = codedPu.mvd(list)[0];
            r.mvd[list][1] = codedPu.mvd(list)[1];
            r.mvpFlag[list] = codedPu.word0().metadata[list].mvp_lX_flag;
        }
        EstimateRateBin<mvp_lX_flag> bin(h, 0);
        r.rateMvpFlag[0] = bin.rate(0).value;
        r.rateMvpFlag[1] = bin.rate(1).value;
        Lambda lambda;
        lambda.set(t.reciprocalSqrtLambda * 0.5);
        r.lambdaHalf = lambda.value;
    }
    havoc_delete_code(code);
}

} // namespace

extern "C" int ref_search_bi_batch(const ref_search_pictures *pics, const void *ref1, intptr_t strideRef1,
                                   const ref_bi_task *tasks, ref_bi_result *results, int count)
{
    if (pics->bps == 1)
    {
        if (pics->lzcnt) runBi<uint8_t, true>(*pics, ref1, strideRef1, tasks, results, count);
        else runBi<uint8_t, false>(*pics, ref1, strideRef1, tasks, results, count);
    }
    else
    {
        if (pics->lzcnt) runBi<uint16_t, true>(*pics, ref1, strideRef1, tasks, results, count);
        else runBi<uint16_t, false>(*pics, ref1, strideRef1, tasks, results, count);
    }
    return 0;
}


/* ---- predictInter's workers + measureSatd (turing/Dsp.h:769-864, turing/Measure.h:96-135) ------------------
 * predictUni / predictBi are free function templates that take the havoc table, the pictures and the vectors, so
 * they are called directly; the chroma rule of Compute<Satd, Rectangle> (Measure.h:156-160: return 0 when the chroma
 * block is not a multiple of four) is the one line restated here. */

extern "C" {

struct ref_pu_planes
{
    const void *p[3];      /* sample (0,0) of Y, Cb, Cr; planes padded by pad (pad/2 for chroma) */
    intptr_t stride[3];
};

struct ref_pu_task
{
    int x0, y0, w, h;
    int predFlag[2];
    int16_t mv[2][2];
};

} // extern "C"

namespace {

template <typename Sample>
std::shared_ptr<Picture<Sample>> makePicture3(const ref_pu_planes &pl, int width, int height, int pad)
{
    std::shared_ptr<Picture<Sample>> p(new Picture<Sample>(width, height, 1, pad, pad, 32));
    for (int c = 0; c < 3; ++c)
    {
        const int sh = c ? 1 : 0, w = width >> sh, h = height >> sh, pd = pad >> sh;
        auto &plane = (*p)[c];
        const Sample *s = static_cast<const Sample *>(pl.p[c]);
        for (int y = -pd; y < h + pd; ++y) std::memcpy(&plane(-pd, y), s + y * pl.stride[c] - pd, sizeof(Sample) * (w + 2 * pd));
    }
    return p;
}

template <typename Sample>
void runPu(const ref_pu_planes *planes /* src, ref L0, ref L1 */, int width, int height, int pad, int isa, int bitDepthY,
           int bitDepthC, const ref_pu_task *tasks, int32_t *satd, Sample *predOut, int count)
{
    auto source = makePicture3<Sample>(planes[0], width, height, pad);
    auto ref0 = makePicture3<Sample>(planes[1], width, height, pad);
    auto ref1 = makePicture3<Sample>(planes[2], width, height, pad);
    Picture<Sample> *refs[2] = {ref0.get(), ref1.get()};

    havoc_code code = havoc_new_code(havoc_instruction_set(isa), 12000000);
    HavocTablePredUni<Sample> tableUni;
    HavocTablePredBi<Sample> tableBi;
    havoc_table_hadamard_satd<Sample> tableSatd;
    havocPopulatePredUni(&tableUni, code);
    havocPopulatePredBi(&tableBi, code);
    havoc_populate_hadamard_satd(&tableSatd, code);

    HAVOC_ALIGN(32, Sample, buffer[3][64 * 64 + 64]);
    for (int i = 0; i < count; ++i)
    {
        const ref_pu_task &t = tasks[i];
        Raster<Sample> pred0{buffer[0], 64}, pred1{buffer[1], 64}, pred2{buffer[2], 64};
        const MotionVector mv0{t.mv[0][0], t.mv[0][1]}, mv1{t.mv[1][0], t.mv[1][1]};
        if (t.predFlag[0] && t.predFlag[1])
            predictBi<Sample>(tableBi, pred0, pred1, pred2, *refs[0], mv0, *refs[1], mv1, t.x0, t.y0, t.w, t.h, bitDepthY, bitDepthC);
        else if (t.predFlag[0])
            predictUni<Sample>(tableUni, pred0, pred1, pred2, *refs[0], mv0, t.x0, t.y0, t.w, t.h, bitDepthY, bitDepthC);
        else
            predictUni<Sample>(tableUni, pred0, pred1, pred2, *refs[1], mv1, t.x0, t.y0, t.w, t.h, bitDepthY, bitDepthC);
        for (int c = 0; c < 3; ++c)
        {
            int w = t.w, h = t.h;
            if (c)
            {
                w >>= 1;
                h >>= 1;
            }
            Raster<Sample> pred{buffer[c], 64};
            satd[3 * i + c] = (c && ((w | h) & 0x3)) ? 0 : measureSatd(&tableSatd, (*source)(t.x0, t.y0, c), pred, w, h);
            if (predOut)
            {
                Sample *o = predOut + (size_t)(3 * i + c) * 64 * 64;
                for (int y = 0; y < h; ++y) std::memcpy(o + y * w, buffer[c] + y * 64, sizeof(Sample) * w);
            }
        }
    }
    havoc_delete_code(code);
}

} // namespace

extern "C" int ref_pu_cost_batch(const ref_pu_planes *planes, int width, int height, int pad, int bps, int isa, int bitDepthY,
                                 int bitDepthC, const ref_pu_task *tasks, int32_t *satd, void *predOut, int count)
{
    if (bps == 1) runPu<uint8_t>(planes, width, height, pad, isa, bitDepthY, bitDepthC, tasks, satd, static_cast<uint8_t *>(predOut), count);
    else runPu<uint16_t>(planes, width, height, pad, isa, bitDepthY, bitDepthC, tasks, satd, static_cast<uint16_t *>(predOut), count);
    return 0;
}
