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
    DecodedPictureStandIn decoded;
    RefPicListStandIn refPicList;
    havoc_table_sad<Sample> tableSad;
    havoc_table_sad_multiref<Sample> tableSad4;
    HavocTablePredUni<Sample> tablePred;
    havoc_table_hadamard_satd<Sample> tableSatd;
    int ctbSize, width, height, x_ctb, y_ctb, partMode, bitDepth;

    Stand()
    {
        void *raw = std::calloc(1, sizeof(StateEncode));
        encode = static_cast<StateEncode *>(raw);
        refPicList.entry.dp = &decoded;
    }
    ~Stand() { std::free(encode); }

    operator Profiler::Timers *() { return &timers; }
    operator StateCodedData *() { return nullptr; } /* declared but unused on this path */
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
    RefPicListStandIn &operator[](RefPicList) { return refPicList; }
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
    h.decoded.reconstructedPicture = recon;
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
