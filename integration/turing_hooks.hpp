// integration/turing_hooks.hpp -- the reference encoder's hot loops on the batched ABI.
//
// This header is compiled INTO the reference encoder (turing/Search.hpp, turing/TaskSao.cpp, turing/TaskEncodeInput.cpp;
// integration/patch_turing.py adds one include and one forwarding line per call site to a build-time copy -- the files
// under /root/reference stay as they are).  Each hook takes the place of a body that calls havoc primitives in a loop:
//
//   searchMotionUni   turing/Search.hpp:1315-1354  -> one hvb_me_task      (integer pattern search + 1/2, 1/4 pel refinement)
//   searchMotionBi    turing/Search.hpp:1498-1653  -> one hvb_me_bi_task
//   measurePuCost     turing/Search.hpp:1669-1683  -> one hvb_pu_cost_task (the prediction + SATD half; the rate half, the
//                                                     merge / uni / bi decision and every side effect stay the reference's)
//   searchIntraPartition's 35-mode loop  :113-142  -> one hvb_intra_sweep_task
//   reconstructInter's transform blocks  turing/Reconstruct.cpp:1237-..., :733-857 -> hvbenc_tu_chain
//   TaskSao::run      turing/TaskSao.cpp:96-160    -> the finished CTU (and the padding it produced) goes up to the device
//   startPictureEncode turing/TaskEncodeInput.cpp:134-250 -> the source picture goes up once
//
// Everything the reference reads from encoder state is gathered here into the task; everything it writes back (coded
// mvd, mvp flag, costMvdZero, mvPreviousInteger2Nx2N) is written here from the result, so the decision code that follows
// cannot tell the difference.  A hook returns false when it does not apply (rate control's per-CTU lambda, weighted
// prediction) and the reference body runs instead.
//
// Speculation: a device result is a pure function of its task, so a worker may issue tasks EARLY, together, and keep the
// results in a per-thread memo keyed by the task's bytes; when the reference's control flow reaches the call, the hook
// finds the answer without a round trip.  Wrong guesses cost device time only, never correctness.
#ifndef INCLUDED_turing_hooks_hpp
#define INCLUDED_turing_hooks_hpp

#include "turing_hooks_state.h"

namespace hvbhooks {

template <class H> struct SampleOf { typedef typename SampleType<H>::Type Type; };

template <class H> hvbenc *sessionOf(H &h)
{
    typedef typename SampleOf<H>::Type Sample;
    return session((int)sizeof(Sample), h[BitDepthY()], h[pic_width_in_luma_samples()], h[pic_height_in_luma_samples()]);
}

template <class H> int inputPicture(H &h)
{
    return pictureId(static_cast<StateEncodePicture *>(h)->docket->picture.get(), false);
}

template <class H> int referencePicture(H &h, int refList, int refIdx)
{
    typedef typename SampleOf<H>::Type Sample;
    auto &rec = static_cast<StateReconstructedPicture<Sample> &>(*h[RefPicList(refList)][refIdx].dp->reconstructedPicture);
    return pictureId(rec.picture.get(), false);
}

template <class H> bool usable(H &h)
{
    StateEncode *stateEncode = h;
    return on() && !stateEncode->useRateControl;
}

#ifdef HVBHOOKS_SEARCH

// what LimitFullPelMv (turing/Search.hpp:1366-1407) clamps to: its members are private, its effect on the extreme
// vectors is not
template <class H> void fullPelLimits(H &h, hvb_mv &lo, hvb_mv &hi)
{
    prediction_unit const *pu = h;
    LimitFullPelMv limit(*pu, h);
    MotionVector a{-32768, -32768}, b{32767, 32767};
    limit(a);
    limit(b);
    lo.x = a[0], lo.y = a[1], hi.x = b[0], hi.y = b[1];
}

template <class H> void fillMeTask(H &h, int refList, hvb_me_task &t)
{
    Speed *speed = h;
    StateEncode *stateEncode = h;
    coding_quadtree const *cqt = h;
    prediction_unit const *pu = h;
    Mvp::Predictors *predictors = h;
    auto *substream = &h[Concrete<StateSubstream>()];
    memset(&t, 0, sizeof(t));
    t.src_pic = (int16_t)inputPicture(h);
    t.ref_pic = (int16_t)referencePicture(h, refList, 0);
    t.x0 = pu->x0, t.y0 = pu->y0, t.w = pu->nPbW, t.h = pu->nPbH;
    EstimateRateBin<mvp_lX_flag> bin(h, 0);
    for (int k = 0; k < 2; ++k)
    {
        t.mvp[k].x = predictors->mvp[0][refList][k][0];
        t.mvp[k].y = predictors->mvp[0][refList][k][1];
        t.rateMvpFlag[k] = bin.rate(k).value;
    }
    Lambda lambda;
    lambda.set(getReciprocalSqrtLambda(h));
    t.lambda = lambda.value;
    fullPelLimits(h, t.limitMin, t.limitMax);
    t.prev2Nx2N.x = substream->mvPreviousInteger2Nx2N[refList][0];
    t.prev2Nx2N.y = substream->mvPreviousInteger2Nx2N[refList][1];
    t.smallSearchWindow = speed->useSmallSearchWindow();
    t.met = stateEncode->met;
    t.log2CbSize = (uint8_t)cqt->log2CbSize;
    t.usePrev2Nx2N = h[PartMode()] != PART_2Nx2N || cqt->cqtDepth != 0;
    t.halfPel = speed->doHalfPelRefinement();
    t.quarterPel = speed->doQuarterPelRefinement();
}

inline bool meLookup(const hvb_me_task &t, hvb_me_result &r)
{
    Memo &m = memo();
    for (int i = 0; i < m.nMe; ++i)
        if (!memcmp(&m.meTask[i], &t, sizeof(t)))
        {
            r = m.meResult[i];
            return true;
        }
    return false;
}

// searchMotionUni (turing/Search.hpp:1315-1354)
template <class H> bool searchMotionUni(H &h, int refList)
{
    if (!usable(h) || !(enabledMask() & 1)) return false;
    {
        prediction_unit const *pu = h;
        if (pu->nPbW * pu->nPbH < meMinArea()) return false;
    }
    StateCodedData *stateCodedData = h;
    auto *substream = &h[Concrete<StateSubstream>()];
    hvb_me_task t;
    fillMeTask(h, refList, t);
    hvb_me_result r;
    if (!meLookup(t, r))
    {
        const int rc = hvbenc_me(sessionOf(h), &t, &r);
        if (rc) fatal("hvbenc_me", rc);
    }
    // fullPelMotionEstimation's side effects (:2151, :2332-2335)
    for (int k = 0; k < 2; ++k)
        if (r.costMvdZero[k]) substream->costMvdZero[refList][k].value = r.costMvdZero[k];
    if (!(r.flags & 1) && h[PartMode()] == PART_2Nx2N)
        substream->mvPreviousInteger2Nx2N[refList] = MotionVector{r.mvInteger.x, r.mvInteger.y};
    stateCodedData->codedPu.mvd(refList) = MotionVector{r.mvd.x, r.mvd.y};
    stateCodedData->codedPu.word0().metadata[refList].mvp_lX_flag = r.mvpFlag;
    return true;
}

// searchMotionBi (turing/Search.hpp:1498-1653)
template <class H> bool searchMotionBi(H &h, int refList)
{
    if (!usable(h) || !(enabledMask() & 2)) return false;
    Speed *speed = h;
    StateCodedData *stateCodedData = h;
    prediction_unit const *pu = h;
    if (pu->nPbW * pu->nPbH < meMinArea()) return false;
    Mvp::Predictors *predictors = h;
    PuData puData;
    setPuDataMvpPredFlags(puData, h, true, true);
    hvb_me_bi_task t;
    memset(&t, 0, sizeof(t));
    t.src_pic = (int16_t)inputPicture(h);
    t.ref_pic = (int16_t)referencePicture(h, refList, puData.refIdx(refList));
    t.other_pic = (int16_t)referencePicture(h, 1 - refList, puData.refIdx(1 - refList));
    t.x0 = pu->x0, t.y0 = pu->y0, t.w = pu->nPbW, t.h = pu->nPbH;
    EstimateRateBin<mvp_lX_flag> bin(h, 0);
    for (int k = 0; k < 2; ++k)
    {
        t.mvp[k].x = predictors->mvp[0][refList][k][0];
        t.mvp[k].y = predictors->mvp[0][refList][k][1];
        t.rateMvpFlag[k] = bin.rate(k).value;
    }
    Lambda lambda;
    lambda.set(getReciprocalSqrtLambda(h) * 0.5);
    t.lambda = lambda.value;
    fullPelLimits(h, t.limitMin, t.limitMax);
    t.mvStart.x = puData.mv(refList)[0], t.mvStart.y = puData.mv(refList)[1];
    t.mvOther.x = puData.mv(1 - refList)[0], t.mvOther.y = puData.mv(1 - refList)[1];
    t.smallWindow = speed->useBiSmallSearchWindow();
    t.halfPel = speed->doHalfPelRefinement();
    t.quarterPel = speed->doQuarterPelRefinement();
    hvb_me_bi_result r;
    const int rc = hvbenc_me_bi(sessionOf(h), &t, &r);
    if (rc) fatal("hvbenc_me_bi", rc);
    stateCodedData->codedPu.mvd(refList) = MotionVector{r.mvd.x, r.mvd.y};
    stateCodedData->codedPu.word0().metadata[refList].mvp_lX_flag = r.mvpFlag;
    return true;
}

// the task of predictInter + SATD for the PU as the cursor's BlockData describes it (turing/Dsp.h:866-915)
template <class H> bool fillPuCostTask(H &h, const prediction_unit &pu, const PuData &puData, hvb_pu_cost_task &t)
{
    memset(&t, 0, sizeof(t));
    t.src_pic = (int16_t)inputPicture(h);
    t.dst_pic = -1;
    for (int list = 0; list < 2; ++list)
    {
        t.ref_pic[list] = -1;
        if (puData.getDpbIndex(list) >= 0)
        {
            t.ref_pic[list] = (int16_t)referencePicture(h, list, puData.refIdx(list));
            t.mvx[list] = puData.mv(list)[0];
            t.mvy[list] = puData.mv(list)[1];
        }
    }
    t.x0 = pu.x0, t.y0 = pu.y0, t.w = pu.nPbW, t.h = pu.nPbH;
    return t.ref_pic[0] >= 0 || t.ref_pic[1] >= 0;
}

inline bool puLookup(const hvb_pu_cost_task &t, int32_t satd[3])
{
    Memo &m = memo();
    for (int i = 0; i < m.nPu; ++i)
        if (!memcmp(&m.puTask[i], &t, sizeof(t)))
        {
            memcpy(satd, m.puResult[i], sizeof(int32_t) * 3);
            return true;
        }
    return false;
}

// the distortion half of measurePuCost (turing/Search.hpp:1669-1683); puData has been set by the caller
template <class H> bool puCost(H &h, const prediction_unit &pu, const PuData &puData, int32_t satd[3])
{
    if (!usable(h) || !(enabledMask() & 4) || h[weightedPredFlag()]) return false;
    if (pu.nPbW * pu.nPbH < puMinArea()) return false;
    hvb_pu_cost_task t;
    if (!fillPuCostTask(h, pu, puData, t)) return false;
    if (puLookup(t, satd)) return true;
    const int rc = hvbenc_pu_cost(sessionOf(h), &t, 1, satd);
    if (rc) fatal("hvbenc_pu_cost", rc);
    return true;
}


// searchMergeModes (turing/Search.hpp:1761-1768) measures the candidates one after the other; their predictions and SATDs
// do not depend on each other, so all of them go to the device in one hand-over, ahead of the loop; the loop's
// measurePuCost calls find their distortion in the memo.
template <class H> void prefetchMergeCosts(H &h, const prediction_unit &pu)
{
    if (!usable(h) || !(enabledMask() & 4) || h[weightedPredFlag()]) return;
    if (pu.nPbW * pu.nPbH < puMinArea()) return;
    Mvp::Predictors *predictors = h;
    Memo &m = memo();
    m.nPu = 0;
    hvb_pu_cost_task tasks[Memo::kPu];
    int n = 0;
    for (int i = 0; i < h[MaxNumMergeCand()] && n < Memo::kPu; ++i)
    {
        hvb_pu_cost_task t;
        if (!fillPuCostTask(h, pu, predictors->merge[i], t)) continue;
        bool seen = false; // candidates may coincide
        for (int k = 0; k < n; ++k) seen |= !memcmp(&tasks[k], &t, sizeof(t));
        if (!seen) tasks[n++] = t;
    }
    if (n < 2) return; // a single candidate gains nothing from going early
    const int rc = hvbenc_pu_cost(sessionOf(h), tasks, n, &m.puResult[0][0]);
    if (rc) fatal("hvbenc_pu_cost(merge candidates)", rc);
    memcpy(m.puTask, tasks, sizeof(hvb_pu_cost_task) * n);
    m.nPu = n;
}

// the 35-mode SATD sweep of searchIntraPartition (turing/Search.hpp:113-142 -> predictIntraLuma, turing/Reconstruct.cpp:630-712)
// for a partition of 4x4 .. 32x32: the reference samples have been substituted by the caller (:54-61); the filtered
// array is derived on the device (turing/IntraReferenceSamples.h:373-419).  Leaves what the loop's last iteration leaves
// in the substream state (predictIntraLuma resets ssd and sets satd, Reconstruct.cpp:1140-1158).
template <class H> bool intraSweep(H &h, IntraPartition const &intraPartition, int32_t distortion[35])
{
    if (!usable(h) || !(enabledMask() & 8)) return false;
    typedef typename SampleOf<H>::Type Sample;
    StateEncodeSubstream<Sample> *stateEncodeSubstream = h;
    const int log2n = intraPartition.log2CbSize - intraPartition.split;
    if (log2n < intraMinLog2() || log2n > 5) return false;
    hvb_intra_sweep_task t;
    memset(&t, 0, sizeof(t));
    t.src.pic = (int16_t)inputPicture(h);
    t.src.cIdx = 0;
    t.src.x = (int16_t)xPositionOf(intraPartition);
    t.src.y = (int16_t)yPositionOf(intraPartition);
    t.log2n = (int8_t)log2n;
    t.cIdx = 0;
    t.strong_intra_smoothing = (int8_t)h[strong_intra_smoothing_enabled_flag()];
    auto &unfiltered = stateEncodeSubstream->unfiltered[0];
    const int rc = hvbenc_intra_sweep(sessionOf(h), &t, &unfiltered(-1, (2 << log2n) - 1), distortion);
    if (rc) fatal("hvbenc_intra_sweep", rc);
    StateEncodeSubstreamBase *base = h;
    base->satd = distortion[34];
    base->ssd[0] = base->ssd[1] = base->ssd[2] = 0;
    return true;
}

#endif // HVBHOOKS_SEARCH

#ifdef HVBHOOKS_RECONSTRUCT

// ---- transform blocks of an inter CU (turing/Reconstruct.cpp:733-857) ------------------------------------------------

inline void snapshotContexts(Contexts &contexts, double lambda, hvb_rdoq_ctx &s)
{
    memset(&s, 0, sizeof(s));
    for (int i = 0; i < 44; ++i) s.sig_coeff_flag[i] = contexts.get<sig_coeff_flag>(i).state;
    for (int i = 0; i < 24; ++i) s.greater1_flag[i] = contexts.get<coeff_abs_level_greater1_flag>(i).state;
    for (int i = 0; i < 6; ++i) s.greater2_flag[i] = contexts.get<coeff_abs_level_greater2_flag>(i).state;
    for (int i = 0; i < 4; ++i) s.coded_sub_block_flag[i] = contexts.get<coded_sub_block_flag>(i).state;
    for (int i = 0; i < 18; ++i) s.last_x_prefix[i] = contexts.get<last_sig_coeff_x_prefix>(i).state;
    for (int i = 0; i < 18; ++i) s.last_y_prefix[i] = contexts.get<last_sig_coeff_y_prefix>(i).state;
    for (int i = 0; i < 2; ++i) s.cbf_luma[i] = contexts.get<cbf_luma>(i).state;
    for (int i = 0; i < 4; ++i) s.cbf_cbcr[i] = contexts.get<cbf_cX>(i).state;
    s.rqt_root_cbf[0] = contexts.get<rqt_root_cbf>(0).state;
    s.lambda = lambda;
}

// the task of one inter transform block, parameters as ReconstructInterBlock::go derives them (:778-790)
template <class H> void fillInterTuTask(H &h, int x0, int y0, int cIdx, int log2n, hvb_tu_task &t)
{
    StateEncode *stateEncode = h;
    memset(&t, 0, sizeof(t));
    t.src.pic = (int16_t)inputPicture(h);
    t.src.cIdx = (int16_t)cIdx;
    t.src.x = (int16_t)(x0 >> (cIdx ? 1 : 0));
    t.src.y = (int16_t)(y0 >> (cIdx ? 1 : 0));
    t.log2n = (int8_t)log2n;
    t.trType = 0;
    t.cIdx = (int8_t)cIdx;
    const int bitDepth = cIdx ? h[BitDepthC()] : h[BitDepthY()];
    int const qpScaled = static_cast<QpState *>(h)->getQp(cIdx);
    int const shiftQuantise = 29 - bitDepth + qpScaled / 6 - log2n;
    int const offsetQuantise = 85 << (shiftQuantise - 9);
    t.qscale = static_cast<QpState *>(h)->getQuantiseScale(cIdx);
    t.qshift = shiftQuantise;
    t.qoffset = offsetQuantise >> (shiftQuantise - 16);
    t.iqscale = static_cast<QpState *>(h)->getScale(cIdx);
    t.iqshift = log2n - 1 + bitDepth - 8;
    t.flags = (int8_t)((stateEncode->rdoq ? 1 : 0) | (h[sign_data_hiding_enabled_flag()] ? 4 : 0));
}

// Issue the given blocks of the current CU in one submission; results into the memo.
template <class H> void issueInterBlocks(H &h, const int (*blocks)[4] /* x0, y0, cIdx, log2n */, int count)
{
    typedef typename SampleOf<H>::Type Sample;
    StateEncodeSubstream<Sample> *stateEncodeSubstream = h;
    Candidate<Sample> *candidate = h;
    coding_quadtree const *cqt = h;
    StateEncode *stateEncode = h;
    TuMemo &m = tuMemo();
    hvb_tu_task tasks[TuMemo::kBlocks];
    hvb_tu_result results[TuMemo::kBlocks];
    const void *pred[TuMemo::kBlocks];
    void *rec[TuMemo::kBlocks];
    intptr_t predStride[TuMemo::kBlocks], recStride[TuMemo::kBlocks];
    int16_t *levels[TuMemo::kBlocks];
    int n = 0;
    for (int i = 0; i < count && m.n + n < TuMemo::kBlocks; ++i)
    {
        const int x0 = blocks[i][0], y0 = blocks[i][1], cIdx = blocks[i][2], log2n = blocks[i][3];
        fillInterTuTask(h, x0, y0, cIdx, log2n, tasks[n]);
        // scanIdx of an inter block is 0 (H.265 7.4.9.11: mode dependent scans are intra only)
        tasks[n].scanIdx = 0;
        auto predPiece = candidate->stateReconstructionCache->components[cIdx].get(stateEncodeSubstream->interPieces[cIdx][0]);
        auto predSamples = predPiece.offset((x0 - cqt->x0) >> (cIdx ? 1 : 0), (y0 - cqt->y0) >> (cIdx ? 1 : 0));
        auto recPiece = candidate->stateReconstructionCache->components[cIdx].get(stateEncodeSubstream->interPieces[cIdx][1 + candidate->rqtdepth]);
        auto recSamples = recPiece.offset((x0 - cqt->x0) >> (cIdx ? 1 : 0), (y0 - cqt->y0) >> (cIdx ? 1 : 0));
        pred[n] = predSamples.p, predStride[n] = predSamples.stride;
        rec[n] = recSamples.p, recStride[n] = recSamples.stride;
        TuBlock &b = m.block[m.n + n];
        b.x0 = x0, b.y0 = y0, b.cIdx = cIdx, b.log2n = log2n;
        levels[n] = b.levels;
        ++n;
    }
    if (!n) return;
    hvb_rdoq_ctx snapshot;
    if (stateEncode->rdoq) snapshotContexts(*static_cast<Contexts *>(h), static_cast<StateEncodePicture *>(h)->lambda, snapshot);
    const int rc = hvbenc_tu_chain(sessionOf(h), tasks, n, stateEncode->rdoq ? &snapshot : nullptr, pred, predStride, rec, recStride, levels, results);
    if (rc) fatal("hvbenc_tu_chain", rc);
    for (int i = 0; i < n; ++i)
    {
        TuBlock &b = m.block[m.n + i];
        b.ssd = results[i].ssd, b.ssdPred = results[i].ssdPred, b.cbf = results[i].cbf;
    }
    m.n += n;
}

// One block of ReconstructInterBlock::go: the reconstruction has been written to the block's piece of the cache
// (interPieces[cIdx][1 + rqtdepth]); returns levels, SSDs and cbf.  Blocks issued ahead by prefetchInterCu are found
// in the memo; anything else costs its own round trip.
template <class H> const TuBlock *interBlock(H &h, residual_coding const &rc)
{
    if (!usable(h) || !(enabledMask() & 16)) return nullptr;
    TuMemo &m = tuMemo();
    if (!m.active) return nullptr;
    for (int pass = 0; pass < 2; ++pass)
    {
        for (int i = 0; i < m.n; ++i)
        {
            TuBlock &b = m.block[i];
            if (b.x0 == rc.x0 && b.y0 == rc.y0 && b.cIdx == rc.cIdx && b.log2n == rc.log2TrafoSize) return &b;
        }
        if (pass) break;
        if (m.n == TuMemo::kBlocks) m.n = 0;
        const int one[1][4] = {{rc.x0, rc.y0, rc.cIdx, rc.log2TrafoSize}};
        issueInterBlocks(h, one, 1);
    }
    return nullptr;
}

// reconstructInter (turing/Reconstruct.cpp:1237-...), before the tree walk: forget the previous CU's blocks
inline void beginInterCu() { tuMemo().n = 0, tuMemo().active = false; }


// ReconstructInter<transform_tree>::go at the root of a CU's transform tree (turing/Reconstruct.cpp:58-96), split flag
// known: every block the walk is about to visit goes to the device in one submission.  The children of a split root
// are assumed not to split again (true unless max_transform_hierarchy_depth_inter > 1); a block that is not found in
// the memo later costs its own round trip, nothing else.
template <class H> void prefetchInterCu(H &h, transform_tree const &tt, bool split)
{
    beginInterCu();
    if (!usable(h) || !(enabledMask() & 16)) return;
    typedef typename SampleOf<H>::Type Sample;
    Candidate<Sample> *candidate = h;
    if (candidate->noresidual) return;
    if (tt.log2TrafoSize < tuMinLog2()) return;
    tuMemo().active = true;
    int blocks[TuMemo::kBlocks][4];
    int n = 0;
    const int log2n = tt.log2TrafoSize;
    auto add = [&](int x0, int y0, int cIdx, int log2) {
        if (log2 == 2 && h[transform_skip_enabled_flag()]) return; // transform skip is decided on the host (:860-1030)
        blocks[n][0] = x0, blocks[n][1] = y0, blocks[n][2] = cIdx, blocks[n][3] = log2;
        ++n;
    };
    if (!split)
    {
        if (log2n > 5) return;
        add(tt.x0, tt.y0, 0, log2n);
        add(tt.x0, tt.y0, 1, log2n - 1);
        add(tt.x0, tt.y0, 2, log2n - 1);
    }
    else
    {
        const int child = log2n - 1;
        if (child > 5 || child < 2) return;
        for (int blkIdx = 0; blkIdx < 4; ++blkIdx)
        {
            const int x = tt.x0 + ((blkIdx & 1) << child), y = tt.y0 + ((blkIdx >> 1) << child);
            add(x, y, 0, child);
            if (child > 2)
            {
                add(x, y, 1, child - 1);
                add(x, y, 2, child - 1);
            }
        }
        if (child == 2)
        {
            // 4x4 luma blocks: the 4x4 chroma blocks of the parent are coded with its last child (H.265 7.3.8.10)
            add(tt.x0, tt.y0, 1, 2);
            add(tt.x0, tt.y0, 2, 2);
        }
    }
    issueInterBlocks(h, blocks, n);
}

// ---- a transform block of an intra candidate (ReconstructIntraBlock::go, turing/Reconstruct.cpp:249-356) ---------------------
// The prediction has been written to the block's piece of the reconstruction cache by the reference's own code (it needs
// the neighbours the previous block of the same candidate reconstructed); residual, forward transform, RDOQ + sign hiding or
// plain quantisation, dequantisation, inverse transform + add (in place, as the reference does) and the SSD come back from
// the device.  Profile of the reference on the bench's job: this chain -- its scalar RDOQ above all -- is the largest single
// consumer of host time (profiles/r02f_host_profile.txt), and a block of 16x16 or more costs the host more than a hand-over.
struct IntraTu
{
    uint32_t ssd;
    int cbf;
};

template <class H, class Sample>
bool intraBlock(H &h, residual_coding const &rc, Sample *samples, intptr_t stride, int trType, int16_t *levels, IntraTu &out)
{
    if (!usable(h) || !(enabledMask() & 32)) return false;
    if (rc.log2TrafoSize < intraTuMinLog2() || rc.log2TrafoSize > 5) return false;
    if (rc.log2TrafoSize == 2 && h[transform_skip_enabled_flag()]) return false; // transform skip is decided on the host (:384-...)
    StateEncode *stateEncode = h;
    QpState *qpState = h;
    auto const &entry = qpState->lookup(rc.cIdx);
    const int bitDepth = rc.cIdx ? h[BitDepthC()] : h[BitDepthY()];
    hvb_tu_task t;
    memset(&t, 0, sizeof(t));
    t.src.pic = (int16_t)inputPicture(h);
    t.src.cIdx = (int16_t)rc.cIdx;
    t.src.x = (int16_t)(rc.x0 >> (rc.cIdx ? 1 : 0));
    t.src.y = (int16_t)(rc.y0 >> (rc.cIdx ? 1 : 0));
    t.log2n = (int8_t)rc.log2TrafoSize;
    t.trType = (int8_t)trType;
    t.cIdx = (int8_t)rc.cIdx;
    t.flags = (int8_t)((stateEncode->rdoq ? 1 : 0) | 2 | (h[sign_data_hiding_enabled_flag()] ? 4 : 0));
    t.qscale = entry.quantiseScale;
    t.qshift = entry.quantizeShift - rc.log2TrafoSize;
    t.qoffset = entry.offsetQuantiseShifted;
    t.iqscale = entry.scale;
    t.iqshift = rc.log2TrafoSize - 1 + bitDepth - 8;
    t.scanIdx = (int8_t)h[scanIdx()];
    hvb_rdoq_ctx snapshot;
    if (stateEncode->rdoq) snapshotContexts(*static_cast<Contexts *>(h), static_cast<StateEncodePicture *>(h)->lambda, snapshot);
    const void *pred = samples;
    void *rec = samples;
    hvb_tu_result result;
    const int rcode = hvbenc_tu_chain(sessionOf(h), &t, 1, stateEncode->rdoq ? &snapshot : nullptr, &pred, &stride, &rec, &stride, &levels, &result);
    if (rcode) fatal("hvbenc_tu_chain(intra)", rcode);
    out.ssd = result.ssd;
    out.cbf = result.cbf;
    return true;
}

#endif // HVBHOOKS_RECONSTRUCT

// ---- pictures ------------------------------------------------------------------------------------------------------

// the source picture, once, before the first CTU of the picture is searched (turing/TaskEncodeInput.cpp:247-249)
template <class Sample, class H> void uploadInput(H &h, PictureWrapper &wrapper)
{
    if (!on()) return;
    auto &picture = static_cast<PictureWrap<Sample> &>(wrapper);
    hvbenc *enc = sessionOf(h);
    const int pic = pictureId(&wrapper, true);
    for (int c = 0; c < 3; ++c)
    {
        const int rc = hvbenc_upload_rect(enc, pic, c, picture[c].p, picture[c].stride, 0, 0, picture[c].width, picture[c].height);
        if (rc) fatal("hvbenc_upload_rect(source)", rc);
    }
}

// A CTU of a reference picture the in-loop filters have finished, with the padding TaskSao / TaskDeblock produced for
// it (turing/TaskSao.cpp:123-153, 80 samples beyond the picture edges), before `saoed` is signalled for it: every
// sample a dependent picture's search may read (turing/TaskEncodeSubstream.cpp:71-93, Search.hpp:1378-1394) is on the
// device before that search can be issued.
template <class Sample, class H> void uploadReconstructedCtu(H &h, Picture<Sample> &picture, int rx, int ry)
{
    if (!on()) return;
    if (!(enabledMask() & 7)) return; // no search / PU-cost hook: nothing on the device reads a reference picture
    if (isSubLayerNonReferencePicture(h[nal_unit_type()])) return;
    hvbenc *enc = sessionOf(h);
    const int pic = pictureId(&picture, false);
    const int pad = 80;
    const bool left = rx == 0, top = ry == 0, right = rx == h[PicWidthInCtbsY()] - 1, bottom = ry == h[PicHeightInCtbsY()] - 1;
    int x0 = rx << h[CtbLog2SizeY()], y0 = ry << h[CtbLog2SizeY()];
    int width = std::min(h[CtbSizeY()], h[pic_width_in_luma_samples()] - x0);
    int height = std::min(h[CtbSizeY()], h[pic_height_in_luma_samples()] - y0);
    hvbenc_rect rects[3];
    for (int c = 0; c < 3; ++c)
    {
        const int sh = c ? 1 : 0, pd = pad >> sh;
        int x = x0 >> sh, y = y0 >> sh, w = width >> sh, hh = height >> sh;
        if (left) x -= pd, w += pd;
        if (right) w += pd;
        if (top) y -= pd, hh += pd;
        if (bottom) hh += pd;
        auto &plane = picture[c];
        rects[c] = hvbenc_rect{c, &plane(x, y), plane.stride, x, y, w, hh};
    }
    const int rc = hvbenc_upload_rects(enc, pic, rects, 3);
    if (rc) fatal("hvbenc_upload_rects(reconstruction)", rc);
}

} // namespace hvbhooks

#endif
