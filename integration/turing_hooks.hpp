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
    hvb_pu_cost_task t;
    if (!fillPuCostTask(h, pu, puData, t)) return false;
    if (puLookup(t, satd)) return true;
    const int rc = hvbenc_pu_cost(sessionOf(h), &t, 1, satd);
    if (rc) fatal("hvbenc_pu_cost", rc);
    return true;
}

#endif // HVBHOOKS_SEARCH

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
    if (isSubLayerNonReferencePicture(h[nal_unit_type()])) return;
    hvbenc *enc = sessionOf(h);
    const int pic = pictureId(&picture, false);
    const int pad = 80;
    const bool left = rx == 0, top = ry == 0, right = rx == h[PicWidthInCtbsY()] - 1, bottom = ry == h[PicHeightInCtbsY()] - 1;
    int x0 = rx << h[CtbLog2SizeY()], y0 = ry << h[CtbLog2SizeY()];
    int width = std::min(h[CtbSizeY()], h[pic_width_in_luma_samples()] - x0);
    int height = std::min(h[CtbSizeY()], h[pic_height_in_luma_samples()] - y0);
    for (int c = 0; c < 3; ++c)
    {
        const int sh = c ? 1 : 0, pd = pad >> sh;
        int x = x0 >> sh, y = y0 >> sh, w = width >> sh, hh = height >> sh;
        if (left) x -= pd, w += pd;
        if (right) w += pd;
        if (top) y -= pd, hh += pd;
        if (bottom) hh += pd;
        auto &plane = picture[c];
        const int rc = hvbenc_upload_rect(enc, pic, c, &plane(x, y), plane.stride, x, y, w, hh);
        if (rc) fatal("hvbenc_upload_rect(reconstruction)", rc);
    }
}

} // namespace hvbhooks

#endif
