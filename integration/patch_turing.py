#!/usr/bin/env python
"""integration/patch_turing.py -- build-time call-site changes of the reference encoder for the batched ABI.

The reference's sources are read where they lie (/root/reference/turing, never edited); the files that need a hook are
written, changed, into a build directory that is neither committed nor shipped (integration/_build/src).  A change is an
anchor (a line of the reference, quoted as narrowly as possible) and the text that goes in front of or behind it; every
anchor must match exactly once, so a reference that moved under the patch fails the build instead of mis-patching.
All logic lives in integration/turing_hooks.hpp; what is inserted here is an include and one forwarding line per call site.

usage: patch_turing.py <reference turing dir> <output dir>"""
import sys
from pathlib import Path

BEFORE, AFTER, REPLACE = "before", "after", "replace"

PATCHES = {
    "Search.hpp": [
        ("#include \"Aps.h\"\n", AFTER,
         "namespace hvbhooks { template <class H> bool intraSweep(H &h, IntraPartition const &intraPartition, int32_t *distortion); }\n"),
        # the 35-mode loop: the candidate bookkeeping of an iteration stays, its prediction + SATD comes from the device
        ("        for (int n = 0; n < 35; ++n)\n        {\n            candidate.resetPieces();\n", BEFORE,
         "        int32_t hvbDistortion[35];\n"
         "        const bool hvbSwept = hvbhooks::intraSweep(h, intraPartition, hvbDistortion);\n"),
        ("                distortion = predictIntraLuma(transform_tree(e.x0 + i, e.y0 + j, e.x0, e.y0, e.log2CbSize - e.split, e.split, e.blkIdx), h);\n", REPLACE,
         "                distortion = hvbSwept ? hvbDistortion[n] : predictIntraLuma(transform_tree(e.x0 + i, e.y0 + j, e.x0, e.y0, e.log2CbSize - e.split, e.split, e.blkIdx), h);\n"),
        # the hooks need LimitFullPelMv / StateMeFullPel (declared above searchMotionBi); searchMotionUni sits above them
        ("template <class H> static void searchMotionUni(H &h, int refList)\n{\n", BEFORE,
         "namespace hvbhooks { template <class H> bool searchMotionUni(H &h, int refList); }\n"),
        ("template <class H> static void searchMotionUni(H &h, int refList)\n{\n", AFTER,
         "    if (hvbhooks::searchMotionUni(h, refList)) return;\n"),
        ("template <class H> static void searchMotionBi(H &h, int refList)\n{\n", BEFORE,
         "#define HVBHOOKS_SEARCH\n#include \"turing_hooks.hpp\"\n"),
        ("template <class H> static void searchMotionBi(H &h, int refList)\n{\n", AFTER,
         "    if (hvbhooks::searchMotionBi(h, refList)) return;\n"),
        # Search<prediction_unit>::go2: results issued ahead for one PU must not outlive it (picture ids are recycled)
        ("            bool debug = false;//pu == prediction_unit(116, 0, 12, 16) && h[PicOrderCntVal()] == 8;\n", AFTER,
         "            hvbhooks::memo().clear();\n"),
        # searchMergeModes: all candidates' predictions + SATDs in one hand-over, ahead of the loop
        ("            populateMergeCandidates(h, pu);\n            for (int i = 0; i < h[MaxNumMergeCand()]; ++i)\n", REPLACE,
         "            populateMergeCandidates(h, pu);\n            hvbhooks::prefetchMergeCosts(h, pu);\n            for (int i = 0; i < h[MaxNumMergeCand()]; ++i)\n"),
        # measurePuCost: the prediction + SATD block
        ("    int32_t satd[3];\n    {\n        StateReconstructedPicture<Sample> *stateReconstructedPicture = h;\n", REPLACE,
         "    int32_t satd[3];\n    if (!hvbhooks::puCost(h, pu, puData, satd))\n    {\n        StateReconstructedPicture<Sample> *stateReconstructedPicture = h;\n"),
    ],
    "Reconstruct.cpp": [
        ("#include \"Rdoq.h\"\n", AFTER, "#define HVBHOOKS_RECONSTRUCT\n#include \"turing_hooks.hpp\"\n"),
        # ReconstructIntraBlock::go: residual .. SSD of one intra transform block from the device (prediction stays here)
        ("        // input picture\n        auto &pictureInput = static_cast<PictureWrap<Sample> &>(*static_cast<StateEncodePicture *>(h)->docket->picture);\n"
         "        auto sourceSamples = pictureInput(rc.x0, rc.y0, rc.cIdx);\n\n        ALIGN(32, int16_t, resSamplesRawBuffer[32 * 32]);\n", AFTER,
         "        ALIGN(32, int16_t, hvbIntraLevels[32 * 32]);\n        hvbhooks::IntraTu hvbIntra;\n"
         "        const bool hvbIntraDone = hvbhooks::intraBlock(h, rc, recSamples.p, (intptr_t)recSamples.stride, nTbS == 4 && rc.cIdx == 0, hvbIntraLevels, hvbIntra);\n"),
        ("        // subtract prediction from input\n        // review: SIMD optimisations, perhaps integrate into forward transform\n        for (int y = 0; y < nTbS; ++y)\n", REPLACE,
         "        // subtract prediction from input\n        // review: SIMD optimisations, perhaps integrate into forward transform\n        if (!hvbIntraDone) for (int y = 0; y < nTbS; ++y)\n"),
        ("        {\n            // forward transform\n            int constexpr bitDepth = 2 * sizeof(Sample) + 6;\n            havoc::table_transform<bitDepth> *table = h;\n", REPLACE,
         "        if (!hvbIntraDone)\n        {\n            // forward transform\n            int constexpr bitDepth = 2 * sizeof(Sample) + 6;\n            havoc::table_transform<bitDepth> *table = h;\n"),
        ("        bool cbf;\n\n        {\n            // review: precompute most of this\n", REPLACE,
         "        bool cbf;\n\n        if (hvbIntraDone)\n        {\n            cbf = hvbIntra.cbf != 0;\n"
         "            memcpy(quantizedCoefficients.p, hvbIntraLevels, sizeof(int16_t) << (2 * rc.log2TrafoSize));\n        }\n        else\n        {\n            // review: precompute most of this\n"),
        ("        {\n            havoc::table_inverse_transform_add<Sample> *table = h;\n            auto *inverseTransformAdd = *havoc::get_inverse_transform_add(table, trType, rc.log2TrafoSize);\n\n"
         "            // inverse transform and add to predicted samples\n", REPLACE,
         "        if (!hvbIntraDone)\n        {\n            havoc::table_inverse_transform_add<Sample> *table = h;\n            auto *inverseTransformAdd = *havoc::get_inverse_transform_add(table, trType, rc.log2TrafoSize);\n\n"
         "            // inverse transform and add to predicted samples\n"),
        ("            backupSSDNoTSkip = stateEncodeSubstream->ssd[rc.cIdx] + ssdFunction(sourceSamples.p, sourceSamples.stride, recSamples.p, recSamples.stride, nTbS, nTbS);\n"
         "            if (!(checkTSkip && cbf))\n", REPLACE,
         "            backupSSDNoTSkip = stateEncodeSubstream->ssd[rc.cIdx] + (hvbIntraDone ? hvbIntra.ssd : ssdFunction(sourceSamples.p, sourceSamples.stride, recSamples.p, recSamples.stride, nTbS, nTbS));\n"
         "            if (!(checkTSkip && cbf))\n"),
        # the root of an inter CU's transform tree: all its blocks in one submission
        ("        stateCodedData->transformTree.word0().split_transform_flag = h[split_transform_flag()];\n\n        Syntax<transform_tree>::go(tt, h);\n", BEFORE,
         "        if (tt.trafoDepth == 0) hvbhooks::prefetchInterCu(h, tt, !!h[split_transform_flag()]);\n"),
        # ReconstructInterBlock::go: transform .. SSD of one block from the device
        ("        bool bUseTSkip = false;\n\n        static int nn = 0;\n", AFTER,
         "        const hvbhooks::TuBlock *hvbTu = (!checkTSkip && candidate->noresidual == 0) ? hvbhooks::interBlock(h, rc) : nullptr;\n"),
        ("        if (candidate->noresidual == 0)\n        {\n            int constexpr bitDepth = 2 * sizeof(Sample) + 6;\n", REPLACE,
         "        if (candidate->noresidual == 0 && !hvbTu)\n        {\n            int constexpr bitDepth = 2 * sizeof(Sample) + 6;\n"),
        ("        bool cbf;\n        if (candidate->noresidual == 0)\n        {\n            if (stateEncode->rdoq)\n", REPLACE,
         "        bool cbf;\n        if (hvbTu)\n        {\n            cbf = hvbTu->cbf != 0;\n"
         "            memcpy(quantizedCoefficients.p, hvbTu->levels, sizeof(int16_t) * n);\n        }\n"
         "        else if (candidate->noresidual == 0)\n        {\n            if (stateEncode->rdoq)\n"),
        ("        if (candidate->noresidual == 0)\n        {\n            // add to prediction\n", REPLACE,
         "        if (candidate->noresidual == 0 && !hvbTu)\n        {\n            // add to prediction\n"),
        ("backupSSDNoTSkip = stateEncodeSubstream->ssd[rc.cIdx] + ssd(sourceSamples.p, sourceSamples.stride, recSamples.p, recSamples.stride, nTbS, nTbS);", REPLACE,
         "backupSSDNoTSkip = stateEncodeSubstream->ssd[rc.cIdx] + (hvbTu ? hvbTu->ssd : ssd(sourceSamples.p, sourceSamples.stride, recSamples.p, recSamples.stride, nTbS, nTbS));"),
        ("stateEncodeSubstream->ssdPrediction[rc.cIdx] += ssd(sourceSamples.p, sourceSamples.stride, predSamples.p, predSamples.stride, nTbS, nTbS);", REPLACE,
         "stateEncodeSubstream->ssdPrediction[rc.cIdx] += (hvbTu ? hvbTu->ssdPred : ssd(sourceSamples.p, sourceSamples.stride, predSamples.p, predSamples.stride, nTbS, nTbS));"),
    ],
    "Search.cpp": [],
    "SearchLzcnt.cpp": [],
    "TaskSao.cpp": [
        ("#include \"Padding.h\"\n", AFTER, "#include \"turing_hooks.hpp\"\n"),
        ("        threadPool->lock();\n        this->syncOut->set(rx, ry);\n", BEFORE,
         "        hvbhooks::uploadReconstructedCtu<Sample>(h, *picture, rx, ry);\n"),
    ],
    "TaskEncodeInput.cpp": [
        ("#include \"SyntaxRbsp.hpp\"\n", AFTER, "#include \"turing_hooks.hpp\"\n"),
        ("    // Enqueue first encoding task and deblocking tasks for theadpool execution\n", BEFORE,
         "    hvbhooks::uploadInput<Sample>(h, *docket->picture);\n"),
    ],
}


def apply(text, patches, name):
    for anchor, mode, insert in patches:
        count = text.count(anchor)
        if count != 1:
            raise SystemExit(f"patch_turing: anchor matches {count} times in {name} (expected 1):\n{anchor}")
        at = text.index(anchor)
        if mode == BEFORE:
            text = text[:at] + insert + text[at:]
        elif mode == AFTER:
            text = text[:at + len(anchor)] + insert + text[at + len(anchor):]
        else:
            text = text[:at] + insert + text[at + len(anchor):]
    return text


def main():
    ref, out = Path(sys.argv[1]), Path(sys.argv[2])
    out.mkdir(parents=True, exist_ok=True)
    for name, patches in PATCHES.items():
        text = (ref / name).read_text()
        (out / name).write_text(apply(text, patches, name))
    print(f"patched {len(PATCHES)} files into {out}")


if __name__ == "__main__":
    main()
