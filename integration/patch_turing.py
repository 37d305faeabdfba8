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
        # the hooks need LimitFullPelMv / StateMeFullPel (declared above searchMotionBi); searchMotionUni sits above them
        ("template <class H> static void searchMotionUni(H &h, int refList)\n{\n", BEFORE,
         "namespace hvbhooks { template <class H> bool searchMotionUni(H &h, int refList); }\n"),
        ("template <class H> static void searchMotionUni(H &h, int refList)\n{\n", AFTER,
         "    if (hvbhooks::searchMotionUni(h, refList)) return;\n"),
        ("template <class H> static void searchMotionBi(H &h, int refList)\n{\n", BEFORE,
         "#define HVBHOOKS_SEARCH\n#include \"turing_hooks.hpp\"\n"),
        ("template <class H> static void searchMotionBi(H &h, int refList)\n{\n", AFTER,
         "    if (hvbhooks::searchMotionBi(h, refList)) return;\n"),
        # measurePuCost: the prediction + SATD block
        ("    int32_t satd[3];\n    {\n        StateReconstructedPicture<Sample> *stateReconstructedPicture = h;\n", REPLACE,
         "    int32_t satd[3];\n    if (!hvbhooks::puCost(h, pu, puData, satd))\n    {\n        StateReconstructedPicture<Sample> *stateReconstructedPicture = h;\n"),
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
