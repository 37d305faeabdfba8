"""turingcodec_b200 -- B200-native pixel hot path of the Turing HEVC encoder.

Only what the path needs: ``csrc/`` (sm_100a CUDA kernels + the C-ABI, libhvb.so) and the
host-side mirrors of the reference interface (``hvb`` for Python, csrc/havoc_b200.cpp for C++).
"""
from . import hvb  # noqa: F401

__all__ = ["hvb"]
