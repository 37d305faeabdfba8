"""The GPU parity tests of the SURVEY 8f kernels (tests/test_gpu_zz_*.py: in-loop filters, SAO statistics, coded-data
records, intra complexity), run in the CPU-only suite on an emulated context.  Those GPU tests have not run on a GPU yet;
executing their own test functions here -- hvb.Context replaced by EmuContext -- checks the tests' logic (task building,
uploads, comparisons) together with the kernels' source, so that their first GPU run can only fail for device reasons."""
import pytest

import emu_context
from turingcodec_b200 import hvb


@pytest.fixture()
def emulated_context(monkeypatch):
    monkeypatch.setattr(hvb, "Context", emu_context.EmuContext)


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_deblock(emulated_context, oracle, bps, bit_depth):
    import test_gpu_zz_loopfilter as lf
    lf.test_deblock_matches_oracle(oracle, bps, bit_depth)


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_sao(emulated_context, oracle, bps, bit_depth):
    import test_gpu_zz_loopfilter as lf
    lf.test_sao_matches_oracle(oracle, bps, bit_depth)


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_sao_statistics(emulated_context, oracle, bps, bit_depth):
    import test_gpu_zz_loopfilter as lf
    lf.test_sao_statistics_match_oracle(oracle, bps, bit_depth)


def test_coded_residual(emulated_context, oracle):
    import test_gpu_zz_codeddata as cd
    cd.test_coded_residual_matches_oracle(oracle)


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_intra_complexity(emulated_context, oracle, bps, bit_depth):
    import test_gpu_zz_preanalysis as pa
    pa.test_intra_complexity_matches_oracle(oracle, bps, bit_depth)
