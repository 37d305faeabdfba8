"""The GPU parity tests themselves, run in the CPU-only suite: the test functions of tests/test_gpu_*.py are handed an
EmuScene (tests/emu_context.py) whose context launches the library's kernels -- their own source -- under the warp-level
host emulator instead of on a B200.  Same task generators, same oracle comparisons, same assertions; what differs is only
where the kernels execute.  (The batch-composition, pipelining, drop-in-encoder and frame-pass tests need the real host
library and stay GPU-only; the motion searches, the PU cost and the metrics have their own emulated tests with smaller
batches in tests/test_host_emulated_*.py.)"""
import pytest

import test_gpu_pred_intra as pred_intra
import test_gpu_tu as tu
from emu_context import EmuScene


@pytest.fixture(scope="module", params=[(1, 8), (2, 10)], ids=["u8", "u16-10bit"])
def scene(request):
    return EmuScene(*request.param)


def test_pred_uni_and_bi(scene, oracle):
    pred_intra.test_pred_uni_and_bi(scene, oracle)


def test_subtract_bi(scene, oracle):
    pred_intra.test_subtract_bi(scene, oracle)


def test_interp_satd(scene, oracle):
    pred_intra.test_interp_satd(scene, oracle)


def test_intra_pred_all_modes(scene, oracle):
    pred_intra.test_intra_pred_all_modes(scene, oracle)


def test_forward_and_inverse_transform(scene, oracle):
    tu.test_forward_and_inverse_transform(scene, oracle)


def test_quantize_and_inverse(scene, oracle):
    tu.test_quantize_and_inverse(scene, oracle)


def test_inverse_transform_add(scene, oracle):
    tu.test_inverse_transform_add(scene, oracle)


def test_rdoq_standalone(scene, oracle):
    tu.test_rdoq_standalone(scene, oracle)


@pytest.mark.parametrize("derive_filtered", [True, False])
def test_intra_satd35_sweep(scene, oracle, derive_filtered):
    pred_intra.test_intra_satd35_sweep(scene, oracle, derive_filtered)


@pytest.mark.parametrize("fused", [False, True], ids=["staged", "fused"])
@pytest.mark.parametrize("use_rdoq", [False, True])
def test_tu_chain(scene, oracle, use_rdoq, fused):
    tu.test_tu_chain(scene, oracle, use_rdoq, fused)


def test_pu_cost_matches_oracle_and_reference(scene, oracle):
    import test_gpu_pu_cost as pu_cost
    pu_cost.test_pu_cost_matches_oracle_and_reference(scene, oracle)


def test_pu_cost_stores_the_prediction(scene, oracle):
    import test_gpu_pu_cost as pu_cost
    pu_cost.test_pu_cost_stores_the_prediction(scene, oracle)


@pytest.fixture(scope="module", params=[(1, 8), (2, 10)], ids=["u8", "u16"])
def metrics_setup(request):
    """test_gpu_metrics' fixture on an EmuContext: two pictures, their padded host copies"""
    import numpy as np

    import test_gpu_metrics as gm
    from emu_context import EmuContext
    from turingcodec_b200 import synth
    bps, bd = request.param
    ctx = EmuContext(0, bps, bd)
    frames = [synth.frame(i, gm.W, gm.H, bd) for i in range(2)]
    if bps == 2:
        frames[1][0][::7, ::5] = 1023
    pics = []
    for f in frames:
        p = ctx.picture_create(gm.W, gm.H, gm.PAD)
        ctx.upload_yuv(p, *f)
        pics.append(p)
    host = [[gm.padded(np.asarray(pl), gm.PAD if c == 0 else gm.PAD // 2) for c, pl in enumerate(f)] for f in frames]
    return ctx, pics, host, bps


def test_padding_matches_edge_replication(metrics_setup):
    import test_gpu_metrics as gm
    gm.test_padding_matches_edge_replication(metrics_setup)


@pytest.mark.parametrize("chroma", [False, True])
def test_sad_ssd_satd(metrics_setup, oracle, chroma):
    import test_gpu_metrics as gm
    gm.test_sad_ssd_satd(metrics_setup, oracle, chroma)


def test_sad4(metrics_setup, oracle):
    import test_gpu_metrics as gm
    gm.test_sad4(metrics_setup, oracle)


def test_me_search_is_independent_of_the_batch():
    """prefixes of a shuffled, size-mixed batch reproduce the full batch's results (four searches per warp with refill from
    the cursor, units of several PUs per warp pass): tests/test_gpu_batching.py's property, at 8 bit"""
    import test_gpu_batching as batching
    batching.test_me_search_is_independent_of_the_batch(EmuScene(1, 8))
