"""libhvb's host side built for the CPU (tests/host_build.py: the csrc files as written, the CUDA runtime replaced by
tests/fake_cuda/fake_cudart.cpp, kernel launches rewritten into launches under the warp-level emulator) and driven through
the real Python binding: hvb_create, pictures, pools, staging and the SURVEY 8f entry points -- deblocking, SAO, SAO
statistics, coded-data records, intra complexity, picture copy -- run the (not yet GPU-run) tests/test_gpu_zz_*.py test
functions end to end without a GPU: ctypes signatures, argument checks, side-information uploads, launch geometry."""
import pytest

import host_build
from turingcodec_b200 import hvb

FILES = ("hvb_context.cu", "hvb_loopfilter.cu", "hvb_codeddata.cu", "hvb_preanalysis.cu")


@pytest.fixture(scope="module")
def host_library(tmp_path_factory):
    # hvb_context.cu's context-snapshot upload calls into hvb_tu.cu, which this subset leaves out
    stubs = "int hvbLaunchRdoqBits(hvb_context *, int, int) { return HVB_OK; }\n"
    return host_build.build(tmp_path_factory.mktemp("host_build"), FILES, stubs)


@pytest.fixture()
def library(host_library, monkeypatch):
    """hvb.py bound to the CPU build for the duration of a test"""
    from pathlib import Path
    monkeypatch.setattr(hvb, "LIB_PATH", Path(host_library._name), raising=False)
    monkeypatch.setattr(hvb, "_lib", None, raising=False)
    yield host_library
    monkeypatch.setattr(hvb, "_lib", None, raising=False)


def test_context_and_pictures(library):
    import numpy as np
    ctx = hvb.Context(0, 1, 8)
    pic, other = ctx.picture_create(64, 32, 16), ctx.picture_create(64, 32, 16)
    rng = np.random.default_rng(1)
    planes = [rng.integers(0, 256, (32, 64), dtype=np.uint8), rng.integers(0, 256, (16, 32), dtype=np.uint8), rng.integers(0, 256, (16, 32), dtype=np.uint8)]
    ctx.upload_yuv(pic, *planes)
    ctx.picture_copy(other, pic)
    for c in range(3):
        assert np.array_equal(ctx.picture_download(other, c, planes[c].shape[1], planes[c].shape[0]), planes[c])
    ctx.picture_destroy(pic)
    with pytest.raises(hvb.HvbError):
        ctx.picture_copy(other, pic)  # destroyed
    ctx.close()


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_in_loop_filters(library, oracle, bps, bit_depth):
    import test_gpu_zz_loopfilter as lf
    lf.test_deblock_matches_oracle(oracle, bps, bit_depth)
    lf.test_sao_matches_oracle(oracle, bps, bit_depth)
    lf.test_sao_statistics_match_oracle(oracle, bps, bit_depth)


def test_coded_residual(library, oracle):
    import test_gpu_zz_codeddata as cd
    cd.test_coded_residual_matches_oracle(oracle)


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_intra_complexity(library, oracle, bps, bit_depth):
    import test_gpu_zz_preanalysis as pa
    pa.test_intra_complexity_matches_oracle(oracle, bps, bit_depth)


@pytest.mark.parametrize("bps,bit_depth", [(1, 8), (2, 10)])
def test_aq_activity_and_scd(library, oracle, bps, bit_depth):
    import test_gpu_zz_preanalysis as pa
    pa.test_aq_activity_and_scd_match_oracle(oracle, bps, bit_depth)
    pa.test_scd_block_stats_refuses_what_the_reference_overreads(oracle)


def test_call_order_errors(library):
    """the entry points refuse, with a message, what cannot work: SAO records before the deblocking records, batches before uploads"""
    import numpy as np
    ctx = hvb.Context(0, 1, 8)
    pic = ctx.picture_create(64, 64, 16)
    with pytest.raises(hvb.HvbError):
        ctx.deblock(np.zeros(1, hvb.deblock_task_t))
    with pytest.raises(hvb.HvbError):
        ctx.sao_info_upload(pic, np.zeros(1, hvb.sao_ctu_t))
    ctx.close()


def test_queue_fast_path_entry_points(library):
    """hvb_picture_reserve, hvb_host_alloc / hvb_signal, hvb_coeff_pool_wrap / hvb_rdoq_contexts_wrap on the CPU build: what the
    submission queue's set-up and completion rest on (hvb_encoder.cpp)"""
    import ctypes as C

    import numpy as np
    ctx = hvb.Context(0, 1, 8)
    # a pool of pictures from one allocation: they behave like any picture; destroying one frees nothing and breaks nothing
    ctx.picture_reserve(64, 32, 16, 3)
    with pytest.raises(hvb.HvbError):
        ctx.picture_reserve(64, 32, 16, 3)  # once per context
    pics = [ctx.picture_create(64, 32, 16) for _ in range(4)]  # the fourth no longer fits the reserve: its own allocation
    rng = np.random.default_rng(3)
    for i, pic in enumerate(pics):
        planes = [rng.integers(0, 256, (32, 64), dtype=np.uint8), rng.integers(0, 256, (16, 32), dtype=np.uint8), rng.integers(0, 256, (16, 32), dtype=np.uint8)]
        ctx.upload_yuv(pic, *planes)
        for c in range(3):
            assert np.array_equal(ctx.picture_download(pic, c, planes[c].shape[1], planes[c].shape[0]), planes[c]), (i, c)
    ctx.picture_destroy(pics[1])
    ctx.picture_destroy(pics[3])
    again = ctx.picture_create(64, 32, 16)
    ctx.upload_yuv(again, *planes)
    assert np.array_equal(ctx.picture_download(again, 0, 64, 32), planes[0])
    # completion flag: stored after the work enqueued before it
    flag = ctx.host_alloc(64)
    word = C.c_int32.from_address(flag)
    word.value = 0
    ctx.signal(flag, 41)
    ctx.sync()
    assert word.value == 41 and ctx.poll() == 1
    # page-locked arrays in place of the context's device arrays: only memory from host_alloc is taken
    pool = ctx.host_alloc(2 * 1024 * 4)
    ctx.coeff_pool_wrap(pool, 1024 * 4)
    ctx.coeff_pool_wrap(None)
    ordinary = np.zeros(1024, np.int16)
    with pytest.raises(hvb.HvbError):
        ctx.coeff_pool_wrap(ordinary.ctypes.data, ordinary.size)
    with pytest.raises(hvb.HvbError):
        ctx.rdoq_contexts_wrap(ordinary.ctypes.data, 4)
    ctx.host_free(pool)
    ctx.host_free(flag)
    ctx.close()
