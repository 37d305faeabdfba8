"""GPU parity: batched SAD / SAD4 / SSD / SATD through the C-ABI vs the oracle, bit-exact."""
import numpy as np
import pytest

from turingcodec_b200 import hvb, synth

pytestmark = pytest.mark.gpu

PU_SIZES = [(64, 64), (64, 48), (64, 32), (64, 16), (48, 64), (32, 64), (32, 32), (32, 24), (32, 16), (32, 8),
            (24, 32), (16, 64), (16, 32), (16, 16), (16, 12), (16, 8), (16, 4), (12, 16), (8, 32), (8, 16),
            (8, 8), (8, 4), (4, 8)]
W, H, PAD = 256, 192, 96


def padded(plane, pad):
    return np.pad(plane, pad, mode="edge")


@pytest.fixture(scope="module", params=[(1, 8), (2, 10)], ids=["u8", "u16"])
def setup(request):
    bps, bd = request.param
    ctx = hvb.Context(0, bps, bd)
    frames = [synth.frame(i, W, H, bd) for i in range(2)]
    if bps == 2:  # use the full 10-bit range with some extreme samples
        frames[1][0][::7, ::5] = 1023
    pics = []
    for f in frames:
        p = ctx.picture_create(W, H, PAD)
        ctx.upload_yuv(p, *f)
        pics.append(p)
    host = [[padded(pl, PAD if c == 0 else PAD // 2) for c, pl in enumerate(f)] for f in frames]
    yield ctx, pics, host, bps
    ctx.close()


def view(host, pic, c_idx, x, y):
    """(array, offset, stride) of sample (x,y) in the padded host copy"""
    pad = PAD if c_idx == 0 else PAD // 2
    a = host[pic][c_idx]
    return a, (int(y) + pad) * a.shape[1] + (int(x) + pad), a.shape[1]


def make_tasks(rng, n, chroma=False):
    t = np.zeros(n, hvb.metric_task_t)
    for i in range(n):
        w, h = PU_SIZES[rng.integers(len(PU_SIZES))]
        c = int(rng.integers(0, 3)) if chroma else 0
        if c:
            w, h = w // 2, h // 2
        pw, ph = (W, H) if c == 0 else (W // 2, H // 2)
        pad = PAD if c == 0 else PAD // 2
        t[i]["w"], t[i]["h"] = w, h
        t[i]["a"]["pic"], t[i]["a"]["cIdx"] = 0, c
        t[i]["a"]["x"] = rng.integers(0, (pw - w) // 4 + 1) * 4 if c == 0 else rng.integers(0, pw - w + 1)
        t[i]["a"]["y"] = rng.integers(0, ph - h + 1)
        t[i]["b"]["pic"], t[i]["b"]["cIdx"] = 1, c
        # candidates reach into the padding, arbitrarily aligned
        t[i]["b"]["x"] = rng.integers(-pad + 1, pw + pad - w - 1)
        t[i]["b"]["y"] = rng.integers(-pad + 1, ph + pad - h - 1)
    return t


def test_padding_matches_edge_replication(setup):
    ctx, pics, host, bps = setup
    # sample the padding through 1x-wide SADs against a known block: simpler, download is unpadded, so
    # compare SAD of a block fully in the padding with the host edge-replicated copy
    rng = np.random.default_rng(0)
    t = make_tasks(rng, 64)
    t["b"]["x"][:32] = -PAD + 1
    t["b"]["y"][32:] = H + PAD - 66
    got = ctx.sad(t)
    from orc import Oracle
    o = Oracle()
    for i in range(t.size):
        a, oa, sa = view(host, 0, 0, t[i]["a"]["x"], t[i]["a"]["y"])
        b, ob, sb = view(host, 1, 0, t[i]["b"]["x"], t[i]["b"]["y"])
        assert got[i] == o.sad(a, oa, sa, b, ob, sb, int(t[i]["w"]), int(t[i]["h"]))


@pytest.mark.parametrize("chroma", [False, True])
def test_sad_ssd_satd(setup, oracle, chroma):
    ctx, pics, host, bps = setup
    rng = np.random.default_rng(11 + chroma)
    t = make_tasks(rng, 600, chroma)
    sad, ssd, satd = ctx.sad(t), ctx.ssd(t), ctx.satd(t)
    for i in range(t.size):
        c = int(t[i]["a"]["cIdx"])
        w, h = int(t[i]["w"]), int(t[i]["h"])
        a, oa, sa = view(host, 0, c, t[i]["a"]["x"], t[i]["a"]["y"])
        b, ob, sb = view(host, 1, c, t[i]["b"]["x"], t[i]["b"]["y"])
        assert sad[i] == oracle.sad(a, oa, sa, b, ob, sb, w, h), (i, w, h)
        assert ssd[i] == oracle.ssd(a, oa, sa, b, ob, sb, w, h), (i, w, h)
        assert satd[i] == oracle.measure_satd(a, oa, sa, b, ob, sb, w, h), (i, w, h)


def test_sad4(setup, oracle):
    ctx, pics, host, bps = setup
    rng = np.random.default_rng(13)
    n = 300
    t = np.zeros(n, hvb.sad4_task_t)
    for i in range(n):
        w, h = PU_SIZES[rng.integers(len(PU_SIZES))]
        t[i]["w"], t[i]["h"] = w, h
        t[i]["src"]["pic"] = 0
        t[i]["src"]["x"] = rng.integers(0, (W - w) // 4 + 1) * 4
        t[i]["src"]["y"] = rng.integers(0, H - h + 1)
        t[i]["ref_pic"] = 1
        t[i]["rx"] = rng.integers(-PAD + 1, W + PAD - w - 1, 4)
        t[i]["ry"] = rng.integers(-PAD + 1, H + PAD - h - 1, 4)
    got = ctx.sad4(t)
    for i in range(n):
        w, h = int(t[i]["w"]), int(t[i]["h"])
        a, oa, sa = view(host, 0, 0, t[i]["src"]["x"], t[i]["src"]["y"])
        offs = [view(host, 1, 0, int(t[i]["rx"][k]), int(t[i]["ry"][k]))[1] for k in range(4)]
        b = host[1][0]
        assert list(got[i]) == oracle.sad4(a, oa, sa, b, offs, b.shape[1], w, h), (i, w, h)


def test_sad_and_sad4_staged_by_tma(setup):
    """hvb_set_tma: the same batches with the blocks staged by cp.async.bulk.tensor (csrc/hvb_metrics_tma.cu) -- every PU
    size, luma and chroma, vectors into the padding, both sample widths -- equal the load/store kernels' results, which the
    tests above check against the oracle."""
    ctx, pics, host, bps = setup
    rng = np.random.default_rng(15)
    n = 400
    t4 = np.zeros(n, hvb.sad4_task_t)
    for i in range(n):
        w, h = PU_SIZES[i % len(PU_SIZES)]
        t4[i]["w"], t4[i]["h"] = w, h
        t4[i]["src"]["pic"] = 0
        t4[i]["src"]["x"] = rng.integers(0, (W - w) // 4 + 1) * 4
        t4[i]["src"]["y"] = rng.integers(0, H - h + 1)
        t4[i]["ref_pic"] = 1
        t4[i]["rx"] = rng.integers(-PAD + 1, W + PAD - w - 1, 4)
        t4[i]["ry"] = rng.integers(-PAD + 1, H + PAD - h - 1, 4)
    singles = [make_tasks(rng, 600), make_tasks(rng, 300, True)]
    singles[0]["b"]["x"][:40] = -PAD + 1  # into the padding
    want4, want = ctx.sad4(t4), [ctx.sad(t) for t in singles]
    ctx.set_tma(True)
    try:
        before = ctx.launch_count
        got4, got = ctx.sad4(t4), [ctx.sad(t) for t in singles]
        assert ctx.launch_count == before + 3
    finally:
        ctx.set_tma(False)
    assert np.array_equal(got4, want4)
    for g, w_ in zip(got, want):
        assert np.array_equal(g, w_)


def test_device_memory_path_and_launch_count(setup, oracle):
    """HVB_DEVICE: tasks and results live in torch tensors; the call only enqueues work."""
    import torch
    ctx, pics, host, bps = setup
    rng = np.random.default_rng(14)
    t = make_tasks(rng, 128)
    want = ctx.sad(t)
    d_tasks = torch.from_numpy(t.view(np.uint8).copy()).cuda()
    d_out = torch.zeros(t.size, dtype=torch.int32, device="cuda")
    before = ctx.launch_count
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    ctx.sad(d_tasks.data_ptr(), t.size, d_out.data_ptr(), hvb.DEVICE)
    torch.cuda.synchronize()
    ctx.set_stream(None)
    assert ctx.launch_count == before + 1
    assert np.array_equal(d_out.cpu().numpy(), want)


def test_empty_batch_and_bad_arguments(setup):
    ctx, pics, host, bps = setup
    assert ctx.sad(np.zeros(0, hvb.metric_task_t)).size == 0
    with pytest.raises(hvb.HvbError):
        ctx.picture_create(0, 16, 0)
    with pytest.raises(hvb.HvbError):
        ctx.picture_upload(200, 0, np.zeros((4, 4), ctx.sample_dtype))
