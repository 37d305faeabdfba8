"""Device time of the in-loop filter passes on one 3840x2160 8-bit picture: deblocking (all vertical edges, then all
horizontal edges), SAO application, SAO statistics of every CTU and component -- CUDA events on the launching stream, median of 10 calls,
next to the bytes each pass has to move.  Content and side information: the generators of
tests/test_oracle_pin_loopfilter.py at 4K.  usage: python tools/time_filters_4k.py [--json out]"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import test_oracle_pin_loopfilter as pin  # noqa: E402
from turingcodec_b200 import hvb  # noqa: E402

pin.W, pin.H = 3840, 2160
W, H, bps, depth = pin.W, pin.H, 1, 8
rng = np.random.default_rng(5)
rows = []
# content and side information first (numpy, seconds at 4K), then the device
planes, blocks, ctu, stride, ctbs = pin.make_case(rng, bps, depth)
src, views, sblocks, sstride, ctus = pin.make_sao_case(rng, bps, depth)
if "--dry" in sys.argv:
    print("generated", [p.shape for p in planes], blocks.shape, len(ctus))
    raise SystemExit(0)
ctx = hvb.Context(0, bps, depth)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)


def on_device(a):
    return torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).cuda()



def timed(label, call, moved_bytes, reset=None):
    ts = []
    for rep in range(13):
        if reset:
            reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        call()
        e1.record(stream)
        torch.cuda.synchronize()
        if rep >= 3:
            ts.append(e0.elapsed_time(e1) * 1000.0)
    us = float(np.median(ts))
    rows.append({"pass": label, "device_us": round(us, 1), "bytes_moved": int(moved_bytes), "GBps": round(moved_bytes / us / 1e3, 1)})
    print(f"{label:46s} {us:9.1f} us   {moved_bytes / 1e6:7.1f} MB   {moved_bytes / us / 1e3:7.1f} GB/s", flush=True)


def records(blocks, ctu):
    b = np.zeros(blocks.shape[:2], hvb.deblock_block_t)
    b["data"], b["packedBs"] = blocks[..., 0].view(np.int8), blocks[..., 1]
    c = np.zeros(ctu.shape[0], hvb.deblock_ctu_t)
    c["tc_offset_div2"], c["beta_offset_div2"] = ctu[:, 0], ctu[:, 1]
    return b, c


picture_bytes = W * H * 3 // 2 * bps
# ---- deblocking
pic = ctx.picture_create(W, H, 16)
spare = ctx.picture_create(W, H, 16)
for c, p in enumerate(planes):
    ctx.picture_upload(spare, c, p)
ctx.deblock_info_upload(pic, *records(blocks, ctu), ctbs[0], ctbs[1], pin.CTB_LOG2)
for edge, name in ((0, "vertical"), (1, "horizontal")):
    t = np.zeros(1, hvb.deblock_task_t)
    t["pic"], t["edgeType"] = pic, edge
    t["xBegin"], t["yBegin"], t["xEnd"], t["yEnd"] = 0, 0, W, H
    d = on_device(t)
    timed(f"deblock, all {name} edges (deblockKernel)", lambda d=d: ctx.deblock(d.data_ptr(), 1, hvb.DEVICE), 2 * picture_bytes + blocks.size,
          reset=lambda: ctx.picture_copy(pic, spare))
# ---- SAO application
visible = [np.ascontiguousarray(a[v]) for a, v in zip(src, views)]
src_pic, dst_pic = ctx.picture_create(W, H, 16), ctx.picture_create(W, H, 16)
for c, p in enumerate(visible):
    ctx.picture_upload(src_pic, c, p)
ctx.picture_pad(src_pic)
ctx.picture_copy(dst_pic, src_pic)
wc, hc = ctbs
b, _ = records(sblocks, np.zeros((wc * hc, 2), np.int8))
ctx.deblock_info_upload(dst_pic, b, np.zeros(wc * hc, hvb.deblock_ctu_t), wc, hc, pin.CTB_LOG2)
ctx.sao_info_upload(dst_pic, np.frombuffer(bytes(ctus), dtype=hvb.sao_ctu_t).copy())
tasks = np.zeros(1, hvb.sao_task_t)
tasks["src_pic"], tasks["dst_pic"] = src_pic, dst_pic
tasks["ctuBegin"], tasks["ctuEnd"] = 0, wc * hc
tasks["lumaFlag"], tasks["chromaFlag"] = 1, 1
d_sao = on_device(tasks)
timed("SAO, whole picture (saoKernel)", lambda: ctx.sao(d_sao.data_ptr(), 1, hvb.DEVICE), 2 * picture_bytes + 42 * wc * hc)
# ---- SAO statistics: every CTU and component
stats, n = [], 1 << pin.CTB_LOG2
for c in range(3):
    nc, w, h = (n, W, H) if c == 0 else (n // 2, W // 2, H // 2)
    for y0 in range(0, h, nc):
        for x0 in range(0, w, nc):
            stats.append((src_pic, dst_pic, c, x0, y0, min(nc, w - x0), min(nc, h - y0), 0))
st = np.array(stats, dtype=hvb.sao_stats_task_t)
d_st = on_device(st)
d_out = torch.zeros(st.size * hvb.sao_stats_t.itemsize, dtype=torch.uint8, device="cuda")
timed(f"SAO statistics, {st.size} CTU x component (saoStatsKernel)", lambda: ctx.sao_stats(d_st.data_ptr(), st.size, d_out.data_ptr(), hvb.DEVICE),
      2 * picture_bytes + st.size * 416)
if "--json" in sys.argv:
    Path(sys.argv[sys.argv.index("--json") + 1]).write_text(json.dumps(rows))
