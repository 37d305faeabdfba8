"""Device time of ONE small batch per kind (what a hand-over of the batched encoder waits for): hvb_*_batch with n = 1, 4, 16, 64 tasks,
device-resident task arrays, CUDA events on the launching stream, median of 30 calls.  usage: python tools/latency_small_batches.py [--json out]"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from turingcodec_b200 import hvb, synth, workload  # noqa: E402

W, H = 1920, 1088
import os
REPS = int(os.environ.get('LAT_REPS', 30))
NS = tuple(int(v) for v in os.environ.get('LAT_NS', '1,4,16,64').split(','))
ONLY = os.environ.get('LAT_ONLY', '')
ctx = hvb.Context(0, 1, 8)
frames = [synth.frame(i, W, H, 8) for i in range(3)]
pics = [ctx.picture_create(W, H, 96) for _ in range(9)]
for pic, f in zip(pics, frames):
    ctx.upload_yuv(pic, *f)
fp = workload.frame_pass(frames[0][0], pics[0], pics[1], (pics[1], pics[2]), tuple(pics[3:9]), n_ctx=64)
ctx.pool_upload(fp.neighbours)
ctx.rdoq_contexts_upload(fp.rdoq_ctx)
ctx.coeff_upload(np.zeros(1, np.int16), fp.coeff_count - 1)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
rows = []


def timed(call, tasks, out_bytes, n):
    d = torch.from_numpy(tasks[:n].copy().view(np.uint8).reshape(-1)).cuda()
    o = torch.zeros(max(16, out_bytes * n), dtype=torch.uint8, device="cuda")
    ts = []
    for rep in range(REPS + 4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        call(d.data_ptr(), n, o.data_ptr(), hvb.DEVICE)
        e1.record(stream)
        torch.cuda.synchronize()
        if rep >= 4:
            ts.append(e0.elapsed_time(e1) * 1000.0)
    return float(np.median(ts))


def sweep(kind, label, call, tasks, out_bytes):
    for n in NS:
        if tasks.size < n or (ONLY and kind not in ONLY):
            continue
        us = timed(call, tasks, out_bytes, n)
        rows.append({"kind": kind, "tasks": label, "n": n, "device_us": round(us, 1)})
        print(f"{kind:12s} {label:22s} n={n:3d}  {us:8.1f} us", flush=True)


me = fp.me
for w in (64, 32, 16, 8):
    sweep("me", f"{w}x{w} PU", ctx.me_search, me[(me["w"] == w) & (me["h"] == w)], hvb.me_result_t.itemsize)
intra = fp.intra
for log2n in (5, 4, 3):
    sweep("intra_sweep", f"{1 << log2n}x{1 << log2n}", ctx.intra_satd35, intra[intra["log2n"] == log2n], 140)
tu = fp.tu
for form, fused_max in (("staged", 0), ("fused", 1 << 20)):
    ctx.set_tu_fused_max(fused_max)
    for log2n in (5, 4, 3):
        sel = tu[(tu["log2n"] == log2n) & (tu["cIdx"] == 0)]
        sweep("tu_chain", f"{form} {1 << log2n}x{1 << log2n} rdoq+sdh", ctx.tu_chain, sel, hvb.tu_result_t.itemsize)
        plain = sel.copy()
        plain["flags"] &= -2
        sweep("tu_chain", f"{form} {1 << log2n}x{1 << log2n} plain", ctx.tu_chain, plain, hvb.tu_result_t.itemsize)
if "--json" in sys.argv:
    Path(sys.argv[sys.argv.index("--json") + 1]).write_text(json.dumps(rows))
