"""Intra sweep time by partition size on the bench's 4K task list.  usage: python tools/intra_breakdown.py"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from turingcodec_b200 import hvb, synth, workload  # noqa: E402

W, H = 3840, 2160
ctx = hvb.Context(0, 1, 8)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
frame = synth.frame(0, W, H, 8)
pic = ctx.picture_create(W, H, 96)
ctx.upload_yuv(pic, *frame)
tasks, pool = workload.intra_tasks(frame[0], pic)
ctx.pool_upload(pool)
for name, mask in [("all", np.ones(tasks.size, bool))] + [(f"n{1 << k}", tasks["log2n"] == k) for k in (5, 4, 3, 2)]:
    sub = tasks[mask]
    d = torch.from_numpy(sub.view(np.uint8).reshape(-1).copy()).cuda()
    o = torch.zeros(sub.size * 35, dtype=torch.int32, device="cuda")
    for _ in range(2):
        ctx.intra_satd35(d.data_ptr(), sub.size, o.data_ptr(), hvb.DEVICE)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5):
        ctx.intra_satd35(d.data_ptr(), sub.size, o.data_ptr(), hvb.DEVICE)
    e1.record(stream)
    torch.cuda.synchronize()
    print(json.dumps({"subset": name, "tasks": int(sub.size), "ms": round(e0.elapsed_time(e1) / 5, 3)}), flush=True)
