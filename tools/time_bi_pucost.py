"""4K-scale timing of the bi-prediction refinement (a17) and the PU cost (a18) kernels: one searchMotionBi call and one
uni- plus one bi-predicted measurePuCost per PU of the bench's task list.  usage: python tools/time_bi_pucost.py"""
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
pics = [ctx.picture_create(W, H, 96) for _ in range(4)]
for i, pic in enumerate(pics[:3]):
    ctx.upload_yuv(pic, *synth.frame(i, W, H, 8))
me = workload.me_tasks(W, H, pics[0], pics[1])
rng = np.random.default_rng(5)

bi = np.zeros(me.size, hvb.me_bi_task_t)
for name in ("src_pic", "ref_pic", "x0", "y0", "w", "h", "mvp", "rateMvpFlag", "limitMin", "limitMax"):
    bi[name] = me[name]
bi["other_pic"] = pics[2]
bi["lambda"] = me["lambda"] // 2
bi["mvStart"]["x"], bi["mvStart"]["y"] = 12 + rng.integers(-4, 5, me.size), 8 + rng.integers(-4, 5, me.size)
bi["mvOther"]["x"], bi["mvOther"]["y"] = 24 + rng.integers(-4, 5, me.size), 16 + rng.integers(-4, 5, me.size)
bi["smallWindow"], bi["halfPel"], bi["quarterPel"] = 0, 1, 1

pu = np.zeros(2 * me.size, hvb.pu_cost_task_t)
for k in range(2):
    t = pu[k::2]
    t["src_pic"], t["dst_pic"] = pics[0], -1
    t["ref_pic"][:, 0], t["ref_pic"][:, 1] = pics[1], (pics[2] if k else -1)
    for name in ("x0", "y0", "w", "h"):
        t[name] = me[name]
    t["mvx"][:, 0], t["mvy"][:, 0] = 12 + rng.integers(-4, 5, me.size), 8 + rng.integers(-4, 5, me.size)
    t["mvx"][:, 1], t["mvy"][:, 1] = 24 + rng.integers(-4, 5, me.size), 16 + rng.integers(-4, 5, me.size)


def timed(fn, tasks, out_bytes):
    d = torch.from_numpy(tasks.view(np.uint8).reshape(-1).copy()).cuda()
    o = torch.zeros(out_bytes, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        fn(d.data_ptr(), tasks.size, o.data_ptr(), hvb.DEVICE)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        fn(d.data_ptr(), tasks.size, o.data_ptr(), hvb.DEVICE)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3


print(json.dumps({"kernel": "meBiSearchKernel", "tasks": int(bi.size), "ms": round(timed(ctx.me_bi_search, bi, bi.size * 32), 3)}))
print(json.dumps({"kernel": "puCostKernel", "tasks": int(pu.size), "ms": round(timed(ctx.pu_cost, pu, pu.size * 12), 3)}))
