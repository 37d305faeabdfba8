"""Where does the motion-search kernel's time go?  Times hvb_me_search_batch on the bench's 4K task list, split by
CU depth and by stage (integer only / + half-pel / + quarter-pel).  usage: python tools/me_breakdown.py [--width W --height H]"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from turingcodec_b200 import hvb, synth, workload  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--width", type=int, default=3840)
p.add_argument("--height", type=int, default=2160)
p.add_argument("--reps", type=int, default=3)
a = p.parse_args()

import torch  # noqa: E402

ctx = hvb.Context(0, 1, 8)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)  # events must be recorded on the stream the kernels are launched on
frames = [synth.frame(i, a.width, a.height, 8) for i in range(2)]
pics = [ctx.picture_create(a.width, a.height, 96) for _ in range(2)]
for pic, f in zip(pics, frames):
    ctx.upload_yuv(pic, *f)
me = workload.me_tasks(a.width, a.height, pics[0], pics[1])


def run(tasks):
    d = torch.from_numpy(tasks.view(np.uint8).reshape(-1).copy()).cuda()
    o = torch.zeros(tasks.size * hvb.me_result_t.itemsize, dtype=torch.uint8, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ctx.me_search(d.data_ptr(), tasks.size, o.data_ptr(), hvb.DEVICE)
    ctx.sync()
    torch.cuda.synchronize()
    ev[0].record(stream)
    for _ in range(a.reps):
        ctx.me_search(d.data_ptr(), tasks.size, o.data_ptr(), hvb.DEVICE)
    ev[1].record(stream)
    torch.cuda.synchronize()
    r = o.cpu().numpy().view(hvb.me_result_t)
    return ev[0].elapsed_time(ev[1]) / a.reps, float(r["nSad"].mean()), float((r["flags"] & 1).mean())


def variant(tasks, half, quarter):
    t = tasks.copy()
    t["halfPel"], t["quarterPel"] = half, quarter
    return t


rows = []
subsets = [("all", np.ones(me.size, bool))]
for depth in range(4):
    subsets.append((f"cu{64 >> depth}", me["log2CbSize"] == 6 - depth))
for name, mask in subsets:
    sub = me[mask]
    for label, half, quarter in (("int", 0, 0), ("int+half", 1, 0), ("full", 1, 1)):
        ms, nsad, early = run(variant(sub, half, quarter))
        rows.append({"subset": name, "stage": label, "tasks": int(sub.size), "ms": round(ms, 3), "mean_nSad": round(nsad, 1),
                     "met_early": round(early, 3)})
        print(json.dumps(rows[-1]), flush=True)
# stream order within the same CU size: 2Nx2N vs the SMP halves
for name, sel in (("cu8_2Nx2N", (me["log2CbSize"] == 3) & (me["w"] == 8) & (me["h"] == 8)),
                  ("cu8_8x4", (me["log2CbSize"] == 3) & (me["h"] == 4)), ("cu8_4x8", (me["log2CbSize"] == 3) & (me["w"] == 4))):
    ms, nsad, early = run(me[sel])
    print(json.dumps({"subset": name, "stage": "full", "tasks": int(sel.sum()), "ms": round(ms, 3), "mean_nSad": round(nsad, 1),
                      "met_early": round(early, 3)}), flush=True)
