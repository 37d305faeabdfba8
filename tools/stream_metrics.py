"""Streaming roofline of the batched SAD / SSD / SATD kernels: >= 10^6 candidates whose
blocks do not overlap, over a working set far larger than L2, so every sample comes from HBM exactly once.
Prints algorithmic GB/s (2*w*h*B per candidate / CUDA-event time) against MEASURED_PEAKS.json.
usage: python tools/stream_metrics.py [--pairs 120] [--block 32] [--json out.json]"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from turingcodec_b200 import hvb  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--pairs", type=int, default=120)
p.add_argument("--block", default="32", help="block size, or a comma-separated list")
p.add_argument("--width", type=int, default=3840)
p.add_argument("--height", type=int, default=2160)
p.add_argument("--reps", type=int, default=5)
p.add_argument("--json", default=None)
a = p.parse_args()

ctx = hvb.Context(0, 1, 8)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
W, H = a.width, a.height
pics = [ctx.picture_create(W, H, 0) for _ in range(2 * a.pairs)]
# fill the luma planes (content is irrelevant to bandwidth; random so that nothing is special)
host = np.random.default_rng(0).integers(0, 256, (H, W), dtype=np.uint8)
for pic in pics:
    ctx.picture_upload(pic, 0, np.roll(host, pic, 1))
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
all_res = []
for n in [int(v) for v in str(a.block).split(",")]:
    xs, ys = np.meshgrid(np.arange(W // n) * n, np.arange(H // n) * n)
    per_pair = xs.size
    tasks = np.zeros(per_pair * a.pairs, hvb.metric_task_t)
    for k in range(a.pairs):
        t = tasks[k * per_pair:(k + 1) * per_pair]
        t["a"]["pic"], t["b"]["pic"] = pics[2 * k], pics[2 * k + 1]
        t["a"]["x"] = t["b"]["x"] = xs.reshape(-1)
        t["a"]["y"] = t["b"]["y"] = ys.reshape(-1)
        t["w"] = t["h"] = n
    d_tasks = torch.from_numpy(tasks.view(np.uint8).reshape(-1).copy()).cuda()
    d_out = torch.zeros(tasks.size, dtype=torch.int32, device="cuda")
    alg_bytes = 2.0 * n * n * tasks.size + 4 * tasks.size + tasks.nbytes
    res = {"candidates": int(tasks.size), "block": n, "working_set_MB": 2 * a.pairs * W * H / 1e6, "peak_GBps": peak}
    for name, fn in (("sad", ctx.sad), ("ssd", ctx.ssd), ("satd", ctx.satd)):
        for _ in range(3):
            fn(d_tasks.data_ptr(), tasks.size, d_out.data_ptr(), hvb.DEVICE)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.reps):
            fn(d_tasks.data_ptr(), tasks.size, d_out.data_ptr(), hvb.DEVICE)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        res[name] = {"ms": ms, "GBps": alg_bytes / ms / 1e6, "frac_of_peak": alg_bytes / ms / 1e6 / peak}
        print(name, res[name])
    print(json.dumps(res))
    all_res.append(res)
if a.json:
    Path(a.json).write_text(json.dumps(all_res if len(all_res) > 1 else all_res[0], indent=1) + "\n")
