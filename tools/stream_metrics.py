"""Streaming roofline of the batched SAD / SAD4 / SSD / SATD kernels: >= 10^6 candidates whose
blocks do not overlap, over a working set far larger than L2, so every sample comes from HBM exactly once.
Reports algorithmic GB/s (2*w*h*B per candidate, 5*w*h*B per SAD4 call, plus the task and result bytes / CUDA-event
time on the launching stream) against MEASURED_PEAKS.json, for 8- and 16-bit samples, with the second operand either
co-located with the first (16-byte aligned blocks) or placed like a motion-search candidate (arbitrary byte alignment).

usage: python tools/stream_metrics.py [--block 64,32,16,8] [--bps 1,2] [--json out.json]
`measure()` is what bench.py's `stream` block calls."""
import argparse
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def peak_gbs():
    path = ROOT / "MEASURED_PEAKS.json"
    return (float(json.loads(path.read_text())["hbm_gbs"]), "measured") if path.exists() else (6650.0, "fallback")


def measure(device=0, blocks=(64, 32, 16, 8), bps_list=(1, 2), layouts=("colocated", "unaligned"), reps=5, side=32704,
            target_candidates=1 << 20, kinds=("sad", "sad4", "ssd", "satd"), log=None, tma=False):
    import torch
    from turingcodec_b200 import hvb
    peak, peak_src = peak_gbs()
    results = []
    for bps in bps_list:
        ctx = hvb.Context(device, bps, 8 if bps == 1 else 10)
        stream = torch.cuda.Stream(device)
        ctx.set_stream(stream.cuda_stream)
        ctx.set_tma(tma)  # SAD / SAD4 staged by cp.async.bulk.tensor (csrc/hvb_metrics_tma.cu)
        # five luma-only-sized pictures (source + four references); content is irrelevant to bandwidth
        pics = [ctx.picture_create(side, side, 0) for _ in range(5)]
        rng = np.random.default_rng(0)
        # (task coordinates are int16: 32704 = 511 * 64 is the largest side; a 64x64 batch is then 261,000 candidates, 2.1 GB)
        row = rng.integers(0, 256 if bps == 1 else 1024, (256, side)).astype(np.uint8 if bps == 1 else np.uint16)
        reps_y = -(-side // 256)
        for pic in pics:  # uploaded in bands of 256 rows: no host copy of the whole plane
            for band in range(reps_y):
                rows_here = min(256, side - band * 256)
                rc = ctx.lib.hvb_picture_upload_rect(ctx.h, pic, 0, hvb._as_ptr(row), side, 0, band * 256, side, rows_here)
                assert rc == 0, rc
        for n in blocks:
            for layout in layouts:
                pitch = n if layout == "colocated" else n + 16
                cols, rows = (side - 16) // pitch, side // n
                count = min(cols * rows, max(target_candidates, 1))
                idx = np.arange(count)
                bx, by = idx % cols, idx // cols
                dx = np.zeros(count, np.int64) if layout == "colocated" else rng.integers(1, 16, count)
                m = np.zeros(count, hvb.metric_task_t)
                m["a"]["pic"], m["b"]["pic"] = pics[0], pics[1]
                m["a"]["x"], m["a"]["y"] = bx * pitch, by * n
                m["b"]["x"], m["b"]["y"] = bx * pitch + dx, by * n
                m["w"] = m["h"] = n
                # SAD4: the four references of a call are four disjoint blocks of one plane (the havoc signature takes one
                # stride), one per quarter of the picture height, each with its own alignment
                rows4 = (side // 4) // n
                count4 = min(count, cols * rows4)
                s4 = np.zeros(count4, hvb.sad4_task_t)
                i4 = np.arange(count4)
                s4["src"]["pic"], s4["src"]["x"], s4["src"]["y"] = pics[0], (i4 % cols) * pitch, (i4 // cols) * n
                s4["ref_pic"], s4["w"], s4["h"] = pics[1], n, n
                for k in range(4):
                    dk = np.zeros(count4, np.int64) if layout == "colocated" else rng.integers(1, 16, count4)
                    s4["rx"][:, k] = (i4 % cols) * pitch + dk
                    s4["ry"][:, k] = (i4 // cols) * n + k * (side // 4)
                d_m = torch.from_numpy(m.view(np.uint8).reshape(-1).copy()).cuda(device)
                d_s4 = torch.from_numpy(s4.view(np.uint8).reshape(-1).copy()).cuda(device)
                d_out = torch.zeros(4 * count, dtype=torch.int32, device=f"cuda:{device}")
                runs = {"sad": (ctx.sad, d_m, m.size, 2.0 * n * n * bps + 4 + 24),
                        "ssd": (ctx.ssd, d_m, m.size, 2.0 * n * n * bps + 4 + 24),
                        "satd": (ctx.satd, d_m, m.size, 2.0 * n * n * bps + 4 + 24),
                        "sad4": (ctx.sad4, d_s4, s4.size, 5.0 * n * n * bps + 16 + 32)}
                for kind in kinds:
                    fn, d_tasks, cnt, per = runs[kind]
                    for _ in range(2):
                        fn(d_tasks.data_ptr(), cnt, d_out.data_ptr(), hvb.DEVICE)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(stream)
                    for _ in range(reps):
                        fn(d_tasks.data_ptr(), cnt, d_out.data_ptr(), hvb.DEVICE)
                    e1.record(stream)
                    torch.cuda.synchronize(device)
                    ms = e0.elapsed_time(e1) / reps
                    gbs = per * cnt / ms / 1e6
                    r = {"kernel": kind + ("(tma)" if tma and kind in ("sad", "sad4") else ""), "block": n, "bytes_per_sample": bps, "layout": layout, "candidates": int(cnt), "ms": round(ms, 4),
                         "GBps": round(gbs, 1), "frac_of_peak": round(gbs / peak, 4)}
                    results.append(r)
                    if log:
                        log(r)
                del d_m, d_s4, d_out
        ctx.close()
    return {"peak_GBps": peak, "peak_source": peak_src, "working_set_MB": 2 * side * side / 1e6, "side": side,
            "how": "CUDA events on the launching stream, %d back-to-back launches after 2 warm-ups; blocks disjoint, every sample read from HBM once; "
                   "algorithmic bytes = (2 (SAD4: 5) w h B + task + result bytes) x candidates" % reps, "rows": results}


if __name__ == "__main__":
    p = argparse.ArgumentParser()
    p.add_argument("--block", default="64,32,16,8")
    p.add_argument("--bps", default="1,2")
    p.add_argument("--layouts", default="colocated,unaligned")
    p.add_argument("--kinds", default="sad,sad4,ssd,satd")
    p.add_argument("--reps", type=int, default=5)
    p.add_argument("--json", default=None)
    p.add_argument("--tma", action="store_true")
    p.add_argument("--side", type=int, default=32704)
    a = p.parse_args()
    res = measure(0, [int(v) for v in a.block.split(",")], [int(v) for v in a.bps.split(",")], a.layouts.split(","), a.reps, side=a.side,
                  kinds=a.kinds.split(","), log=lambda r: print(json.dumps(r), flush=True), tma=a.tma)
    if a.json:
        Path(a.json).write_text(json.dumps(res, indent=1) + "\n")
