"""Throw-away runner of the last GPU call of round 1: PU-cost timing, SATD staging depth sweep; --ncu: one launch of
each kernel on a reduced working set for an ncu capture."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from turingcodec_b200 import hvb, synth, workload  # noqa: E402

NCU = "--ncu" in sys.argv
ctx = hvb.Context(0, 1, 8)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
W, H = 3840, 2160


def timed(fn, d_tasks, n, d_out, reps=5, warm=2):
    for _ in range(warm):
        fn(d_tasks.data_ptr(), n, d_out.data_ptr(), hvb.DEVICE)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn(d_tasks.data_ptr(), n, d_out.data_ptr(), hvb.DEVICE)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


# ---- PU cost on the bench's PU list: one uni- and one bi-predicted evaluation per PU
pics = [ctx.picture_create(W, H, 96) for _ in range(4)]
for i, pic in enumerate(pics[:3]):
    ctx.upload_yuv(pic, *synth.frame(i, W, H, 8))
me = workload.me_tasks(W, H, pics[0], pics[1])
rng = np.random.default_rng(5)
pu = np.zeros(2 * me.size, hvb.pu_cost_task_t)
for k in range(2):
    t = pu[k::2]
    t["src_pic"], t["dst_pic"] = pics[0], -1
    t["ref_pic"][:, 0], t["ref_pic"][:, 1] = pics[1], (pics[2] if k else -1)
    for name in ("x0", "y0", "w", "h"):
        t[name] = me[name]
    t["mvx"][:, 0], t["mvy"][:, 0] = 12 + rng.integers(-4, 5, me.size), 8 + rng.integers(-4, 5, me.size)
    t["mvx"][:, 1], t["mvy"][:, 1] = 24 + rng.integers(-4, 5, me.size), 16 + rng.integers(-4, 5, me.size)
d = torch.from_numpy(pu.view(np.uint8).reshape(-1).copy()).cuda()
o = torch.zeros(pu.size * 3, dtype=torch.int32, device="cuda")
if NCU:
    ctx.pu_cost(d.data_ptr(), pu.size, o.data_ptr(), hvb.DEVICE)
    torch.cuda.synchronize()
else:
    ms = timed(ctx.pu_cost, d, pu.size, o, reps=3)
    print(json.dumps({"kernel": "puCostKernel", "tasks": int(pu.size), "ms": round(ms, 3), "checksum": int(o.sum().item())}), flush=True)
    for sel, name in ((slice(0, None, 2), "uni"), (slice(1, None, 2), "bi")):
        part = np.ascontiguousarray(pu[sel])
        dp = torch.from_numpy(part.view(np.uint8).reshape(-1).copy()).cuda()
        ms = timed(ctx.pu_cost, dp, part.size, o, reps=3)
        print(json.dumps({"kernel": "puCostKernel", "list": name, "tasks": int(part.size), "ms": round(ms, 3)}), flush=True)
for pic in pics:
    ctx.picture_destroy(pic)

# ---- streaming SATD: staging depth sweep
pairs = 30 if NCU else 120
spics = [ctx.picture_create(W, H, 0) for _ in range(2 * pairs)]
host = np.random.default_rng(0).integers(0, 256, (H, W), dtype=np.uint8)
for pic in spics:
    ctx.picture_upload(pic, 0, np.roll(host, pic, 1))
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
for n in ((64,) if NCU else (64, 32)):
    xs, ys = np.meshgrid(np.arange(W // n) * n, np.arange(H // n) * n)
    per_pair = xs.size
    tasks = np.zeros(per_pair * pairs, hvb.metric_task_t)
    for k in range(pairs):
        t = tasks[k * per_pair:(k + 1) * per_pair]
        t["a"]["pic"], t["b"]["pic"] = spics[2 * k], spics[2 * k + 1]
        t["a"]["x"] = t["b"]["x"] = xs.reshape(-1)
        t["a"]["y"] = t["b"]["y"] = ys.reshape(-1)
        t["w"] = t["h"] = n
    d_tasks = torch.from_numpy(tasks.view(np.uint8).reshape(-1).copy()).cuda()
    d_out = torch.zeros(tasks.size, dtype=torch.int32, device="cuda")
    if NCU:
        ctx.satd(d_tasks.data_ptr(), tasks.size, d_out.data_ptr(), hvb.DEVICE)
        torch.cuda.synchronize()
        break
    alg = 2.0 * n * n * tasks.size + 4 * tasks.size + tasks.nbytes
    ref = None
    for stages in (4, 2, 3, 6, 8):
        os.environ["HVB_SATD_STAGES"] = str(stages)
        ms = timed(ctx.satd, d_tasks, tasks.size, d_out)
        got = d_out.cpu().numpy()
        if ref is None:
            ref = got
        print(json.dumps({"kernel": "satd", "block": n, "stages": stages, "ms": round(ms, 4), "GBps": round(alg / ms / 1e6, 1),
                          "frac": round(alg / ms / 1e6 / peak, 4), "same": bool(np.array_equal(ref, got))}), flush=True)
    os.environ.pop("HVB_SATD_STAGES", None)
