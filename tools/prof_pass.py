"""Run one frame pass (or one of its three kernels) for profiling under ncu.
usage: python tools/prof_pass.py [--width W --height H] [--only me|intra|tu] [--reps N]"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from turingcodec_b200 import hvb, synth, workload  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--width", type=int, default=1920)
p.add_argument("--height", type=int, default=1080)
p.add_argument("--only", default="all")
p.add_argument("--reps", type=int, default=2)
a = p.parse_args()

import torch  # noqa: E402

ctx = hvb.Context(0, 1, 8)
frames = [synth.frame(i, a.width, a.height, 8) for i in range(3)]
pics = [ctx.picture_create(a.width, a.height, 96) for _ in range(9)]
for pic, f in zip(pics, frames):
    ctx.upload_yuv(pic, *f)
fp = workload.frame_pass(frames[0][0], pics[0], pics[1], (pics[1], pics[2]), tuple(pics[3:9]))
ctx.pool_upload(fp.neighbours)
ctx.rdoq_contexts_upload(fp.rdoq_ctx)
ctx.coeff_upload(np.zeros(1, np.int16), fp.coeff_count - 1)
dev = lambda x: torch.from_numpy(x.view(np.uint8).reshape(-1).copy()).cuda()
d_me, d_intra, d_tu = dev(fp.me), dev(fp.intra), dev(fp.tu)
o_me = torch.zeros(fp.me.size * 56, dtype=torch.uint8, device="cuda")
o_intra = torch.zeros(fp.intra.size * 35, dtype=torch.int32, device="cuda")
o_tu = torch.zeros(fp.tu.size * 16, dtype=torch.uint8, device="cuda")
for _ in range(a.reps):
    if a.only in ("all", "me"):
        ctx.me_search(d_me.data_ptr(), fp.me.size, o_me.data_ptr(), hvb.DEVICE)
    if a.only in ("all", "intra"):
        ctx.intra_satd35(d_intra.data_ptr(), fp.intra.size, o_intra.data_ptr(), hvb.DEVICE)
    if a.only in ("all", "tu"):
        ctx.tu_chain(d_tu.data_ptr(), fp.tu.size, o_tu.data_ptr(), hvb.DEVICE)
ctx.sync()
r = o_me.cpu().numpy().view(hvb.me_result_t)
print("units", fp.units, "mean nSad", float(r["nSad"].mean()), "early", float((r["flags"] & 1).mean()))
