"""Event-time hvb_tu_chain_batch per TU size / plane on one 4K frame pass (device-resident tasks)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from turingcodec_b200 import hvb, synth, workload  # noqa: E402

W, H = 3840, 2160
ctx = hvb.Context(0, 1, 8)
frames = [synth.frame(i, W, H, 8) for i in range(3)]
pics = [ctx.picture_create(W, H, 96) for _ in range(9)]
for pic, f in zip(pics, frames):
    ctx.upload_yuv(pic, *f)
tu, coeff_count = workload.tu_tasks(W, H, pics[0], (pics[1], pics[2]), tuple(pics[3:9]), 64)
ctx.rdoq_contexts_upload(workload.rdoq_contexts(64))
ctx.coeff_upload(np.zeros(1, np.int16), coeff_count - 1)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
for flags_name, flags_mask in (("rdoq+sdh", None), ("plain quant", -2)):
    for log2n in (5, 4, 3, 2):
        for c_idx in (0, 1):
            sel = tu[(tu["log2n"] == log2n) & ((tu["cIdx"] == 0) == (c_idx == 0))].copy()
            if not sel.size:
                continue
            if flags_mask is not None:
                sel["flags"] &= flags_mask
            d = torch.from_numpy(sel.view(np.uint8).reshape(-1).copy()).cuda()
            o = torch.zeros(sel.size * 16, dtype=torch.uint8, device="cuda")
            for _ in range(2):
                ctx.tu_chain(d.data_ptr(), sel.size, o.data_ptr(), hvb.DEVICE)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.tu_chain(d.data_ptr(), sel.size, o.data_ptr(), hvb.DEVICE)
            e1.record(stream)
            torch.cuda.synchronize()
            r = o.cpu().numpy().view(hvb.tu_result_t)
            print(f"{flags_name:12s} n={1 << log2n:2d} {'luma' if c_idx == 0 else 'chroma':6s} tus={sel.size:7d} "
                  f"ms={e0.elapsed_time(e1):7.3f} ns/tu={e0.elapsed_time(e1) * 1e6 / sel.size:8.1f} cbf={r['cbf'].mean():.3f}")
