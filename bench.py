#!/usr/bin/env python
"""bench.py -- 4K encode frames per second of the batched B200 build of the Turing encoder, next to the reference on the host.

Metric and configuration are BASELINE.json's: 3840x2160 YUV420 8-bit, `--speed medium`, synthetic frames.
What runs is the reference encoder with its hot loops -- motion search (integer + sub-pel, uni and bi), PU cost, the
35-mode intra sweep and the transform / RDOQ / reconstruction blocks -- on libhvb.so's sm_100a kernels through the
submission queue of include/hvb_encoder.h (integration/), IDR-segment-parallel (integration/segments_main.cpp); its
bitstream is byte-identical to `turing_ref encode --asm 0` with the same options, which the run checks and reports
(`bitstream_md5_equals_asm0`).  Identity configuration: `--speed medium --no-sao` (with SAO the reference itself is not
reproducible from run to run, profiles/r02a_asm0_asm1_experiment.txt).

  step    one IDR segment of --segment-frames (8) pictures; the timed region is ONE job of K segments, sharded over the ranks
          (rank r encodes segments r, r + N, ...: strong scaling -- the host's cores, which run the reference's decision code
          for every rank, are a fixed resource of the box)
  value   frames/s from the encoder's own clock: first picture submitted to last bitstream byte, device session set up
          and the source clip in the page cache before the clock starts (max over ranks)
  e2e     frames/s of the whole `turing_b200_segments` process by wall clock: start-up, CUDA context and session, reading the
          YUV file, every host->device byte (source pictures, finished CTUs of reference pictures, tasks, predictions) and
          device->host byte (results, reconstructions, levels) inside the timed region
  roofline  the kernel group that took most device time during the timed encode (CUDA events per batch, HVB_PROFILE):
          algorithmic bytes / device time against the measured HBM copy bandwidth -- the batches of a host-driven encoder
          are a few tasks deep, so this is far from the roofline; `hot_path_pass` gives the same kernels at saturating
          batch sizes (one whole frame's candidates per launch: round 1's frame pass, verified against the oracle at 4K)
          and `stream` the streaming SAD / SAD4 / SSD / SATD kernels against the HBM roofline
  cpu_baseline  `turing_ref encode --asm 1` (AVX2/xbyak JIT, all host threads) on a bounded clip of the same content and options

`--impl reference` times the unmodified reference encoder alone (rank 0 only under torchrun).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "4K YUV420 8-bit encode fps (medium)"
PASS_METRIC = "hot-path frame passes per second (uni + bi motion search, PU costs, intra sweep, TU/RDOQ for every candidate of one frame, inputs resident)"
UNIT = "frames/s"
# the unmodified reference encoder (test infrastructure, built by oracle/Makefile): the CPU arm and the identity check
REFERENCE_ENCODER = ROOT / "oracle" / "_ref" / "turing_ref"
N_PICS = 9  # source, reference, second prediction source, six reconstruction targets


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=12, help="IDR segments encoded in the timed region (per rank)")
    p.add_argument("--warmup", type=int, default=3, help="segments encoded before it")
    p.add_argument("--segment-frames", type=int, default=8)
    p.add_argument("--parallel-segments", type=int, default=0, help="segments in flight per rank (0: chosen from the host's cores)")
    p.add_argument("--threads", type=int, default=0, help="pool threads per encoder instance (0: chosen from the host's cores)")
    p.add_argument("--concurrent-frames", type=int, default=8)
    p.add_argument("--engines", type=int, default=0, help="dispatcher engines of the submission queue (0: default)")
    p.add_argument("--clip-frames", type=int, default=32, help="distinct frames in the synthetic clip (segments wrap around it)")
    p.add_argument("--no-identity", action="store_true", help="skip the md5 comparison with turing_ref --asm 0")
    p.add_argument("--no-pass", action="store_true", help="skip the hot_path_pass block")
    p.add_argument("--no-stream", action="store_true", help="skip the stream block")
    p.add_argument("--pass-steps", type=int, default=10)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--width", type=int, default=3840)
    p.add_argument("--height", type=int, default=2160)
    p.add_argument("--bit-depth", type=int, default=8, choices=[8, 10],
                   help="10: the 16-bit sample path (BASELINE.json configs[3]); the headline metric is the 8-bit pass")
    p.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    p.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--workdir", default="")
    return p.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def pinned_empty(shape, dtype):
    import torch
    t = torch.empty(int(np.prod(shape)) * np.dtype(dtype).itemsize, dtype=torch.uint8, pin_memory=True)
    return t.numpy().view(dtype).reshape(shape), t


def bit_depth_of(args):
    return int(getattr(args, "bit_depth", 8))


def build_inputs(width, height, bit_depth=8):
    from turingcodec_b200 import synth
    dtype = np.uint8 if bit_depth == 8 else np.uint16
    frames = [[pl.astype(dtype) for pl in synth.frame(i, width, height, bit_depth)] for i in range(3)]
    return frames


class GpuArm:
    def __init__(self, args, device):
        import torch
        from turingcodec_b200 import hvb, workload
        self.torch, self.hvb = torch, hvb
        self.device = device
        torch.cuda.set_device(device)
        bd = bit_depth_of(args)
        self.bps, self.dtype = (1, np.uint8) if bd == 8 else (2, np.uint16)
        self.ctx = hvb.Context(device, self.bps, bd)
        self.stream = torch.cuda.Stream(device)
        self.ctx.set_stream(self.stream.cuda_stream)
        w, h = args.width, args.height
        self.frames = build_inputs(w, h, bd)
        # pictures: 0 source, 1 reference, 2 second prediction source, 3..8 reconstruction targets
        self.pics = [self.ctx.picture_create(w, h, 96) for _ in range(N_PICS)]
        self.pinned = []
        for pic, f in zip(self.pics[:3], self.frames):
            planes = []
            for c, pl in enumerate(f):
                arr, keep = pinned_empty(pl.shape, self.dtype)
                arr[...] = pl
                planes.append((arr, keep))
                self.ctx.picture_upload(pic, c, arr)
            self.ctx.picture_pad(pic)
            self.pinned.append(planes)
        self.fp = workload.frame_pass(self.frames[0][0], self.pics[0], self.pics[1], (self.pics[1], self.pics[2]),
                                      tuple(self.pics[3:9]), bit_depth=bd)
        fp = self.fp
        self.ctx.pool_upload(fp.neighbours)
        self.ctx.rdoq_contexts_upload(fp.rdoq_ctx)
        self.ctx.coeff_upload(np.zeros(1, np.int16), fp.coeff_count - 1)  # size the level pool
        dev = torch.device("cuda", device)

        def to_dev(a):
            return torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(dev)

        self.d_me, self.d_intra, self.d_tu = to_dev(fp.me), to_dev(fp.intra), to_dev(fp.tu)
        self.o_me = torch.zeros(fp.me.size * hvb.me_result_t.itemsize, dtype=torch.uint8, device=dev)
        self.o_intra = torch.zeros(fp.intra.size * 35, dtype=torch.int32, device=dev)
        self.o_tu = torch.zeros(fp.tu.size * hvb.tu_result_t.itemsize, dtype=torch.uint8, device=dev)
        # e2e path: task arrays, neighbours and result buffers live in page-locked host memory (the encoder's arena),
        # so the C-ABI copies from / to them directly
        self._keep = []

        def pin(arr):
            view, keep = pinned_empty(arr.shape, arr.dtype)
            view[...] = arr
            self._keep.append(keep)
            return view

        self.p_me, self.p_intra, self.p_tu, self.p_nb = pin(fp.me), pin(fp.intra), pin(fp.tu), pin(fp.neighbours)
        self.h_me = pin(np.zeros(fp.me.size, hvb.me_result_t))
        self.h_intra = pin(np.zeros((fp.intra.size, 35), np.int32))
        self.h_tu = pin(np.zeros(fp.tu.size, hvb.tu_result_t))
        # The pipelined e2e loop double-buffers what it uploads every step (an encoder uploads frame k+1 while frame k is
        # searched): a second source / reference picture pair, a second region of the neighbour pool, and the task
        # arrays that name them.  Set 0 is the resident run's own.
        self.sets = [dict(pics=self.pics[:2], pool_offset=0, me=self.p_me, intra=self.p_intra, tu=self.p_tu)]
        if getattr(args, "double_buffer", True):
            pics_b = [self.ctx.picture_create(w, h, 96) for _ in range(2)]
            me_b, intra_b, tu_b = fp.me.copy(), fp.intra.copy(), fp.tu.copy()
            for name in ("src_pic", "ref_pic"):
                pic = me_b[name]
                me_b[name] = np.where(pic == self.pics[0], pics_b[0], np.where(pic == self.pics[1], pics_b[1], pic))
            for arr, names in ((intra_b, ("src",)), (tu_b, ("src", "pred", "rec"))):
                for name in names:
                    pic = arr[name]["pic"]
                    arr[name]["pic"] = np.where(pic == self.pics[0], pics_b[0], np.where(pic == self.pics[1], pics_b[1], pic))
            intra_b["nb_unfiltered"] += fp.neighbours.size
            self.sets.append(dict(pics=pics_b, pool_offset=int(fp.neighbours.size), me=pin(me_b), intra=pin(intra_b), tu=pin(tu_b)))
        self.kernel_ms = {"me": 0.0, "intra": 0.0, "tu": 0.0}
        torch.cuda.synchronize(device)
        # The bi-prediction refinement and the PU costs of the same PUs (turing/Search.hpp:1796-1827, :1656-1706): seeded with
        # the vectors the uni-directional search has just found, as Search<prediction_unit>::go2 seeds them; PUs of 8x4 / 4x8
        # are uni-predicted only (:1886).  One searchMotionBi and one uni + one bi measurePuCost per PU.
        self.ctx.me_search(self.d_me.data_ptr(), fp.me.size, self.o_me.data_ptr(), hvb.DEVICE)
        torch.cuda.synchronize(device)
        found = self.o_me.cpu().numpy().view(hvb.me_result_t)
        me = fp.me
        both = (me["w"].astype(np.int32) + me["h"]) != 12
        bi = np.zeros(int(both.sum()), hvb.me_bi_task_t)
        for name in ("src_pic", "ref_pic", "x0", "y0", "w", "h", "mvp", "rateMvpFlag", "limitMin", "limitMax"):
            bi[name] = me[name][both]
        bi["other_pic"] = self.pics[2]
        bi["lambda"] = me["lambda"][both] // 2
        bi["mvStart"], bi["mvOther"] = found["mv"][both], found["mv"][both]
        bi["smallWindow"], bi["halfPel"], bi["quarterPel"] = 0, 1, 1
        pu = np.zeros(me.size + bi.size, hvb.pu_cost_task_t)
        pu["src_pic"], pu["dst_pic"] = self.pics[0], -1
        for name in ("x0", "y0", "w", "h"):
            pu[name][:me.size], pu[name][me.size:] = me[name], bi[name]
        pu["ref_pic"][:, 0], pu["ref_pic"][:me.size, 1], pu["ref_pic"][me.size:, 1] = self.pics[1], -1, self.pics[2]
        pu["mvx"][:me.size, 0], pu["mvy"][:me.size, 0] = found["mv"]["x"], found["mv"]["y"]
        for k in range(2):
            pu["mvx"][me.size:, k], pu["mvy"][me.size:, k] = bi["mvStart"]["x"], bi["mvStart"]["y"]
        self.bi, self.pu = bi, pu
        self.d_bi, self.d_pu = to_dev(bi), to_dev(pu)
        self.o_bi = torch.zeros(bi.size * hvb.me_bi_result_t.itemsize, dtype=torch.uint8, device=dev)
        self.o_pu = torch.zeros(pu.size * 3, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(device)

    # resident step: three launches
    def step_resident(self, events=None):
        hvb, fp = self.hvb, self.fp
        if events is not None:
            events[0].record(self.stream)
        self.ctx.me_search(self.d_me.data_ptr(), fp.me.size, self.o_me.data_ptr(), hvb.DEVICE)
        if events is not None:
            events[1].record(self.stream)
        self.ctx.intra_satd35(self.d_intra.data_ptr(), fp.intra.size, self.o_intra.data_ptr(), hvb.DEVICE)
        if events is not None:
            events[2].record(self.stream)
        self.ctx.tu_chain(self.d_tu.data_ptr(), fp.tu.size, self.o_tu.data_ptr(), hvb.DEVICE)
        if events is not None:
            events[3].record(self.stream)
        self.ctx.me_bi_search(self.d_bi.data_ptr(), self.bi.size, self.o_bi.data_ptr(), hvb.DEVICE)
        if events is not None:
            events[4].record(self.stream)
        self.ctx.pu_cost(self.d_pu.data_ptr(), self.pu.size, self.o_pu.data_ptr(), hvb.DEVICE)
        if events is not None:
            events[5].record(self.stream)

    # host-facing step: upload pictures + neighbours, host task arrays in, host result arrays out
    def step_e2e(self, k: int = 0):
        st = self.sets[k % len(self.sets)]
        for pic, planes in zip(st["pics"], self.pinned[:2]):
            for c, (arr, _) in enumerate(planes):
                self.ctx.picture_upload(pic, c, arr)
            self.ctx.picture_pad(pic)
        self.ctx.pool_upload(self.p_nb, st["pool_offset"])
        self.ctx.me_search(st["me"], out=self.h_me)
        self.ctx.intra_satd35(st["intra"], out=self.h_intra)
        self.ctx.tu_chain(st["tu"], out=self.h_tu)

    def e2e_bytes(self):
        fp = self.fp
        frames = sum(arr.nbytes for planes in self.pinned[:2] for arr, _ in planes)
        h2d = frames + fp.neighbours.nbytes + fp.me.nbytes + fp.intra.nbytes + fp.tu.nbytes
        d2h = self.h_me.nbytes + self.h_intra.nbytes + self.h_tu.nbytes
        return int(h2d), int(d2h)


# ------------------------------------------------------------------------------------------------
# CPU arm (oracle / reference tables)
# ------------------------------------------------------------------------------------------------
class CpuArm:
    """Runs a bounded sample of the same frame pass through oracle_bench.c on the host threads."""

    def __init__(self, args, fp=None, frames=None):
        sys.path.insert(0, str(ROOT / "tests"))
        import orc
        from turingcodec_b200 import hvb, workload
        self.hvb = hvb
        self.oracle = orc.Oracle()
        L = self.oracle.lib
        self.kind = "port"
        L.orc_bench_use_port()
        self.ref = None
        if orc.have_ref():
            try:
                self.ref = orc.Ref(use_asm=True)
                R = self.ref.lib
                ptr = lambda f: C.cast(f, C.c_void_p)
                L.orc_bench_use_reference.argtypes = [C.c_void_p] * 9
                L.orc_bench_use_reference(self.ref.h, ptr(R.ref_sad), ptr(R.ref_hadamard_satd), ptr(R.ref_pred_uni),
                                          ptr(R.ref_pred_intra), ptr(R.ref_transform_fwd), ptr(R.ref_inverse_transform_add),
                                          ptr(R.ref_quantize_inverse), ptr(R.ref_ssd))
                self.kind = "reference"
            except OSError:
                self.ref = None
        self.threads = int(L.orc_bench_threads())
        w, h = args.width, args.height
        self.bd = bit_depth_of(args)
        self.bps, dtype = (1, np.uint8) if self.bd == 8 else (2, np.uint16)
        self.frames = frames if frames is not None else build_inputs(w, h, self.bd)
        pad = 96
        # host pictures 0..4 mirroring the GPU arm's ids; planes padded like the device ones, 64-byte aligned rows
        self.planes = (C.c_void_p * 0)()
        self.host = []
        table = np.zeros((N_PICS * 3, 2), np.int64)
        for pic in range(N_PICS):
            for c in range(3):
                pw, ph, pd = (w, h, pad) if c == 0 else (w // 2, h // 2, pad // 2)
                stride = (pw + 2 * pd + 63) // 64 * 64 + 64
                raw = np.zeros((ph + 2 * pd + 2) * stride + 64, dtype)
                off = ((-raw.ctypes.data) % 64) // self.bps
                buf = raw[off:off + (ph + 2 * pd + 2) * stride].reshape(ph + 2 * pd + 2, stride)
                x0 = (pd + 63) // 64 * 64
                if pic < 3:
                    buf[pd:pd + ph, x0:x0 + pw] = self.frames[pic][c]
                    buf[pd:pd + ph, x0 - pd:x0] = buf[pd:pd + ph, x0:x0 + 1]
                    buf[pd:pd + ph, x0 + pw:x0 + pw + pd] = buf[pd:pd + ph, x0 + pw - 1:x0 + pw]
                    buf[:pd] = buf[pd]
                    buf[pd + ph:pd + ph + pd] = buf[pd + ph - 1]
                self.host.append((raw, buf))
                table[pic * 3 + c] = (buf.ctypes.data + (pd * stride + x0) * self.bps, stride)
        self.table = table
        self.fp = fp if fp is not None else workload.frame_pass(self.frames[0][0], 0, 1, (1, 2), tuple(range(3, 9)), bit_depth=self.bd)
        self.levels = np.zeros(self.fp.coeff_count, np.int16)
        L.orc_bench_me.restype = C.c_double
        L.orc_bench_intra.restype = C.c_double
        L.orc_bench_tu.restype = C.c_double
        L.orc_bench_me.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.orc_bench_intra.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.orc_bench_tu.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]

    def run_fraction(self, frac: float):
        """run the first `frac` of every task list (tasks are ordered depth-major, so strided sampling keeps the mix)"""
        L, fp, hvb = self.oracle.lib, self.fp, self.hvb
        stride = max(1, int(round(1.0 / frac)))
        me = np.ascontiguousarray(fp.me[::stride])
        intra = np.ascontiguousarray(fp.intra[::stride])
        tu = np.ascontiguousarray(fp.tu[::stride])
        o_me = np.zeros(me.size, hvb.me_result_t)
        o_intra = np.zeros((intra.size, 35), np.int32)
        o_tu = np.zeros(tu.size, hvb.tu_result_t)
        p = self.table.ctypes.data
        t = L.orc_bench_me(p, me.ctypes.data, me.size, o_me.ctypes.data, self.bps, self.bd)
        t += L.orc_bench_intra(p, fp.neighbours.ctypes.data, intra.ctypes.data, intra.size, o_intra.ctypes.data, self.bps, self.bd)
        t += L.orc_bench_tu(p, p, fp.rdoq_ctx.ctypes.data, tu.ctypes.data, tu.size, self.levels.ctypes.data, o_tu.ctypes.data, self.bps, self.bd)
        return t, 1.0 / stride, (o_me, o_intra, o_tu, me, intra, tu)

    def measure(self, target_seconds: float):
        # calibrate on 1/64 of a frame, then size the sample for ~target_seconds of wall time: a strided fraction of
        # one frame pass on a slow host, several whole passes back to back on a fast one
        t, f, _ = self.run_fraction(1 / 64)
        per_frame = t / f
        if per_frame > target_seconds:
            frac = max(1 / 512, target_seconds / per_frame)
            t, f, _ = self.run_fraction(frac)
            sample = f"every {int(round(1 / f))}-th task of one frame pass ({f:.4f} frame)"
        else:
            t = f = 0.0
            reps = 0
            while t < target_seconds and reps < 400:
                ti, fi, _ = self.run_fraction(1.0)
                t, f, reps = t + ti, f + fi, reps + 1
            sample = f"{reps} whole frame passes back to back"
        fps = f / t
        return {"value": fps, "unit": UNIT, "cores": self.threads, "kind": self.kind,
                "sample": f"{sample}, {t:.1f} s on {self.threads} threads"}


# ------------------------------------------------------------------------------------------------
def peaks():
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        return float(json.loads(path.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_from_profile(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture (profiles/traffic.json)"""
    path = ROOT / "profiles" / "traffic.json"
    if not path.exists():
        return None
    table = json.loads(path.read_text())
    parts = [table.get(k) for k in kernel.split("+")]
    return None if any(p is None for p in parts) else float(sum(parts))


def ncu_evidence():
    """What the committed ncu capture (profiles/*_kernels.csv, newest) says about each kernel of the pass: issue-slot use,
    DRAM throughput and occupancy -- the pass is issue / latency bound, so these explain `roofline.frac` (an HBM fraction)."""
    import csv
    import re
    files = sorted((ROOT / "profiles").glob("*_4k_kernels.csv"))
    if not files:
        return None
    rows = list(csv.reader(files[-1].open()))
    hdr = rows[0]
    want = {"gpu__time_duration.sum": "ms", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct"}
    out = {"source": files[-1].name}
    for r in rows[2:]:
        name = re.sub(r"[<(].*", "", r[0].split("::")[-1]).strip()
        out[name] = {v: float(r[hdr.index(k)].replace(",", "")) for k, v in want.items() if k in hdr}
    return out


def reference_encoder_fps(width, height, frames=6):
    """The reference's real encoder (oracle/_ref/turing_ref, built unmodified by oracle/Makefile `encoder`) on the same
    synthetic content at `--speed medium`, all host threads: context for the hot-path numbers, not their baseline."""
    exe = ROOT / "oracle" / "_ref" / "turing_ref"
    if not exe.exists():
        return None
    import tempfile
    from turingcodec_b200 import synth
    with tempfile.TemporaryDirectory() as tmp:
        clip = Path(tmp) / "clip.yuv"
        base = synth.frame(0, width, height, 8)
        with open(clip, "wb") as f:
            for i in range(frames):
                for c, plane in enumerate(base):
                    sh = (2 * i, 3 * i) if c == 0 else (i, (3 * i) // 2)
                    f.write(np.roll(plane, sh, (0, 1)).tobytes())
        cmd = [str(exe), "encode", "--input-res", f"{width}x{height}", "--frame-rate", "30", "--frames", str(frames), "--speed", "medium",
               "-o", str(Path(tmp) / "out.bit"), str(clip)]
        t0 = time.time()
        res = subprocess.run(cmd, capture_output=True, text=True)
        wall = time.time() - t0
        if res.returncode != 0:
            return {"error": (res.stdout + res.stderr)[-300:]}
        return {"fps": frames / wall, "frames": frames, "wall_s": wall, "threads": os.cpu_count(),
                "cmd": "turing_ref encode --speed medium (AVX2/xbyak JIT, threads auto, concurrent-frames 4)"}


# ------------------------------------------------------------------------------------------------
# the encoders
# ------------------------------------------------------------------------------------------------
# The submission queue and the hooks as the bench runs them, chosen by end-to-end frames per second on the 16-core B200 box
# (profiles/r02f_segments_matrix_c*.jsonl): the encoder's pool threads are fiber schedulers (integration/fiber_pool.cpp), so an instance needs
# three threads, not dozens, and twelve instances (segments) are in flight; engines are grouped by kind.  HVB_HOOKS / thresholds: which blocks
# leave the host -- the 32x32 transform blocks of intra candidates and the transform trees of 64x64 inter CUs (DCT + RDOQ + reconstruction;
# the transform blocks are 56 % of the reference's host time, profiles/r02f_host_profile.txt).  Wider settings (16x16 blocks, the motion
# search, the sweeps: tools/segments_matrix.py) are bit-identical too and slower end to end: a hand-over's latency is on the row's critical
# path and the serial RDOQ walk of one block is slower on a device thread than on a host core.
QUEUE_ENV = {"HVB_ENGINES": "20", "HVB_ENGINE_SHARES": "1,1,1,1,1,15", "HVB_FIBERS": "128", "HVB_HOOKS": "48", "HVB_INTRA_TU_MIN_LOG2": "5",
             "HVB_TU_MIN_LOG2": "6"}


def host_plan(args, world):
    """instances / pool threads for this host: a pool thread is a fiber scheduler that carries many CTU rows, so threads ~ cores, not rows"""
    cores = os.cpu_count() or 16
    per_rank = max(4, cores // max(1, world))
    parallel = args.parallel_segments or max(2, min(12, (3 * per_rank) // 4, args.steps))
    threads = args.threads or max(2, min(8, (9 * per_rank) // (4 * parallel)))
    return cores, parallel, threads


def encoder_options(args):
    from turingcodec_b200 import encoder
    return [*encoder.MEDIUM, "--concurrent-frames", str(args.concurrent_frames), "--segment", str(args.segment_frames), "--verbosity", "0"]


def workdir(args):
    base = Path(args.workdir) if args.workdir else Path("/dev/shm" if Path("/dev/shm").exists() else "/tmp") / f"hvb_bench_{os.environ.get('MASTER_PORT', os.getpid())}"
    base.mkdir(parents=True, exist_ok=True)
    return base


def run_segments(args, clip, frames, out_dir, tag, parallel, threads, rank=0, device=0, profile=False, ranks=1, extra_env=None):
    """one run of integration/_build/turing_b200_segments; returns (wall seconds, encoder-clock seconds, queue stats, md5)"""
    import re
    from turingcodec_b200 import encoder
    bit = out_dir / f"{tag}.bit"
    cmd = [str(encoder.SEGMENTS), "--parallel-segments", str(parallel), "--clip-frames", str(args.clip_frames)]
    if ranks > 1:
        cmd += ["--segment-rank", str(rank), "--segment-ranks", str(ranks)]
    cmd += ["--input-res", f"{args.width}x{args.height}",
           "--frame-rate", "30", "--frames", str(frames), "--threads", str(threads), "-o", str(bit)]
    if args.bit_depth != 8:
        cmd += ["--bit-depth", str(args.bit_depth)]
    cmd += [*encoder_options(args), str(clip)]
    env = dict(os.environ, HVB_STATS="1", HVB_DEVICE=str(device), HVB_PROFILE="1" if profile else "0")
    env["LD_LIBRARY_PATH"] = str(encoder.LIB_DIR) + ":" + env.get("LD_LIBRARY_PATH", "")
    for name, value in QUEUE_ENV.items():
        env.setdefault(name, value)  # (the caller's environment wins: tools/ sweeps these)
    if args.engines:
        env["HVB_ENGINES"] = str(args.engines)
    if extra_env:
        env.update(extra_env)
    t0 = time.perf_counter()
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    wall = time.perf_counter() - t0
    if res.returncode != 0:
        raise RuntimeError(f"turing_b200_segments failed ({res.returncode}): {res.stdout[-800:]}\n{res.stderr[-2500:]}")
    m = re.search(r"segments wall: ([\d.]+) s", res.stderr)
    inner = float(m.group(1)) if m else wall
    m = re.search(r"hvbenc stats: (\{.*\})", res.stderr)
    stats = json.loads(m.group(1)) if m else {}
    md5, size = None, 0
    if ranks == 1:
        md5 = encoder.md5_file(bit)
        size = bit.stat().st_size
        bit.unlink()
    return wall, inner, stats, md5, size, " ".join(cmd[1:-1])


KIND_KERNELS = {"me": "meSearchSmallKernel+meSearchKernel+meSubpelKernel", "me_bi": "meBiSearchKernel+meSubpelKernel<BI>", "pu_cost": "puCostKernel",
                "intra_sweep": "intraSweepKernel8", "tu_chain": "tuFrontKernel+tuOrderKernel+tuRdoqKernel+tuBackKernel"}


def encode_roofline(stats):
    """the kernel group with the most device time in the timed encode: algorithmic bytes / device time vs the HBM copy peak"""
    kinds = [k for k in KIND_KERNELS if stats.get(k, {}).get("device_ms", 0) > 0]
    if not kinds:
        return None
    peak, peak_src = peaks()
    top = max(kinds, key=lambda k: stats[k]["device_ms"])
    s = stats[top]
    launches = max(1, s["batches"])
    achieved = s["algorithmic_bytes"] / (s["device_ms"] * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": KIND_KERNELS[top], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": s["algorithmic_bytes"] / launches,
            "ms_per_launch": s["device_ms"] / launches, "tasks_per_launch": s["tasks"] / launches,
            "all_kernels": {KIND_KERNELS[k]: {"device_ms": stats[k]["device_ms"], "launches": stats[k]["batches"], "tasks": stats[k]["tasks"],
                                              "algorithmic_GBps": stats[k]["algorithmic_bytes"] / max(stats[k]["device_ms"], 1e-9) / 1e6,
                                              "mean_round_trip_us": stats[k]["mean_wait_us"]} for k in kinds},
            "note": "a host-driven encoder hands the device a few tasks at a time (tasks_per_launch): the kernels are latency-bound here by "
                    "construction; hot_path_pass has the same kernels at a whole frame per launch, stream the SAD/SATD kernels against HBM"}


def issue_ruler():
    """The other ruler for a latency-bound kernel: issue slots of the one warp a block's RDOQ walk runs on, from the
    committed ncu capture of tuFusedKernel -- the kernel behind every transform-block hand-over of the encode (profiles/r02f_tufused_ncu.csv)."""
    import csv
    path = ROOT / "profiles" / "r02f_tufused_ncu.csv"
    if not path.exists():
        return None
    rows = list(csv.reader(open(path)))
    hdr, first = rows[0], rows[2]
    def val(name):
        return float(first[hdr.index(name)]) if name in hdr else None
    issue = val("smsp__issue_active.avg.pct_of_peak_sustained_active")
    return {"bound": "issue slots of a single warp per scheduler (one block's serial RDOQ walk; 16 dense 32x32 blocks, one warp each)",
            "kernel": "tuFusedKernel", "achieved": issue / 100.0, "peak": 1.0, "unit": "instructions/cycle/scheduler", "frac": issue / 100.0,
            "us": val("gpu__time_duration.sum"), "threads_per_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "stall_cycles_per_issue": {"wait": val("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
                                       "long_scoreboard": val("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
                                       "short_scoreboard": val("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio")},
            "dram_kbytes": val("dram__bytes_read.sum"), "source": "profiles/r02f_tufused_ncu.csv (ncu --set full --clock-control none, GPU call 23)",
            "reading": "the walk is a chain of dependent instructions on one thread: a quarter of the issue slots of its scheduler, stalls are "
                       "fixed-latency waits, DRAM traffic 140 KB -- neither HBM nor issue width bounds it, the dependency chain does"}


def reference_encode(args, clip, frames, out_dir, tag, asm, threads=None):
    from turingcodec_b200 import encoder
    opts = ["--asm", str(asm), *encoder_options(args)]
    if args.bit_depth != 8:
        opts += ["--bit-depth", str(args.bit_depth)]
    return encoder.encode(REFERENCE_ENCODER, clip, args.width, args.height, frames, opts, out_dir, tag, threads=threads, dump_reconstruction=False)


def run_reference(args, rank):
    """--impl reference: the unmodified reference encoder (AVX2/xbyak JIT, all host threads) on the same content and options"""
    if rank != 0:
        return
    from turingcodec_b200 import encoder
    if not REFERENCE_ENCODER.exists():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/turing_ref not built (make -C oracle encoder needs /root/reference)"}))
        return
    out = workdir(args)
    seg = args.segment_frames
    clip_frames = args.clip_frames
    clip = out / f"clip_{args.width}x{args.height}_{clip_frames}.yuv"
    if not clip.exists():
        encoder.write_clip(clip, args.width, args.height, clip_frames, args.bit_depth)
    t0 = time.time()
    # bounded: the reference cannot wrap around its input, so a step is a segment of the clip; K steps are timed in runs of
    # at most clip_frames pictures
    for _ in range(min(args.warmup, 1)):
        reference_encode(args, clip, min(seg, clip_frames), out, "refwarm", 1)
    frames_left, wall, inner, runs = args.steps * seg, 0.0, 0.0, 0
    last = None
    while frames_left > 0:
        n = min(frames_left, clip_frames)
        last = reference_encode(args, clip, n, out, "ref", 1)
        wall += last["wall_s"]
        inner += last.get("encoder_wall_s", last["wall_s"])
        frames_left -= n
        runs += 1
    frames = args.steps * seg
    fps = frames / inner
    cores = os.cpu_count()
    line = {"metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * inner / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8" if args.bit_depth == 8 else "u16", "data": "synthetic", "impl": "reference",
            "config": workload_config(args, 1, None, None),
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": f"turing_ref encode --asm 1 ({last['cmd']}), {frames} frames in {runs} run(s) of <= {clip_frames} frames, "
                                       f"{inner:.1f} s by the encoder's clock on {cores} host threads"},
            "e2e": {"value": frames / wall, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0}
    print(json.dumps(line))


def workload_config(args, world, parallel, threads):
    cfg = {"workload": f"{args.width}x{args.height} YUV420 {args.bit_depth}-bit, turing encode --speed medium --no-sao --segment {args.segment_frames} "
                       f"--concurrent-frames {args.concurrent_frames} (BASELINE.json configs[2]; identity configuration, see bench.py docstring), "
                       f"one step = one IDR segment of {args.segment_frames} frames, synthetic clip of {args.clip_frames} frames",
           "frames_per_step": args.segment_frames,
           "l2": "n/a for the encode (working set: 16+ pictures of 12.4 MB per instance, far above L2); hot_path_pass: per-step working set ~370 MB > 126 MB L2"}
    return cfg


def hot_path_pass(args, local, steps):
    """round 1's measurement, kept as the kernels' own benchmark: one launch per kind over every candidate of a 4K frame, inputs
    resident; checked here against the oracle's frame pass on the same inputs (the CPU arm's port) when it is built"""
    import torch
    from turingcodec_b200 import hvb, workload
    arm = GpuArm(args, local)
    for _ in range(3):
        arm.step_resident()
    torch.cuda.synchronize(local)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(steps)]
    for k in range(steps):
        arm.step_resident(ev[k])
    torch.cuda.synchronize(local)
    total_ms = ev[0][0].elapsed_time(ev[-1][5])
    kern = {"me": 0.0, "intra": 0.0, "tu": 0.0, "bi": 0.0, "pu": 0.0}
    for k in range(steps):
        for i, name in enumerate(("me", "intra", "tu", "bi", "pu")):
            kern[name] += ev[k][i].elapsed_time(ev[k][i + 1]) / steps
    fp = arm.fp
    me_out = arm.o_me.cpu().numpy().view(hvb.me_result_t)
    n_sad_samples = int((me_out["nSad"].astype(np.int64) * fp.me["w"].astype(np.int64) * fp.me["h"].astype(np.int64)).sum())
    ab = workload.algorithmic_bytes(fp, 0, arm.bps)
    # SURVEY.md 8d / DESIGN.md section 4: bi search  whB + 121 whB + 18 ((w+7)(h+7) + wh) B per PU; PU cost  3/2 ((w+7)(h+7) + wh) B per list
    B = arm.bps
    wh_bi = arm.bi["w"].astype(np.int64) * arm.bi["h"]
    sup_bi = (arm.bi["w"].astype(np.int64) + 7) * (arm.bi["h"].astype(np.int64) + 7) + wh_bi
    sup_pu = (arm.pu["w"].astype(np.int64) + 7) * (arm.pu["h"].astype(np.int64) + 7) + arm.pu["w"].astype(np.int64) * arm.pu["h"]
    lists = (arm.pu["ref_pic"] >= 0).sum(axis=1)
    alg = {"me": ab["me_fixed"] + n_sad_samples * arm.bps, "intra": ab["intra"], "tu": ab["tu"],
           "bi": int((122 * wh_bi + 18 * sup_bi).sum()) * B, "pu": int((3 * sup_pu * lists).sum()) * B // 2}
    peak, peak_src = peaks()
    names = {"me": "meSearchSmallKernel+meSearchKernel+meSubpelKernel", "intra": "intraSweepKernel8", "tu": "tuFrontKernel+tuRdoqKernel+tuBackKernel",
             "bi": "meBiSearchKernel+meSubpelKernel<BI>", "pu": "puCostKernel"}
    units = dict(fp.units, bi_searches=int(arm.bi.size), pu_costs=int(arm.pu.size))
    out = {"metric": PASS_METRIC, "value": 1000.0 * steps / total_ms, "unit": "passes/s", "steps": steps, "units_per_step": units,
           "kernels": {names[k]: {"ms": kern[k], "algorithmic_GBps": alg[k] / (kern[k] * 1e-3) / 1e9, "frac_of_hbm_peak": alg[k] / (kern[k] * 1e-3) / 1e9 / peak}
                       for k in kern},
           "peak_GBps": peak, "peak_source": peak_src, "ncu": ncu_evidence()}
    # parity at the benchmarked scale: the whole 4K pass against the oracle's port on the host
    verified = None
    try:
        cpu = CpuArm(args, frames=arm.frames)
        cpu.oracle.lib.orc_bench_use_port()
        _, _, (o_me, o_intra, o_tu, *_rest) = cpu.run_fraction(1.0)
        g_me = me_out
        g_intra = arm.o_intra.cpu().numpy().reshape(-1, 35)
        g_tu = arm.o_tu.cpu().numpy().view(hvb.tu_result_t)
        levels = arm.ctx.coeff_download(fp.coeff_count)
        verified = bool(all(np.array_equal(g_me[n], o_me[n]) for n in ("mv", "mvd", "mvInteger", "mvpFlag", "cost", "subpelCost", "nSad", "flags"))
                        and np.array_equal(g_intra, o_intra) and all(np.array_equal(g_tu[n], o_tu[n]) for n in ("ssd", "ssdPred", "cbf", "sadQuad"))
                        and np.array_equal(levels, cpu.levels))
    except (OSError, ImportError, AttributeError) as e:
        out["verify_error"] = str(e)[:200]
    out["verified_vs_oracle"] = verified
    arm.ctx.close()
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from turingcodec_b200 import build, encoder
    if not build.LIB.exists():
        build.build()
    if not encoder.SEGMENTS.exists():
        raise SystemExit("integration/_build/turing_b200_segments is not built (python -c 'import __graft_entry__ as g; g.build()' where /root/reference exists); "
                         "there is no fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(local)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cores, parallel, threads = host_plan(args, world)
    out = workdir(args)
    seg = args.segment_frames
    clip = out / f"clip_{args.width}x{args.height}_{args.clip_frames}.yuv"
    if rank == 0 and not clip.exists():
        encoder.write_clip(clip, args.width, args.height, args.clip_frames, args.bit_depth)
    barrier()
    rank_dir = out / f"rank{rank}"
    rank_dir.mkdir(exist_ok=True)

    # ---- warm-up: W segments (page cache, driver, clocks) -----------------------------------------
    if args.warmup > 0:
        run_segments(args, clip, args.warmup * seg, rank_dir, "warm", parallel, threads, rank, local, profile=False)
    barrier()
    # ---- timed region: one job of K segments, rank r takes segments r, r + world, ... (strong scaling) ------------------
    frames = args.steps * seg
    job_dir = out / "job"
    if rank == 0:
        job_dir.mkdir(exist_ok=True)
    with ClockSampler(local) as clocks:
        barrier()
        wall, inner, stats, _md5, _size, cmd = run_segments(args, clip, frames, job_dir if world > 1 else rank_dir, "timed", parallel, threads,
                                                            rank, local, profile=True, ranks=world)
        barrier()
    if world > 1 and rank == 0:
        parts = [job_dir / f"timed.bit.seg{k}" for k in range(args.steps)]
        from turingcodec_b200 import sharding
        sharding.concat_segments(parts, job_dir / "timed.bit")
        for part in parts:
            part.unlink()
    wall, inner = max_over_ranks(wall), max_over_ranks(inner)
    value = frames / inner
    my_steps = max(1, len(range(rank, args.steps, world)))
    e2e = {"value": frames / wall, "unit": UNIT, "h2d_bytes_per_step": stats.get("h2d_bytes", 0) / my_steps,
           "d2h_bytes_per_step": stats.get("d2h_bytes", 0) / my_steps, "steps": args.steps,
           "timing": "wall clock around the turing_b200_segments process of each rank (max over ranks): process start, CUDA context and session, "
                     "YUV read, encode, bitstream write; h2d = source pictures + finished CTUs of reference pictures + tasks + predictions, "
                     "d2h = results + reconstructions + levels (counted by the submission queue)"}

    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1000.0 * inner / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "u8" if args.bit_depth == 8 else "u16", "data": "synthetic",
                "config": workload_config(args, world, parallel, threads), "host_cores": cores, "cmd": cmd,
                "parallelism": f"{world} GPU(s), per rank up to {parallel} segments in flight x {threads} pool threads; segments sharded over ranks, no collective",
                "roofline": encode_roofline(stats), "issue_ruler": issue_ruler(), "queue": stats, "e2e": None if args.no_e2e else e2e,
                "gpu_launches": int(stats.get("kernel_launches", 0)), "clocks": clocks.summary()}
    # ---- identity: one more run on a clip the reference can encode in one go, against --asm 0 --------------------------------------------
    if rank == 0 and world == 1 and not args.no_identity and REFERENCE_ENCODER.exists():
        n = min(args.clip_frames, 2 * seg + 1)
        _, _, _, md5, size, _ = run_segments(args, clip, n, rank_dir, "ident", parallel, threads, rank, local, profile=False)
        ref0 = reference_encode(args, clip, n, rank_dir, "ref0", 0)
        line["bitstream_md5_equals_asm0"] = bool(md5 == ref0["bitstream_md5"])
        line["identity"] = {"frames": n, "bitstream_bytes": size, "md5": md5, "reference_asm0_md5": ref0["bitstream_md5"],
                            "reference_asm0_fps": ref0["fps"], "cmd": ref0["cmd"]}
    # ---- the same driver with every hook off (the reference's JIT path in 12 instances): what the segment parallelism alone is worth -----
    if rank == 0 and world == 1 and not args.no_cpu:
        n = frames  # the same job: fewer segments would leave instances idle and flatter the device
        w0, i0, _s, _m, _z, _c = run_segments(args, clip, n, rank_dir, "hostonly", parallel, threads, rank, local, extra_env={"HVB_BATCHED": "0"})
        line["host_only_same_driver"] = {"value": n / i0, "e2e": n / w0, "unit": UNIT, "frames": n,
                                         "note": "turing_b200_segments with HVB_BATCHED=0: no device work at all, the reference's AVX2 path in "
                                                 f"{parallel} instances x {threads} threads -- the share of `value` that is the driver's, not the device's"}
    # ---- the reference on the host cores (reported baseline) -----------------------------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu and REFERENCE_ENCODER.exists():
        n = min(args.clip_frames, 3 * seg)
        ref1 = reference_encode(args, clip, n, rank_dir, "ref1", 1)
        line["cpu_baseline"] = {"value": n / ref1.get("encoder_wall_s", ref1["wall_s"]), "unit": UNIT, "cores": cores, "kind": "reference",
                                "sample": f"turing_ref encode --asm 1 on the first {n} frames of the same clip, same options, all host threads "
                                          f"({ref1.get('encoder_wall_s', ref1['wall_s']):.1f} s by the encoder's clock)"}
    # ---- the kernels at saturating batch sizes ---------------------------------------------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_pass:
        line["hot_path_pass"] = hot_path_pass(args, local, args.pass_steps)
    if rank == 0 and world == 1 and not args.no_stream:
        sys.path.insert(0, str(ROOT / "tools"))
        import stream_metrics
        st = stream_metrics.measure(local, blocks=(64, 32), reps=3)
        line["stream"] = st
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
