"""turing_b200_segments on the bench's job (3840x2160 medium --no-sao, --segment 8, 96 frames wrapping a 32-frame clip) under different
queue / threshold / instance settings; one JSON line per run.  usage: python tools/segments_matrix.py OUT.jsonl tag:parallel:threads:ENV=V,ENV=V ..."""
import json
import os
import re
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from turingcodec_b200 import encoder  # noqa: E402

W, H, CLIP, FRAMES, SEG = 3840, 2160, 32, int(os.environ.get("MATRIX_FRAMES", 96)), 8
tmp = Path("/dev/shm/hvb_matrix")
tmp.mkdir(exist_ok=True)
clip = tmp / "clip.yuv"
if not clip.exists():
    encoder.write_clip(clip, W, H, CLIP)
out = open(sys.argv[1], "a")
for spec in sys.argv[2:]:
    tag, parallel, threads, envs = (spec.split(":") + [""])[:4]
    env = dict(os.environ, HVB_STATS="1", HVB_PROFILE="0")
    env["LD_LIBRARY_PATH"] = str(encoder.LIB_DIR) + ":" + env.get("LD_LIBRARY_PATH", "")
    env.update(dict(kv.split("=", 1) for kv in re.split(r",(?=[A-Z_0-9]+=)", envs) if kv))
    bit = tmp / "out.bit"
    cmd = [str(encoder.SEGMENTS), "--parallel-segments", parallel, "--clip-frames", str(CLIP), "--input-res", f"{W}x{H}", "--frame-rate", "30",
           "--frames", str(FRAMES), "--threads", threads, "-o", str(bit), *encoder.MEDIUM, "--concurrent-frames", "8", "--segment", str(SEG),
           "--verbosity", "0", str(clip)]
    t0 = time.perf_counter()
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    wall = time.perf_counter() - t0
    ru = os.times()
    row = {"tag": tag, "parallel": int(parallel), "threads": int(threads), "env": envs, "rc": res.returncode, "wall_s": wall, "fps_e2e": FRAMES / wall}
    if res.returncode == 0:
        m = re.search(r"segments wall: ([\d.]+) s", res.stderr)
        row["fps"] = FRAMES / float(m.group(1)) if m else None
        m = re.search(r"hvbenc stats: (\{.*\})", res.stderr)
        q = json.loads(m.group(1)) if m else {}
        m = re.search(r"hvbenc set-up: (.*)", res.stderr)
        row["setup"] = m.group(1) if m else None
        row["md5"] = encoder.md5_file(bit)
        row["dispatches"], row["busy_s"] = q.get("dispatches"), q.get("engine_busy_s")
        row["kinds"] = {k: (q[k]["requests"], q[k]["batches"], round(q[k]["mean_wait_us"])) for k in ("me", "me_bi", "pu_cost", "intra_sweep", "tu_chain") if k in q}
        row["phases_us"] = {k: (round(q[k].get("mean_queue_us", 0)), round(q[k].get("mean_batch_us", 0)), round(q[k].get("mean_wake_us", 0)))
                            for k in ("me", "me_bi", "pu_cost", "intra_sweep", "tu_chain") if k in q and q[k]["requests"]}
        row["cpu_s"] = ru.children_user + ru.children_system
    else:
        row["err"] = res.stderr[-600:]
    out.write(json.dumps(row) + "\n")
    out.flush()
    print(json.dumps(row)[:600], flush=True)
