"""Turn an ncu report into the small, reviewable artefacts kept under profiles/:
  <out>_kernels.csv   one row per captured launch with the metrics DESIGN.md / bench.py cite
  traffic.json        kernel name -> dram bytes (read + write) per launch, read by bench.py's roofline.traffic
usage: python tools/summarize_profile.py report.ncu-rep profiles/r01_4k"""
import csv
import json
import re
import subprocess
import sys
from pathlib import Path

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
KEEP = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
keep = [k for k in KEEP if k in idx]


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(value.replace(",", "")) * scale


traffic = {}
with open(out + "_kernels.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(keep)
    w.writerow([units[idx[k]] for k in keep])
    for r in rows[2:]:
        w.writerow([r[idx[k]] for k in keep])
        name = re.sub(r"[<(].*", "", r[idx["Kernel Name"]].split("::")[-1]).strip()
        total = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
            to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        traffic.setdefault(name, []).append(total)
tj = Path(out).parent / "traffic.json"
old = json.loads(tj.read_text()) if tj.exists() else {}
old.update({k: sum(v) / len(v) for k, v in traffic.items()})
tj.write_text(json.dumps(old, indent=1, sort_keys=True) + "\n")
print(open(out + "_kernels.csv").read()[:3000])
