"""Aggregate an ncu report's per-line samples / executed instructions by source file and by line ranges.
usage: python tools/ncu_by_region.py report.ncu-rep kernel-regex [file:lo-hi=name ...]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep, regex = sys.argv[1], sys.argv[2]
regions = []
for spec in sys.argv[3:]:
    loc, name = spec.split("=")
    f, rng = loc.split(":")
    lo, hi = map(int, rng.split("-"))
    regions.append((f, lo, hi, name))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", f"regex:{regex}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr = None, None
agg = defaultdict(lambda: [0, 0])
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":
        try:
            smp, ins, ln = int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]), int(r[0])
        except ValueError:
            continue
        key = cur_file
        for f, lo, hi, name in regions:
            if f == cur_file and lo <= ln <= hi:
                key = name
                break
        agg[key][0] += smp
        agg[key][1] += ins
ts = sum(v[0] for v in agg.values()) or 1
ti = sum(v[1] for v in agg.values()) or 1
for k, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{100 * s / ts:5.1f}% smp {100 * i / ti:5.1f}% inst  {k}")
