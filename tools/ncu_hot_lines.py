"""Summarise an ncu report's per-source-line warp-stall samples for one kernel.
usage: python tools/ncu_hot_lines.py report.ncu-rep kernel-regex [top]"""
import csv
import subprocess
import sys

rep, regex = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", f"regex:{regex}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
lines, cur_file, hdr = [], None, None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":  # a source line row (Address == "-")
        try:
            lines.append((int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]), cur_file, r[0], r[1].strip()[:100]))
        except ValueError:
            pass
total = sum(x[0] for x in lines) or 1
inst = sum(x[1] for x in lines) or 1
print(f"total samples {total}, warp instructions {inst}")
for smp, ins, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{100 * smp / total:5.1f}% smp {100 * ins / inst:5.1f}% inst  {f}:{ln:>4}  {src}")
