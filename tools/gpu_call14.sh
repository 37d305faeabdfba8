#!/bin/bash
# round 2, GPU call 14: what one small batch costs on the device, per kind and size (the latency a hand-over waits for), and the kernels inside a
# transform-block batch of 1 and 16 blocks (ncu launch list)
set -x
mkdir -p gpurun_out/c14
timeout 600 python tools/latency_small_batches.py --json gpurun_out/c14/latency.json > gpurun_out/c14/latency.log 2>&1
cat gpurun_out/c14/latency.log | tail -n 60
LAT_REPS=1 LAT_NS=1,16 LAT_ONLY=tu_chain timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c14/tu_launches.csv \
    python tools/latency_small_batches.py > gpurun_out/c14/ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/c14/tu_launches.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size') if 'Grid Size' in hdr else None
out=[]
for r in rows[1:]:
    out.append((r[ki][:60], r[gi] if gi is not None else '', r[vi]))
# the last launches: print the tail grouped
for o in out[-140:]: print(o)
PY
