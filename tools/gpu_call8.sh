#!/bin/bash
# round 2, GPU call 8: full GPU suite (AQ / SCD, decoder drop-in, TMA fallback, asynchronous unaligned SATD rows),
# streaming roofline on 2 GB batches (LSU and TMA), ncu captures of the streaming kernels
set -x
mkdir -p gpurun_out/c8
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c8/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c8/pytest_gpu.log
tail -n 12 gpurun_out/c8/pytest_gpu.log | cut -c1-300
timeout 600 python tools/stream_metrics.py --block 64,32 --json gpurun_out/c8/stream.json > gpurun_out/c8/stream.log 2>&1
cut -c1-200 gpurun_out/c8/stream.log | tail -n 34
timeout 300 python tools/stream_metrics.py --block 64,32 --kinds sad,sad4 --layouts colocated --tma --json gpurun_out/c8/stream_tma.json > gpurun_out/c8/stream_tma.log 2>&1
cut -c1-200 gpurun_out/c8/stream_tma.log | tail -n 10
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sadKernel|sad4Kernel|satdMmaKernel|satdKernel' -o gpurun_out/c8/stream_kernels \
    python tools/stream_metrics.py --side 16384 --block 64,32 --bps 1,2 --kinds sad,sad4,satd --reps 1 > gpurun_out/c8/ncu_stream.log 2>&1
ncu -i gpurun_out/c8/stream_kernels.ncu-rep --page raw --csv > gpurun_out/c8/stream_kernels_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/c8/stream_kernels_raw.csv')))
hdr=rows[0]
g=lambda r,k: r[hdr.index(k)] if k in hdr else ''
for r in rows[2:]:
    try:
        t=float(g(r,'gpu__time_duration.sum')); rd=float(g(r,'dram__bytes_read.sum')); wr=float(g(r,'dram__bytes_write.sum'))
        print(g(r,'Kernel Name')[:60], 'us',round(t,1), 'dramMB', round(rd+wr,1), g(r,'dram__bytes_read.sum') and hdr and rows[1][hdr.index('dram__bytes_read.sum')], 'issue%', g(r,'smsp__issue_active.avg.pct_of_peak_sustained_active')[:5])
    except Exception as e: print('ERR', e)
PY
