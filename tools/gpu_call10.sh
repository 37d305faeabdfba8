#!/bin/bash
# round 2, GPU call 10: GPU suite, streaming roofline after the SATD sign restructuring / 16-bit tensor-core SATD / aligned SAD fast path,
# ncu of the streaming kernels (report kept on the box, CSV summary back)
set -x
mkdir -p gpurun_out/c10
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c10/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c10/pytest_gpu.log
tail -n 25 gpurun_out/c10/pytest_gpu.log | cut -c1-250
timeout 600 python tools/stream_metrics.py --block 64,32 --json gpurun_out/c10/stream.json > gpurun_out/c10/stream.log 2>&1
cut -c1-200 gpurun_out/c10/stream.log | tail -n 34
timeout 900 ncu --set full --clock-control none -k regex:'sadKernel|sad4Kernel|ssdKernel|satdMma' --launch-skip 0 -o /tmp/stream_kernels \
    python tools/stream_metrics.py --side 16384 --block 64,32 --bps 1,2 --kinds sad,sad4,ssd,satd --reps 1 > gpurun_out/c10/ncu_stream.log 2>&1
ncu -i /tmp/stream_kernels.ncu-rep --page raw --csv > /tmp/stream_kernels_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('/tmp/stream_kernels_raw.csv')))
hdr,units=rows[0],rows[1]
keep=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active',
      'sm__warps_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active','sm__pipe_tensor_op_imma_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','launch__registers_per_thread',
      'smsp__average_warp_latency_issue_stalled_wait.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
      'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
      'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
idx=[hdr.index(k) for k in keep if k in hdr]
with open('gpurun_out/c10/stream_kernels_ncu.csv','w',newline='') as f:
    w=csv.writer(f); w.writerow([hdr[i] for i in idx]); w.writerow([units[i] for i in idx])
    for r in rows[2:]: w.writerow([r[i] for i in idx])
for r in rows[2:]:
    g=lambda k: r[hdr.index(k)] if k in hdr else ''
    print(g('Kernel Name')[:48], 'us',g('gpu__time_duration.sum')[:6], 'rd',g('dram__bytes_read.sum')[:7],units[hdr.index('dram__bytes_read.sum')], 'issue',g('smsp__issue_active.avg.pct_of_peak_sustained_active')[:5],'imma',g('sm__pipe_tensor_op_imma_cycles_active.avg.pct_of_peak_sustained_active')[:5],'regs',g('launch__registers_per_thread'))
PY
ls -la /tmp/stream_kernels.ncu-rep
