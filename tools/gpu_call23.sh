#!/bin/bash
# round 2, GPU call 23: streaming roofline after the SAD/SSD kernel split, bench with 15 transform-block engines, ncu: launch list of the small-batch
# tool (every kind, both TU forms) and a full capture of tuFusedKernel / sadSmallKernel
set -x
mkdir -p gpurun_out/c23
timeout 300 python -m pytest tests/test_gpu_metrics.py -m gpu -x -q > gpurun_out/c23/pytest_metrics.log 2>&1; echo "rc=$?" >> gpurun_out/c23/pytest_metrics.log; tail -n 3 gpurun_out/c23/pytest_metrics.log
timeout 600 python tools/stream_metrics.py --block 64,32 --json gpurun_out/c23/stream.json > gpurun_out/c23/stream.log 2>&1
cut -c1-160 gpurun_out/c23/stream.log | tail -n 34
( time timeout 1500 python bench.py --no-stream > gpurun_out/c23/bench.json 2> gpurun_out/c23/bench.err ) 2> gpurun_out/c23/bench.time
tail -n 4 gpurun_out/c23/bench.err gpurun_out/c23/bench.time; head -c 300 gpurun_out/c23/bench.json; echo
LAT_REPS=1 LAT_NS=1,16 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/c23/latency_launches.csv \
    python tools/latency_small_batches.py > gpurun_out/c23/ncu_launches.log 2>&1
LAT_REPS=1 LAT_NS=16 LAT_ONLY=tu_chain timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tuFusedKernel' -c 3 -o gpurun_out/c23/tufused \
    python tools/latency_small_batches.py > gpurun_out/c23/ncu_tufused.log 2>&1
ncu -i gpurun_out/c23/tufused.ncu-rep --page raw --csv > gpurun_out/c23/tufused_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:'sadSmallKernel|ssdSmallKernel' -c 4 -o gpurun_out/c23/sadsmall \
    python tools/stream_metrics.py --side 16384 --block 32 --bps 1 --kinds sad,ssd --reps 1 > gpurun_out/c23/ncu_sadsmall.log 2>&1
ncu -i gpurun_out/c23/sadsmall.ncu-rep --page raw --csv > gpurun_out/c23/sadsmall_raw.csv 2>/dev/null
ls -la gpurun_out/c23/
