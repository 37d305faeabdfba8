#!/bin/bash
# round 2, GPU call 6: full GPU suite (TMA parity included), TMA streaming roofline + ncu, default bench run
set -x
mkdir -p gpurun_out/c6
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/c6/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c6/pytest_gpu.log
tail -n 6 gpurun_out/c6/pytest_gpu.log
timeout 600 python tools/stream_metrics.py --kinds sad,sad4 --tma --json gpurun_out/c6/stream_tma.json > gpurun_out/c6/stream_tma.log 2>&1
tail -n 34 gpurun_out/c6/stream_tma.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sadTmaKernel' -c 8 -o gpurun_out/c6/tma_kernels \
    python tools/stream_metrics.py --block 64,32 --bps 1,2 --layouts unaligned --kinds sad,sad4 --tma --reps 1 > gpurun_out/c6/ncu_tma.log 2>&1
ncu -i gpurun_out/c6/tma_kernels.ncu-rep --page raw --csv > gpurun_out/c6/tma_kernels_raw.csv 2>/dev/null
( time timeout 1500 python bench.py > gpurun_out/c6/bench.json 2> gpurun_out/c6/bench.err ) 2> gpurun_out/c6/bench.time
tail -n 4 gpurun_out/c6/bench.err gpurun_out/c6/bench.time; head -c 700 gpurun_out/c6/bench.json
