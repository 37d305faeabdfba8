#!/bin/bash
# round 2, GPU call 1: parity suite, sanitizer passes, streaming roofline + ncu captures of the streaming kernels
set -x
mkdir -p gpurun_out/c1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1/smi.txt 2>&1
nproc > gpurun_out/c1/nproc.txt; lscpu | head -20 >> gpurun_out/c1/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c1/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1/pytest_gpu.log
timeout 900 python tools/stream_metrics.py --json gpurun_out/c1/stream.json > gpurun_out/c1/stream.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sadKernel|sad4Kernel|ssdKernel|satdMma' -c 24 -o gpurun_out/c1/stream_kernels \
    python tools/stream_metrics.py --block 64,32 --bps 1,2 --layouts unaligned --reps 1 > gpurun_out/c1/ncu_stream.log 2>&1
ncu -i gpurun_out/c1/stream_kernels.ncu-rep --page raw --csv > gpurun_out/c1/stream_kernels_raw.csv 2>/dev/null
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_pu_cost.py tests/test_gpu_metrics.py tests/test_gpu_tu.py tests/test_gpu_zz_codeddata.py tests/test_gpu_zz_loopfilter.py tests/test_gpu_zz_preanalysis.py -m gpu -x -q > gpurun_out/c1/sanitizer_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/c1/sanitizer_$tool.log
done
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/c1/bench.json 2> gpurun_out/c1/bench.err
tail -3 gpurun_out/c1/pytest_gpu.log; tail -2 gpurun_out/c1/sanitizer_*.log; cat gpurun_out/c1/bench.json | head -c 600
