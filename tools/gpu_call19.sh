#!/bin/bash
# round 2, GPU call 19: the TU batch of the queue as ONE launch (snapshots read from, levels written to page-locked arrays by the kernel; tables derived
# inside the kernel) plus the completion flag
set -x
mkdir -p gpurun_out/c19
timeout 900 python -m pytest tests/test_gpu_tu.py tests/test_gpu_batching.py tests/test_gpu_batched_encoder.py -m gpu -x -q > gpurun_out/c19/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c19/pytest_gpu.log
tail -n 6 gpurun_out/c19/pytest_gpu.log | cut -c1-250
E=HVB_ENGINES=32,HVB_FIBERS=128
timeout 1500 python tools/segments_matrix.py gpurun_out/c19/matrix.jsonl \
  itu4:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4 \
  itu4_tu27:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4,HVB_ENGINE_SHARES=1,1,1,1,1,27 \
  tu4_tu27:12:2:$E,HVB_HOOKS=48,HVB_INTRA_TU_MIN_LOG2=4,HVB_TU_MIN_LOG2=4,HVB_ENGINE_SHARES=1,1,1,1,1,27 \
  itu4_tu27_p16:16:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4,HVB_ENGINE_SHARES=1,1,1,1,1,27 \
  itu3_tu27:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=3,HVB_ENGINE_SHARES=1,1,1,1,1,27 \
  > gpurun_out/c19/matrix.log 2> gpurun_out/c19/matrix.err
cut -c1-200 gpurun_out/c19/matrix.log; tail -n 5 gpurun_out/c19/matrix.err
