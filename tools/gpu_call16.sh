#!/bin/bash
# round 2, GPU call 16: one-launch transform-block chain (tuFusedKernel) and completion flags instead of stream queries: parity, small-batch
# latency of both forms, and the segment driver again
set -x
mkdir -p gpurun_out/c16
timeout 900 python -m pytest tests/test_gpu_tu.py tests/test_gpu_batching.py tests/test_gpu_batched_encoder.py tests/test_gpu_frame_pass.py -m gpu -x -q > gpurun_out/c16/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c16/pytest_gpu.log
tail -n 6 gpurun_out/c16/pytest_gpu.log | cut -c1-250
LAT_ONLY=tu_chain timeout 600 python tools/latency_small_batches.py --json gpurun_out/c16/latency_tu.json > gpurun_out/c16/latency_tu.log 2>&1
tail -n 50 gpurun_out/c16/latency_tu.log
E=HVB_ENGINES=32,HVB_FIBERS=128
BIG=HVB_ME_MIN_AREA=1024,HVB_PU_MIN_AREA=1024,HVB_INTRA_MIN_LOG2=5,HVB_TU_MIN_LOG2=5
timeout 1500 python tools/segments_matrix.py gpurun_out/c16/matrix.jsonl \
  itu4:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4 \
  itu4_tu27:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4,HVB_ENGINE_SHARES=1,1,1,1,1,27 \
  itu4_staged:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4,HVB_TU_FUSED_MAX=0 \
  itu4_noflag:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4,HVB_DONE_FLAG=0 \
  tu4:12:2:$E,HVB_HOOKS=48,HVB_INTRA_TU_MIN_LOG2=4,HVB_TU_MIN_LOG2=4,HVB_ENGINE_SHARES=1,1,1,1,1,27 \
  big_itu4:12:2:$E,$BIG,HVB_INTRA_TU_MIN_LOG2=4 \
  all:12:2:$E \
  > gpurun_out/c16/matrix.log 2> gpurun_out/c16/matrix.err
cut -c1-330 gpurun_out/c16/matrix.log; tail -n 5 gpurun_out/c16/matrix.err
