#!/bin/bash
# round 2, GPU call 13: intra transform blocks (residual .. RDOQ .. reconstruction) on the device: which hooks pay on the bench's job
set -x
mkdir -p gpurun_out/c13
E=HVB_ENGINES=32,HVB_FIBERS=128
BIG=HVB_ME_MIN_AREA=1024,HVB_PU_MIN_AREA=1024,HVB_INTRA_MIN_LOG2=5,HVB_TU_MIN_LOG2=5
timeout 1500 python tools/segments_matrix.py gpurun_out/c13/matrix.jsonl \
  off_p12:12:2:HVB_BATCHED=0 \
  itu4:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4 itu5:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=5 itu3:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=3 \
  itu4_spin:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4,HVB_POLL_SLEEP_US=0 \
  tu4:12:2:$E,HVB_HOOKS=48,HVB_INTRA_TU_MIN_LOG2=4,HVB_TU_MIN_LOG2=4 \
  big_itu4:12:2:$E,$BIG,HVB_INTRA_TU_MIN_LOG2=4 \
  itu4_p16:16:1:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4 \
  itu4_sweep5:12:2:$E,HVB_HOOKS=40,HVB_INTRA_TU_MIN_LOG2=4,HVB_INTRA_MIN_LOG2=5 \
  > gpurun_out/c13/matrix.log 2> gpurun_out/c13/matrix.err
cut -c1-300 gpurun_out/c13/matrix.log; tail -n 5 gpurun_out/c13/matrix.err
