#!/bin/bash
# round 2, GPU call 21: the host-bound end of the range: only the 32x32 transform blocks leave the host; how many pool threads keep the cores busy
set -x
mkdir -p gpurun_out/c21
E=HVB_ENGINES=32,HVB_FIBERS=128,HVB_ENGINE_SHARES=1,1,1,1,1,27
timeout 1500 python tools/segments_matrix.py gpurun_out/c21/matrix.jsonl \
  itu5_t3:12:3:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=5 \
  itu5_t4:12:4:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=5 \
  tu5_t4:12:4:$E,HVB_HOOKS=48,HVB_INTRA_TU_MIN_LOG2=5,HVB_TU_MIN_LOG2=6 \
  tu5_p16_t3:16:3:$E,HVB_HOOKS=48,HVB_INTRA_TU_MIN_LOG2=5,HVB_TU_MIN_LOG2=6 \
  off_p12_t4:12:4:HVB_BATCHED=0 \
  > gpurun_out/c21/matrix.log 2> gpurun_out/c21/matrix.err
cut -c1-200 gpurun_out/c21/matrix.log; tail -n 5 gpurun_out/c21/matrix.err
