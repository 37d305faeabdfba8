#!/bin/bash
# round 2, GPU call 15: is the intra-TU configuration bound by the number of engines that serve transform blocks?
set -x
mkdir -p gpurun_out/c15
timeout 1200 python tools/segments_matrix.py gpurun_out/c15/matrix.jsonl \
  itu4_tu27:12:2:HVB_ENGINES=32,HVB_FIBERS=128,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4,HVB_ENGINE_SHARES=1,1,1,1,1,27 \
  itu4_tu59:12:2:HVB_ENGINES=64,HVB_FIBERS=128,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4,HVB_ENGINE_SHARES=1,1,1,1,1,59 \
  itu4_tu59_p16:16:1:HVB_ENGINES=64,HVB_FIBERS=128,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4,HVB_ENGINE_SHARES=1,1,1,1,1,59 \
  tu4_tu59:12:2:HVB_ENGINES=64,HVB_FIBERS=128,HVB_HOOKS=48,HVB_INTRA_TU_MIN_LOG2=4,HVB_TU_MIN_LOG2=4,HVB_ENGINE_SHARES=1,1,1,1,1,59 \
  > gpurun_out/c15/matrix.log 2> gpurun_out/c15/matrix.err
cut -c1-330 gpurun_out/c15/matrix.log; tail -n 5 gpurun_out/c15/matrix.err
