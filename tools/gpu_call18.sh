#!/bin/bash
# round 2, GPU call 18: notification without the pool mutex, one notification per scheduler and batch, no spinning; completion by flag + completion
# thread against dispatcher threads asleep in the driver's blocking wait
set -x
mkdir -p gpurun_out/c18
E=HVB_ENGINES=32,HVB_FIBERS=128
timeout 1500 python tools/segments_matrix.py gpurun_out/c18/matrix.jsonl \
  itu4:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4 \
  itu4_block:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4,HVB_BLOCKING_SYNC=1,HVB_POLLER=0 \
  itu4_block_tu27:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4,HVB_BLOCKING_SYNC=1,HVB_POLLER=0,HVB_ENGINE_SHARES=1,1,1,1,1,27 \
  itu4_p16:16:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4 \
  itu5_block:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=5,HVB_BLOCKING_SYNC=1,HVB_POLLER=0 \
  > gpurun_out/c18/matrix.log 2> gpurun_out/c18/matrix.err
cut -c1-200 gpurun_out/c18/matrix.log; tail -n 5 gpurun_out/c18/matrix.err
