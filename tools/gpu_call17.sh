#!/bin/bash
# round 2, GPU call 17: where a hand-over's time goes (queue / batch / wake-up phases), after the per-pool nudge counters
set -x
mkdir -p gpurun_out/c17
E=HVB_ENGINES=32,HVB_FIBERS=128
timeout 1500 python tools/segments_matrix.py gpurun_out/c17/matrix.jsonl \
  itu4:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4 \
  itu4_t3:12:3:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4 \
  itu4_spin0:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4,HVB_FIBER_SPIN_US=0 \
  itu4_spin200:12:2:$E,HVB_HOOKS=32,HVB_INTRA_TU_MIN_LOG2=4,HVB_FIBER_SPIN_US=200 \
  tu4:12:2:$E,HVB_HOOKS=48,HVB_INTRA_TU_MIN_LOG2=4,HVB_TU_MIN_LOG2=4 \
  > gpurun_out/c17/matrix.log 2> gpurun_out/c17/matrix.err
cut -c1-200 gpurun_out/c17/matrix.log; tail -n 5 gpurun_out/c17/matrix.err
