#!/bin/bash
# round 2, GPU call 20: service threads (a batch is answered and the next one issued by the thread that sees the completion flag) against a
# dispatcher thread per engine + completion thread
set -x
mkdir -p gpurun_out/c20
E=HVB_ENGINES=32,HVB_FIBERS=128
timeout 1500 python tools/segments_matrix.py gpurun_out/c20/matrix.jsonl \
  tu4_svc4:12:2:$E,HVB_HOOKS=48,HVB_INTRA_TU_MIN_LOG2=4,HVB_TU_MIN_LOG2=4,HVB_ENGINE_SHARES=1,1,1,1,1,27 \
  tu4_svc8:12:2:$E,HVB_HOOKS=48,HVB_INTRA_TU_MIN_LOG2=4,HVB_TU_MIN_LOG2=4,HVB_ENGINE_SHARES=1,1,1,1,1,27,HVB_SERVICE_THREADS=8 \
  tu4_svc2_e16:12:2:HVB_ENGINES=16,HVB_FIBERS=128,HVB_HOOKS=48,HVB_INTRA_TU_MIN_LOG2=4,HVB_TU_MIN_LOG2=4,HVB_ENGINE_SHARES=1,1,1,1,1,11,HVB_SERVICE_THREADS=2 \
  tu4_svc0:12:2:$E,HVB_HOOKS=48,HVB_INTRA_TU_MIN_LOG2=4,HVB_TU_MIN_LOG2=4,HVB_ENGINE_SHARES=1,1,1,1,1,27,HVB_SERVICE_THREADS=0 \
  all_svc4:12:2:$E \
  > gpurun_out/c20/matrix.log 2> gpurun_out/c20/matrix.err
cut -c1-200 gpurun_out/c20/matrix.log; tail -n 5 gpurun_out/c20/matrix.err
