#!/bin/bash
# round 2, GPU call 12: the segment driver with fiber pools: instances in flight x size thresholds x engines (by kind), on the bench's job
set -x
mkdir -p gpurun_out/c12
BIG=HVB_ME_MIN_AREA=1024,HVB_PU_MIN_AREA=1024,HVB_INTRA_MIN_LOG2=5,HVB_TU_MIN_LOG2=5
MID=HVB_ME_MIN_AREA=256,HVB_PU_MIN_AREA=1024,HVB_INTRA_MIN_LOG2=4,HVB_TU_MIN_LOG2=4
E=HVB_ENGINES=32,HVB_FIBERS=128
timeout 1500 python tools/segments_matrix.py gpurun_out/c12/matrix.jsonl \
  off_p4:4:4:HVB_BATCHED=0 off_p8:8:2:HVB_BATCHED=0 off_p12:12:2:HVB_BATCHED=0 \
  all_p8_e32:8:2:$E mid_p8_e32:8:2:$E,$MID big_p8_e32:8:2:$E,$BIG \
  mid_p12_e32:12:2:$E,$MID big_p12_e32:12:2:$E,$BIG big_p12_e8:12:2:HVB_ENGINES=8,HVB_FIBERS=128,$BIG \
  mid_p12_e32_t1:12:1:$E,$MID all_p12_e32:12:2:$E \
  > gpurun_out/c12/matrix.log 2> gpurun_out/c12/matrix.err
cut -c1-420 gpurun_out/c12/matrix.log; tail -n 5 gpurun_out/c12/matrix.err
free -g | head -2
