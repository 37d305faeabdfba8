#!/bin/bash
# round 2, GPU call 25: after the set-up diet (one page-locked arena per engine, buffers by kind, one device allocation for the picture pool):
# encoder parity tests, the bench's configuration through the matrix tool (set-up times), the bench line
set -x
mkdir -p gpurun_out/c25
timeout 600 python -m pytest tests/test_gpu_batched_encoder.py tests/test_gpu_tu.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/c25/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c25/pytest_gpu.log
tail -n 4 gpurun_out/c25/pytest_gpu.log | cut -c1-250
timeout 400 python tools/segments_matrix.py gpurun_out/c25/matrix.jsonl \
  bench:12:3:HVB_ENGINES=20,HVB_ENGINE_SHARES=1,1,1,1,1,15,HVB_FIBERS=128,HVB_HOOKS=48,HVB_INTRA_TU_MIN_LOG2=5,HVB_TU_MIN_LOG2=6 > gpurun_out/c25/matrix.log 2> gpurun_out/c25/matrix.err
cut -c1-400 gpurun_out/c25/matrix.log; tail -n 3 gpurun_out/c25/matrix.err
( time timeout 900 python bench.py --no-stream --no-pass > gpurun_out/c25/bench.json 2> gpurun_out/c25/bench.err ) 2> gpurun_out/c25/bench.time
tail -n 4 gpurun_out/c25/bench.err gpurun_out/c25/bench.time; head -c 300 gpurun_out/c25/bench.json; echo
