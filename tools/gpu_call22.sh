#!/bin/bash
# round 2, GPU call 22: the driver's round-end sequence on the current tree: full GPU suite, smoke, bench (both arms), launch list of the bench
set -x
mkdir -p gpurun_out/c22
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c22/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c22/pytest_gpu.log
tail -n 6 gpurun_out/c22/pytest_gpu.log | cut -c1-250
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c22/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c22/smoke.log; tail -n 4 gpurun_out/c22/smoke.log
( time timeout 1500 python bench.py > gpurun_out/c22/bench.json 2> gpurun_out/c22/bench.err ) 2> gpurun_out/c22/bench.time
tail -n 4 gpurun_out/c22/bench.err gpurun_out/c22/bench.time; head -c 900 gpurun_out/c22/bench.json; echo
( time timeout 900 python bench.py --impl reference > gpurun_out/c22/bench_ref.json 2> gpurun_out/c22/bench_ref.err ) 2> gpurun_out/c22/bench_ref.time
tail -n 3 gpurun_out/c22/bench_ref.time; head -c 400 gpurun_out/c22/bench_ref.json; echo
