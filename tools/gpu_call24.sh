#!/bin/bash
# round 2, GPU call 24: the driver's round-end sequence on the tree as committed: full GPU suite, smoke, both bench arms; TMA form of SAD / SAD4 measured once more
set -x
mkdir -p gpurun_out/c24
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c24/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c24/pytest_gpu.log
tail -n 5 gpurun_out/c24/pytest_gpu.log | cut -c1-250
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c24/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/c24/smoke.log; tail -n 3 gpurun_out/c24/smoke.log
( time timeout 900 python bench.py --impl reference > gpurun_out/c24/bench_ref.json 2> gpurun_out/c24/bench_ref.err ) 2> gpurun_out/c24/bench_ref.time
tail -n 3 gpurun_out/c24/bench_ref.time; head -c 300 gpurun_out/c24/bench_ref.json; echo
( time timeout 1500 python bench.py > gpurun_out/c24/bench.json 2> gpurun_out/c24/bench.err ) 2> gpurun_out/c24/bench.time
tail -n 4 gpurun_out/c24/bench.err gpurun_out/c24/bench.time; head -c 300 gpurun_out/c24/bench.json; echo
timeout 300 python tools/stream_metrics.py --kinds sad,sad4 --tma --json gpurun_out/c24/stream_tma.json > gpurun_out/c24/stream_tma.log 2>&1
cut -c1-170 gpurun_out/c24/stream_tma.log | tail -n 18
