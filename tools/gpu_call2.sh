#!/bin/bash
# round 2, GPU call 2: first runs of the batched encoder on a B200
set -x
mkdir -p gpurun_out/c2
timeout 900 python -m pytest tests/test_gpu_batched_encoder.py -m gpu -x -q > gpurun_out/c2/pytest_batched.log 2>&1; echo "rc=$?" >> gpurun_out/c2/pytest_batched.log
tail -n 5 gpurun_out/c2/pytest_batched.log
timeout 900 python tools/encode_compare.py 1920x1080 17 --threads 32,96 > gpurun_out/c2/enc_1080p.jsonl 2> gpurun_out/c2/enc_1080p.err
cut -c1-400 gpurun_out/c2/enc_1080p.jsonl
timeout 1500 python tools/encode_compare.py 3840x2160 9 --threads 64,160 > gpurun_out/c2/enc_4k.jsonl 2> gpurun_out/c2/enc_4k.err
cut -c1-400 gpurun_out/c2/enc_4k.jsonl
tail -n 5 gpurun_out/c2/*.err
