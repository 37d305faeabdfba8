#!/bin/bash
# round 2, GPU call 4: completion poller, merged uploads, segment-parallel driver
set -x
mkdir -p gpurun_out/c4
timeout 900 python -m pytest tests/test_gpu_batched_encoder.py -m gpu -x -q > gpurun_out/c4/pytest_batched.log 2>&1; echo "rc=$?" >> gpurun_out/c4/pytest_batched.log
tail -n 3 gpurun_out/c4/pytest_batched.log
O="--speed medium --no-sao --concurrent-frames 16"
timeout 900 python tools/encode_compare.py 3840x2160 17 --threads 64,128 --env HVB_ENGINES=8 --opts "$O" > gpurun_out/c4/enc_4k_e8.jsonl 2> gpurun_out/c4/enc_4k_e8.err
timeout 600 python tools/encode_compare.py 3840x2160 17 --threads 128 --no-asm0 --no-asm1 --env HVB_ENGINES=16 --opts "$O" > gpurun_out/c4/enc_4k_e16.jsonl 2> gpurun_out/c4/enc_4k_e16.err
timeout 600 python tools/encode_compare.py 3840x2160 17 --threads 128 --no-asm0 --no-asm1 --env HVB_ENGINES=16,HVB_HOOKS=7 --opts "$O" > gpurun_out/c4/enc_4k_e16_mask7.jsonl 2> gpurun_out/c4/enc_4k_e16_mask7.err
timeout 600 python tools/encode_compare.py 3840x2160 17 --threads 128 --no-asm0 --no-asm1 --env HVB_ENGINES=16,HVB_HOOKS=23 --opts "$O" > gpurun_out/c4/enc_4k_e16_mask23.jsonl 2> gpurun_out/c4/enc_4k_e16_mask23.err
S="--speed medium --no-sao --concurrent-frames 8 --segment 8"
timeout 1500 python tools/encode_compare.py 3840x2160 33 --threads 32,64 --segments 5 --env HVB_ENGINES=16 --opts "$S" > gpurun_out/c4/seg_4k.jsonl 2> gpurun_out/c4/seg_4k.err
tail -n 3 gpurun_out/c4/*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c4/*.jsonl')):
    for l in open(f):
        d=json.loads(l); print(f.split('/')[-1], d['run'], round(d['fps'],3), d.get('identical_to_asm0'), json.dumps(d.get('queue'))[:1000])
PY
