#!/bin/bash
# round 2, GPU call 3: multi-engine queue, all hooks; latency statistics
set -x
mkdir -p gpurun_out/c3
timeout 900 python -m pytest tests/test_gpu_batched_encoder.py -m gpu -x -q > gpurun_out/c3/pytest_batched.log 2>&1; echo "rc=$?" >> gpurun_out/c3/pytest_batched.log
tail -n 5 gpurun_out/c3/pytest_batched.log
for e in 1 8; do
timeout 900 python tools/encode_compare.py 1920x1080 17 --threads 96 --env HVB_ENGINES=$e --no-asm0 --no-asm1 --opts "--speed medium --no-sao --concurrent-frames 16" > gpurun_out/c3/enc_1080p_e$e.jsonl 2> gpurun_out/c3/enc_1080p_e$e.err
done
timeout 900 python tools/encode_compare.py 1920x1080 17 --threads 48,160 --env HVB_ENGINES=16 --opts "--speed medium --no-sao --concurrent-frames 16" > gpurun_out/c3/enc_1080p.jsonl 2> gpurun_out/c3/enc_1080p.err
timeout 1500 python tools/encode_compare.py 3840x2160 17 --threads 128 --env HVB_ENGINES=16 --opts "--speed medium --no-sao --concurrent-frames 16" > gpurun_out/c3/enc_4k.jsonl 2> gpurun_out/c3/enc_4k.err
timeout 600 python tools/encode_compare.py 3840x2160 17 --threads 128 --no-asm0 --no-asm1 --env HVB_ENGINES=16,HVB_HOOKS=7 --opts "--speed medium --no-sao --concurrent-frames 16" > gpurun_out/c3/enc_4k_mask7.jsonl 2> gpurun_out/c3/enc_4k_mask7.err
tail -n 3 gpurun_out/c3/*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c3/*.jsonl')):
    for l in open(f):
        d=json.loads(l); print(f.split('/')[-1], d['run'], round(d['fps'],3), d.get('identical_to_asm0'), json.dumps(d.get('queue'))[:900])
PY
