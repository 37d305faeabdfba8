#!/bin/bash
# round 2, GPU call 11: GPU suite at HEAD; streaming roofline after the turn-structured SAD/SSD loops and the two-accumulator SATD;
# the batched encoder with its pool threads as fiber schedulers (integration/fiber_pool.cpp) against a thread per row
set -x
mkdir -p gpurun_out/c11
nproc > gpurun_out/c11/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/c11/host.txt; free -g >> gpurun_out/c11/host.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c11/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c11/pytest_gpu.log
tail -n 8 gpurun_out/c11/pytest_gpu.log | cut -c1-250
timeout 600 python tools/stream_metrics.py --block 64,32 --json gpurun_out/c11/stream.json > gpurun_out/c11/stream.log 2>&1
cut -c1-200 gpurun_out/c11/stream.log | tail -n 34

O="--speed medium --no-sao --concurrent-frames 16"
run() { # tag env threads
  timeout 300 python tools/encode_compare.py 3840x2160 17 --threads $3 --no-asm0 --no-asm1 --env $2 --opts "$O" > gpurun_out/c11/$1.jsonl 2> gpurun_out/c11/$1.err
}
timeout 400 python tools/encode_compare.py 3840x2160 17 --threads 16 --env HVB_ENGINES=8 --opts "$O" > gpurun_out/c11/fib48_t16.jsonl 2> gpurun_out/c11/fib48_t16.err
run thr_t48 HVB_ENGINES=8,HVB_FIBERS=0 48
run fib48_t8 HVB_ENGINES=8 8
run fib96_t16_e4 HVB_ENGINES=4,HVB_FIBERS=96 16
run fib96_t16_e16 HVB_ENGINES=16,HVB_FIBERS=96 16
run fib48_t12_mask7 HVB_ENGINES=8,HVB_HOOKS=7 12
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c11/*.jsonl')):
    for l in open(f):
        d=json.loads(l); q=d.get('queue') or {}
        print(f.split('/')[-1], d.get('run'), 'fps',round(d['fps'],3), 'enc_wall',d.get('encoder_wall_s'),'user',d.get('host_user_s'),'sys',d.get('host_system_s'),'ident',d.get('identical_to_asm0'),'md5',d['bitstream_md5'][:8],
              'disp',q.get('dispatches'),'busy',q.get('engine_busy_s'),
              {k:(q[k]['requests'],q[k]['batches'],round(q[k]['mean_wait_us'])) for k in ('me','me_bi','pu_cost','intra_sweep','tu_chain') if k in q})
PY
tail -n 5 gpurun_out/c11/*.err | cut -c1-300
timeout 300 python tools/encode_compare.py 3840x2160 32 --threads 4 --segments 4 --no-asm0 --no-asm1 --opts "--speed medium --no-sao --concurrent-frames 8 --segment 8" --env HVB_ENGINES=8 > gpurun_out/c11/seg4_t4.jsonl 2> gpurun_out/c11/seg4_t4.err
cat gpurun_out/c11/seg4_t4.jsonl | cut -c1-1500; tail -n 3 gpurun_out/c11/seg4_t4.err | cut -c1-300
free -g | head -2
