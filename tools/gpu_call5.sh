#!/bin/bash
# round 2, GPU call 5: tuning matrix of the queue + first run of the new bench.py
set -x
mkdir -p gpurun_out/c5
O="--speed medium --no-sao --concurrent-frames 16"
run() { # tag env threads
  timeout 600 python tools/encode_compare.py 3840x2160 17 --threads $3 --no-asm0 --no-asm1 --env $2 --opts "$O" > gpurun_out/c5/$1.jsonl 2> gpurun_out/c5/$1.err
}
run e4_t32 HVB_ENGINES=4 32
run e4_t64 HVB_ENGINES=4 64
run e8_t32 HVB_ENGINES=8 32
run e8_t48 HVB_ENGINES=8 48
run e16_t48 HVB_ENGINES=16 48
run e8_t48_spin HVB_ENGINES=8,HVB_POLLER=0 48
run e8_t48_intra16 HVB_ENGINES=8,HVB_INTRA_MIN_LOG2=4 48
run e8_t48_mask7 HVB_ENGINES=8,HVB_HOOKS=7 48
run e2_t32 HVB_ENGINES=2 32
timeout 1200 python bench.py --steps 4 --warmup 1 > gpurun_out/c5/bench.json 2> gpurun_out/c5/bench.err
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/c5/bench_ref.json 2> gpurun_out/c5/bench_ref.err
tail -n 5 gpurun_out/c5/bench.err gpurun_out/c5/bench_ref.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c5/e*.jsonl')):
    for l in open(f):
        d=json.loads(l); q=d.get('queue') or {}
        print(f.split('/')[-1], 'fps',round(d['fps'],3), 'busy',q.get('engine_busy_s'),'disp',q.get('dispatches'), {k:(q[k]['requests'],round(q[k]['mean_wait_us'])) for k in ('me','me_bi','pu_cost','intra_sweep','tu_chain') if k in q})
for f in ('gpurun_out/c5/bench.json','gpurun_out/c5/bench_ref.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, {k:d.get(k) for k in ('value','e2e','bitstream_md5_equals_asm0','cpu_baseline','ms_per_step','gpu_launches','clocks')}); print(json.dumps(d.get('roofline'))[:1500])
    except Exception as e: print(f, 'ERR', e)
PY
