#!/bin/bash
# round 2, GPU call 7: full GPU suite (TMA parity included), TMA streaming roofline + ncu, size-threshold matrix of the
# batched encoder (which blocks are worth a hand-over), default bench run of both arms
set -x
mkdir -p gpurun_out/c7
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c7/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/c7/pytest_gpu.log
tail -n 6 gpurun_out/c7/pytest_gpu.log
timeout 300 python tools/stream_metrics.py --kinds sad,sad4 --tma --json gpurun_out/c7/stream_tma.json > gpurun_out/c7/stream_tma.log 2>&1
tail -n 34 gpurun_out/c7/stream_tma.log | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'sadTmaKernel' -c 8 -o gpurun_out/c7/tma_kernels \
    python tools/stream_metrics.py --block 64,32 --bps 1,2 --layouts unaligned --kinds sad,sad4 --tma --reps 1 > gpurun_out/c7/ncu_tma.log 2>&1
ncu -i gpurun_out/c7/tma_kernels.ncu-rep --page raw --csv > gpurun_out/c7/tma_kernels_raw.csv 2>/dev/null
nproc > gpurun_out/c7/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/c7/host.txt

O="--speed medium --no-sao --concurrent-frames 16"
run() { # tag env threads [extra encode_compare flags]
  timeout 300 python tools/encode_compare.py 3840x2160 17 --threads $3 --no-asm0 --no-asm1 --env $2 --opts "$O" $4 > gpurun_out/c7/$1.jsonl 2> gpurun_out/c7/$1.err
}
timeout 300 python tools/encode_compare.py 3840x2160 17 --threads 48 --env HVB_ENGINES=8 --opts "$O" > gpurun_out/c7/base.jsonl 2> gpurun_out/c7/base.err
run off HVB_BATCHED=0 16
run me1024 HVB_ENGINES=8,HVB_ME_MIN_AREA=1024,HVB_PU_MIN_AREA=1024,HVB_INTRA_MIN_LOG2=5,HVB_TU_MIN_LOG2=5 32
run me4096 HVB_ENGINES=8,HVB_ME_MIN_AREA=4096,HVB_PU_MIN_AREA=4096,HVB_INTRA_MIN_LOG2=6,HVB_TU_MIN_LOG2=6 32
run me1024_only HVB_ENGINES=8,HVB_ME_MIN_AREA=1024,HVB_PU_MIN_AREA=100000,HVB_INTRA_MIN_LOG2=6,HVB_TU_MIN_LOG2=7 32
run me256_only HVB_ENGINES=8,HVB_ME_MIN_AREA=256,HVB_PU_MIN_AREA=100000,HVB_INTRA_MIN_LOG2=6,HVB_TU_MIN_LOG2=7 32
run tu5_only HVB_ENGINES=8,HVB_ME_MIN_AREA=100000,HVB_PU_MIN_AREA=100000,HVB_INTRA_MIN_LOG2=6,HVB_TU_MIN_LOG2=5 32
run tu4_only HVB_ENGINES=8,HVB_ME_MIN_AREA=100000,HVB_PU_MIN_AREA=100000,HVB_INTRA_MIN_LOG2=6,HVB_TU_MIN_LOG2=4 32
run intra5_only HVB_ENGINES=8,HVB_ME_MIN_AREA=100000,HVB_PU_MIN_AREA=100000,HVB_INTRA_MIN_LOG2=5,HVB_TU_MIN_LOG2=7 32
run me1024_tu5 HVB_ENGINES=8,HVB_ME_MIN_AREA=1024,HVB_PU_MIN_AREA=100000,HVB_INTRA_MIN_LOG2=6,HVB_TU_MIN_LOG2=5 32
run me1024_tu5_t64 HVB_ENGINES=8,HVB_ME_MIN_AREA=1024,HVB_PU_MIN_AREA=100000,HVB_INTRA_MIN_LOG2=6,HVB_TU_MIN_LOG2=5 64
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c7/*.jsonl')):
    for l in open(f):
        d=json.loads(l); q=d.get('queue') or {}
        print(f.split('/')[-1], d.get('run'), 'fps',round(d['fps'],3), 'enc_wall',d.get('encoder_wall_s'),'user',d.get('host_user_s'),'sys',d.get('host_system_s'),'ident',d.get('identical_to_asm0'),'md5',d['bitstream_md5'][:8],
              {k:(q[k]['requests'],round(q[k]['mean_wait_us'])) for k in ('me','me_bi','pu_cost','intra_sweep','tu_chain') if k in q})
PY
( time timeout 1200 python bench.py > gpurun_out/c7/bench.json 2> gpurun_out/c7/bench.err ) 2> gpurun_out/c7/bench.time
tail -n 4 gpurun_out/c7/bench.err gpurun_out/c7/bench.time; head -c 700 gpurun_out/c7/bench.json
( time timeout 600 python bench.py --impl reference > gpurun_out/c7/bench_ref.json 2> gpurun_out/c7/bench_ref.err ) 2> gpurun_out/c7/bench_ref.time
tail -n 3 gpurun_out/c7/bench_ref.time; head -c 400 gpurun_out/c7/bench_ref.json
