#!/bin/bash
# round 2, multi-GPU call: usage  STEPS="check bench nccl ivf loop" bash tools/r2_multi.sh N [full]
# every command carries its own timeout: a hang must not eat the GPU budget
set -u
N=${1:-2}; MODE=${2:-dry}; STEPS=${STEPS:-"check bench ivf loop"}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$MODE" = "full" ]; then ROWS=12500000; LROWS=10000000; else ROWS=1000000; LROWS=1000000; fi
for s in $STEPS; do case $s in
  ivf)   # config 4: IVF-4096 over N x ROWS x 768
    timeout 170 $TR --master-port 29516 tools/ivf_bench.py --rows $ROWS > gpurun_out/m${N}_ivf.json 2> gpurun_out/m${N}_ivf.err
    cat gpurun_out/m${N}_ivf.json; tail -2 gpurun_out/m${N}_ivf.err | cut -c1-300;;
  loop)  # config 5: probe -> retrieve -> rerank loop
    timeout 170 $TR --master-port 29517 tools/loop_bench.py --rows $LROWS > gpurun_out/m${N}_loop.json 2> gpurun_out/m${N}_loop.err
    cat gpurun_out/m${N}_loop.json; tail -2 gpurun_out/m${N}_loop.err | cut -c1-300;;
  check) # parity of the sharded search on every rank (peer exchange), headline shapes included
    timeout 120 $TR --master-port 29511 tools/gpu_check_dist.py > gpurun_out/m${N}_check_peer.log 2>&1
    cp gpurun_out/check_dist_w${N}.json gpurun_out/m${N}_check_peer.json 2>/dev/null
    grep RESULT gpurun_out/m${N}_check_peer.log | cut -c1-1500;;
  bench|nccl)
    if [ $s = nccl ]; then export CRAG_EXCHANGE=nccl; X="--no-encode"; else unset CRAG_EXCHANGE; X=""; fi
    timeout 170 $TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 $X > gpurun_out/m${N}_bench_$s.json 2> gpurun_out/m${N}_bench_$s.err
    unset CRAG_EXCHANGE
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/m${N}_bench_$s.json')); e=d.pop('encode',None)
    print('$s', d['n_gpus'], round(d['value']), round(d['ms_per_step']*1000,1),'us/step e2e', round(d['e2e']['value']), 'scan', round(d['roofline']['kernel_ms']*1000,1), 'parity', d['parity']['mismatches'], d['config']['exchange'])
    if e: print('   encode', round(e['value']), e['probe_batch']['ms'])
except Exception as ex: print('$s failed', ex)
PY
    tail -2 gpurun_out/m${N}_bench_$s.err | cut -c1-300;;
esac; done
