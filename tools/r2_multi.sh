#!/bin/bash
# round 2, multi-GPU call: usage  bash tools/r2_multi.sh N [full]
set -u
N=${1:-2}; MODE=${2:-dry}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python -m pytest tests/test_search_gpu.py tests/test_encoder_gpu.py -q -m gpu --tb=short 2>&1 | tail -15 > gpurun_out/m${N}_pytest.log
# 1. parity of the sharded search on every rank, both exchange formulations, headline shapes included
$TR --master-port 29511 tools/gpu_check_dist.py > gpurun_out/m${N}_check_peer.log 2>&1
CRAG_EXCHANGE=nccl $TR --master-port 29512 tools/gpu_check_dist.py > gpurun_out/m${N}_check_nccl.log 2>&1
cp gpurun_out/check_dist_w${N}.json gpurun_out/m${N}_check_nccl.json 2>/dev/null
# 2. bench at N (peer exchange, then the NCCL formulation, then without the CUDA graph)
$TR --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/m${N}_bench_peer.json 2> gpurun_out/m${N}_bench_peer.err
CRAG_EXCHANGE=nccl $TR --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 --no-encode > gpurun_out/m${N}_bench_nccl.json 2> gpurun_out/m${N}_bench_nccl.err
$TR --master-port 29515 bench.py --gpus $N --steps 20 --warmup 3 --no-encode --no-graph > gpurun_out/m${N}_bench_nograph.json 2> gpurun_out/m${N}_bench_nograph.err
if [ "$MODE" = "full" ]; then ROWS=12500000; LROWS=10000000; else ROWS=2000000; LROWS=2000000; fi
# 3. config 4: IVF-4096 over N x ROWS x 768
$TR --master-port 29516 tools/ivf_bench.py --rows $ROWS > gpurun_out/m${N}_ivf.json 2> gpurun_out/m${N}_ivf.err
# 4. config 5: probe -> retrieve -> rerank loop
$TR --master-port 29517 tools/loop_bench.py --rows $LROWS > gpurun_out/m${N}_loop.json 2> gpurun_out/m${N}_loop.err
tail -5 gpurun_out/m${N}_pytest.log; grep RESULT gpurun_out/m${N}_check_peer.log gpurun_out/m${N}_check_nccl.log | cut -c1-1500
for f in peer nccl nograph; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/m${N}_bench_$f.json')); e=d.pop('encode',None)
    print('$f', d['n_gpus'], round(d['value']), round(d['ms_per_step']*1000,1),'us/step e2e', round(d['e2e']['value']), 'scan', round(d['roofline']['kernel_ms']*1000,1), 'parity', d['parity']['mismatches'], d['config']['exchange'])
    if e: print('   encode', round(e['value']), e['probe_batch']['ms'])
except Exception as ex: print('$f failed', ex)
PY
done
tail -2 gpurun_out/m${N}_bench_peer.err; cat gpurun_out/m${N}_ivf.json; tail -2 gpurun_out/m${N}_ivf.err; cat gpurun_out/m${N}_loop.json; tail -2 gpurun_out/m${N}_loop.err
