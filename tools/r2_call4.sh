#!/bin/bash
# round 2, GPU call 4 (one GPU): tests, timings, ncu of the short scan (1.25M rows = one rank's shard at N=8)
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "it/s\|^Batch\|^NER\|^Extract\|^Process\|Loading weights" | tail -60 > gpurun_out/r2d_pytest.log
ROWS=1250000,10000000 K=10 python tools/gpu_time_search.py 2>&1 | grep rows > gpurun_out/r2d_time.log
ROWS=1250000,10000000 K=100 python tools/gpu_time_search.py 2>&1 | grep rows >> gpurun_out/r2d_time.log
ROWS=1250000 K=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_topk_kernel -s 6 -c 1 -f -o gpurun_out/r2d_scan_1p25m python tools/gpu_time_search.py > gpurun_out/r2d_ncu.log 2>&1
ROWS=1250000 K=100 timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_topk_kernel -s 6 -c 1 -f -o gpurun_out/r2d_scan_1p25m_k100 python tools/gpu_time_search.py >> gpurun_out/r2d_ncu.log 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
tail -25 gpurun_out/r2d_pytest.log; cat gpurun_out/r2d_time.log; tail -3 gpurun_out/r2d_ncu.log; ls -la gpurun_out/*.ncu-rep; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench.json'))
e=d.pop('encode'); print(json.dumps(d)[:1200]); print({k:e[k] for k in ('value','ms_per_step','mixed_length','probe_batch')})
PY
