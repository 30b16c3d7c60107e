#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "it/s\|^Batch\|^NER\|^Extract\|^Process\|Loading weights" | tail -60 > gpurun_out/r2e_pytest.log
ROWS=1250000,10000000 K=10 python tools/gpu_time_search.py 2>&1 | grep rows > gpurun_out/r2e_time.log
ROWS=1250000,10000000 K=100 python tools/gpu_time_search.py 2>&1 | grep rows >> gpurun_out/r2e_time.log
ROWS=1250000 K=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_topk_kernel -s 6 -c 1 -f -o gpurun_out/r2e_scan_1p25m python tools/gpu_time_search.py > gpurun_out/r2e_ncu.log 2>&1
timeout 400 python tools/ivf_bench.py > gpurun_out/r2e_ivf_12m.json 2> gpurun_out/r2e_ivf_12m.err
tail -25 gpurun_out/r2e_pytest.log; cat gpurun_out/r2e_time.log; cat gpurun_out/r2e_ivf_12m.json
