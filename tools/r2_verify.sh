#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "it/s\|^Batch\|^NER\|^Extract\|^Process\|Loading weights" | tail -40 > gpurun_out/r2j_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2j_smoke.log 2>&1
ROWS=1250000,10000000 K=100 timeout 200 python tools/gpu_time_search.py 2>&1 | grep rows > gpurun_out/r2j_time.log
ROWS=1250000 K=10 timeout 200 python tools/gpu_time_search.py 2>&1 | grep rows >> gpurun_out/r2j_time.log
timeout 300 python tools/ivf_bench.py > gpurun_out/r2j_ivf_12m.json 2> gpurun_out/r2j_ivf_12m.err
tail -12 gpurun_out/r2j_pytest.log; tail -2 gpurun_out/r2j_smoke.log; cat gpurun_out/r2j_time.log; cat gpurun_out/r2j_ivf_12m.json; tail -3 gpurun_out/r2j_ivf_12m.err
