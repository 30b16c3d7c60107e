#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "it/s\|^Batch\|^NER\|^Extract\|^Process\|Loading weights" | tail -30 > gpurun_out/r2l_pytest.log
ROWS=1250000,10000000 K=100 timeout 200 python tools/gpu_time_search.py 2>&1 | grep rows > gpurun_out/r2l_time.log
ROWS=1250000,10000000 K=10 timeout 200 python tools/gpu_time_search.py 2>&1 | grep rows >> gpurun_out/r2l_time.log
ROWS=1250000 K=32 timeout 200 python tools/gpu_time_search.py 2>&1 | grep rows >> gpurun_out/r2l_time.log
timeout 300 python tools/ivf_bench.py > gpurun_out/r2l_ivf_12m.json 2> gpurun_out/r2l_ivf_12m.err
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-encode > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
tail -8 gpurun_out/r2l_pytest.log; cat gpurun_out/r2l_time.log; cat gpurun_out/r2l_ivf_12m.json; cut -c1-300 gpurun_out/r2l_bench.json; tail -2 gpurun_out/r2l_bench.err
