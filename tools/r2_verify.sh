#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_search_gpu.py tests/test_ivf_gpu.py -q -m gpu --tb=short 2>&1 | tail -12 > gpurun_out/r2k_pytest.log
ROWS=1250000,10000000 K=100 timeout 200 python tools/gpu_time_search.py 2>&1 | grep rows > gpurun_out/r2k_time.log
ROWS=1250000 K=10 timeout 200 python tools/gpu_time_search.py 2>&1 | grep rows >> gpurun_out/r2k_time.log
ROWS=1250000 K=32 timeout 200 python tools/gpu_time_search.py 2>&1 | grep rows >> gpurun_out/r2k_time.log
timeout 300 python tools/ivf_bench.py > gpurun_out/r2k_ivf_12m.json 2> gpurun_out/r2k_ivf_12m.err
tail -8 gpurun_out/r2k_pytest.log; cat gpurun_out/r2k_time.log; cat gpurun_out/r2k_ivf_12m.json; tail -3 gpurun_out/r2k_ivf_12m.err
