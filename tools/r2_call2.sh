#!/bin/bash
# round 2, GPU call 2 (one GPU): everything written since call 1
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/r2b_pytest.log
ROWS=1250000,10000000 K=10 python tools/gpu_time_search.py > gpurun_out/r2b_time_k10.log 2>&1
ROWS=1250000,10000000 K=100 python tools/gpu_time_search.py > gpurun_out/r2b_time_k100.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
timeout 400 python tools/ivf_bench.py > gpurun_out/r2b_ivf_12m.json 2> gpurun_out/r2b_ivf_12m.err
tail -30 gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_time_k10.log gpurun_out/r2b_time_k100.log | grep rows; cat gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err; cat gpurun_out/r2b_ivf_12m.json
