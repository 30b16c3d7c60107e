#!/bin/bash
# round 2, GPU call 3 (one GPU): full GPU test suite, A/B of the tile permutation / pooled floor, IVF, bench
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "it/s\|^Batch\|^NER\|^Extract\|^Process\|Loading weights" | tail -120 > gpurun_out/r2c_pytest.log
ab() { # name, shift, pool, K
  CRAG_SEARCH_PERM_SHIFT=$2 CRAG_SEARCH_POOL=$3 ROWS=1250000,10000000 K=$4 python tools/gpu_time_search.py 2>&1 | grep rows | sed "s/^/$1 /" >> gpurun_out/r2c_ab.log
}
: > gpurun_out/r2c_ab.log
ab "k10 shift3 pool1" 3 1 10
ab "k10 natural pool1" -1 1 10
ab "k10 shift0 pool1" 0 1 10
ab "k10 shift3 pool0" 3 0 10
ab "k10 natural pool0" -1 0 10
ab "k100 shift3 pool1" 3 1 100
ab "k100 natural pool1" -1 1 100
ab "k100 shift3 pool0" 3 0 100
timeout 400 python tools/ivf_bench.py > gpurun_out/r2c_ivf_12m.json 2> gpurun_out/r2c_ivf_12m.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -40 gpurun_out/r2c_pytest.log; cat gpurun_out/r2c_ab.log; cat gpurun_out/r2c_ivf_12m.json; tail -3 gpurun_out/r2c_ivf_12m.err; cat gpurun_out/r2c_bench.json | cut -c1-1500
