#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "it/s\|^Batch\|^NER\|^Extract\|^Process\|Loading weights" | tail -40 > gpurun_out/r2h_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2h_smoke.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
tail -25 gpurun_out/r2h_pytest.log; tail -2 gpurun_out/r2h_smoke.log; cut -c1-400 gpurun_out/r2h_bench.json; tail -3 gpurun_out/r2h_bench.err
