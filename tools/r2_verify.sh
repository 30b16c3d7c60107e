#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "it/s\|^Batch\|^NER\|^Extract\|^Process\|Loading weights" | tail -30 > gpurun_out/r2m_pytest.log
ROWS=1250000 K=100 timeout 100 python tools/gpu_time_search.py 2>&1 | grep rows > gpurun_out/r2m_time.log
ROWS=1250000,10000000 K=10 timeout 100 python tools/gpu_time_search.py 2>&1 | grep rows >> gpurun_out/r2m_time.log
tail -8 gpurun_out/r2m_pytest.log; cat gpurun_out/r2m_time.log
