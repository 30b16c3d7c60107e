#!/bin/bash
# round 2, final one-GPU call: the full GPU suite, smoke(), the bench line, and the ncu captures profiles/ cites
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "it/s\|^Batch\|^NER\|^Extract\|^Process\|Loading weights" | tail -30 > gpurun_out/r2g_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2g_smoke.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 1 --encode-steps 1 --no-cpu-baseline > gpurun_out/r2g_launches_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:search_topk_kernel -s 4 -c 1 -f -o gpurun_out/r2g_scan_10m python bench.py --steps 2 --warmup 1 --no-encode --no-cpu-baseline > gpurun_out/r2g_ncu_scan.log 2>&1
tail -12 gpurun_out/r2g_pytest.log; tail -2 gpurun_out/r2g_smoke.log; cut -c1-900 gpurun_out/r2g_bench.json; tail -2 gpurun_out/r2g_bench.err; wc -l gpurun_out/r2g_launches.csv; ls -la gpurun_out/r2g_scan_10m.ncu-rep
