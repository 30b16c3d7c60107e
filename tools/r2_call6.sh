#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -v "it/s\|^Batch\|^NER\|^Extract\|^Process\|Loading weights" | tail -40 > gpurun_out/r2f_pytest.log
ROWS=1250000,10000000 K=10 python tools/gpu_time_search.py 2>&1 | grep rows > gpurun_out/r2f_time.log
ROWS=1250000,10000000 K=100 python tools/gpu_time_search.py 2>&1 | grep rows >> gpurun_out/r2f_time.log
ROWS=1250000 K=100 timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_topk_kernel -s 6 -c 1 -f -o gpurun_out/r2f_scan_1p25m_k100 python tools/gpu_time_search.py > gpurun_out/r2f_ncu.log 2>&1
timeout 400 python tools/ivf_bench.py > gpurun_out/r2f_ivf_12m.json 2> gpurun_out/r2f_ivf_12m.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
CRAG_GEMM_SMALL_M=0 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_nosmallm.json 2> gpurun_out/r2f_bench_nosmallm.err
tail -25 gpurun_out/r2f_pytest.log; cat gpurun_out/r2f_time.log; cat gpurun_out/r2f_ivf_12m.json; python - <<'PY'
import json
for f in ('gpurun_out/r2f_bench.json','gpurun_out/r2f_bench_nosmallm.json'):
    d=json.load(open(f)); e=d.pop('encode'); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['parity']['mismatches'])
    print({k:e[k] for k in ('value','ms_per_step')}, e['mixed_length']['chunks_per_s'], e['probe_batch'])
PY
