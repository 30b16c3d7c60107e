#!/bin/bash
# ncu evidence for the round (run on the GPU box through gpurun; one GPU).  Outputs land in gpurun_out/.
#   1. launch list of the bench command (our kernels only): per-launch device time, cold-cache + serialised
#   2. --set full capture of the dominant kernel (search scan) and of the encoder GEMMs / attention
set -u
mkdir -p gpurun_out
K='regex:search_topk_kernel|merge_topk_kernel|finalize_exchange_kernel|hist_kernel|scan_kernel|scatter_kernel|gemm2?_bf16_kernel|attention_(tc_)?kernel|layernorm_kernel|pool_normalize_kernel'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --encode-steps 1 --no-cpu-baseline \
    > gpurun_out/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:search_topk_kernel -s 4 -c 1 \
    -f -o gpurun_out/prof_search python bench.py --steps 1 --warmup 3 --no-encode --no-cpu-baseline \
    > gpurun_out/prof_search.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:gemm2?_bf16_kernel|attention_(tc_)?kernel' -s 28 -c 5 \
    -f -o gpurun_out/prof_encoder python bench.py --rows 300000 --steps 1 --warmup 3 --encode-steps 1 --no-cpu-baseline \
    > gpurun_out/prof_encoder.log 2>&1
ls -la gpurun_out/
