#!/bin/bash
# First GPU call of the next round (one GPU): everything written after round 1's GPU budget ended gets measured here.
#   gpurun --timeout 900 -- 'bash tools/round2_first_call.sh'
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/r2_pytest.log
# 1. GEMM A/B variants (bit 2 = pipelined epilogue): correctness + TFLOP/s per shape, one subprocess per case
python tools/gpu_check_encoder.py --only gemm > gpurun_out/r2_gemm_cases.log 2>&1
# 2. whole bge-large forward with and without the variant
python tools/gpu_check_encoder.py --kind perf --case 0 > gpurun_out/r2_perf_default.log 2>&1
CRAG_GEMM_VARIANT=4 python tools/gpu_check_encoder.py --kind perf --case 0 > gpurun_out/r2_perf_variant4.log 2>&1
CRAG_GEMM_VARIANT=4 python tools/gpu_check_encoder.py --kind enc --case 4 > gpurun_out/r2_enc_variant4.log 2>&1
CRAG_GEMM_VARIANT=12 python tools/gpu_check_encoder.py --kind perf --case 0 > gpurun_out/r2_perf_variant12.log 2>&1
CRAG_GEMM_VARIANT=12 python tools/gpu_check_encoder.py --kind enc --case 4 > gpurun_out/r2_enc_variant12.log 2>&1
# 2b. attention SPLIT variant (8 softmax warps, two threads per query row): correctness + timing, then whole forward
CRAG_ATTN_VARIANT=1 python tools/gpu_check_encoder.py --only attn > gpurun_out/r2_attn_split_cases.log 2>&1
python tools/gpu_check_encoder.py --only attn > gpurun_out/r2_attn_default_cases.log 2>&1
CRAG_ATTN_VARIANT=1 python tools/gpu_check_encoder.py --kind perf --case 0 > gpurun_out/r2_perf_attn_split.log 2>&1
CRAG_ATTN_VARIANT=1 CRAG_GEMM_VARIANT=4 python tools/gpu_check_encoder.py --kind perf --case 0 > gpurun_out/r2_perf_both.log 2>&1
CRAG_ATTN_VARIANT=1 python tools/gpu_check_encoder.py --kind enc --case 4 > gpurun_out/r2_enc_attn_split.log 2>&1
# 3. IVF (config 4): one rank's share of 100M x 768, then a smaller one
timeout 600 python tools/ivf_bench.py > gpurun_out/r2_ivf_12m.json 2> gpurun_out/r2_ivf_12m.err
timeout 300 python tools/ivf_bench.py --rows 2000000 --nlist 1024 --nprobe 16 > gpurun_out/r2_ivf_2m.json 2> gpurun_out/r2_ivf_2m.err
# 4. config-5 cycle
timeout 300 python tools/loop_bench.py > gpurun_out/r2_loop.json 2> gpurun_out/r2_loop.err
tail -3 gpurun_out/r2_pytest.log; grep -hE '"variant": (4|8|12)' gpurun_out/r2_gemm_cases.log; grep -h '"tc": 1' gpurun_out/r2_attn_split_cases.log gpurun_out/r2_attn_default_cases.log; cat gpurun_out/r2_perf_*.log gpurun_out/r2_enc_*.log gpurun_out/r2_ivf_*.json gpurun_out/r2_loop.json
