#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_hift_gpu.py tests/test_e2e_gpu.py tests/test_c2_gpu.py tests/test_flow_gpu.py tests/test_streaming_gpu.py -m gpu -q -rf --timeout 600 -s > gpurun_out/r2b_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_tests.log
tail -4 gpurun_out/r2b_tests.log
STEPS=3 timeout -k 10 300 python scripts/stage_bench.py flow > gpurun_out/r2b_stage.log 2>&1
tail -5 gpurun_out/r2b_stage.log
KV32=1 timeout -k 10 600 python scripts/prof_llm_batch.py 32 4 288 > gpurun_out/r2b_llm_prof_kv32.log 2>&1
tail -30 gpurun_out/r2b_llm_prof_kv32.log
timeout -k 10 600 python scripts/prof_llm_batch.py 32 4 288 > gpurun_out/r2b_llm_prof_bf16.log 2>&1
tail -24 gpurun_out/r2b_llm_prof_bf16.log
