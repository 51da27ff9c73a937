#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_llm_gpu.py tests/test_e2e_gpu.py tests/test_c2_gpu.py -m gpu -q -rf --timeout 600 > gpurun_out/r2h_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_tests.log
tail -6 gpurun_out/r2h_tests.log
KV32=1 timeout -k 10 600 python scripts/prof_llm_batch.py 32 4 288 > gpurun_out/r2h_llm_prof_kv32.log 2>&1
grep -E "B=|hvx::|Self CUDA time total" gpurun_out/r2h_llm_prof_kv32.log | cut -c1-76,150-250 | head -14
timeout -k 10 600 python scripts/prof_llm_batch.py 32 4 288 > gpurun_out/r2h_llm_prof_bf16.log 2>&1
grep -E "B=|llm_attn|Self CUDA time total" gpurun_out/r2h_llm_prof_bf16.log | cut -c1-76,150-250 | head -5
HVX_ATTN_DECODE=seq KV32=1 timeout -k 10 600 python scripts/time_llm_batch.py 32 4 64 2>&1 | tail -1
KV32=1 timeout -k 10 600 python scripts/time_llm_batch.py 32 4 64 2>&1 | tail -1
