#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_llm_gpu.py tests/test_e2e_gpu.py tests/test_c2_gpu.py -m gpu -q -rf --timeout 600 > gpurun_out/r2g_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_tests.log
tail -4 gpurun_out/r2g_tests.log
KV32=1 timeout -k 10 600 python scripts/prof_llm_batch.py 32 4 288 > gpurun_out/r2g_llm_prof_kv32.log 2>&1
grep -E "B=|hvx::|Self CUDA time total" gpurun_out/r2g_llm_prof_kv32.log | cut -c1-76,150-250
