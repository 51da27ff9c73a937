#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_llm_gpu.py tests/test_e2e_gpu.py tests/test_c2_gpu.py tests/test_streaming_gpu.py -m gpu -q --timeout 300 2>&1 | tail -3 | tee gpurun_out/r2i1_tests.log
KV32=1 timeout -k 10 600 python scripts/prof_llm_batch.py 32 4 288 > gpurun_out/r2i1_llm_prof_kv32.log 2>&1
grep -E "B=|hvx::|Self CUDA time total" gpurun_out/r2i1_llm_prof_kv32.log | cut -c1-76,150-250 | head -14
HVX_LLM_QKV_SPLITK=1 HVX_LLM_FUSE_NORM=0 KV32=1 timeout -k 10 600 python scripts/prof_llm_batch.py 32 4 288 2>&1 | grep -E "Self CUDA time total"
