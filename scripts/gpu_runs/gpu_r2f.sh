#!/bin/bash
mkdir -p gpurun_out
HVX_LLM_NO_GRAPH=1 HVX_GEMM_TIMELINE=1 KV32=1 timeout -k 10 300 python scripts/time_llm_batch.py 32 4 4 > gpurun_out/r2f_incontext_timeline.log 2>&1
grep -c "gemm timeline" gpurun_out/r2f_incontext_timeline.log
grep "gemm timeline" gpurun_out/r2f_incontext_timeline.log | awk '{print $3,$4,$5,$8,$9,$13,$14}' | sort | uniq -c | sort -rn | head -30
timeout -k 10 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -rf --timeout 300 -s -k "three_term" > gpurun_out/r2f_tests_gemm3.log 2>&1
tail -3 gpurun_out/r2f_tests_gemm3.log
