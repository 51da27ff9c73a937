#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -rf --timeout 600 --maxfail 20 > gpurun_out/r2k1_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k1_tests.log
tail -3 gpurun_out/r2k1_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1
timeout -k 10 600 python scripts/stage_bench.py flow 2>&1 | tail -4 | tee gpurun_out/r2k1_stage.log
timeout -k 10 600 python scripts/first_audio.py 8 2>&1 | grep "^\[lib" | tee gpurun_out/r2k1_first_audio.log
KV32=1 timeout -k 10 600 python scripts/prof_llm_batch.py 32 4 288 2>&1 | grep -E "Self CUDA time total"
KV32=1 timeout -k 10 600 python scripts/time_llm_batch.py 32 4 288 2>&1 | tail -1 | tee gpurun_out/r2k1_llm_time.log
