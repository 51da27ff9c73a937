#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python scripts/first_audio.py 10 2>&1 | grep "^\[lib" | tee gpurun_out/r2j1_first_audio.log
HVX_LIB_PATH=$PWD/flowmirror_hydravox_b200/libhydravox_b200_oldgemm.so timeout -k 10 600 python scripts/first_audio.py 10 2>&1 | grep "^\[lib" | tee -a gpurun_out/r2j1_first_audio.log
timeout -k 10 600 python scripts/first_audio.py 10 2>&1 | grep "^\[lib" | tee -a gpurun_out/r2j1_first_audio.log
timeout -k 10 600 python -m pytest tests/test_llm_gpu.py tests/test_e2e_gpu.py -m gpu -q --timeout 300 2>&1 | tail -2
KV32=1 timeout -k 10 600 python scripts/prof_llm_batch.py 32 4 288 2>&1 | grep -E "Self CUDA time total"
HVX_LLM_QKV_SPLITK=1 KV32=1 timeout -k 10 600 python scripts/prof_llm_batch.py 32 4 288 2>&1 | grep -E "Self CUDA time total"
