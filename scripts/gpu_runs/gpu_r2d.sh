#!/bin/bash
mkdir -p gpurun_out
KV32=1 timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 600 -c 6 -o gpurun_out/r2d_llm_gemm -f python scripts/prof_llm_batch.py 32 4 32 > gpurun_out/r2d_ncu.log 2>&1
tail -5 gpurun_out/r2d_ncu.log
ls -la gpurun_out/r2d_llm_gemm.ncu-rep
