#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python scripts/prof_step.py c2 > gpurun_out/r2t_kernels_c2.txt 2> gpurun_out/r2t_c2.err; tail -2 gpurun_out/r2t_c2.err; head -14 gpurun_out/r2t_kernels_c2.txt | cut -c1-150
timeout -k 10 900 python scripts/prof_step.py c3 > gpurun_out/r2t_kernels_c3.txt 2> gpurun_out/r2t_c3.err; tail -2 gpurun_out/r2t_c3.err; head -14 gpurun_out/r2t_kernels_c3.txt | cut -c1-150
timeout -k 10 600 python -m pytest tests/test_llm_gpu.py tests/test_e2e_gpu.py tests/test_c2_gpu.py -m gpu -q --timeout 300 2>&1 | tail -3 | tee gpurun_out/r2t_tests.log
KV32=1 timeout -k 10 600 python scripts/prof_llm_batch.py 32 4 288 > gpurun_out/r2t_llm_prof_kv32.log 2>&1
grep -E "B=|hvx::|Self CUDA time total" gpurun_out/r2t_llm_prof_kv32.log | cut -c1-76,150-250 | head -8
./scripts/ubench/ex2 | tee gpurun_out/r2t_ex2.log
python __graft_entry__.py smoke 2>&1 | tail -1
