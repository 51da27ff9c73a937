#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python scripts/gemm_timeline.py > gpurun_out/r2e_gemm_timeline.log 2>&1
grep -E "CUPTI|total" gpurun_out/r2e_gemm_timeline.log | cut -c1-160 | tail -24
timeout -k 10 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -rf --timeout 300 -s -k "three_term" > gpurun_out/r2e_tests_gemm3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_tests_gemm3.log
tail -25 gpurun_out/r2e_tests_gemm3.log | cut -c1-200
timeout -k 10 600 python -m pytest tests/test_flow_gpu.py tests/test_c2_gpu.py tests/test_unet_gpu.py -m gpu -q -rf --timeout 300 -s > gpurun_out/r2e_tests_flow.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_tests_flow.log
tail -6 gpurun_out/r2e_tests_flow.log | cut -c1-200
STEPS=3 timeout -k 10 300 python scripts/stage_bench.py flow > gpurun_out/r2e_stage.log 2>&1
grep parity gpurun_out/r2e_stage.log
HVX_NO_PAIR=1 STEPS=3 timeout -k 10 300 python scripts/stage_bench.py flow > gpurun_out/r2e_stage_nopair.log 2>&1
grep parity gpurun_out/r2e_stage_nopair.log
