#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gemm_gpu.py tests/test_flow_gpu.py tests/test_unet_gpu.py tests/test_unet_nc_gpu.py tests/test_c2_gpu.py -m gpu -q --timeout 300 2>&1 | tail -3 | tee gpurun_out/r2w_tests.log
timeout -k 10 600 python scripts/stage_bench.py flow 2>&1 | tail -4 | sed 's/^/[halves 2] /' | tee gpurun_out/r2w_stage.log
HVX_ATTN_HALVES=1 timeout -k 10 600 python scripts/stage_bench.py flow 2>&1 | tail -4 | sed 's/^/[halves 1] /' | tee -a gpurun_out/r2w_stage.log
