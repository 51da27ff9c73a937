#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gemm_gpu.py tests/test_flow_gpu.py tests/test_unet_gpu.py tests/test_unet_nc_gpu.py tests/test_c2_gpu.py tests/test_hift_gpu.py tests/test_llm_gpu.py -m gpu -q --timeout 300 2>&1 | tail -3 | tee gpurun_out/r2g2_tests.log
timeout -k 10 600 python scripts/stage_bench.py all 2>&1 | tail -6 | tee gpurun_out/r2g2_stage.log
