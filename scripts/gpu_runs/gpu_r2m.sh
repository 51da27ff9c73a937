#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python scripts/stage_bench.py flow 2>&1 | tail -4 | tee gpurun_out/r2m_stage.log
timeout -k 10 900 python -m pytest tests/test_flow_gpu.py tests/test_gemm_gpu.py tests/test_c2_gpu.py -m gpu -q -rf --timeout 600 2>&1 | tail -5 | tee gpurun_out/r2m_tests.log
timeout -k 10 600 python scripts/time_e2e.py c3 1 2>&1 | tail -1 | tee gpurun_out/r2m_e2e_overlap_prio.log
HVX_PIPE_OVERLAP=0 timeout -k 10 600 python scripts/time_e2e.py c3 1 2>&1 | tail -1 | tee gpurun_out/r2m_e2e_serial.log
