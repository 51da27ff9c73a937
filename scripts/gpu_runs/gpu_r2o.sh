#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_e2e_gpu.py tests/test_flow_gpu.py tests/test_gemm_gpu.py -m gpu -q --timeout 300 2>&1 | tail -3 | tee gpurun_out/r2o_tests.log
for r in 12 0 32; do
HVX_PIPE_RESERVE_SMS=$r timeout -k 10 600 python scripts/time_e2e.py c3 1 2>&1 | tail -1 | sed "s/^/[reserve $r] /" | tee -a gpurun_out/r2o_e2e.log
done
HVX_LLM_STREAM_PRIO=0 HVX_PIPE_RESERVE_SMS=12 timeout -k 10 600 python scripts/time_e2e.py c3 1 2>&1 | tail -1 | sed "s/^/[reserve 12, no stream priority] /" | tee -a gpurun_out/r2o_e2e.log
