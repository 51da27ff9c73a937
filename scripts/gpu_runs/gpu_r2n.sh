#!/bin/bash
mkdir -p gpurun_out
for r in 12 24 4; do
HVX_PIPE_RESERVE_SMS=$r timeout -k 10 600 python scripts/time_e2e.py c3 1 2>&1 | tail -1 | sed "s/^/[reserve $r] /" | tee -a gpurun_out/r2n_e2e.log
done
timeout -k 10 600 python -m pytest tests/test_attention_gpu.py tests/test_flow_gpu.py -m gpu -q --timeout 600 2>&1 | tail -3 | tee gpurun_out/r2n_tests.log
