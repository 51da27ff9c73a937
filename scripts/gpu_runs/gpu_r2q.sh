#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_flow_gpu.py tests/test_streaming_gpu.py tests/test_gemm_gpu.py tests/test_unet_gpu.py -m gpu -q -rf -s --timeout 300 > gpurun_out/r2q_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2q_tests.log
grep -E "flow stream|passed|failed|rc=|Error|error|FAILED" gpurun_out/r2q_tests.log | tail -30
