#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_unet_nc_gpu.py tests/test_unet_gpu.py tests/test_streaming_gpu.py tests/test_e2e_gpu.py -m gpu -q -rf --timeout 600 --maxfail 30 -s > gpurun_out/r2l_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_tests.log
grep -E "^\[unet_nc|passed|failed|rc=|Error|error" gpurun_out/r2l_tests.log | tail -40
HVX_PIPE_OVERLAP=0 timeout -k 10 600 python scripts/time_e2e.py c3 1 2>&1 | tail -2 | tee gpurun_out/r2l_e2e_serial.log
timeout -k 10 600 python scripts/time_e2e.py c3 2 2>&1 | tail -3 | tee gpurun_out/r2l_e2e_overlap.log
HVX_PIPE_MIN_FRAMES=1 timeout -k 10 600 python scripts/time_e2e.py c3 1 2>&1 | tail -2 | tee gpurun_out/r2l_e2e_overlap_min1.log
