#!/bin/bash
# BASELINE configs[0] end-to-end parity at full dims against the chained reference modules (tests/test_c1_gpu.py)
mkdir -p gpurun_out
timeout -k 5 85 python -m pytest tests/test_c1_gpu.py -m gpu -q -s --timeout 80 > gpurun_out/r2z1_c1.log 2>&1
echo "rc=$?" >> gpurun_out/r2z1_c1.log
grep -E "^\[c1\]|passed|failed|rc=|Error|assert" gpurun_out/r2z1_c1.log | tail -12
