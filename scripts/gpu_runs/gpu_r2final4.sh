#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_llm_gpu.py tests/test_e2e_gpu.py tests/test_c2_gpu.py tests/test_streaming_gpu.py -m gpu -q --timeout 300 2>&1 | tail -2 | tee gpurun_out/r2f4_tests.log
timeout -k 10 1500 python bench.py > gpurun_out/r2f4_bench.json 2> gpurun_out/r2f4_bench.err
echo "bench rc=$?"
tail -c 200 gpurun_out/r2f4_bench.json
tail -2 gpurun_out/r2f4_bench.err
