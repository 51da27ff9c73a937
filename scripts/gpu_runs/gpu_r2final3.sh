#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1500 python bench.py > gpurun_out/r2f3_bench.json 2> gpurun_out/r2f3_bench.err
echo "bench rc=$?"
tail -c 300 gpurun_out/r2f3_bench.json
tail -3 gpurun_out/r2f3_bench.err
