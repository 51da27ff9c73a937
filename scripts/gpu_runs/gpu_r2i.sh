#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1500 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --first-audio-runs 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
echo "bench rc=$?"
tail -c 2500 gpurun_out/r2i_bench.json
tail -5 gpurun_out/r2i_bench.err
