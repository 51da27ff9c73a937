#!/bin/bash
# round 2, first GPU call: parity tests, stage micro-benchmarks, a short default bench run
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout -k 10 900 python -m pytest tests -m gpu -q -rf --timeout 600 --maxfail 40 -s > gpurun_out/r2a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
timeout -k 10 600 python scripts/stage_bench.py all > gpurun_out/r2a_stage.log 2>&1
echo "stage rc=$?" >> gpurun_out/r2a_stage.log
cat gpurun_out/r2a_stage.log | tail -12
timeout -k 10 1200 python bench.py --steps 1 --warmup 1 --e2e-steps 1 --first-audio-runs 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r2a_bench.json
tail -5 gpurun_out/r2a_bench.err
