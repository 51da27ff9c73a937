#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python scripts/first_audio.py 8 2>&1 | grep "^\[lib" | tee gpurun_out/r2j3_first_audio.log
timeout -k 10 600 python scripts/stage_bench.py flow 2>&1 | tail -4 | tee gpurun_out/r2j3_stage.log
