#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python scripts/bench_stream_flow.py 1000 10 2>&1 | tail -4 | tee gpurun_out/r2r_stream_flow.log
timeout -k 10 900 python scripts/bench_stream_flow.py 3000 10 2>&1 | tail -4 | tee -a gpurun_out/r2r_stream_flow.log
