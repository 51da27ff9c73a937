#!/bin/bash
# end of round: full parity suite + smoke + default bench with the final binary
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -rf --timeout 600 --maxfail 40 > gpurun_out/r2f2_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f2_tests.log
tail -3 gpurun_out/r2f2_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/r2f2_smoke.log
timeout -k 10 1500 python bench.py > gpurun_out/r2f2_bench.json 2> gpurun_out/r2f2_bench.err
echo "bench rc=$?"
tail -c 300 gpurun_out/r2f2_bench.json
tail -3 gpurun_out/r2f2_bench.err
