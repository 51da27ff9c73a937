#!/bin/bash
# end-of-round evidence, 1 GPU: full parity suite, launch list of a bounded bench step (ncu), the default bench line
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -rf --timeout 600 --maxfail 40 > gpurun_out/r2s_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s_tests.log
tail -4 gpurun_out/r2s_tests.log
R=/tmp/ncu; mkdir -p $R
HVX_PROFILE=1 timeout -k 10 1100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $R/launches.csv python bench.py --workload c2 --n-text 32 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-extras --first-audio-runs 0 > gpurun_out/r2s_bench_under_ncu.log 2>&1
python scripts/ncu_summary.py $R/launches.csv 40 > gpurun_out/r2s_launches_summary.txt 2>&1
grep -m1 -n '"ID"' $R/launches.csv | head -1
awk 'NR>1' $R/launches.csv | grep '^"' | head -401 > gpurun_out/r2s_launches_head400.csv
head -12 gpurun_out/r2s_launches_summary.txt
timeout -k 10 1500 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r2s_bench.json
tail -3 gpurun_out/r2s_bench.err
