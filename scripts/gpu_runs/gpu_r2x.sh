#!/bin/bash
# end-of-round check: GPU eager comparator leg of bench.py alone, the full GPU test suite, smoke()
mkdir -p gpurun_out
timeout -k 5 140 python - > gpurun_out/r2x_eager.log 2>&1 <<'P'
import sys, json, time
sys.argv = ["bench.py"]
import bench
a = bench.parse()
t0 = time.time()
print(json.dumps(bench.gpu_eager_baseline(a, "cuda:0", 1024)))
print("wall", time.time() - t0)
P
tail -c 1500 gpurun_out/r2x_eager.log
timeout -k 5 150 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -3 | tee gpurun_out/r2x_tests.log
timeout -k 5 60 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/r2x_smoke.log
