#!/bin/bash
# 2-GPU check: NCCL scatter/gather test + the bench's sharded e2e leg
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout -k 10 600 python -m pytest tests/test_parallel_gpu.py -m gpu -q --timeout 300 -rs 2>&1 | tail -4 | tee gpurun_out/r2z_parallel_test.log
timeout -k 10 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 1 --warmup 1 --e2e-steps 1 --no-extras --no-cpu-baseline --first-audio-runs 0 > gpurun_out/r2z_bench_n2.json 2> gpurun_out/r2z_bench_n2.err
echo "bench rc=$?"; tail -c 1800 gpurun_out/r2z_bench_n2.json; tail -3 gpurun_out/r2z_bench_n2.err
