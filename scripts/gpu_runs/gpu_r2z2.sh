#!/bin/bash
# the sharded-synthesis entry under torchrun with ONE rank: NCCL single-rank communicator, packed scatter / gather, engine
mkdir -p gpurun_out
timeout -k 3 28 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29534 scripts/run_sharded.py > gpurun_out/r2z2_sharded1.log 2>&1
echo "rc=$?" >> gpurun_out/r2z2_sharded1.log
grep -E "SHARDED|rc=|Error" gpurun_out/r2z2_sharded1.log | tail -5
