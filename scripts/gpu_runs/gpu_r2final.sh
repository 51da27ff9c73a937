#!/bin/bash
# end-of-round evidence, 1 GPU: full parity suite, smoke, ncu --set full of the attention kernel as it ends the round, default bench
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -m gpu -q -rf --timeout 600 --maxfail 40 > gpurun_out/r2f_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_tests.log
tail -3 gpurun_out/r2f_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/r2f_smoke.log
R=/tmp/ncu; mkdir -p $R
HVX_FLOW_PRECISE=1 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"dit_attention_v5|gemm_pair3" -s 60 -c 6 -o $R/flow_final -f python scripts/prof_flow.py 1 > gpurun_out/r2f_ncu.log 2>&1
python scripts/ncu_full_summary.py $R/flow_final.ncu-rep > gpurun_out/r2f_ncu_flow_summary.txt 2>&1
KV32=1 timeout -k 10 600 ncu --set full --clock-control none -k regex:"llm_attn_mma" -s 3000 -c 2 -o $R/llm_final -f python scripts/prof_llm_batch.py 32 4 288 > gpurun_out/r2f_ncu_llm.log 2>&1
python scripts/ncu_full_summary.py $R/llm_final.ncu-rep > gpurun_out/r2f_ncu_llm_summary.txt 2>&1
timeout -k 10 1500 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r2f_bench.json
tail -3 gpurun_out/r2f_bench.err
