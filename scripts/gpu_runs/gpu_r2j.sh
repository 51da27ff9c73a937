#!/bin/bash
# round 2 evidence call: full GPU parity suite, per-class stage timings, ncu --set full of the top kernels of each stage
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -m gpu -q -rf --timeout 600 --maxfail 40 > gpurun_out/r2j_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_tests.log
tail -4 gpurun_out/r2j_tests.log
timeout -k 10 600 python scripts/stage_bench.py all > gpurun_out/r2j_stage.log 2>&1
tail -12 gpurun_out/r2j_stage.log
# flow, parity mode: persistent / pair three-term GEMMs + DiT attention + layernorm (2 NFE at T=2298)
HVX_FLOW_PRECISE=1 timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_pair3|gemm_persist|dit_attention_v5|dit_ln" -s 40 -c 14 -o gpurun_out/r2j_flow -f python scripts/prof_flow.py 1 > gpurun_out/r2j_ncu_flow.log 2>&1
tail -3 gpurun_out/r2j_ncu_flow.log
# LLM batched decode (32 sequences x 4 heads, fp32 KV): attention + GEMMs + norm + split-K reduce
KV32=1 timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"llm_attn_mma|gemm_bf16_kernel|llm_norm16|llm_splitk" -s 3000 -c 16 -o gpurun_out/r2j_llm -f python scripts/prof_llm_batch.py 32 4 64 > gpurun_out/r2j_ncu_llm.log 2>&1
tail -3 gpurun_out/r2j_ncu_llm.log
# HiFT: tensor-core convolutions
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_kernel|gemm_persist|hift_act_up|conv1d" -s 30 -c 14 -o gpurun_out/r2j_hift -f python scripts/stage_bench.py hift > gpurun_out/r2j_ncu_hift.log 2>&1
tail -3 gpurun_out/r2j_ncu_hift.log
ls -la gpurun_out/r2j_*
