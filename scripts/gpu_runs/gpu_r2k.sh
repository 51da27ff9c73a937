#!/bin/bash
# round 2 ncu evidence: --set full captures of the top kernels of each stage, summarised ON the box (the .ncu-rep files are
# too large to travel back), plus the launch list of one bounded default-workload bench step
mkdir -p gpurun_out
R=/tmp/ncu; mkdir -p $R
HVX_FLOW_PRECISE=1 timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_pair3|gemm_persist|gemm_bf16|dit_attention_v5|dit_ln" -s 60 -c 12 -o $R/flow_parity -f python scripts/prof_flow.py 1 > gpurun_out/r2k_ncu_flow_parity.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_persist|dit_attention_v5" -s 30 -c 6 -o $R/flow_serving -f python scripts/prof_flow.py 1 > gpurun_out/r2k_ncu_flow_serving.log 2>&1
KV32=1 timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"llm_attn_mma|gemm_bf16_kernel|llm_norm16|llm_splitk" -s 3000 -c 12 -o $R/llm_b32 -f python scripts/prof_llm_batch.py 32 4 64 > gpurun_out/r2k_ncu_llm.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"llm_gemv|llm_attn|llm_sampler" -s 2000 -c 12 -o $R/llm_b1 -f python scripts/prof_llm_batch.py 1 2 32 > gpurun_out/r2k_ncu_llm_b1.log 2>&1
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_kernel|gemm_persist|hift_act_up|stft|istft" -s 20 -c 12 -o $R/hift -f python scripts/stage_bench.py hift > gpurun_out/r2k_ncu_hift.log 2>&1
for n in flow_parity flow_serving llm_b32 llm_b1 hift; do
  python scripts/ncu_full_summary.py $R/$n.ncu-rep > gpurun_out/r2k_ncu_${n}_summary.txt 2>&1
  ncu -i $R/$n.ncu-rep --page details --csv 2>/dev/null | grep -iE "stall|Issue Slots Busy|No Eligible|Eligible Warps|Theoretical Occupancy|Achieved Occupancy|Bank conflicts|L2 Hit|Mem Busy|Max Bandwidth" | cut -c1-400 | head -400 > gpurun_out/r2k_ncu_${n}_details.csv
done
# launch list of one bounded bench step (c2 workload: 1 utterance) for the kernels' SHARES
HVX_PROFILE=1 timeout -k 10 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $R/launches.csv python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-extras --first-audio-runs 0 > gpurun_out/r2k_bench_under_ncu.log 2>&1
python scripts/ncu_summary.py $R/launches.csv > gpurun_out/r2k_launches_c2_summary.txt 2>&1
head -400 $R/launches.csv > gpurun_out/r2k_launches_c2_head400.csv
ls -la $R gpurun_out/r2k_*
