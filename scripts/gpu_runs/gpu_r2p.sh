#!/bin/bash
mkdir -p gpurun_out
R=/tmp/ncu; mkdir -p $R
KV32=1 timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:llm_attn_mma -s 400 -c 2 -o $R/attn -f python scripts/prof_llm_batch.py 32 4 288 > gpurun_out/r2p_ncu.log 2>&1
tail -3 gpurun_out/r2p_ncu.log
python scripts/ncu_full_summary.py $R/attn.ncu-rep > gpurun_out/r2p_attn_summary.txt 2>&1
ncu -i $R/attn.ncu-rep --page details --csv 2>/dev/null | cut -d, -f5,12- | cut -c1-600 > gpurun_out/r2p_attn_details.csv
ncu -i $R/attn.ncu-rep --page source --csv --print-source sass 2>/dev/null | head -3000 > gpurun_out/r2p_attn_source.csv
ls -la gpurun_out/r2p_*
