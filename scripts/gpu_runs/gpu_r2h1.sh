#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_flow_gpu.py tests/test_streaming_gpu.py tests/test_e2e_gpu.py tests/test_c2_gpu.py -m gpu -q --timeout 300 2>&1 | tail -3 | tee gpurun_out/r2h1_tests.log
for g in 640 0; do
HVX_FLOW_GRAPH_FRAMES=$g timeout -k 10 900 python bench.py --workload c2 --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --first-audio-runs 10 > gpurun_out/r2h1_bench_c2_g$g.json 2> gpurun_out/r2h1_bench_c2_g$g.err
python - <<P
import json
d=json.loads([l for l in open('gpurun_out/r2h1_bench_c2_g$g.json') if l.startswith('{')][-1])
print("graph_frames=$g", "e2e", round(d['e2e']['value']), d['e2e']['stage_ms_per_step_rank0'], "first_audio", {k:round(v,1) for k,v in d['first_audio'].items() if isinstance(v,float)})
P
done
