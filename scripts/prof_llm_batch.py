"""Kernel-time table of a batched decode (torch.profiler / CUPTI; B sequences x head_k rows per step):
python scripts/prof_llm_batch.py [B] [head_k] [n_text]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmirror_hydravox_b200 import dims as D, synth, _lib as L
from flowmirror_hydravox_b200.llm import NativeLLM
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n_text = int(sys.argv[3]) if len(sys.argv) > 3 else 16
ld = D.LLM_FULL
kv32 = os.environ.get("KV32") == "1"
e = L.Engine(ld=ld, max_ctx=2 + 16 + n_text + 125 + 8 * n_text + 64, max_seqs=B, kv_f32=kv32); m = NativeLLM(e)
m.load_state_dict(synth.llm_state_dict(ld, 0, dtype=torch.bfloat16, eos_scale=0.0))
reqs = []
for i in range(B):
    u = synth.utterance(ld, D.FLOW_FULL, n_text, seed=100 + i)
    reqs.append(dict(text=u["text"], prompt_text=u["prompt_text"], prompt_speech=u["prompt_speech"]))
uu = torch.rand(B, 4 * 8 * n_text + 1024, generator=torch.Generator().manual_seed(0))
sp = dict(top_p=0.9, top_k=10, win_size=24, tau_r=0.2)
m.generate_batch(reqs, head_k=K, sampling=sp, min_ratio=8, max_ratio=8, u=uu)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    out = m.generate_batch(reqs, head_k=K, sampling=sp, min_ratio=8, max_ratio=8, u=uu)
    torch.cuda.synchronize()
print(f"B={B} head_k={K} n_text={n_text} kv32={kv32}: {sum(len(o) for o in out)} tokens, {len(out[0]) // K} decode steps")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
