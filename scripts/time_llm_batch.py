"""Wall time of a batched LLM stage (prefill + decode) : python scripts/time_llm_batch.py B head_k n_text
(compare HVX_LLM_SPLITK=0/1 and HVX_GEMV_MAX_ROWS=8/32 across runs)"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmirror_hydravox_b200 import dims as D, synth, _lib as L
from flowmirror_hydravox_b200.llm import NativeLLM
B, K, n_text = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ld = D.LLM_FULL
e = L.Engine(ld=ld, max_ctx=4096, max_seqs=B); m = NativeLLM(e)
m.load_state_dict(synth.llm_state_dict(ld, 0, dtype=torch.bfloat16, eos_scale=0.0))
reqs = []
for i in range(B):
    u = synth.utterance(ld, D.FLOW_FULL, n_text, seed=100 + i)
    reqs.append(dict(text=u["text"], prompt_text=u["prompt_text"], prompt_speech=u["prompt_speech"]))
uu = torch.rand(B, 8192, generator=torch.Generator().manual_seed(0))
sp = dict(top_p=0.9, top_k=10, win_size=24, tau_r=0.2)
out = m.generate_batch(reqs, head_k=K, sampling=sp, min_ratio=8, max_ratio=8, u=uu)
torch.cuda.synchronize()
t0 = time.perf_counter()
out2 = m.generate_batch(reqs, head_k=K, sampling=sp, min_ratio=8, max_ratio=8, u=uu)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
n = sum(len(o) for o in out2)
print(f"B={B} K={K} n_text={n_text} splitk={os.environ.get('HVX_LLM_SPLITK', '1')} gemv_max_rows={os.environ.get('HVX_GEMV_MAX_ROWS', '8')}: "
      f"{dt * 1e3:.1f} ms for {n} tokens ({len(out2[0]) // K} steps, {dt * 1e6 / max(1, len(out2[0]) // K):.0f} us/step incl. prefill), same as first run: {out == out2}")
