"""Short LLM-only run for ncu: full dims, BASELINE config-2 prompt, a few decode steps."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmirror_hydravox_b200 import dims as D, synth, _lib as L
from flowmirror_hydravox_b200.llm import NativeLLM
ld = D.LLM_FULL
n_tok = int(sys.argv[1]) if len(sys.argv) > 1 else 16
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
head_k = int(sys.argv[3]) if len(sys.argv) > 3 else 2
e = L.Engine(ld=ld, max_ctx=2048, max_seqs=batch); m = NativeLLM(e)
m.load_state_dict(synth.llm_state_dict(ld, 0, dtype=torch.bfloat16, eos_scale=0.0))
reqs = []
for i in range(batch):
    u = synth.utterance(ld, D.FLOW_FULL, 128, seed=1986 + i)
    reqs.append(dict(text=u["text"], prompt_text=u["prompt_text"], prompt_speech=u["prompt_speech"]))
ratio = n_tok / 128.0
sp = dict(top_p=0.9, top_k=10, win_size=24, tau_r=0.2)
for it in range(2):
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    out = m.generate_batch(reqs, head_k=head_k, sampling=sp, min_ratio=ratio, max_ratio=ratio, u=torch.rand(batch, 4096, generator=torch.Generator().manual_seed(1)))
    t1.record(); torch.cuda.synchronize()
    print("tokens", [len(o) for o in out], "ms", t0.elapsed_time(t1))
