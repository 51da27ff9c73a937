import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmirror_hydravox_b200 import dims as D, synth, _lib as L
from flowmirror_hydravox_b200.llm import NativeLLM
from oracle import llm_ref
ld = D.LLM_TINY
e = L.Engine(ld=ld, max_ctx=512, max_seqs=4); m = NativeLLM(e)
sd = synth.llm_state_dict(ld, 0, eos_scale=0.0); m.load_state_dict(sd)
sdb = {k: (v.to(torch.bfloat16).float() if v.ndim >= 2 else v.float()) for k, v in sd.items()}
g = torch.Generator().manual_seed(3)
reqs = [dict(text=torch.randint(0, ld.text_vocab, (n,), generator=g), prompt_text=torch.randint(0, ld.text_vocab, (3,), generator=g),
             prompt_speech=torch.randint(0, ld.speech_token_size, (p,), generator=g)) for n, p in ((5, 0), (9, 4), (7, 11))]
u = torch.rand(3, 2048, generator=g)
sp = dict(top_p=0.9, top_k=10, win_size=24, tau_r=0.2)
batch = m.generate_batch(reqs, head_k=2, u=u, sampling=sp, min_ratio=4, max_ratio=4)
for i, r in enumerate(reqs):
    single = m.generate_batch([r], head_k=2, u=u[i:i + 1], sampling=sp, min_ratio=4, max_ratio=4)[0]
    ref = llm_ref.inference(sdb, ld, r["text"], r["prompt_text"], r["prompt_speech"], u[i], head_k=2, sp=sp, min_ratio=4, max_ratio=4, kv_dtype=torch.bfloat16)
    f = lambda a, b: next((k for k, (x, y) in enumerate(zip(a, b)) if x != y), min(len(a), len(b)))
    print(i, "len", len(batch[i]), len(single), len(ref), "batch~ref", f(batch[i], ref), "single~ref", f(single, ref), "batch~single", f(batch[i], single))
    print("  batch ", batch[i][:8]); print("  single", single[:8]); print("  ref   ", ref[:8])
