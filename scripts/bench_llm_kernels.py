import sys, os, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmirror_hydravox_b200 import dims as D, synth, _lib as L
from flowmirror_hydravox_b200.llm import NativeLLM
ld = D.LLM_FULL
e = L.Engine(ld=ld, max_ctx=2048, max_seqs=1); m = NativeLLM(e)
m.load_state_dict(synth.llm_state_dict(ld, 0, dtype=torch.bfloat16, eos_scale=0.0))
names = ["whole step (no sampler)", "qkv gemv x24", "attention x24", "o-proj x24", "gate-up x24", "down x24", "MTP heads+logits", "whole step, fused persistent kernel"]
ms = (C.c_float * 1)()
for which in range(8):
    L.check(L.lib().hvx_llm_bench_kernels(e.h, 1, 2, 800, which, 20, ms))
    n = 1 if which in (0, 6, 7) else 24
    print(f"{names[which]:28s} {ms[0]*1e3:8.1f} us per rep  ({ms[0]*1e3/n:6.2f} us per launch)" if n > 1 else f"{names[which]:28s} {ms[0]*1e3:8.1f} us")
