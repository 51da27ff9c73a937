"""Device-side timeline of the fused persistent decode step (CTA 0): HVX_FUSED_TIMELINE=1 python scripts/fused_timeline.py [ctx]"""
import sys, os, torch, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("HVX_FUSED_TIMELINE", "1")
from flowmirror_hydravox_b200 import dims as D, synth, _lib as L
from flowmirror_hydravox_b200.llm import NativeLLM
ld = D.LLM_FULL
ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 800
e = L.Engine(ld=ld, max_ctx=4200, max_seqs=1); m = NativeLLM(e)
m.load_state_dict(synth.llm_state_dict(ld, 0, dtype=torch.bfloat16, eos_scale=0.0))
ms = (C.c_float * 1)()
paces = [int(x) for x in os.environ.get("PACES", "").split(",") if x]
for which in (0, 7, 8):
    for pace in (paces if which == 7 and paces else [None]):
        if pace is not None:
            os.environ["HVX_FUSED_PACE"] = str(pace)
        L.check(L.lib().hvx_llm_bench_kernels(e.h, 1, 2, ctx, which, 20, ms))
        print(f"which={which} ctx={ctx} pace={pace}: {ms[0]*1e3:.1f} us per step")
