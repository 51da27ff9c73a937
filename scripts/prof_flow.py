"""Short flow-only run for ncu: full dims, BASELINE config-2 frames (T=2298), a few Euler steps."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmirror_hydravox_b200 import dims as D, synth, _lib as L
from flowmirror_hydravox_b200.flow import NativeFlow
fd = D.FLOW_FULL
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
import os
e = L.Engine(fd=fd, flow_precise=os.environ.get("HVX_FLOW_PRECISE") == "1"); f = NativeFlow(e); f.load_state_dict(synth.flow_state_dict(fd, 0))
u = synth.utterance(D.LLM_FULL, fd, 128, seed=1986)
tok = torch.randint(0, fd.vocab, (1, 1024), generator=torch.Generator().manual_seed(3))
for it in range(2):
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    mel, _ = f.inference(token=tok, embedding=u["embedding"][None], prompt_token=u["prompt_speech"][None], prompt_feat=u["prompt_feat"][None], n_timesteps=steps)
    t1.record(); torch.cuda.synchronize()
    print("frames", mel.shape[2], "steps", steps, "ms", t0.elapsed_time(t1), "ms/NFE", t0.elapsed_time(t1) / steps)
