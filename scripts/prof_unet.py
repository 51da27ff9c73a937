"""Short U-Net-estimator run for ncu: full dims, T=2298 (the bench utterance), eager launches (no graph replay)."""
import os, sys, torch
os.environ["HVX_UNET_NO_GRAPH"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmirror_hydravox_b200 import dims as D, synth, _lib as L
from flowmirror_hydravox_b200.flow import NativeUNetEstimator
T = int(sys.argv[1]) if len(sys.argv) > 1 else 2298
ud = D.UNET_FULL
e = L.Engine(ud=ud, flow_precise=os.environ.get("HVX_FLOW_PRECISE") == "1")
m = NativeUNetEstimator(e).load_state_dict(synth.unet_state_dict(ud, 0))
g = torch.Generator().manual_seed(0)
x, mu, cond = (torch.randn(2, ud.mel, T, generator=g).cuda() for _ in range(3))
spks, t = torch.randn(2, ud.mel, generator=g).cuda(), torch.tensor([0.5, 0.5]).cuda()
for _ in range(2):
    m(x, None, mu, t, spks, cond)
torch.cuda.synchronize()
