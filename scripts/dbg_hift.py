import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmirror_hydravox_b200 import dims as D, synth, _lib as L
from flowmirror_hydravox_b200.hift import NativeHiFT
from oracle import hift_ref
hd = D.HIFT_FULL
e = L.Engine(hd=hd); h = NativeHiFT(e); sd = synth.hift_state_dict(hd, 0); h.load_state_dict(sd)
g = torch.load("tests/golden/hift_full.pt", weights_only=False)
T = g["T"]; print("T", T)
h.set_sine_table(synth.hift_sine_table(hd, max(T, 2048)))
wav, src, f0 = h.inference(g["mel"], return_f0=True)
print("f0 rel", ((f0.cpu() - g["f0"][0]).abs() / (g["f0"][0].abs() + 1)).max().item())
wav, src = h.inference(g["mel"], f0=g["f0"])
wav2, _ = h.inference(g["mel"], f0=g["f0"])
print("determinism", (wav - wav2).abs().max().item())
err = (wav.cpu() - g["wav"]).abs()
print("src err", (src.cpu() - g["src"]).abs().max().item())
print("wav rms", err.pow(2).mean().sqrt().item(), "max", err.max().item(), "argmax", err.argmax().item(), "n", err.numel())
fe = err.reshape(-1, 480).max(1).values
print("per-frame max err", [round(x, 5) for x in fe.tolist()[:40]])
wav_s, _ = h.inference(g["mel"], finalize=False, f0=g["f0_stream"])
print("stream rms", (wav_s.cpu() - g["wav_stream"]).pow(2).mean().sqrt().item())
