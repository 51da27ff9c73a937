"""Stage-level timings with the live per-class profiler (hvx_profile_*): python scripts/stage_bench.py [flow|hift|all]
  flow: one CFM solve at T = 2298 (C2) and one 3-utterance group of ~2700 frames (a C3 group), serving and parity mode
  hift: one vocoder pass at 2048 frames, tensor-core decode vs HVX_HIFT_FP32=1
Prints per-class ms / TFLOP/s so a kernel change can be judged from one gpurun call."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmirror_hydravox_b200 import dims as D, synth, _lib as L
from flowmirror_hydravox_b200.flow import NativeFlow
from flowmirror_hydravox_b200.hift import NativeHiFT

what = sys.argv[1] if len(sys.argv) > 1 else "all"
steps = int(os.environ.get("STEPS", "5"))


def report(tag, e, ms_total):
    prof = e.profile_collect()
    parts = []
    for k, (ms, work, n) in prof.items():
        if n:
            unit = f"{work / ms / 1e9:.0f} TFLOP/s" if k in ("gemm", "attention", "hift_conv") else f"{work / ms / 1e6:.0f} GB/s"
            parts.append(f"{k} {ms:.1f} ms ({100 * ms / ms_total:.0f} %, {unit}, {n} sites)")
    print(f"[{tag}] total {ms_total:.1f} ms | " + " | ".join(parts), flush=True)


def timed(fn):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b)


if what in ("flow", "all"):
    fd = D.FLOW_FULL
    sd = synth.flow_state_dict(fd, 0)
    u = synth.utterance(D.LLM_FULL, fd, 128, seed=1986)
    g = torch.Generator().manual_seed(3)
    for precise in (False, True):
        e = L.Engine(fd=fd, flow_precise=precise)
        f = NativeFlow(e); f.load_state_dict(sd)
        tok = torch.randint(0, fd.vocab, (1, 1024), generator=g)
        kw = dict(token=tok, embedding=u["embedding"][None], prompt_token=u["prompt_speech"][None], prompt_feat=u["prompt_feat"][None], n_timesteps=steps)
        f.inference(**kw)
        e.profile(True); e.profile_collect()
        ms = timed(lambda: f.inference(**kw))
        e.profile_collect()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f.inference(**kw); b.record(); torch.cuda.synchronize()
        report(f"flow {'parity' if precise else 'serving'} T=2298 x{steps} NFE ({a.elapsed_time(b) / steps:.2f} ms/NFE)", e, a.elapsed_time(b))
        reqs = [dict(token=torch.randint(0, fd.vocab, (1, n), generator=g), embedding=u["embedding"][None], prompt_token=u["prompt_speech"][None],
                     prompt_feat=u["prompt_feat"][None]) for n in (1240, 1180, 1100)]
        f.inference_batch(reqs, n_timesteps=steps)
        e.profile_collect()
        a.record(); f.inference_batch(reqs, n_timesteps=steps); b.record(); torch.cuda.synchronize()
        report(f"flow {'parity' if precise else 'serving'} group 3 x ~2700 frames x{steps} NFE ({a.elapsed_time(b) / steps:.2f} ms/NFE)", e, a.elapsed_time(b))
        e.close()

if what in ("hift", "all"):
    hd = D.HIFT_FULL
    T = int(os.environ.get("HIFT_T", "2048"))
    e = L.Engine(hd=hd)
    v = NativeHiFT(e, sine_table=synth.hift_sine_table(hd, T))
    v.load_state_dict(synth.hift_state_dict(hd, 0))
    mel = (torch.rand(1, hd.mel, T, generator=torch.Generator().manual_seed(5)) * 6 - 6).cuda()
    for fp32 in ("0", "1"):
        os.environ["HVX_HIFT_FP32"] = fp32
        v.inference(mel); v.inference(mel)
        e.profile(True); e.profile_collect()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); v.inference(mel); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        report(f"hift {'fp32 CUDA-core' if fp32 == '1' else 'tensor-core'} decode T={T} ({2 * 336.9e6 * T / ms / 1e9:.1f} TFLOP/s algorithmic)", e, ms)
        e.profile(False)
    e.close()
