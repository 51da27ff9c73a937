"""First-audio latency of the streaming path (BASELINE configs[4] shape): python scripts/first_audio.py [runs]"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from flowmirror_hydravox_b200 import dims as D, synth
from flowmirror_hydravox_b200.model_manager import ModelManager
from flowmirror_hydravox_b200.streaming import StreamingSynthesizer
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 10
parity = os.environ.get("MODE", "parity") == "parity"
ld, fd, hd = D.LLM_FULL, D.FLOW_FULL, D.HIFT_FULL
mm = ModelManager(hd=hd, fd=fd, ld=ld, device="cuda:0", max_ctx=2048, max_seqs=1, n_timesteps=10, sine_seconds=60, kv_f32=parity, flow_precise=parity)
mm.load_state_dicts(synth.llm_state_dict(ld, 0, dtype=torch.bfloat16, eos_scale=0.0), synth.flow_state_dict(fd, 0), synth.hift_state_dict(hd, 0))
rq = synth.utterance(ld, fd, 64, seed=77)
for inc in (False, True):
    ss = StreamingSynthesizer(mm, incremental=inc)
    lat, tot, parts = [], [], []
    for i in range(runs + 2):
        dbg = {}
        t0 = time.perf_counter()
        n = sum(c["tts_speech"].shape[1] for c in ss.tts(rq, head_k=2, sampling=bench.SAMPLING, n_timesteps=10, min_ratio=8, max_ratio=8, debug=dbg))
        if i >= 2:
            lat.append(dbg["first_audio_ms"]); tot.append((time.perf_counter() - t0) * 1e3)
    lat.sort(); tot.sort()
    print(f"[lib {os.path.basename(os.environ.get('HVX_LIB_PATH', 'default'))} incremental={inc}] first audio p50 {lat[len(lat) // 2]:.1f} ms (min {lat[0]:.1f}), "
          f"whole stream p50 {tot[len(tot) // 2]:.0f} ms for {n / hd.sr:.1f} s of audio", flush=True)
