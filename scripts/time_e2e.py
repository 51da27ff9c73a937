"""Wall time of ModelManager.synthesize_batch (hvx_synthesize_host) on a bench workload: python scripts/time_e2e.py [workload] [reps]
Prints tokens/s and the stage stream times; used for A/B runs of pipeline switches (HVX_PIPE_OVERLAP, HVX_PIPE_MIN_FRAMES)."""
import os, sys, time, types, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from flowmirror_hydravox_b200 import dims as D, synth
from flowmirror_hydravox_b200.model_manager import ModelManager

wlname = sys.argv[1] if len(sys.argv) > 1 else "c3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
parity = os.environ.get("MODE", "parity") == "parity"
a = types.SimpleNamespace(workload=wlname, batch=0, head_k=0, cfm_steps=25)
batch, (lo, hi), head_k, _ = bench.wl(a)
reqs = bench.make_requests(a, batch)
max_tok = int(hi * bench.RATIO)
mm = ModelManager(hd=D.HIFT_FULL, fd=D.FLOW_FULL, ld=D.LLM_FULL, device="cuda:0", max_ctx=2 + bench.P_TEXT + hi + bench.P_TOK + max_tok + 64,
                  max_seqs=batch, n_timesteps=25, sine_seconds=max_tok / 25 + 2, kv_f32=parity, flow_precise=parity)
mm.load_state_dicts(synth.llm_state_dict(D.LLM_FULL, 0, dtype=torch.bfloat16, eos_scale=0.0), synth.flow_state_dict(D.FLOW_FULL, 0),
                    synth.hift_state_dict(D.HIFT_FULL, 0))
u = torch.zeros(len(reqs), 4 * max_tok + 1024)
for i, r in enumerate(reqs):
    u[i, : r["u"].numel()] = r["u"]
run = lambda: mm.synthesize_batch(reqs, head_k=head_k, sampling=bench.SAMPLING, n_timesteps=25, min_ratio=bench.RATIO, max_ratio=bench.RATIO, u=u,
                                  return_tokens=True)
wavs, toks = run()
torch.cuda.synchronize()
ref = [w.clone() for w in wavs]
for _ in range(reps):
    t0 = time.perf_counter()
    wavs, toks = run()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    n = sum(len(t) for t in toks)
    same = all(torch.equal(x, y) for x, y in zip(wavs, ref))
    print(f"[{wlname} overlap={os.environ.get('HVX_PIPE_OVERLAP', '1')} min_frames={os.environ.get('HVX_PIPE_MIN_FRAMES', '4096')}] {n} tokens in {dt * 1e3:.0f} ms = "
          f"{n / dt:.0f} tokens/s, RTF {25 * dt / n:.5f}; stage stream ms {dict((k, round(v)) for k, v in mm.last_stage_ms.items())}; "
          f"bitwise equal to the first run: {same}", flush=True)
