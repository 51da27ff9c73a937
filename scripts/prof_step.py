"""Kernel-time table of ONE bench step (CUPTI through torch.profiler: real, overlapping-free in-stream durations, not serialised
like ncu): python scripts/prof_step.py [workload] [n_text]   -> per-kernel share of the step's kernel time, launches, mean us."""
import os, sys, types, collections, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from flowmirror_hydravox_b200 import dims as D, synth
from flowmirror_hydravox_b200.model_manager import ModelManager
from torch.profiler import profile, ProfilerActivity

wlname = sys.argv[1] if len(sys.argv) > 1 else "c2"
a = types.SimpleNamespace(workload=wlname, batch=0, head_k=0, cfm_steps=25, n_text=int(sys.argv[2]) if len(sys.argv) > 2 else 0)
parity = os.environ.get("MODE", "parity") == "parity"
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
run(); run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    wavs, toks = run()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type is None or "cuda" not in str(ev.device_type).lower():
        continue
    name = ev.name.split("(")[0].replace("void ", "").replace("hvx::", "")
    t = agg.setdefault(name, [0, 0.0])
    t[0] += 1; t[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values())
n_tok = sum(len(t) for t in toks)
print(f"# {bench.workload_name(a)}; mode {'parity' if parity else 'serving'}; one hvx_synthesize_host step: {n_tok} speech tokens, "
      f"{sum(v[0] for v in agg.values())} device activities, {tot / 1e3:.1f} ms of device time; stage stream ms {mm.last_stage_ms}")
print(f"{'share':>7} {'total ms':>10} {'count':>8} {'mean us':>9}  kernel / activity")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{100 * v[1] / tot:6.2f}% {v[1] / 1e3:10.2f} {v[0]:8d} {v[1] / v[0]:9.2f}  {k[:110]}")
