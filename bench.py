#!/usr/bin/env python
"""bench.py — HydraVox hot path (multi-head AR decode -> CFM Euler -> HiFT) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--head-k K] [--n-text T] [--cfm-steps S]
  python bench.py --impl reference ...      # the reference algorithm's CPU path (oracle port) on the host cores

A step = one pass of the whole hot path over one batch of synthetic utterances (BASELINE.json configs[1] by
default: one zero-shot utterance, 16+128 text tokens, 125 prompt speech tokens + 250 prompt mel frames,
inference_head_num=2, 25 CFM Euler steps, fixed-length protocol min=max ratio 8 -> 1024 speech tokens = 40.96 s
of audio).  Metric: speech tokens per second through the whole path (RTF = 25 / value, also printed).

  value      inputs resident in HBM, stage calls through the C-ABI with device pointers, CUDA-event timed
  e2e        ModelManager.synthesize_batch -> hvx_synthesize_host with pinned HOST buffers: H2D of the request,
             three stages, D2H of tokens + waveform inside the timed region
Under torchrun (N>1) every rank runs its own batch (utterances are independent): weak scaling, no
data-path collective; NCCL is used for the barrier and the max-over-ranks reduction only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from flowmirror_hydravox_b200 import dims as D, synth  # noqa: E402

SAMPLING = dict(top_p=0.9, top_k=10, win_size=24, tau_r=0.2)       # server /zero-shot defaults (router.py:22-44)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="utterances per GPU per step")
    ap.add_argument("--head-k", type=int, default=2)
    ap.add_argument("--n-text", type=int, default=128)
    ap.add_argument("--cfm-steps", type=int, default=25)
    ap.add_argument("--ratio", type=float, default=8.0, help="speech tokens per text token (min=max, SURVEY 8d)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--first-audio-runs", type=int, default=20, help="streaming first-audio latency samples (0 = skip)")
    ap.add_argument("--parity-mode-steps", type=int, default=2, help="also time the parity mode (fp32 KV cache, split-fp16 flow GEMMs); 0 = skip")
    ap.add_argument("--no-variants", action="store_true", help="skip the timing of the alternate estimator / vocoders (SURVEY 8 a7', a12')")
    ap.add_argument("--cpu-tokens", type=int, default=128, help="speech tokens in the CPU baseline sample (~10-20 s of CPU work)")
    return ap.parse_args()


def workload_name(a):
    return (f"zero-shot utterance x{a.batch}/GPU: 16+{a.n_text} text tokens, 125 prompt speech tokens (+250 prompt mel frames), "
            f"inference_head_num={a.head_k}, {a.cfm_steps} CFM Euler steps, {int(a.n_text * a.ratio)} speech tokens "
            f"({a.n_text * a.ratio / 25:.2f} s of 24 kHz audio) per utterance")


# --------------------------------------------------------------------------- algorithmic work
def llm_step_bytes(ld: D.LlmDims, head_k: int, ctx: float, n_seq: int) -> float:
    """bytes one decode step must read: bf16 weights once + the live KV cache."""
    layer = ld.hidden * (ld.q_heads + 2 * ld.kv_heads) * ld.head_dim + ld.hidden * ld.hidden + 3 * ld.hidden * ld.inter
    mtp = 2 * ld.hidden * ld.hidden + 3 * ld.hidden * ld.mtp_inter           # v, o, gate, up, down (q/k are dead)
    w = 2.0 * (ld.layers * layer + head_k * mtp + ld.speech_vocab * ld.hidden)
    kv = n_seq * ld.layers * 2 * ld.kv_heads * ld.head_dim * 2.0 * (ctx + head_k)
    return w + kv


def flow_nfe_flops(fd: D.FlowDims, T: int) -> float:
    inner = fd.heads * fd.dim_head
    per_row = fd.depth * (4 * fd.dim * inner + 2 * fd.dim * fd.dim * fd.ff_mult) + 4 * fd.mel * fd.dim \
        + 2 * fd.dim * (fd.dim // fd.pos_groups) * fd.pos_k + fd.dim * fd.mel
    return 2.0 * (2 * T) * per_row + 2.0 * fd.depth * 4 * T * T * inner      # linear + attention (2 CFG rows)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        self.th.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6 or not f[0].isdigit():
                continue
            sm.append(int(f[0])); mx = int(f[1])
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arm (oracle port)
_CPU_SDS = {}


def cpu_oracle_run(a, n_tokens: int, threads: int):
    """The reference algorithm restated on the CPU (oracle/, fp32 torch): the same three stages on a bounded sample of
    the workload — one utterance without prompt whose fixed length is n_tokens speech tokens."""
    from oracle import flow_ref, hift_ref, llm_ref
    torch.set_num_threads(threads)
    ld, fd, hd = D.LLM_FULL, D.FLOW_FULL, D.HIFT_FULL
    n_text = max(1, int(round(n_tokens / a.ratio)))
    if not _CPU_SDS:
        _CPU_SDS.update(llm=synth.llm_state_dict(ld, 0, eos_scale=0.0), flow=synth.flow_state_dict(fd, 0), hift=synth.hift_state_dict(hd, 0))
    llm_sd, flow_sd, hift_sd = _CPU_SDS["llm"], _CPU_SDS["flow"], _CPU_SDS["hift"]
    u = synth.utterance(ld, fd, n_text, seed=1986, zero_shot=False)
    us = torch.rand(8 * n_tokens + 64, generator=torch.Generator().manual_seed(7))
    noise = synth.flow_noise(fd)
    t0 = time.perf_counter()
    with torch.no_grad():
        toks = llm_ref.inference(llm_sd, ld, u["text"], u["prompt_text"], u["prompt_speech"], us, head_k=a.head_k, sp=SAMPLING,
                                 min_ratio=a.ratio, max_ratio=a.ratio)
        t1 = time.perf_counter()
        mel = flow_ref.inference(flow_sd, torch.tensor(toks)[None], u["embedding"][None], noise, fd, a.cfm_steps)
        t2 = time.perf_counter()
        table = synth.hift_sine_table(hd, mel.shape[2])
        wav, _ = hift_ref.inference(hift_sd, mel, table, hd)
        t3 = time.perf_counter()
    total = t3 - t0
    return dict(tokens=len(toks), seconds=total, stage_s=dict(llm=t1 - t0, flow=t2 - t1, hift=t3 - t2),
                tokens_per_s=len(toks) / total, rtf=total / (wav.shape[1] / hd.sr),
                sample=f"1 utterance, {n_text} text tokens, no prompt, {len(toks)} speech tokens, head_k={a.head_k}, "
                       f"{a.cfm_steps} CFM steps, full model dims, fp32 oracle port (KV-cached decode), {threads} threads")


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    res = []
    for i in range(max(1, min(a.steps, 2)) + (1 if a.warmup else 0)):
        r = cpu_oracle_run(a, a.cpu_tokens, threads)
        if i or not a.warmup:
            res.append(r)
    secs = sum(r["seconds"] for r in res) / len(res)
    tps = res[0]["tokens"] / secs
    line = {"impl": "reference", "metric": "speech_tokens_per_s", "value": tps, "unit": "tokens/s", "n_gpus": a.gpus, "steps": len(res),
            "warmup": 1 if a.warmup else 0, "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(a), "timed_sample": res[0]["sample"]},
            "rtf": 25.0 / tps, "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": res[0]["sample"],
                                                "stage_s": res[0]["stage_s"]},
            "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------- native arm
def native_arm(a):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from flowmirror_hydravox_b200 import _lib as L
    from flowmirror_hydravox_b200.model_manager import ModelManager

    ld, fd, hd = D.LLM_FULL, D.FLOW_FULL, D.HIFT_FULL
    n_tok = int(a.n_text * a.ratio)
    max_ctx = 2 + 16 + a.n_text + 125 + n_tok + 64
    mm = ModelManager(hd=hd, fd=fd, ld=ld, device=f"cuda:{local}", max_ctx=max_ctx, max_seqs=max(a.batch, 1), n_timesteps=a.cfm_steps,
                      sine_seconds=n_tok / 25 + 2)
    mm.load_state_dicts(synth.llm_state_dict(ld, 0, dtype=torch.bfloat16, eos_scale=0.0), synth.flow_state_dict(fd, 0),
                        synth.hift_state_dict(hd, 0))
    llm, flow, hift = mm.models["llm"], mm.models["flow"], mm.models["hift"]
    reqs = [synth.utterance(ld, fd, a.n_text, seed=1986 + rank * 1000 + i) for i in range(a.batch)]
    for r in reqs:
        r["min_ratio"] = r["max_ratio"] = a.ratio
    u_all = torch.rand(a.batch, 4 * n_tok + 1024, generator=torch.Generator().manual_seed(rank))
    # device-resident copies for the `value` leg
    dreqs = [{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in r.items()} for r in reqs]
    u_dev = u_all.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)        # > L2 (126 MB)

    def step_device():
        toks = llm.generate_batch(dreqs, head_k=a.head_k, u=u_dev, sampling=SAMPLING, min_ratio=a.ratio, max_ratio=a.ratio)
        # flow in groups of similar length, as hvx_synthesize_host forms them (csrc/pipeline.cu: <= 8192 padded frames, <= 20 % padding)
        items = [dict(token=torch.tensor(t, device=dev, dtype=torch.int32)[None], embedding=r["embedding"][None],
                      prompt_token=r["prompt_speech"][None], prompt_feat=r["prompt_feat"][None]) for r, t in zip(dreqs, toks)]
        items.sort(key=lambda it: -(it["token"].shape[1] + it["prompt_token"].shape[1]))
        fr = lambda it: 2 * (it["token"].shape[1] + it["prompt_token"].shape[1])
        g0 = 0
        while g0 < len(items):
            g1 = g0 + 1
            while g1 < len(items) and g1 - g0 < 16 and (g1 - g0 + 1) * fr(items[g0]) <= 8192 and fr(items[g1]) * 5 >= fr(items[g0]) * 4:
                g1 += 1
            if g1 - g0 == 1:
                it = items[g0]
                mels = [flow.inference(token=it["token"], embedding=it["embedding"], prompt_token=it["prompt_token"],
                                       prompt_feat=it["prompt_feat"], n_timesteps=a.cfm_steps)[0]]
            else:
                mels = flow.inference_batch(items[g0:g1], n_timesteps=a.cfm_steps)
            for mel in mels:
                hift.inference(speech_feat=mel)
            g0 = g1
        return sum(len(t) for t in toks)

    def step_e2e():
        wavs, toks = mm.synthesize_batch(reqs, head_k=a.head_k, sampling=SAMPLING, n_timesteps=a.cfm_steps, min_ratio=a.ratio,
                                         max_ratio=a.ratio, u=u_all, return_tokens=True)
        return sum(len(t) for t in toks), sum(w.shape[1] for w in wavs)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        units = 0
        for s in range(steps):
            flush.zero_()                                               # L2 flush between timed iterations
            ev[s][0].record()
            r = fn()
            ev[s][1].record()
            units += r if isinstance(r, int) else r[0]
        torch.cuda.synchronize()
        return sum(x.elapsed_time(y) for x, y in ev), units

    profiling = os.environ.get("HVX_PROFILE") == "1"       # ncu launch-list runs: one warm-up pass, numbers not reported
    for _ in range(1 if profiling else max(a.warmup, 3)):
        flush.zero_(); step_device()
    for _ in range(0 if profiling else 2):
        flush.zero_(); step_e2e()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    l0 = mm.engine.launches()
    ms_dev, tok_dev = timed(step_device, a.steps)
    launches = mm.engine.launches() - l0
    barrier()
    stage_acc = dict(llm=0.0, flow=0.0, hift=0.0)

    def e2e_once():
        r = step_e2e()
        for k in stage_acc:
            stage_acc[k] += mm.last_stage_ms[k]
        return r
    ms_e2e, tok_e2e = timed(e2e_once, a.steps)
    barrier()
    clk = clocks.stop()
    t = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
    cnt = torch.tensor([tok_dev, tok_e2e, launches], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_dev, ms_e2e = t.tolist()
    tok_dev, tok_e2e, launches = cnt.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = tok_dev / (ms_dev / 1e3)
    e2e = tok_e2e / (ms_e2e / 1e3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    # stage rooflines from the device-timed stage splits of the e2e leg (CUDA events inside hvx_synthesize_host)
    n_steps_llm = a.steps * (n_tok / a.head_k)
    ctx_avg = 2 + 16 + a.n_text + 125 + n_tok / 2
    llm_bytes = llm_step_bytes(ld, a.head_k, ctx_avg, a.batch)
    llm_gbs = llm_bytes * n_steps_llm / (stage_acc["llm"] / 1e3) / 1e9
    T = 2 * (125 + n_tok)
    flow_tf = flow_nfe_flops(fd, T) * a.cfm_steps * a.batch * a.steps / (stage_acc["flow"] / 1e3) / 1e12
    hift_tf = 2 * 336.9e6 * 2 * n_tok * a.batch * a.steps / (stage_acc["hift"] / 1e3) / 1e12
    tot = sum(stage_acc.values())
    # dominant kernel = llm_gemv_tma_kernel (the decode linears' weight stream; share of the step in profiles/): each
    # class of its launches (24 layers back to back, same PDL launch path as the decode graph) is CUDA-event timed on
    # the engine's stream right after the timed region; achieved = bf16 weight bytes of those launches / that time
    import ctypes as C
    ms1 = (C.c_float * 1)()
    H, I, QKV = ld.hidden, ld.inter, (ld.q_heads + 2 * ld.kv_heads) * ld.head_dim
    classes = {"qkv": (1, 2.0 * QKV * H), "o_proj": (3, 2.0 * H * H), "gate_up": (4, 2.0 * 2 * I * H), "down": (5, 2.0 * H * I)}
    kern, kb, kt = {}, 0.0, 0.0
    rows = a.batch * a.head_k
    if rows <= 8:            # the single-pass weight-streaming GEMV step (above 8 rows the step runs on tcgen05 GEMMs)
        for name, (which, nbytes) in classes.items():
            L.check(L.lib().hvx_llm_bench_kernels(mm.engine.h, a.batch, a.head_k, int(ctx_avg), which, 10, ms1))
            us = ms1[0] * 1e3 / ld.layers
            kern[name] = {"us_per_launch": us, "bytes_per_launch": nbytes, "gbs": nbytes / us / 1e3}
            kb += nbytes * ld.layers
            kt += ms1[0] * 1e-3
        L.check(L.lib().hvx_llm_bench_kernels(mm.engine.h, a.batch, a.head_k, int(ctx_avg), 2, 10, ms1))
        kern["kv_attention"] = {"us_per_launch": ms1[0] * 1e3 / ld.layers}
        L.check(L.lib().hvx_llm_bench_kernels(mm.engine.h, a.batch, a.head_k, int(ctx_avg), 0, 10, ms1))
        kern["whole_decode_step_no_sampler_us"] = ms1[0] * 1e3
    gemv_gbs = kb / kt / 1e9 if kt else None
    roof = {"bound": "hbm", "achieved": gemv_gbs if gemv_gbs else llm_gbs, "peak": hbm_peak, "unit": "GB/s",
            "frac": (gemv_gbs if gemv_gbs else llm_gbs) / hbm_peak, "traffic": 8.72e6 + 3.4e6,
            "kernel": "llm_gemv_tma_kernel<R> (decode linears): bf16 weight bytes of the 96 per-step launches / their CUDA-event time, "
                      "measured per class back to back on the engine stream after the timed region; traffic = dram read+write of the "
                      "gate_up launch from profiles/r1_ncu_full_summary.txt (algorithmic 8.72 MB)",
            "peak_source": peak_src, "per_class": kern,
            "stages": {"llm": {"share": stage_acc["llm"] / tot, "bound": "hbm", "achieved_gbs": llm_gbs, "frac": llm_gbs / hbm_peak,
                               "note": "algorithmic bytes per decode step x steps / device time of the whole LLM stage (prefill, sampler, launch gaps included)"},
                       "flow": {"share": stage_acc["flow"] / tot, "bound": "tensor", "achieved_tflops": flow_tf, "frac": flow_tf / tf_peak},
                       "hift": {"share": stage_acc["hift"] / tot, "bound": "fp32 cuda cores (algorithmic conv flops)", "achieved_tflops": hift_tf}}}
    line = {"metric": "speech_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(a), "l2": "256 MiB flush between timed steps; per-step weight traffic (2.3 GB) exceeds L2",
                       "precision": "llm: bf16 weights + bf16 KV cache, fp32 activations/accumulate; flow: fp16 operands, fp32 accumulate/state; hift: fp32",
                       "utterances_per_gpu": a.batch},
            "rtf": 25.0 / value if value else None, "rtf_per_gpu": 25.0 * world / value if value else None,
            "e2e": {"value": e2e, "unit": "tokens/s", "h2d_bytes_per_step": mm.h2d_bytes, "d2h_bytes_per_step": mm.d2h_bytes,
                    "ms_per_step": ms_e2e / a.steps, "rtf": 25.0 * world / e2e if e2e else None,
                    "stage_ms_per_step": {k: v / a.steps for k, v in stage_acc.items()}},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof}
    # the same step in parity mode (the mode whose mel error is <= 1e-3 against the fp32 reference, tests/test_flow_gpu.py)
    if a.parity_mode_steps > 0 and world == 1:
        del flush
        torch.cuda.empty_cache()
        pm = ModelManager(hd=hd, fd=fd, ld=ld, device=f"cuda:{local}", max_ctx=max_ctx, max_seqs=max(a.batch, 1), n_timesteps=a.cfm_steps,
                          sine_seconds=n_tok / 25 + 2, kv_f32=True, flow_precise=True)
        pm.load_state_dicts(synth.llm_state_dict(ld, 0, dtype=torch.bfloat16, eos_scale=0.0), synth.flow_state_dict(fd, 0),
                            synth.hift_state_dict(hd, 0))
        def pm_step():
            w, t = pm.synthesize_batch(reqs, head_k=a.head_k, sampling=SAMPLING, n_timesteps=a.cfm_steps, min_ratio=a.ratio,
                                       max_ratio=a.ratio, u=u_all, return_tokens=True)
            return sum(len(x) for x in t)
        pm_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_pm = sum(pm_step() for _ in range(a.parity_mode_steps))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        line["parity_mode"] = {"e2e_value": n_pm / dt, "unit": "tokens/s", "rtf": 25.0 * dt / n_pm, "steps": a.parity_mode_steps,
                               "stage_ms": dict(pm.last_stage_ms),
                               "what": "fp32 KV cache + three-term split-fp16 flow GEMMs (mel max-abs 2e-4 vs the fp32 reference); "
                                       "the headline numbers use the serving mode (bf16 KV cache, fp16 flow operands = reference precision)"}
        pm.engine.close()
        del pm
    # BASELINE config 5: first-audio latency of the streaming path (AR decode overlapped with chunked flow + vocoder)
    if a.first_audio_runs > 0:
        from flowmirror_hydravox_b200.streaming import StreamingSynthesizer
        ss = StreamingSynthesizer(mm)
        rq = synth.utterance(ld, fd, 64, seed=77)
        lat, tot_ms = [], []
        for i in range(a.first_audio_runs + 2):
            dbg = {}
            t0 = time.perf_counter()
            n_s = sum(c["tts_speech"].shape[1] for c in ss.tts(rq, head_k=2, sampling=SAMPLING, n_timesteps=10, min_ratio=a.ratio,
                                                                 max_ratio=a.ratio, debug=dbg))
            if i >= 2:
                lat.append(dbg["first_audio_ms"]); tot_ms.append((time.perf_counter() - t0) * 1e3)
        lat.sort(); tot_ms.sort()
        line["first_audio"] = {"p50_ms": lat[len(lat) // 2], "min_ms": lat[0], "max_ms": lat[-1], "runs": len(lat),
                               "total_ms_p50": tot_ms[len(tot_ms) // 2], "audio_s": n_s / hd.sr,
                               "config": "batch=1, 16+64 text tokens, 125-token prompt, inference_head_num=2, 10 CFM steps, first chunk = "
                                         "25+3 tokens; request -> first waveform chunk on the host (wall clock)"}
    # SURVEY 8 a7' / a12': the alternate estimator (U-Net) and the transposed-conv vocoders on the same frame count, CUDA events
    if not a.no_variants and world == 1:
        line["variants"] = variants(mm.engine.device, 2 * (n_tok + 125), a.cfm_steps, tf_peak)
    if not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        r = cpu_oracle_run(a, a.cpu_tokens, threads)
        line["cpu_baseline"] = {"value": r["tokens_per_s"], "unit": "tokens/s", "cores": threads, "kind": "port", "sample": r["sample"],
                                "stage_s": r["stage_s"], "rtf": r["rtf"]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def variants(device, T, cfm_steps, tf_peak):
    """CUDA-event timings of hvx_cfm_solve_unet, hvx_hifigan_vocode and hvx_hift_t_vocode at T mel frames (inputs resident)."""
    import torch
    from flowmirror_hydravox_b200 import _lib as L, dims as D, synth
    from flowmirror_hydravox_b200.flow import NativeUNetCFM
    from flowmirror_hydravox_b200.hift import NativeHiFiGAN, NativeHiFTTransposed

    def timed(fn, reps=3):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = {}
    g = torch.Generator().manual_seed(5)
    ud = D.UNET_FULL
    C, inner, ff = ud.ch, ud.heads * ud.head_dim, ud.ch * ud.ff_mult
    res = lambda cin: 3 * cin * C + 3 * C * C + cin * C
    mac = res(ud.in_ch) + ud.n_mid * res(C) + res(2 * C) + ud.n_res * ud.n_blocks * (C * 3 * inner + inner * C + 2 * C * ff) + 9 * C * C + C * ud.mel
    flops = (2 * (2 * T) * mac + 2 * ud.n_res * ud.n_blocks * 4 * T * T * inner) * cfm_steps
    e = L.Engine(ud=ud, device=str(device))
    cfm = NativeUNetCFM(e); cfm.load_state_dict(synth.unet_state_dict(ud, 0))
    mu, cond = (torch.randn(1, ud.mel, T, generator=g).to(device) for _ in range(2))
    spk = torch.randn(1, ud.mel, generator=g).to(device)
    ms = timed(lambda: cfm(mu, None, cfm_steps, spks=spk, cond=cond))
    tf = flops / ms / 1e9
    out["unet_cfm_solve"] = {"ms": ms, "frames": T, "euler_steps": cfm_steps, "tflops": tf, "frac_of_bf16_peak": tf / tf_peak,
                             "what": "hvx_cfm_solve_unet: CausalConditionalCFM.forward over the U-Net estimator (71.3 M params), fp16 operands"}
    e.close()
    for name, hd, cls, sdf in (("hifigan_v1", D.HIFIGAN_V1, NativeHiFiGAN, synth.hifigan_state_dict),
                               ("hift_transposed", D.HIFT_FULL, NativeHiFTTransposed, synth.hift_t_state_dict)):
        e = L.Engine(hd=hd, device=str(device))
        v = cls(e); v.load_state_dict(sdf(hd, 0))
        mel = (torch.rand(1, hd.mel, T, generator=g) * 6 - 6).to(device)
        ms = timed((lambda: v(mel)) if cls is NativeHiFiGAN else (lambda: v.inference(mel)))
        audio_s = T * hd.frame_samples / hd.sr
        out[name] = {"ms": ms, "frames": T, "audio_s": audio_s, "rtf": ms / 1e3 / audio_s}
        e.close()
    return out


def main():
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        native_arm(a)


if __name__ == "__main__":
    main()
