#!/usr/bin/env python
"""bench.py — HydraVox hot path (multi-head AR decode -> CFM Euler -> HiFT) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|b8|c4] [--mode parity|serving]
  python bench.py --impl reference ...      # the reference algorithm's CPU path (oracle port) on the host cores

A step = one pass of the whole hot path over one batch of synthetic utterances.  Default workload = BASELINE.json
configs[2] ("c3"): 32 zero-shot utterances per GPU of MIXED length — n_text ~ U{64..512} text tokens (seed 1986), each with a
16-token prompt text, 125 prompt speech tokens and 250 prompt mel frames — inference_head_num=4, 25 CFM Euler steps,
fixed-length protocol min = max ratio 8: 512..4096 speech tokens = 20..164 s of 24 kHz audio per utterance.
Other workloads: c2 = configs[1] (one 16+128-token utterance, K=2), b8 = 8 mixed-length utterances, c4 = configs[3] shape
(32 chunks per GPU, n_text ~ U{12..40}, K=2).  Metric: speech tokens per second through the whole path (RTF = 25 / value).

  value   inputs resident in HBM, stage calls through the C-ABI with device pointers, CUDA-event timed
  e2e     the public call with HOST buffers.  1 GPU: ModelManager.synthesize_batch -> hvx_synthesize_host (H2D of the requests,
          three stages, D2H of tokens + waveforms inside the timed region).  N GPUs: rank 0 owns the N*batch requests;
          parallel.synthesize_sharded deals them (one packed NCCL scatter), every rank synthesises its shard, the waveforms
          return to rank 0 (one NCCL gather) — all inside the timed region, scatter / gather milliseconds reported.
Both legs run in the mode whose parity tests/ assert (default `parity`: fp32 KV cache, three-term split-fp16 flow GEMMs:
mel <= 1e-3 max-abs against the fp32 reference); `serving_mode` reports the reference's own serving precision next to it.
Baselines on the same line (never the product path): `cpu_baseline` = the oracle port of the reference algorithm on the host cores,
`gpu_eager_baseline` = the same eager PyTorch code with its tensors on the B200 (SURVEY 8d's GPU comparator), one utterance each.
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
RATIO = 8.0                                                         # speech tokens per text token, min = max
P_TOK, P_TEXT = 125, 16                                             # 5 s zero-shot prompt
WORKLOADS = {   # name: (utterances per GPU, (n_text lo, hi), inference_head_num, BASELINE.json config it realises)
    "c3": (32, (64, 512), 4, "configs[2]"),
    "c2": (1, (128, 128), 2, "configs[1]"),
    "b8": (8, (64, 512), 4, "configs[2] at batch 8"),
    "c4": (32, (12, 40), 2, "configs[3] (32 text chunks per GPU)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="parity", choices=["parity", "serving"])
    ap.add_argument("--batch", type=int, default=0, help="override utterances per GPU")
    ap.add_argument("--head-k", type=int, default=0, help="override inference_head_num")
    ap.add_argument("--n-text", type=int, default=0, help="override the text length of every utterance (profiling runs: short utterances)")
    ap.add_argument("--cfm-steps", type=int, default=25)
    ap.add_argument("--e2e-steps", type=int, default=3, help="timed steps of the end-to-end leg (<= --steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip serving-mode / batch-1 / first-audio side measurements")
    ap.add_argument("--first-audio-runs", type=int, default=10, help="streaming first-audio latency samples (0 = skip)")
    ap.add_argument("--variants", action="store_true", help="also time the alternate estimator / vocoders (SURVEY 8 a7', a12')")
    ap.add_argument("--cpu-tokens", type=int, default=128, help="speech tokens of the CPU sample utterance (~20-30 s of CPU work)")
    ap.add_argument("--wire", default="f32", choices=["f32", "s16"], help="waveform gather format of the sharded e2e leg")
    return ap.parse_args()


def wl(a):
    b, rng, k, cfg = WORKLOADS[a.workload]
    if getattr(a, "n_text", 0):
        rng, cfg = (a.n_text, a.n_text), cfg + f" with n_text overridden to {a.n_text}"
    return (a.batch or b), rng, (a.head_k or k), cfg


def workload_name(a, world=1):
    b, (lo, hi), k, cfg = wl(a)
    length = f"{lo}" if lo == hi else f"U{{{lo}..{hi}}} (seed 1986)"
    return (f"BASELINE {cfg}: {b} zero-shot utterances per GPU x {world} GPU, n_text = {length} text tokens + {P_TEXT} prompt text tokens, "
            f"{P_TOK} prompt speech tokens (+{2 * P_TOK} prompt mel frames), inference_head_num={k}, {a.cfm_steps} CFM Euler steps, "
            f"min=max token/text ratio {RATIO:g} -> {int(lo * RATIO)}..{int(hi * RATIO)} speech tokens per utterance")


def make_requests(a, n_total):
    """the synthetic batch: request i has n_text drawn from U{lo..hi} by a generator seeded 1986, contents seeded 1986+1+i"""
    _, (lo, hi), _, _ = wl(a)
    g = torch.Generator().manual_seed(1986)
    n_texts = torch.randint(lo, hi + 1, (n_total,), generator=g).tolist()
    reqs = []
    for i, n in enumerate(n_texts):
        r = synth.utterance(D.LLM_FULL, D.FLOW_FULL, n, seed=1987 + i, prompt_tokens=P_TOK, prompt_text=P_TEXT)
        r["min_ratio"] = r["max_ratio"] = RATIO
        r["u"] = torch.rand(4 * int(n * RATIO) + 1024, generator=torch.Generator().manual_seed(7000 + i))
        reqs.append(r)
    return reqs


# --------------------------------------------------------------------------- algorithmic work
def llm_step_bytes(ld: D.LlmDims, head_k: int, ctx: float, n_seq: float) -> float:
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
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU arm (oracle port)
_CPU_SDS = {}


def cpu_oracle_run(a, n_tokens: int, threads: int, device: str = "cpu"):
    """The reference algorithm restated in eager PyTorch (oracle/, fp32) on a BOUNDED sample of the workload: ONE zero-shot
    utterance of the workload's shape — same prompt (16 prompt text tokens, 125 prompt speech tokens, 250 prompt mel frames), same
    inference_head_num, same number of CFM steps, full model dims — whose text is shortened so the fixed-length protocol emits
    n_tokens speech tokens.  Throughput per token on the CPU does not improve with length (attention grows with T^2), so the
    short sample flatters the CPU arm.  device="cuda:i" runs the very same eager code with its tensors on the GPU: the
    `gpu_eager_baseline` comparator of SURVEY 8(d) (a baseline leg like the CPU one — never the product path)."""
    from oracle import flow_ref, hift_ref, llm_ref
    torch.set_num_threads(threads)
    ld, fd, hd = D.LLM_FULL, D.FLOW_FULL, D.HIFT_FULL
    _, _, head_k, _ = wl(a)
    n_text = max(1, int(round(n_tokens / RATIO)))
    if not _CPU_SDS:
        _CPU_SDS.update(llm=synth.llm_state_dict(ld, 0, eos_scale=0.0), flow=synth.flow_state_dict(fd, 0), hift=synth.hift_state_dict(hd, 0))
    on_gpu = device != "cpu"
    key = lambda k: k + "@" + device if on_gpu else k
    if on_gpu and key("llm") not in _CPU_SDS:
        for k in ("llm", "flow", "hift"):
            _CPU_SDS[key(k)] = {n: v.float().to(device) for n, v in _CPU_SDS[k].items()}
    llm_sd, flow_sd, hift_sd = _CPU_SDS[key("llm")], _CPU_SDS[key("flow")], _CPU_SDS[key("hift")]
    u = synth.utterance(ld, fd, n_text, seed=1987, prompt_tokens=P_TOK, prompt_text=P_TEXT)
    u = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in u.items()}
    us = torch.rand(8 * n_tokens + 64, generator=torch.Generator().manual_seed(7))
    noise = synth.flow_noise(fd).to(device)
    sync = (lambda: torch.cuda.synchronize(device)) if on_gpu else (lambda: None)
    sync()
    t0 = time.perf_counter()
    ctx = lambda: torch.device(device)                   # the oracle's factory calls (arange, zeros, ...) follow the context
    with torch.no_grad():
        with ctx():
            toks = llm_ref.inference(llm_sd, ld, u["text"], u["prompt_text"], u["prompt_speech"], us, head_k=head_k, sp=SAMPLING,
                                     min_ratio=RATIO, max_ratio=RATIO)
            sync(); t1 = time.perf_counter()
            mel = flow_ref.inference(flow_sd, torch.tensor(toks)[None], u["embedding"][None], noise, fd, a.cfm_steps,
                                     u["prompt_speech"][None].long(), u["prompt_feat"][None])
            sync(); t2 = time.perf_counter()
        table = synth.hift_sine_table(hd, mel.shape[2]).to(device)
        with ctx():
            wav, _ = hift_ref.inference(hift_sd, mel, table, hd)
        wav = wav.cpu()
        t3 = time.perf_counter()
    total = t3 - t0
    where = (f"eager PyTorch fp32 on {torch.cuda.get_device_name(device)} (TF32 matmul/conv "
             f"{'on' if torch.backends.cuda.matmul.allow_tf32 else 'off'})") if on_gpu else f"{threads} threads"
    return dict(tokens=len(toks), seconds=total, stage_s=dict(llm=t1 - t0, flow=t2 - t1, hift=t3 - t2),
                tokens_per_s=len(toks) / total, rtf=total / (wav.shape[1] / hd.sr),
                sample=f"1 zero-shot utterance of the workload's shape: {n_text}+{P_TEXT} text tokens, {P_TOK} prompt speech tokens (+{2 * P_TOK} prompt "
                       f"mel frames), inference_head_num={head_k}, {a.cfm_steps} CFM steps, {len(toks)} speech tokens ({len(toks) / 25:.2f} s of audio), "
                       f"full model dims, fp32 oracle port of the reference algorithm (KV-cached decode, i.e. faster than the reference's "
                       f"own no-cache loop), {where}")


def gpu_eager_baseline(a, device: str, n_tokens: int):
    """SURVEY 8(d)'s honest GPU comparator: the reference algorithm as eager PyTorch kernels (cuBLAS / cuDNN / ATen; the
    oracle port with its tensors on the B200, batch 1 like the reference) on one utterance of the workload's shape.  TF32 is
    switched on for the baseline's matmuls and convolutions (the reference serves in bf16 / fp16; fp32-without-TF32 would
    flatter the engine).  Any failure is reported, never raised: this leg must not cost the bench its line."""
    prev = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
        cpu_oracle_run(a, 32, 1, device)                                  # warm-up: cuBLAS / cuDNN handles and heuristics
        r = cpu_oracle_run(a, n_tokens, os.cpu_count() or 1, device)
        return {"value": r["tokens_per_s"], "unit": "tokens/s", "rtf": r["rtf"], "stage_s": r["stage_s"], "sample": r["sample"],
                "kind": "oracle port of the reference algorithm, eager PyTorch, batch 1 (the reference has no batched path)"}
    except Exception as e:                                                # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev
        for k in [k for k in _CPU_SDS if "@" in k]:
            del _CPU_SDS[k]
        torch.cuda.empty_cache()


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
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a, max(a.gpus, 1)), "timed_sample": res[0]["sample"],
                       "note": "the CPU arm times a bounded sample (one short utterance per step) of the same workload shape; tokens/s is "
                               "per host, not per utterance, so it compares with the GPU arm's whole-batch tokens/s"},
            "rtf": 25.0 / tps, "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": res[0]["sample"],
                                                "stage_s": res[0]["stage_s"]},
            "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------- native arm
def flow_groups(frames, limit=8192, max_u=16):
    """the grouping rule of hvx_synthesize_host (csrc/pipeline.cu): longest first, <= limit padded frames, <= 20 % padding"""
    order = sorted(range(len(frames)), key=lambda i: -frames[i])
    groups, g0 = [], 0
    while g0 < len(order):
        g1 = g0 + 1
        while g1 < len(order) and g1 - g0 < max_u and (g1 - g0 + 1) * frames[order[g0]] <= limit and frames[order[g1]] * 5 >= frames[order[g0]] * 4:
            g1 += 1
        groups.append(order[g0:g1])
        g0 = g1
    return groups


def native_arm(a):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from flowmirror_hydravox_b200 import parallel
    from flowmirror_hydravox_b200.model_manager import ModelManager

    ld, fd, hd = D.LLM_FULL, D.FLOW_FULL, D.HIFT_FULL
    batch, (lo, hi), head_k, _ = wl(a)
    max_tok = int(hi * RATIO)
    max_ctx = 2 + P_TEXT + hi + P_TOK + max_tok + 64
    parity = a.mode == "parity"

    def build_mm(parity_mode, seqs):
        mm_ = ModelManager(hd=hd, fd=fd, ld=ld, device=f"cuda:{local}", max_ctx=max_ctx, max_seqs=max(seqs, 1), n_timesteps=a.cfm_steps,
                           sine_seconds=max_tok / 25 + 2, kv_f32=parity_mode, flow_precise=parity_mode)
        mm_.load_state_dicts(synth.llm_state_dict(ld, 0, dtype=torch.bfloat16, eos_scale=0.0), synth.flow_state_dict(fd, 0),
                             synth.hift_state_dict(hd, 0))
        return mm_

    # the whole job's requests; every rank derives the same list and the same LPT deal, so the device-resident leg runs on
    # exactly the shards the sharded e2e leg scatters
    all_reqs = make_requests(a, batch * world)
    shards = parallel.shard_requests(all_reqs, world)
    reqs = [all_reqs[i] for i in shards[rank]]
    mm = build_mm(parity, len(reqs))
    llm, flow, hift = mm.models["llm"], mm.models["flow"], mm.models["hift"]
    u_all = torch.zeros(len(reqs), 4 * max_tok + 1024)
    for i, r in enumerate(reqs):
        u_all[i, : r["u"].numel()] = r["u"]
    dreqs = [{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in r.items() if k != "u"} for r in reqs]
    u_dev = u_all.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)        # > L2 (126 MB)

    def step_device():
        toks = llm.generate_batch(dreqs, head_k=head_k, u=u_dev, sampling=SAMPLING, min_ratio=RATIO, max_ratio=RATIO)
        items = [dict(token=torch.tensor(t, device=dev, dtype=torch.int32)[None], embedding=r["embedding"][None],
                      prompt_token=r["prompt_speech"][None], prompt_feat=r["prompt_feat"][None]) for r, t in zip(dreqs, toks)]
        frames = [2 * (it["token"].shape[1] + it["prompt_token"].shape[1]) for it in items]
        for grp in flow_groups(frames):
            if len(grp) == 1:
                it = items[grp[0]]
                mels = [flow.inference(token=it["token"], embedding=it["embedding"], prompt_token=it["prompt_token"],
                                       prompt_feat=it["prompt_feat"], n_timesteps=a.cfm_steps)[0]]
            else:
                mels = flow.inference_batch([items[i] for i in grp], n_timesteps=a.cfm_steps)
            for mel in mels:
                hift.inference(speech_feat=mel)
        return sum(len(t) for t in toks)

    frame = hd.frame_samples
    e2e_stats = {}

    def step_e2e():
        if world == 1:
            wavs, toks = mm.synthesize_batch(reqs, head_k=head_k, sampling=SAMPLING, n_timesteps=a.cfm_steps, min_ratio=RATIO,
                                             max_ratio=RATIO, u=u_all, return_tokens=True)
            return sum(len(t) for t in toks)

        def synth_fn(mine):
            u = torch.zeros(len(mine), 4 * max_tok + 1024)
            for i, r in enumerate(mine):
                u[i, : r["u"].numel()] = r["u"]
            return mm.synthesize_batch(mine, head_k=head_k, sampling=SAMPLING, n_timesteps=a.cfm_steps, min_ratio=RATIO, max_ratio=RATIO, u=u)
        st = {}
        out = parallel.synthesize_sharded(synth_fn, all_reqs if rank == 0 else None, wire=a.wire, stats=st)
        for k, v in st.items():
            e2e_stats[k] = e2e_stats.get(k, 0.0) + v
        return sum(w.shape[1] for w in out) // (2 * frame) if rank == 0 else 0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    prof_tot = {}

    def timed(fn, steps, profile=False):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        units = 0
        for s in range(steps):
            flush.zero_()                                               # L2 flush between timed iterations
            ev[s][0].record()
            units += fn()
            ev[s][1].record()
            if profile:                                                 # synchronises AFTER the step's closing event
                for k, (ms, work, n) in mm.engine.profile_collect().items():
                    t = prof_tot.setdefault(k, [0.0, 0.0, 0])
                    t[0] += ms; t[1] += work; t[2] += n
        torch.cuda.synchronize()
        return sum(x.elapsed_time(y) for x, y in ev), units

    profiling = os.environ.get("HVX_PROFILE") == "1"       # ncu launch-list runs: one warm-up pass, numbers not reported
    n_warm = 1 if profiling else max(a.warmup, 3)
    for _ in range(n_warm):
        flush.zero_(); step_device()
    for _ in range(0 if profiling else 1):
        flush.zero_(); step_e2e()
    e2e_stats.clear()
    clocks = ClockSampler(local)
    barrier()
    clocks.start()
    mm.engine.profile(True)
    mm.engine.profile_collect()
    l0 = mm.engine.launches()
    ms_dev, tok_dev = timed(step_device, a.steps, profile=True)
    launches = mm.engine.launches() - l0
    mm.engine.profile(False)
    barrier()
    stage_acc = dict(llm=0.0, flow=0.0, hift=0.0)
    n_e2e = max(1, min(a.e2e_steps, a.steps))

    def e2e_once():
        r = step_e2e()
        for k in stage_acc:
            stage_acc[k] += mm.last_stage_ms[k]
        return r
    ms_e2e, tok_e2e = timed(e2e_once, n_e2e)
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
            dist.barrier()
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
    peak_src = "measured (MEASURED_PEAKS.json, bf16_tflops_sustained / hbm_gbs)" if peaks else "fallback (B200_PROFILING.md)"

    # ---- roofline: per kernel class, CUDA events around every launch of the class on its own stream DURING the timed steps
    n_tok_rank = [int(r["text"].numel() * RATIO) for r in reqs]
    steps_llm = a.steps * (max(n_tok_rank) / head_k)                      # the longest sequence decides the number of decode steps
    live_seq = sum(n_tok_rank) / max(n_tok_rank)                          # sequences alive per step on average
    ctx_avg = 2 + P_TEXT + (lo + hi) / 2 + P_TOK + sum(n_tok_rank) / len(n_tok_rank) / 2
    per_class = {}
    for k, (ms, work, n) in prof_tot.items():
        if n == 0:
            continue
        d = {"ms_per_step": ms / a.steps, "call_sites_per_step": n / a.steps, "share_of_step": ms / ms_dev}
        if k in ("gemm", "attention", "hift_conv"):
            tf = work / ms / 1e9 if ms else 0.0
            d.update(bound="tensor", algorithmic_tflop_per_step=work / a.steps / 1e12, achieved_tflops=tf, frac=tf / tf_peak)
            if k == "gemm" and parity:
                d["tensor_pipe_tflops"] = 3.0 * tf       # three split-fp16 products per algorithmic product
        elif k == "layernorm":
            gbs = work / ms / 1e6 if ms else 0.0
            d.update(bound="hbm", achieved_gbs=gbs, frac=gbs / hbm_peak)
        elif k == "llm_step":
            byt = llm_step_bytes(ld, head_k, ctx_avg, live_seq)
            gbs = byt * work / ms / 1e6 if ms else 0.0
            d.update(bound="hbm", steps=work / a.steps, us_per_step=1e3 * ms / work if work else None, bytes_per_step=byt,
                     achieved_gbs=gbs, frac=gbs / hbm_peak)
        per_class[k] = d
    dom = max((k for k in per_class if "frac" in per_class[k]), key=lambda k: per_class[k]["ms_per_step"], default=None)
    kernel_names = {"gemm": "gemm_persist_kernel / gemm_bf16_kernel (tcgen05 + TMA; DiT linears, 3 split-fp16 products each in parity mode)",
                    "attention": "dit_attention_v5_kernel (tcgen05; S, P and O in TMEM)", "hift_conv": "HiFT convolutions",
                    "llm_step": "decode step (CUDA graph: tcgen05 GEMMs above 8 rows, TMA GEMV below)", "layernorm": "dit_ln_mod_kernel"}
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
        traffic = tr.get(dom, {}).get("dram_bytes_per_launch")
    except Exception:
        tr = {}
    roof = None
    if dom:
        dc = per_class[dom]
        tensor = dc["bound"] == "tensor"
        roof = {"bound": dc["bound"], "achieved": dc["achieved_tflops"] if tensor else dc["achieved_gbs"], "peak": tf_peak if tensor else hbm_peak,
                "unit": "TFLOP/s" if tensor else "GB/s", "frac": dc["frac"], "traffic": traffic,
                "kernel": f"{kernel_names.get(dom, dom)}: dominant kernel class of the timed steps ({100 * dc['share_of_step']:.0f} % of the step); achieved = "
                          f"algorithmic work of its launches / summed CUDA-event duration of those launches, events recorded around every launch "
                          f"on the launching stream inside the timed region",
                "traffic_source": (f"{tr[dom]['source']}; one representative launch: {tr[dom]['kernel']}, algorithmic "
                                   f"{tr[dom]['algorithmic_bytes_per_launch']} B") if traffic else None,
                "peak_source": peak_src, "per_class": per_class}
    tot = sum(stage_acc.values()) or 1.0
    line = {"metric": "speech_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world, "steps": a.steps, "warmup": n_warm,
            "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(a, world), "mode": a.mode,
                       "l2": "256 MiB flush between timed steps; per-step weight + activation traffic (> 2 GB) exceeds L2",
                       "precision": ("llm: bf16 weights, fp32 KV cache, fp32 activations/accumulate; flow: three-term split-fp16 tensor-core products "
                                     "(A_hi W_hi + A_lo W_hi + A_hi W_lo), fp32 accumulate/state — the mode tests/test_flow_gpu.py::"
                                     "test_flow_parity_mode_meets_north_star and tests/test_c2_gpu.py::test_flow_c2_size_parity assert at <= 1e-3 max-abs on mel; hift: fp32"
                                     if parity else
                                     "llm: bf16 weights + bf16 KV cache; flow: fp16 operands (the reference's serving precision), fp32 accumulate/state; hift: fp32"),
                       "utterances_per_gpu": batch, "tokens_per_step": tok_dev / a.steps,
                       "legs": "value: the three stage calls through the C-ABI with device-resident inputs, flow groups as the pipeline forms them; "
                               "e2e: one hvx_synthesize_host call with host buffers (H2D requests, decode, flow + vocoder groups, D2H waveforms)"},
            "rtf": 25.0 / value if value else None, "rtf_per_gpu": 25.0 * world / value if value else None,
            "e2e": {"value": e2e, "unit": "tokens/s", "steps": n_e2e, "ms_per_step": ms_e2e / n_e2e, "rtf": 25.0 / e2e if e2e else None,
                    "h2d_bytes_per_step": mm.h2d_bytes * (world if world > 1 else 1), "d2h_bytes_per_step": mm.d2h_bytes * (world if world > 1 else 1),
                    "stage_ms_per_step_rank0": {k: v / n_e2e for k, v in stage_acc.items()},
                    "stage_share": {k: v / tot for k, v in stage_acc.items()},
                    "api": "ModelManager.synthesize_batch -> hvx_synthesize_host" if world == 1 else
                           "parallel.synthesize_sharded(ModelManager.synthesize_batch): rank 0 host requests -> packed NCCL scatter -> per-rank synthesis -> NCCL gather -> rank 0 host waveforms"},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roof}
    if world > 1:
        line["e2e"]["collective_ms_per_step_rank0"] = {k: e2e_stats.get(k, 0.0) / n_e2e for k in ("scatter_ms", "synth_ms", "gather_ms")}
        line["e2e"]["note"] = "bytes are rank 0's shard x N (every rank copies its shard H2D and its waveforms D2H; rank 0 additionally gathers all waveforms)"
        line["e2e"]["wire"] = a.wire
    # ---- side measurements (bounded; none of them feeds `value`)
    if not a.no_extras and world == 1:
        del flush
        torch.cuda.empty_cache()
        other = build_mm(not parity, len(reqs))

        def other_step():
            w, t_ = other.synthesize_batch(reqs, head_k=head_k, sampling=SAMPLING, n_timesteps=a.cfm_steps, min_ratio=RATIO, max_ratio=RATIO,
                                           u=u_all, return_tokens=True)
            return sum(len(x) for x in t_)
        other_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_o = other_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        line["serving_mode" if parity else "parity_mode"] = {
            "e2e_value": n_o / dt, "unit": "tokens/s", "rtf": 25.0 * dt / n_o, "steps": 1, "stage_ms": dict(other.last_stage_ms),
            "what": ("bf16 KV cache + single-product fp16 flow GEMMs = the reference's own serving precision (mel 2-3e-3 max-abs vs the fp32 reference)"
                     if parity else "fp32 KV cache + three-term split-fp16 flow GEMMs (mel <= 1e-3 max-abs vs the fp32 reference)")}
        # BASELINE configs[1]: one zero-shot utterance, 16+128 text tokens, inference_head_num=2 (the batch-1 latency case)
        r1 = synth.utterance(ld, fd, 128, seed=1986, prompt_tokens=P_TOK, prompt_text=P_TEXT)
        u1 = torch.rand(1, 4 * 1024 + 1024, generator=torch.Generator().manual_seed(0))
        for m_, key in ((mm, "batch1_c2"), (other, "batch1_c2_other_mode")):
            f1 = lambda: m_.synthesize_batch([r1], head_k=2, sampling=SAMPLING, n_timesteps=a.cfm_steps, min_ratio=RATIO, max_ratio=RATIO, u=u1,
                                             return_tokens=True)
            f1(); f1()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n1 = sum(len(f1()[1][0]) for _ in range(3))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            line[key] = {"e2e_value": n1 / dt, "unit": "tokens/s", "rtf": 25.0 * dt / n1, "steps": 3, "stage_ms": dict(m_.last_stage_ms),
                         "mode": a.mode if m_ is mm else ("serving" if parity else "parity"),
                         "workload": "BASELINE configs[1]: 1 zero-shot utterance, 16+128 text tokens, inference_head_num=2, 25 CFM steps, 1024 speech tokens"}
        other.engine.close()
        del other
    # BASELINE configs[4]: first-audio latency of the streaming path (AR decode overlapped with chunked flow + vocoder)
    if a.first_audio_runs > 0 and not a.no_extras and world == 1:
        from flowmirror_hydravox_b200.streaming import StreamingSynthesizer
        ss = StreamingSynthesizer(mm)
        rq = synth.utterance(ld, fd, 64, seed=77)
        lat, tot_ms = [], []
        for i in range(a.first_audio_runs + 2):
            dbg = {}
            t0 = time.perf_counter()
            n_s = sum(c["tts_speech"].shape[1] for c in ss.tts(rq, head_k=2, sampling=SAMPLING, n_timesteps=10, min_ratio=RATIO, max_ratio=RATIO, debug=dbg))
            if i >= 2:
                lat.append(dbg["first_audio_ms"]); tot_ms.append((time.perf_counter() - t0) * 1e3)
        lat.sort(); tot_ms.sort()
        # the same request with the incremental flow session (hvx_flow_stream_*: only the new 50-frame chunk of every non-final hop
        # is evaluated; bit-identical mel): total time of the stream
        ss_inc = StreamingSynthesizer(mm, incremental=True)
        tot_inc = []
        for i in range(3):
            t0 = time.perf_counter()
            for _c in ss_inc.tts(rq, head_k=2, sampling=SAMPLING, n_timesteps=10, min_ratio=RATIO, max_ratio=RATIO):
                pass
            tot_inc.append((time.perf_counter() - t0) * 1e3)
        line["first_audio"] = {"p50_ms": lat[len(lat) // 2], "min_ms": lat[0], "max_ms": lat[-1], "runs": len(lat),
                               "total_ms_p50": tot_ms[len(tot_ms) // 2], "total_ms_incremental_flow": sorted(tot_inc)[1],
                               "audio_s": n_s / hd.sr, "mode": a.mode,
                               "config": "BASELINE configs[4]: batch=1, 16+64 text tokens, 125-token prompt, inference_head_num=2, 10 CFM steps, "
                                         "first chunk = 25+3 tokens; request -> first waveform chunk on the host (wall clock)"}
    if a.variants and world == 1:
        line["variants"] = variants(mm.engine.device, 2 * (1024 + P_TOK), a.cfm_steps, tf_peak)
    if not a.no_extras and world == 1:
        line["gpu_eager_baseline"] = gpu_eager_baseline(a, f"cuda:{local}", 1024)
    if not a.no_cpu_baseline and world == 1:                 # rank 0 at N = 1 only
        threads = os.cpu_count() or 1
        r = cpu_oracle_run(a, a.cpu_tokens, threads)
        line["cpu_baseline"] = {"value": r["tokens_per_s"], "unit": "tokens/s", "cores": threads, "kind": "port", "sample": r["sample"],
                                "stage_s": r["stage_s"], "rtf": r["rtf"]}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def variants(device, T, cfm_steps, tf_peak):
    """CUDA-event timings of hvx_cfm_solve_unet, hvx_hifigan_vocode and hvx_hift_t_vocode at T mel frames (inputs resident)."""
    from flowmirror_hydravox_b200 import _lib as L
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
