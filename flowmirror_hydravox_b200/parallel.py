"""Utterance-sharded data parallelism over the GPUs of one box (BASELINE configs[3]).

The reference scales by running one worker process per GPU that pull requests from a shared queue
(server/worker.py:31,116-127; app_server.py:58) — no collective at all.  Here one process per GPU is launched by
torchrun; rank 0 owns the request batch, deals it to the ranks by predicted work, every rank runs the full three-stage
engine on its shard, and the waveforms come back to rank 0.  The only communication is that scatter and gather (NCCL
over NVLink on the GPU box, gloo in the CPU tests); the data path itself has no collective because utterances are
independent.

Wire format: a shard travels as ONE flat int32 tensor (ids as they are, fp32 payloads bit-cast) through one `dist.scatter`;
the waveforms of a shard come back as ONE flat buffer — concatenated without per-utterance padding, fp32 or 16-bit PCM —
through one `dist.gather`, preceded by one small gather of (request index, length) pairs.  No pickling, no per-row copies.

Documents whose segments chain (`last_prompt=True`: segment i is zero-shot-prompted by the audio of segment i-1,
server/model_utils/infer_speech_model.py:392-413) are serial inside a document: `shard_documents` deals whole documents.
"""
from __future__ import annotations

import time
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

_I32 = ("text", "prompt_text", "prompt_speech")            # int32 id arrays of a request
_F32 = ("prompt_feat", "embedding", "u")                    # fp32 arrays (optional: prompt_feat, u)
_SCALARS = ("min_ratio", "max_ratio", "speed")              # fp32 scalars, NaN = absent


def predicted_work(req: Dict) -> float:
    """Speech tokens dominate every stage's cost: N ~ ratio * n_text (llm_multi_head_v3.py:955-956); the flow's attention adds a
    term quadratic in the frame count."""
    n_text = int(req["text"].numel())
    ratio = float(req.get("max_ratio", 20.0)) if req.get("min_ratio") == req.get("max_ratio") and req.get("max_ratio") else 8.0
    n_tok = n_text * ratio + 0.25 * int(req.get("prompt_speech", torch.zeros(0)).numel())
    return n_tok * (1.0 + n_tok / 4200.0)        # 0.756*T + 1.8e-4*T^2 GFLOP per NFE with T = 2*n_tok


def _lpt(weights: Sequence[float], world: int) -> List[List[int]]:
    order = sorted(range(len(weights)), key=lambda i: -weights[i])
    load = [0.0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += weights[i]
    return shards


def shard_requests(requests: Sequence[Dict], world: int) -> List[List[int]]:
    """Longest-processing-time-first deal: indices of `requests` per rank, balanced by predicted work."""
    return _lpt([predicted_work(r) for r in requests], world)


def shard_documents(documents: Sequence[Sequence[Dict]], world: int) -> List[List[int]]:
    """Chained segmentation (`last_prompt=True`, infer_speech_model.py:392-413): the segments of a document form a serial chain
    (segment i needs segment i-1's audio as its prompt), so whole documents are dealt; returns document indices per rank."""
    return _lpt([sum(predicted_work(r) for r in doc) for doc in documents], world)


# ------------------------------------------------------------------------------------------------ wire format
def pack_requests(requests: Sequence[Dict], indices: Sequence[int]) -> torch.Tensor:
    """-> flat int32 tensor: [n, then per request: index, 3 id lengths, 3 fp32 lengths, prompt_feat columns, 3 scalars, payloads]."""
    parts = [np.array([len(requests)], dtype=np.int32)]
    for r, gi in zip(requests, indices):
        ids = [np.ascontiguousarray(r[k].reshape(-1).to(torch.int32).numpy()) if r.get(k) is not None else np.zeros(0, np.int32) for k in _I32]
        fls = [np.ascontiguousarray(r[k].reshape(-1).to(torch.float32).numpy()) if r.get(k) is not None else np.zeros(0, np.float32) for k in _F32]
        pf_cols = int(r["prompt_feat"].shape[-1]) if r.get("prompt_feat") is not None else 0
        sc = np.array([float(r[k]) if r.get(k) is not None else np.nan for k in _SCALARS], dtype=np.float32)
        head = np.array([gi] + [a.size for a in ids] + [a.size for a in fls] + [pf_cols], dtype=np.int32)
        parts += [head, sc.view(np.int32)] + ids + [a.view(np.int32) for a in fls]
    return torch.from_numpy(np.concatenate(parts))


def unpack_requests(words: torch.Tensor) -> (List[Dict], List[int]):
    w = words.cpu().numpy()
    n, p = int(w[0]), 1
    out, idx = [], []
    for _ in range(n):
        head = w[p: p + 8]; p += 8
        sc = w[p: p + 3].view(np.float32); p += 3
        r: Dict = {}
        for k, ln in zip(_I32, head[1:4]):
            r[k] = torch.from_numpy(w[p: p + ln].copy()); p += int(ln)
        for k, ln in zip(_F32, head[4:7]):
            a = torch.from_numpy(w[p: p + ln].view(np.float32).copy()); p += int(ln)
            if k == "prompt_feat":
                r[k] = a.view(-1, int(head[7])) if ln else None
            elif k == "u":
                if ln:
                    r[k] = a
            else:
                r[k] = a
        for k, v in zip(_SCALARS, sc):
            if not np.isnan(v):
                r[k] = float(v)
        out.append(r); idx.append(int(head[0]))
    return out, idx


def _dev(group=None) -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def scatter_requests(requests: Optional[Sequence[Dict]], group=None) -> (List[Dict], List[int], int):
    """rank 0 passes the full list, the others None; returns this rank's requests, their global indices and the total count.
    One broadcast of (total, padded shard words) + one scatter of the packed shards."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = _dev(group)
    packed = None
    meta = torch.zeros(2, dtype=torch.int64, device=dev)
    if rank == 0:
        shards = shard_requests(requests, world)
        packed = [pack_requests([requests[i] for i in s], s) for s in shards]
        meta = torch.tensor([len(requests), max(int(p.numel()) for p in packed)], dtype=torch.int64, device=dev)
    dist.broadcast(meta, src=0, group=group)
    n_total, words = int(meta[0]), int(meta[1])
    recv = torch.empty(words, dtype=torch.int32, device=dev)
    send = None
    if rank == 0:
        send = []
        for p in packed:
            b = torch.zeros(words, dtype=torch.int32, device=dev)
            b[: p.numel()] = p.to(dev, non_blocking=True)
            send.append(b)
    dist.scatter(recv, send, src=0, group=group)
    mine, idx = unpack_requests(recv)
    return mine, idx, n_total


def gather_waveforms(wavs: Sequence[torch.Tensor], idx: Sequence[int], n_total: int, group=None, wire: str = "f32") -> Optional[List[torch.Tensor]]:
    """Each rank contributes its (1, n_i) waveforms; rank 0 gets all n_total in request order.  wire: "f32" (bit-exact) or "s16"
    (16-bit PCM, what the server encodes in the end: router.py:262-283 — half the bytes)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = _dev(group)
    assert wire in ("f32", "s16")
    # 1) (count, total samples, then (index, length) pairs): one small all_gather so every rank knows the padded flat size
    max_shard = n_total
    meta = torch.zeros(2 + 2 * max_shard, dtype=torch.int64)
    meta[0], meta[1] = len(wavs), sum(int(w.shape[-1]) for w in wavs)
    for k, (w, i) in enumerate(zip(wavs, idx)):
        meta[2 + 2 * k], meta[3 + 2 * k] = i, w.shape[-1]
    meta = meta.to(dev)
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    metas = torch.stack(metas).cpu()
    flat_len = (max(int(metas[:, 1].max()), 1) + 1) & ~1
    # 2) one flat buffer per rank, utterances back to back (16-bit samples travel as int32 words: neither NCCL nor gloo
    #    has an int16 type)
    dt = torch.float32 if wire == "f32" else torch.int16
    flat = torch.zeros(flat_len, dtype=dt, device=dev)
    if wavs:
        cat = torch.cat([w.reshape(-1) for w in wavs]).to(dev, non_blocking=True)
        flat[: cat.numel()] = cat if wire == "f32" else (cat.clamp(-1.0, 1.0) * 32767.0).round().to(torch.int16)
    send = flat if wire == "f32" else flat.view(torch.int32)
    gathered = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
    dist.gather(send, gathered, dst=0, group=group)
    if rank != 0:
        return None
    host = torch.stack(gathered).cpu()                                     # one D2H
    if wire == "s16":
        host = host.view(torch.int16)
    out: List[Optional[torch.Tensor]] = [None] * n_total
    for r in range(world):
        off = 0
        for k in range(int(metas[r, 0])):
            i, n = int(metas[r, 2 + 2 * k]), int(metas[r, 3 + 2 * k])
            seg = host[r, off: off + n]
            out[i] = (seg if wire == "f32" else seg.to(torch.float32) / 32767.0).unsqueeze(0)
            off += n
    return out


def synthesize_sharded(synth_fn: Callable[[List[Dict]], List[torch.Tensor]], requests: Optional[Sequence[Dict]], group=None,
                       wire: str = "f32", stats: Optional[Dict] = None):
    """rank 0: list of requests in, list of waveforms out (request order); other ranks pass None and get None.
    `synth_fn` is ModelManager.synthesize_batch bound to this rank's engine.  stats (optional dict) receives this rank's
    wall-clock split: scatter_ms, synth_ms, gather_ms, n_mine."""
    sync = (lambda: torch.cuda.synchronize()) if dist.get_backend(group) == "nccl" else (lambda: None)
    t0 = time.perf_counter()
    mine, idx, n_total = scatter_requests(requests, group)
    sync(); t1 = time.perf_counter()
    wavs = synth_fn(mine) if mine else []
    sync(); t2 = time.perf_counter()
    out = gather_waveforms(wavs, idx, n_total, group, wire)
    sync(); t3 = time.perf_counter()
    if stats is not None:
        stats.update(scatter_ms=(t1 - t0) * 1e3, synth_ms=(t2 - t1) * 1e3, gather_ms=(t3 - t2) * 1e3, n_mine=len(mine))
    return out
