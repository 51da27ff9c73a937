"""Utterance-sharded data parallelism over the GPUs of one box.

The reference scales by running one worker process per GPU that pull requests from a shared queue
(server/worker.py:31,122-127; app_server.py:58) — no collective at all.  Here one process per GPU is launched by
torchrun; rank 0 owns the request batch, deals it to the ranks by predicted work, every rank runs the full three-stage
engine on its shard, and the waveforms come back to rank 0.  The only communication is that scatter and gather (NCCL
over NVLink on the GPU box, gloo in the CPU tests); the data path itself has no collective because utterances are
independent.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


def predicted_work(req: Dict) -> float:
    """Speech tokens dominate every stage's cost: N ~ ratio * n_text (llm_multi_head_v3.py:955-956)."""
    n_text = int(req["text"].numel())
    ratio = float(req.get("max_ratio", 20.0)) if req.get("min_ratio") == req.get("max_ratio") and req.get("max_ratio") else 8.0
    return n_text * ratio + 0.25 * int(req.get("prompt_speech", torch.zeros(0)).numel())


def shard_requests(requests: Sequence[Dict], world: int) -> List[List[int]]:
    """Longest-processing-time-first deal: indices of `requests` per rank, balanced by predicted work."""
    order = sorted(range(len(requests)), key=lambda i: -predicted_work(requests[i]))
    load = [0.0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += predicted_work(requests[i])
    return shards


def _dev(group=None) -> torch.device:
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def scatter_requests(requests: Optional[Sequence[Dict]], group=None) -> (List[Dict], List[int]):
    """rank 0 passes the full list, the others None; returns this rank's requests and their global indices."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    payload = None
    if rank == 0:
        shards = shard_requests(requests, world)
        payload = [([requests[i] for i in idx], idx) for idx in shards]
    out = [None]
    dist.scatter_object_list(out, payload, src=0, group=group)
    mine, idx = out[0]
    return list(mine), list(idx)


def gather_waveforms(wavs: Sequence[torch.Tensor], idx: Sequence[int], n_total: int, group=None) -> Optional[List[torch.Tensor]]:
    """Each rank contributes its (1, n_i) waveforms; rank 0 gets all n_total in request order."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = _dev(group)
    # 1) shard sizes and lengths: one small all_gather
    max_shard = (n_total + world - 1) // world + n_total            # upper bound on a shard's size
    meta = torch.full((2 * max_shard + 1,), -1, dtype=torch.int64, device=dev)
    meta[0] = len(wavs)
    for k, (w, i) in enumerate(zip(wavs, idx)):
        meta[1 + 2 * k], meta[2 + 2 * k] = i, w.shape[-1]
    metas = [torch.empty_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    max_len = max([int(m[2 + 2 * k]) for m in metas for k in range(int(m[0]))] + [1])
    max_cnt = max(int(m[0]) for m in metas)
    # 2) padded payload: one gather
    buf = torch.zeros(max(max_cnt, 1), max_len, dtype=torch.float32, device=dev)
    for k, w in enumerate(wavs):
        buf[k, : w.shape[-1]] = w.reshape(-1).to(dev)
    gathered = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0, group=group)
    if rank != 0:
        return None
    out: List[Optional[torch.Tensor]] = [None] * n_total
    for r, m in enumerate(metas):
        for k in range(int(m[0])):
            i, n = int(m[1 + 2 * k]), int(m[2 + 2 * k])
            out[i] = gathered[r][k, :n].cpu().unsqueeze(0)
    return out


def synthesize_sharded(synth_fn: Callable[[List[Dict]], List[torch.Tensor]], requests: Optional[Sequence[Dict]], group=None):
    """rank 0: list of requests in, list of waveforms out (request order); other ranks pass None and get None.
    `synth_fn` is ModelManager.synthesize_batch bound to this rank's engine."""
    n_total = torch.tensor([len(requests) if dist.get_rank(group) == 0 else 0], dtype=torch.int64, device=_dev(group))
    dist.broadcast(n_total, src=0, group=group)
    mine, idx = scatter_requests(requests, group)
    wavs = synth_fn(mine) if mine else []
    return gather_waveforms(wavs, idx, int(n_total), group)
