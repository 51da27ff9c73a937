"""world_size-2 gloo test of the utterance sharding plumbing: no GPU, a stub stands in for the engine."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flowmirror_hydravox_b200 import parallel


def _reqs(n):
    g = torch.Generator().manual_seed(0)
    return [dict(text=torch.randint(0, 100, (int(torch.randint(3, 40, (1,), generator=g)),), generator=g),
                 prompt_speech=torch.zeros(0, dtype=torch.int32), min_ratio=8.0, max_ratio=8.0) for _ in range(n)]


def _stub_synth(reqs):
    """deterministic 'waveform': length and content derived from the request only."""
    return [(torch.arange(r["text"].numel() * 8 * 3, dtype=torch.float32) * 1e-3 + float(r["text"].sum())).unsqueeze(0) for r in reqs]


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    reqs = _reqs(n) if rank == 0 else None
    stats = {}
    out = parallel.synthesize_sharded(_stub_synth, reqs, stats=stats)
    out16 = parallel.synthesize_sharded(lambda rq: [w * 1e-4 for w in _stub_synth(rq)], reqs, wire="s16")
    if rank == 0:
        ref = _stub_synth(_reqs(n))
        ok = len(out) == n and all(torch.equal(a, b) for a, b in zip(out, ref))
        ok = ok and all((a - (b * 1e-4).clamp(-1, 1)).abs().max().item() <= 0.5 / 32767 + 1e-7 for a, b in zip(out16, ref))
        ok = ok and {"scatter_ms", "synth_ms", "gather_ms", "n_mine"} <= set(stats)
        q.put(ok)
    else:
        assert out is None
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("n", [1, 7])
def test_sharded_synthesis_two_ranks(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_lpt_sharding_balances():
    reqs = _reqs(33)
    for world in (1, 2, 4, 8):
        shards = parallel.shard_requests(reqs, world)
        assert sorted(i for s in shards for i in s) == list(range(33))
        loads = [sum(parallel.predicted_work(reqs[i]) for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(parallel.predicted_work(r) for r in reqs)


def test_wire_format_roundtrip():
    """a shard travels as one flat int32 tensor: ids as they are, fp32 payloads bit-cast, optional fields preserved"""
    g = torch.Generator().manual_seed(1)
    reqs = [dict(text=torch.randint(0, 1000, (5,), generator=g, dtype=torch.int32), prompt_text=torch.randint(0, 1000, (3,), generator=g, dtype=torch.int32),
                 prompt_speech=torch.randint(0, 6561, (4,), generator=g, dtype=torch.int32), prompt_feat=torch.randn(8, 80, generator=g),
                 embedding=torch.rand(192, generator=g), u=torch.rand(64, generator=g), min_ratio=8.0, max_ratio=8.0, speed=1.25),
            dict(text=torch.randint(0, 1000, (2,), generator=g, dtype=torch.int32), prompt_text=torch.zeros(0, dtype=torch.int32),
                 prompt_speech=torch.zeros(0, dtype=torch.int32), prompt_feat=None, embedding=torch.rand(192, generator=g))]
    words = parallel.pack_requests(reqs, [7, 3])
    assert words.dtype == torch.int32 and words.dim() == 1
    back, idx = parallel.unpack_requests(torch.cat([words, torch.zeros(11, dtype=torch.int32)]))     # scatter pads the shard
    assert idx == [7, 3] and len(back) == 2
    for a, b in zip(reqs, back):
        for k in ("text", "prompt_text", "prompt_speech", "embedding"):
            assert torch.equal(a[k], b[k])
        assert (a["prompt_feat"] is None and b["prompt_feat"] is None) or torch.equal(a["prompt_feat"], b["prompt_feat"])
        assert ("u" in a) == ("u" in b) and ("u" not in a or torch.equal(a["u"], b["u"]))
        for k in ("min_ratio", "max_ratio", "speed"):
            assert a.get(k) == b.get(k)


def test_documents_shard_whole_and_chain_in_lockstep():
    """last_prompt=True chains are serial inside a document (infer_speech_model.py:392-413): whole documents are dealt to
    ranks, and on a rank the chains advance in lock-step, one batched call per chain position."""
    import random
    from flowmirror_hydravox_b200 import output
    docs = [[dict(text=torch.arange(n), tag=(d, k)) for k, n in enumerate(lens)] for d, lens in enumerate(([5, 9, 4], [7], [3, 3], [12, 2, 2, 6]))]
    for world in (1, 2, 3):
        sh = parallel.shard_documents(docs, world)
        assert sorted(i for s in sh for i in s) == [0, 1, 2, 3]

    calls = []

    class MM:
        configs = {"sample_rate": 24000}

        def synthesize_batch(self, reqs, **kw):
            calls.append([r["tag"] for r in reqs])
            return [torch.full((1, 10 * int(r["text"].numel())), float(r.get("prompted_by", -1))) for r in reqs]

    def reprompt(req, prev_req, prev_wav):
        assert prev_wav.shape[1] == 10 * int(prev_req["text"].numel())       # the audio of the segment before it
        req["prompted_by"] = prev_req["tag"][1]
        return req

    outs = output.synthesize_documents(MM(), docs, reprompt=reprompt, last_prompt=True, rng=random.Random(0))
    assert calls == [[(0, 0), (1, 0), (2, 0), (3, 0)], [(0, 1), (2, 1), (3, 1)], [(0, 2), (3, 2)], [(3, 3)]]
    assert len(outs) == 4 and outs[1].shape == (1, 70)
    r = random.Random(0)
    p0 = [int(r.uniform(50, 150) * 24) for _ in range(2)]
    assert outs[0].shape[1] == 180 + sum(p0) and outs[0][0, 50 + p0[0]] == 0.0 and outs[0][0, -1] == 1.0   # segment 2 prompted by segment 1
    calls.clear()
    flat = output.synthesize_documents(MM(), docs, last_prompt=False, rng=random.Random(0))
    assert len(calls) == 1 and len(calls[0]) == 10 and len(flat) == 4
    one = output.synthesize_chained(MM(), docs[3], reprompt, rng=random.Random(1))
    assert one.shape[1] > 220
