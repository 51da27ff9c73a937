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
    out = parallel.synthesize_sharded(_stub_synth, reqs)
    if rank == 0:
        ref = _stub_synth(_reqs(n))
        ok = len(out) == n and all(torch.equal(a, b) for a, b in zip(out, ref))
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
