"""The innermost boundary: the reference's own ConditionalCFM.forward_estimator / solve_euler
(cosyvoice/flow/flow_matching.py:71-153), unmodified, driven through flowmirror_hydravox_b200.flow.NativeEstimatorPool
in this CPU container.  The pool's execute step is replaced by the reference's nn.Module estimator acting on the raw
addresses the seam handed over (there is no GPU here); everything else — acquire/release, tensor names, the shape and
address protocol, the in-place result in `x` — is the product code.  Skipped where /root/reference does not exist."""
import contextlib
import ctypes
import os

import pytest
import torch

from flowmirror_hydravox_b200 import dims as D, synth

REF = os.environ.get("HVX_REFERENCE_ROOT", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (GPU box)")


class _FakeStream:
    cuda_stream = 0

    def synchronize(self):
        pass


def _view(addr, shape):
    n = 1
    for s in shape:
        n *= s
    buf = (ctypes.c_float * n).from_address(addr)
    return torch.frombuffer(buf, dtype=torch.float32).view(*shape)


def test_reference_cfm_drives_the_pool(monkeypatch):
    from oracle import refshim
    from flowmirror_hydravox_b200.flow import NativeEstimatorPool, _SeamContext
    fd = D.FLOW_TINY
    flow = refshim.build_flow(fd)
    flow.load_state_dict(synth.flow_state_dict(fd, 0), strict=True)
    cfm = flow.decoder
    module = cfm.estimator
    calls = []

    class Ctx(_SeamContext):
        def set_input_shape(self, name, shape):
            calls.append(("shape", name, tuple(shape)))
            return super().set_input_shape(name, shape)

        def set_tensor_address(self, name, addr):
            calls.append(("addr", name))
            return super().set_tensor_address(name, addr)

        def execute_async_v3(self, handle):
            T, a = self.bound()                       # the product's validation of what the reference bound
            mel = self.pool.mel
            x, mu, cond = (_view(a[k], (2, mel, T)) for k in ("x", "mu", "cond"))
            mask, t, spks = _view(a["mask"], (2, 1, T)), _view(a["t"], (2,)), _view(a["spks"], (2, mel))
            with torch.no_grad():
                y = module(x.clone(), mask, mu, t, spks, cond, streaming=self.pool.streaming)
            _view(a["estimator_out"], (2, mel, T)).copy_(y)
            calls.append(("execute", T, a["estimator_out"] == a["x"]))
            return True

    class Owner:                                      # what the pool reads from a NativeFlow
        engine = type("E", (), {"h": None, "device": "cpu"})()
        dims = fd

    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _FakeStream())
    g = torch.Generator().manual_seed(5)
    T, n_steps = 46, 3
    mu = torch.randn(1, fd.mel, T, generator=g)
    cond = torch.randn(1, fd.mel, T, generator=g)
    spks = torch.randn(1, fd.mel, generator=g)
    mask = torch.ones(1, 1, T)
    with torch.no_grad():
        want, _ = cfm(mu=mu, mask=mask, n_timesteps=n_steps, spks=spks, cond=cond)           # nn.Module branch (:127-128)
    del cfm.estimator                                                                        # cli/model.py:86
    cfm.estimator = NativeEstimatorPool(Owner(), dtype=torch.float32, context_cls=Ctx, stream_factory=contextlib.nullcontext)
    assert not isinstance(cfm.estimator, torch.nn.Module)
    with torch.no_grad():
        got, _ = cfm(mu=mu, mask=mask, n_timesteps=n_steps, spks=spks, cond=cond)           # pool branch (:129-153)
    assert torch.equal(got, want)
    # the protocol the reference spoke, per Euler step: 6 shapes, 7 addresses in engine order, one execute, in place into x
    per_step = 6 + 7 + 1
    assert len(calls) == n_steps * per_step
    step0 = calls[:per_step]
    assert [c[1] for c in step0 if c[0] == "shape"] == ["x", "mask", "mu", "t", "spks", "cond"]
    assert [c[1] for c in step0 if c[0] == "addr"] == ["x", "mask", "mu", "t", "spks", "cond", "estimator_out"]
    assert step0[-1] == ("execute", T, True)
    assert cfm.estimator.trt_context_pool.qsize() == 1                                       # released after every call


def test_pool_rejects_what_the_seam_cannot_take():
    from flowmirror_hydravox_b200.flow import NativeEstimatorPool, _SeamContext
    from flowmirror_hydravox_b200._lib import HvxError

    class Owner:
        engine = type("E", (), {"h": None, "device": "cpu"})()
        dims = D.FLOW_TINY

    pool = NativeEstimatorPool(Owner(), context_cls=_SeamContext, stream_factory=contextlib.nullcontext, min_T=4, max_T=3000)
    (ctx, stream), eng = pool.acquire_estimator()
    assert [eng.get_tensor_name(i) for i in range(7)] == ["x", "mask", "mu", "t", "spks", "cond", "estimator_out"]
    with pytest.raises(ValueError):
        ctx.set_input_shape("y", (1,))
    mel = D.FLOW_TINY.mel
    for n, shp in (("x", (2, mel, 2)), ("mask", (2, 1, 2)), ("mu", (2, mel, 2)), ("t", (2,)), ("spks", (2, mel)), ("cond", (2, mel, 2))):
        ctx.set_input_shape(n, shp)
    with pytest.raises(HvxError, match="not bound"):
        ctx.bound()
    for i in range(7):
        ctx.set_tensor_address(eng.get_tensor_name(i), 4096)
    with pytest.raises(HvxError, match="profile"):            # T = 2 < min_T (cli/model.py:93-98: dynamic T in [4, 3000])
        ctx.bound()
    ctx.set_input_shape("x", (1, mel, 8))
    with pytest.raises(HvxError, match="CFG batch"):
        ctx.bound()
    pool.release_estimator(ctx, stream)
    with pytest.raises(ValueError):
        NativeEstimatorPool(Owner(), dtype=torch.float64, context_cls=_SeamContext, stream_factory=contextlib.nullcontext)
