"""NativeFlow — drop-in for models['flow'] (CausalMaskedDiffWithDiT, cosyvoice/flow/flow.py:283-430)."""
from __future__ import annotations

import torch

from . import _lib as L
from .weights import pack_flow, pack_unet, pack_unet_nc


def rand_noise(mel: int = 80, frames: int = 15000) -> torch.Tensor:
    """CausalConditionalCFM.rand_noise (flow_matching.py:200-201): randn(1, 80, 50*300) drawn right after
    set_all_random_seed(0); a CPU generator seeded 0 reproduces it bit for bit (checked against the
    reference module by oracle/make_golden.py)."""
    g = torch.Generator().manual_seed(0)
    return torch.randn(1, mel, 50 * 300, generator=g)[:, :, :frames].contiguous()


class NativeFlow:
    def __init__(self, engine: "L.Engine", noise: torch.Tensor | None = None, n_timesteps: int = 10):
        self.engine = engine
        self.dims = engine.fd
        self.bf16 = False          # attributes ModelManager.load_models sets (infer_speech_model.py:105-117)
        self.fp16 = True
        self.n_timesteps = n_timesteps     # the reference hard-codes 10 (flow.py:425)
        n = rand_noise(self.dims.mel, self.dims.noise_frames) if noise is None else noise
        self.noise = n.reshape(self.dims.mel, -1).to(engine.device, torch.float32).contiguous()
        assert self.noise.shape[1] == self.dims.noise_frames

    def load_state_dict(self, sd, strict=True):
        self.engine.set_tensors(L.STAGE_FLOW, pack_flow(sd, self.dims, precise=getattr(self.engine, "flow_precise", False)))
        self.engine.finalize(L.STAGE_FLOW)
        return self

    def eval(self):
        return self

    def cuda(self):
        return self

    def half(self):
        return self

    def to(self, *a, **k):
        return self

    @torch.no_grad()
    def inference(self, token, token_len=None, embedding=None, finalize=True, prompt_token=None, prompt_token_len=None,
                  prompt_feat=None, prompt_feat_len=None, streaming=False, n_timesteps=None):
        """token (1,N) int, embedding (1,spk_in), prompt_token (1,P), prompt_feat (1,2P,mel)
        -> (mel fp32 (1, mel, 2N'), None)   [flow.py:367-430]"""
        assert token.shape[0] == 1, "reference asserts batch 1 (flow.py:387)"
        d, dev = self.dims, self.engine.device
        n_tok = int(token.shape[1])
        n_prompt = 0 if prompt_token is None else int(prompt_token.shape[1])
        toks = token.reshape(-1) if n_prompt == 0 else torch.cat([prompt_token.reshape(-1), token.reshape(-1)])
        toks = toks.to(dev, torch.int32).contiguous()
        emb = embedding.reshape(-1).to(dev, torch.float32).contiguous()
        pf = None
        if n_prompt:
            pf = prompt_feat.reshape(-1, d.mel).to(dev, torch.float32).contiguous()
            assert pf.shape[0] == 2 * n_prompt, "prompt_feat must hold token_mel_ratio frames per prompt token"
        n_out = 2 * (n_tok if finalize else n_tok - 3)
        mel = torch.empty(1, d.mel, n_out, device=dev, dtype=torch.float32)
        steps = int(n_timesteps or self.n_timesteps)
        L.check(L.lib().hvx_flow_inference(self.engine.h, L.ptr(toks), n_prompt, n_tok, L.ptr(emb), L.ptr(pf),
                                           L.ptr(self.noise), steps, int(bool(streaming)), int(bool(finalize)),
                                           L.ptr(mel), L.stream_ptr()))
        return mel, None

    @torch.no_grad()
    def inference_batch(self, requests, finalize=True, streaming=False, n_timesteps=None):
        """Several utterances in one solve (hvx_flow_inference_batch).  requests: dicts with the keyword arguments of `inference`
        (token, embedding, prompt_token, prompt_feat).  Returns the list of mel tensors `inference` would return one by one."""
        import ctypes as C
        d, dev = self.dims, self.engine.device
        U = len(requests)
        keep, toks, embs, pfs, outs, n_p, n_t = [], [], [], [], [], [], []
        for r in requests:
            token, pt = r["token"], r.get("prompt_token")
            ntok = int(token.shape[1])
            npr = 0 if pt is None else int(pt.shape[1])
            t = (token.reshape(-1) if npr == 0 else torch.cat([pt.reshape(-1), token.reshape(-1)])).to(dev, torch.int32).contiguous()
            emb = r["embedding"].reshape(-1).to(dev, torch.float32).contiguous()
            pf = None
            if npr:
                pf = r["prompt_feat"].reshape(-1, d.mel).to(dev, torch.float32).contiguous()
                assert pf.shape[0] == 2 * npr
            mel = torch.empty(1, d.mel, 2 * (ntok if finalize else ntok - 3), device=dev, dtype=torch.float32)
            keep += [t, emb, pf]; outs.append(mel)
            toks.append(t.data_ptr()); embs.append(emb.data_ptr()); pfs.append(0 if pf is None else pf.data_ptr())
            n_p.append(npr); n_t.append(ntok)
        vp, ip = C.c_void_p * U, C.c_int * U
        L.check(L.lib().hvx_flow_inference_batch(self.engine.h, U, vp(*toks), ip(*n_p), ip(*n_t), vp(*embs), vp(*pfs), L.ptr(self.noise),
                                                 int(n_timesteps or self.n_timesteps), int(bool(streaming)), int(bool(finalize)),
                                                 vp(*[m.data_ptr() for m in outs]), L.stream_ptr()))
        return outs

    # ---- incremental streaming: the repeated `inference(token=all so far, streaming=True, finalize=False)[..., offset:]` calls of
    # CosyVoice2Model.token2wav (cli/model.py:279-297) as a session that evaluates only the new chunk(s)
    def stream_begin(self, n_timesteps=None, max_frames=None):
        steps = int(n_timesteps or self.n_timesteps)
        frames = int(max_frames or min(self.dims.noise_frames, 4096))
        L.check(L.lib().hvx_flow_stream_begin(self.engine.h, steps, frames, L.stream_ptr()))
        return self

    @torch.no_grad()
    def stream_append(self, token, embedding, prompt_token=None, prompt_feat=None):
        """token (1, N): every new token so far incl. the 3 look-ahead tokens -> the mel frames that are new since the previous
        call, (1, mel, n_new) == inference(token, ..., streaming=True, finalize=False)[0][:, :, frames_already_returned:]"""
        import ctypes as C
        d, dev = self.dims, self.engine.device
        n_tok = int(token.shape[1])
        n_prompt = 0 if prompt_token is None else int(prompt_token.shape[1])
        toks = token.reshape(-1) if n_prompt == 0 else torch.cat([prompt_token.reshape(-1), token.reshape(-1)])
        toks = toks.to(dev, torch.int32).contiguous()
        emb = embedding.reshape(-1).to(dev, torch.float32).contiguous()
        pf = None
        if n_prompt:
            pf = prompt_feat.reshape(-1, d.mel).to(dev, torch.float32).contiguous()
            assert pf.shape[0] == 2 * n_prompt, "prompt_feat must hold token_mel_ratio frames per prompt token"
        cap = 2 * (n_tok - 3)
        mel = torch.empty(d.mel * max(cap, 1), device=dev, dtype=torch.float32)
        n_new = C.c_int(0)
        L.check(L.lib().hvx_flow_stream_append(self.engine.h, L.ptr(toks), n_prompt, n_tok, L.ptr(emb), L.ptr(pf), L.ptr(self.noise),
                                               L.ptr(mel), C.byref(n_new), L.stream_ptr()))
        n = int(n_new.value)
        return mel[: d.mel * n].view(1, d.mel, n)

    def stream_end(self):
        L.check(L.lib().hvx_flow_stream_end(self.engine.h))

    @torch.no_grad()
    def estimator(self, x, mask, mu, t, spks, cond, streaming=False):
        """The TensorRT seam of ConditionalCFM.forward_estimator (flow_matching.py:126-153): (2,mel,T) tensors."""
        dev = self.engine.device
        T = int(x.shape[2])
        args = [a.to(dev, torch.float32).contiguous() for a in (x, mu, t, spks, cond)]
        out = torch.empty(2, self.dims.mel, T, device=dev, dtype=torch.float32)
        L.check(L.lib().hvx_dit_estimator(self.engine.h, *[L.ptr(a) for a in args], T, int(bool(streaming)), L.ptr(out),
                                          L.stream_ptr()))
        return out


class NativeUNetEstimator(torch.nn.Module):
    """Drop-in for `CausalConditionalDecoder` (cosyvoice/flow/decoder.py:294-494) as ConditionalCFM.estimator: the call
    `estimator(x, mask, mu, t, spks, cond, streaming=...)` of flow_matching.py:128 on (2, mel, T) CFG-stacked tensors.
    It is an `nn.Module` (without parameters: the weights live in the engine), so it can be assigned over the registered
    submodule `flow.decoder.estimator` and `forward_estimator` takes its nn.Module branch (flow_matching.py:127-128).
    Build the engine with `ud=dims.UNET_FULL`; `Engine(flow_precise=True)` selects the three-term split-fp16 parity mode."""

    def __init__(self, engine: "L.Engine"):
        super().__init__()
        if engine.ud is None:
            raise L.HvxError("engine was created without U-Net dims (Engine(ud=dims.UNET_FULL))")
        self.engine = engine
        self.dims = engine.ud

    @property
    def noncausal(self) -> bool:
        """True: the engine holds the non-causal multi-level ConditionalDecoder (decoder.py:88-291; Engine(ud=dims.UNET_NC_FULL))"""
        return hasattr(self.dims, "channels")

    def load_state_dict(self, sd, strict=True):
        pack = pack_unet_nc if self.noncausal else pack_unet
        self.engine.set_tensors(L.STAGE_UNET, pack(sd, self.dims, precise=getattr(self.engine, "flow_precise", False)))
        self.engine.finalize(L.STAGE_UNET)
        return self

    def eval(self):
        return self

    def cuda(self):
        return self

    def half(self):
        return self

    def to(self, *a, **k):
        return self

    @torch.no_grad()
    def forward(self, x, mask, mu, t, spks=None, cond=None, streaming=False, out=None, _dump=None):
        dev = self.engine.device
        if x.dim() != 3 or x.shape[0] != 2 or x.shape[1] != self.dims.mel:
            raise ValueError(f"estimator input must be (2, {self.dims.mel}, T) — the CFG batch of solve_euler; got {tuple(x.shape)}")
        T = int(x.shape[2])
        mask_d = None
        if mask is not None and not bool((mask != 0).all()):
            if not self.noncausal:
                raise ValueError("the causal variant takes no padded batch: the seam's mask is all-true at inference (flow_matching.py:104-111)")
            mask_d = (mask != 0).reshape(2, T).to(dev, torch.float32).contiguous()
            if not bool((mask_d[:, 1:] <= mask_d[:, :-1]).all()):
                raise ValueError("padding masks are prefix masks (make_pad_mask): a hole inside the valid frames is not supported")
        args = [a.to(dev, torch.float32).contiguous() for a in (x, mu, t.reshape(-1).expand(2) if t.numel() == 1 else t, spks, cond)]
        if out is None:            # `out=` (and device-resident fp32 inputs) keeps every pointer stable: the CUDA-graph replay path
            out = torch.empty(2, self.dims.mel, T, device=dev, dtype=torch.float32)
        if self.noncausal and (mask_d is not None or _dump is not None):
            L.check(L.lib().hvx_unet_estimator_masked(self.engine.h, L.ptr(args[0]), L.ptr(mask_d), *[L.ptr(a) for a in args[1:]], T,
                                                      L.ptr(out), L.ptr(_dump), 0 if _dump is None else int(_dump.shape[0]), L.stream_ptr()))
        elif _dump is None:
            L.check(L.lib().hvx_unet_estimator(self.engine.h, *[L.ptr(a) for a in args], T, int(bool(streaming)), L.ptr(out),
                                               L.stream_ptr()))
        else:
            L.check(L.lib().hvx_unet_estimator_debug(self.engine.h, *[L.ptr(a) for a in args], T, int(bool(streaming)), L.ptr(out),
                                                     L.ptr(_dump), int(_dump.shape[0]), L.stream_ptr()))
        return out if out.dtype == x.dtype or x.dtype not in (torch.float16, torch.bfloat16) else out.to(x.dtype)   # the seam's dtype (spks.dtype)



class NativeUNetCFM:
    """Drop-in for `CausalConditionalCFM` whose estimator is the U-Net (cosyvoice/flow/flow_matching.py:197-228 + solve_euler
    :71-124): `forward(mu, mask, n_timesteps, temperature, spks, cond, streaming) -> (mel fp32 (1, mel, T), None)` in one
    hvx_cfm_solve_unet call — noise slice, cosine schedule, CFG staging, estimator, Euler update all on the device."""

    def __init__(self, engine: "L.Engine", noise: torch.Tensor | None = None):
        import ctypes as C
        self._cf = C.c_float
        self.engine = engine
        self.estimator = NativeUNetEstimator(engine)
        self.dims = self.estimator.dims
        n = rand_noise(self.dims.mel, 50 * 300) if noise is None else noise
        self.rand_noise = n.reshape(self.dims.mel, -1).to(engine.device, torch.float32).contiguous()

    def load_state_dict(self, sd, strict=True):
        """accepts the CFM module's state_dict (keys `estimator.*`) or the estimator's own"""
        est = {k[len("estimator."):]: v for k, v in sd.items() if k.startswith("estimator.")}
        self.estimator.load_state_dict(est or sd)
        return self

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    @torch.no_grad()
    def forward(self, mu, mask=None, n_timesteps=10, temperature=1.0, spks=None, cond=None, streaming=False):
        d, dev = self.dims, self.engine.device
        if mu.dim() != 3 or mu.shape[0] != 1 or mu.shape[1] != d.mel:
            raise ValueError(f"mu must be (1, {d.mel}, T); got {tuple(mu.shape)}")
        if mask is not None and not bool((mask != 0).all()):
            raise ValueError("padded batches are not built (the reference solves one utterance at a time)")
        T = int(mu.shape[2])
        f = lambda a: None if a is None else a.to(dev, torch.float32).contiguous()
        mu_d, spks_d, cond_d = f(mu), f(spks), f(cond)
        out = torch.empty(1, d.mel, T, device=dev, dtype=torch.float32)
        L.check(L.lib().hvx_cfm_solve_unet(self.engine.h, L.ptr(mu_d), L.ptr(spks_d), L.ptr(cond_d), L.ptr(self.rand_noise),
                                           int(self.rand_noise.shape[1]), T, int(n_timesteps), self._cf(float(temperature)),
                                           int(bool(streaming)), L.ptr(out), L.stream_ptr()))
        return out, None

    __call__ = forward


class NativeConditionalCFM(NativeUNetCFM):
    """Drop-in for the non-causal `ConditionalCFM` (cosyvoice/flow/flow_matching.py:22-69) over the multi-level ConditionalDecoder
    (Engine(ud=dims.UNET_NC_FULL)): `forward(mu, mask, n_timesteps, temperature, spks, cond, prompt_len, cache) -> (mel, cache)`.
    The noise is drawn per call (`torch.randn_like(mu)`, :52) and the prompt / 34-frame overlap of z and mu are carried in
    `cache` exactly as :53-60; the Euler solve is the same hvx_cfm_solve_unet call."""

    def __init__(self, engine: "L.Engine"):
        super().__init__(engine, noise=torch.zeros(engine.ud.mel, 1))

    @torch.no_grad()
    def forward(self, mu, mask=None, n_timesteps=10, temperature=1.0, spks=None, cond=None, prompt_len=0, cache=None, z=None):
        d, dev = self.dims, self.engine.device
        if mu.dim() != 3 or mu.shape[0] != 1 or mu.shape[1] != d.mel:
            raise ValueError(f"mu must be (1, {d.mel}, T); got {tuple(mu.shape)}")
        if mask is not None and not bool((mask != 0).all()):
            raise ValueError("solve_euler takes one utterance (flow_matching.py:99-104): its mask is all-true")
        T = int(mu.shape[2])
        mu = mu.to(dev, torch.float32).clone()
        z = (torch.randn_like(mu) if z is None else z.to(dev, torch.float32)) * temperature       # z: test hook (pinned noise)
        if cache is not None and cache.shape[2] != 0:
            cs = cache.shape[2]
            z[:, :, :cs] = cache[:, :, :, 0].to(dev, torch.float32)
            mu[:, :, :cs] = cache[:, :, :, 1].to(dev, torch.float32)
        z_cache = torch.cat([z[:, :, :prompt_len], z[:, :, -34:]], dim=2)
        mu_cache = torch.cat([mu[:, :, :prompt_len], mu[:, :, -34:]], dim=2)
        cache = torch.stack([z_cache, mu_cache], dim=-1)
        f = lambda a: None if a is None else a.to(dev, torch.float32).contiguous()
        z, mu_d, spks_d, cond_d = z.contiguous(), mu.contiguous(), f(spks), f(cond)
        out = torch.empty(1, d.mel, T, device=dev, dtype=torch.float32)
        L.check(L.lib().hvx_cfm_solve_unet(self.engine.h, L.ptr(mu_d), L.ptr(spks_d), L.ptr(cond_d), L.ptr(z), T, T, int(n_timesteps),
                                           self._cf(1.0), 0, L.ptr(out), L.stream_ptr()))
        return out, cache

    __call__ = forward


# ------------------------------------------------------------------------------------------------------------------
# The reference's innermost plug-in seam (SURVEY 8b "innermost"): ConditionalCFM.forward_estimator
# (cosyvoice/flow/flow_matching.py:126-153) treats an estimator that is NOT an nn.Module as a pool of TensorRT execution
# contexts (TrtContextWrapper, cosyvoice/utils/common.py:198-213; installed by CosyVoice2Model.load_trt,
# cosyvoice/cli/model.py:82-98):
#     [context, stream], engine = pool.acquire_estimator()
#     with stream: context.set_input_shape(name, shape) x6; context.set_tensor_address(engine.get_tensor_name(i), ptr) x7;
#                  assert context.execute_async_v3(cuda_stream_handle) is True
#     pool.release_estimator(context, stream)        # result is read from x: the 7th address is x.data_ptr()
# NativeEstimatorPool is that object backed by hvx_estimator_seam.
_SEAM_NAMES = ("x", "mask", "mu", "t", "spks", "cond", "estimator_out")     # bin/export_onnx.py:77-78 input/output names


class _SeamEngine:
    """the part of a tensorrt.ICudaEngine forward_estimator touches"""
    num_io_tensors = len(_SEAM_NAMES)

    def get_tensor_name(self, i: int) -> str:
        return _SEAM_NAMES[i]


class _SeamContext:
    """the part of a tensorrt.IExecutionContext forward_estimator touches"""

    def __init__(self, pool: "NativeEstimatorPool"):
        self.pool = pool
        self.shapes, self.addrs = {}, {}

    def set_input_shape(self, name: str, shape) -> bool:
        if name not in _SEAM_NAMES[:6]:
            raise ValueError(f"unknown estimator input {name!r}")
        self.shapes[name] = tuple(int(s) for s in shape)
        return True

    def set_tensor_address(self, name: str, addr: int) -> bool:
        if name not in _SEAM_NAMES:
            raise ValueError(f"unknown estimator tensor {name!r}")
        self.addrs[name] = int(addr)
        return True

    def bound(self):
        """validate what forward_estimator bound; returns (T, addresses by name)"""
        p = self.pool
        mel = p.mel
        shp = self.shapes
        missing = [n for n in _SEAM_NAMES if n not in self.addrs] + [n for n in _SEAM_NAMES[:6] if n not in shp]
        if missing:
            raise L.HvxError(f"estimator seam: tensors not bound: {sorted(set(missing))}")
        T = shp["x"][2]
        if shp["x"] != (2, mel, T) or shp["mu"] != (2, mel, T) or shp["cond"] != (2, mel, T) or shp["mask"] != (2, 1, T) \
                or shp["t"] != (2,) or shp["spks"] != (2, mel):
            raise L.HvxError(f"estimator seam: shapes {shp} are not the CFG batch of solve_euler (flow_matching.py:99-104)")
        if not p.min_T <= T <= p.max_T:               # the TensorRT profile of cli/model.py:93-98
            raise L.HvxError(f"estimator seam: T={T} outside the profile [{p.min_T}, {p.max_T}]")
        return T, self.addrs

    def execute_async_v3(self, stream_handle: int) -> bool:
        import ctypes as C
        p = self.pool
        T, a = self.bound()
        vp = C.c_void_p
        L.check(L.lib().hvx_estimator_seam(p.engine.h, p.kind, vp(a["x"]), vp(a["mu"]), vp(a["t"]), vp(a["spks"]), vp(a["cond"]),
                                           vp(a["estimator_out"]), int(T), L._DT[p.dtype], int(bool(p.streaming)),
                                           vp(int(stream_handle))))
        return True


class NativeEstimatorPool:
    """Drop-in for `TrtContextWrapper` (cosyvoice/utils/common.py:198-213): assign it to `flow.decoder.estimator` after deleting
    the nn.Module (`del flow.decoder.estimator`, exactly as load_trt does at cli/model.py:86) and the reference's own
    `ConditionalCFM.forward_estimator` drives the B200 estimator through raw pointers, in place into `x`.

    owner: a NativeFlow (DiT estimator) or NativeUNetEstimator whose weights are loaded.  dtype: the dtype of the seam tensors =
    the flow module's dtype (`spks.dtype`; the reference serves in fp16, infer_speech_model.py:105-117).  The TensorRT seam does
    not carry `streaming` (the mask is baked into the engine); set `pool.streaming = True` for the chunk-mask estimator."""

    def __init__(self, owner, dtype: torch.dtype = torch.float32, trt_concurrent: int = 1, streaming: bool = False,
                 min_T: int = 1, max_T: int = 15000, context_cls=None, stream_factory=None):
        """context_cls / stream_factory: test hooks (tests/test_seam_cpu.py drives the reference's forward_estimator through
        this pool on a CPU box with a context whose execute step is the reference's own nn.Module)."""
        import queue
        self.engine = owner.engine
        self.kind = 1 if isinstance(owner, NativeUNetEstimator) else 0
        self.mel = owner.dims.mel
        if dtype not in (torch.float32, torch.float16, torch.bfloat16):
            raise ValueError(f"estimator seam dtype {dtype}: fp32, fp16 or bf16")
        self.dtype, self.streaming, self.min_T, self.max_T = dtype, streaming, min_T, max_T
        self.trt_engine = _SeamEngine()
        self.trt_context_pool = queue.Queue(maxsize=trt_concurrent)
        context_cls = context_cls or _SeamContext
        stream_factory = stream_factory or (lambda: torch.cuda.stream(torch.cuda.Stream(self.engine.device)))
        for _ in range(trt_concurrent):
            self.trt_context_pool.put([context_cls(self), stream_factory()])

    def acquire_estimator(self):
        return self.trt_context_pool.get(), self.trt_engine

    def release_estimator(self, context, stream):
        self.trt_context_pool.put([context, stream])
