"""NativeFlow — drop-in for models['flow'] (CausalMaskedDiffWithDiT, cosyvoice/flow/flow.py:283-430)."""
from __future__ import annotations

import torch

from . import _lib as L
from .weights import pack_flow, pack_unet


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


class NativeUNetEstimator:
    """Drop-in for `CausalConditionalDecoder` (cosyvoice/flow/decoder.py:294-494) as ConditionalCFM.estimator: the call
    `estimator(x, mask, mu, t, spks, cond, streaming=...)` of flow_matching.py:128 on (2, mel, T) CFG-stacked tensors.
    Build the engine with `ud=dims.UNET_FULL`; `Engine(flow_precise=True)` selects the three-term split-fp16 parity mode."""

    def __init__(self, engine: "L.Engine"):
        if engine.ud is None:
            raise L.HvxError("engine was created without U-Net dims (Engine(ud=dims.UNET_FULL))")
        self.engine = engine
        self.dims = engine.ud

    def load_state_dict(self, sd, strict=True):
        self.engine.set_tensors(L.STAGE_UNET, pack_unet(sd, self.dims, precise=getattr(self.engine, "flow_precise", False)))
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
        if mask is not None and not bool((mask != 0).all()):
            raise ValueError("padded batches are not built: the seam's mask is all-true at inference (flow_matching.py:104-111)")
        T = int(x.shape[2])
        args = [a.to(dev, torch.float32).contiguous() for a in (x, mu, t.reshape(-1).expand(2) if t.numel() == 1 else t, spks, cond)]
        if out is None:            # `out=` (and device-resident fp32 inputs) keeps every pointer stable: the CUDA-graph replay path
            out = torch.empty(2, self.dims.mel, T, device=dev, dtype=torch.float32)
        if _dump is None:
            L.check(L.lib().hvx_unet_estimator(self.engine.h, *[L.ptr(a) for a in args], T, int(bool(streaming)), L.ptr(out),
                                               L.stream_ptr()))
        else:
            L.check(L.lib().hvx_unet_estimator_debug(self.engine.h, *[L.ptr(a) for a in args], T, int(bool(streaming)), L.ptr(out),
                                                     L.ptr(_dump), int(_dump.shape[0]), L.stream_ptr()))
        return out

    __call__ = forward


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
