"""NativeLLM — drop-in for models['llm'] (CosyVoice3LM, cosyvoice/llm/llm_multi_head_v3.py:622-960)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Generator, List, Optional, Sequence

import torch

from . import _lib as L
from .weights import pack_llm

DEFAULT_SAMPLING = dict(top_p=0.8, top_k=25, win_size=10, tau_r=0.1)      # ras_sampling defaults (common.py:138)


class NativeLLM:
    def __init__(self, engine: "L.Engine", seed: int = 0):
        self.engine = engine
        self.dims = engine.ld
        self.sampling = None               # server/worker.py:57-63 binds functools.partial(ras_sampling, top_p=..., ...)
        self.inference_head_num = 1        # server/worker.py:64-65
        self.bf16, self.fp16 = True, False
        self._gen = torch.Generator().manual_seed(seed)
        self._u_bufs = {}

    def load_state_dict(self, sd, strict=True):
        self.engine.set_tensors(L.STAGE_LLM, pack_llm(sd, self.dims))
        self.engine.finalize(L.STAGE_LLM)
        return self

    def eval(self):
        return self

    def cuda(self):
        return self

    def half(self):
        return self

    def to(self, *a, **k):
        return self

    # ---- sampler parameters exactly as the worker binds them
    def _sampler(self, override: Optional[Dict] = None) -> "L.Sampler":
        kw = dict(DEFAULT_SAMPLING)
        s = self.sampling
        if s is not None and getattr(s, "keywords", None):
            kw.update({k: v for k, v in s.keywords.items() if k in kw})
        if override:
            kw.update(override)
        return L.Sampler(float(kw["top_p"]), float(kw["tau_r"]), int(kw["top_k"]), int(kw["win_size"]))

    @torch.no_grad()
    def generate_batch(self, requests: Sequence[Dict], head_k: Optional[int] = None, u: Optional[torch.Tensor] = None,
                       sampling: Optional[Dict] = None, min_ratio: float = 2.0, max_ratio: float = 20.0,
                       out: Optional[torch.Tensor] = None, cnt: Optional[torch.Tensor] = None) -> List[List[int]]:
        """requests: dicts with 1-D int tensors text, prompt_text, prompt_speech.  One u-stream row per request.
        `out` (n, max_out) / `cnt` (n,) int32 device tensors may be passed in so that a streaming consumer can watch the
        tokens appear while this call is still running (streaming.py)."""
        e, dev, d = self.engine, self.engine.device, self.dims
        n = len(requests)
        head_k = int(self.inference_head_num if head_k is None else head_k)
        keep = []
        max_out = 1
        for s, r in enumerate(requests):
            text = torch.cat([r["prompt_text"].reshape(-1), r["text"].reshape(-1)]).to(dev, torch.int32).contiguous()
            ps = r["prompt_speech"].reshape(-1).to(dev, torch.int32).contiguous()
            n_new = int(r["text"].numel())
            keep += [text, ps]
            L.check(L.lib().hvx_llm_begin(e.h, s, L.ptr(text), int(text.numel()), n_new, L.ptr(ps) if ps.numel() else None,
                                          int(ps.numel()), C.c_float(r.get("min_ratio", min_ratio)),
                                          C.c_float(r.get("max_ratio", max_ratio))))
            max_out = max(max_out, int(n_new * r.get("max_ratio", max_ratio)) + 8)
        if u is None:
            u = torch.rand(n, 4 * max_out + 1024, generator=self._gen)
        u = u.reshape(n, -1)
        if u.device != dev:                         # keep one device buffer per shape: a stable pointer lets the engine reuse its decode graph
            key = tuple(u.shape)
            buf = self._u_bufs.get(key)
            if buf is None:
                buf = self._u_bufs[key] = torch.empty(key, device=dev, dtype=torch.float32)
            buf.copy_(u.to(torch.float32))
            u = buf
        u = u.to(torch.float32).contiguous()
        if out is None:
            out = torch.zeros(n, max_out, device=dev, dtype=torch.int32)
        if cnt is None:
            cnt = torch.zeros(n, device=dev, dtype=torch.int32)
        max_out = int(out.shape[1])
        sp = self._sampler(sampling)
        L.check(L.lib().hvx_llm_generate(e.h, n, head_k, C.byref(sp), L.ptr(u), int(u.shape[1]), L.ptr(out), max_out,
                                         L.ptr(cnt), L.stream_ptr()))
        cnt_h = cnt.cpu().tolist()
        out_h = out.cpu()
        return [out_h[s, : cnt_h[s]].tolist() for s in range(n)]

    @torch.no_grad()
    def inference(self, text, text_len=None, prompt_text=None, prompt_text_len=None, prompt_speech_token=None,
                  prompt_speech_token_len=None, embedding=None, sampling: int = 25, max_token_text_ratio: float = 20,
                  min_token_text_ratio: float = 2, uuid: str = "") -> Generator[int, None, None]:
        """Same signature as CosyVoice3LM.inference (:926-939); yields python ints."""
        z = torch.zeros(0, dtype=torch.int32)
        req = dict(text=text.reshape(-1), prompt_text=z if prompt_text is None else prompt_text.reshape(-1),
                   prompt_speech=z if prompt_speech_token is None else prompt_speech_token.reshape(-1))
        toks = self.generate_batch([req], min_ratio=min_token_text_ratio, max_ratio=max_token_text_ratio)[0]
        for t in toks:
            yield t

    @torch.no_grad()
    def probe(self, text, prompt_text, prompt_speech):
        """Prefill + MTP heads on the last prompt row: (final-normed hidden (H,), head log-probs (heads, vocab))."""
        e, dev, d = self.engine, self.engine.device, self.dims
        t = torch.cat([prompt_text.reshape(-1), text.reshape(-1)]).to(dev, torch.int32).contiguous()
        ps = prompt_speech.reshape(-1).to(dev, torch.int32).contiguous()
        L.check(L.lib().hvx_llm_begin(e.h, 0, L.ptr(t), int(t.numel()), int(text.numel()), L.ptr(ps) if ps.numel() else None,
                                      int(ps.numel()), C.c_float(2.0), C.c_float(20.0)))
        hid = torch.empty(d.hidden, device=dev, dtype=torch.float32)
        lp = torch.empty(d.mtp_heads, d.speech_vocab, device=dev, dtype=torch.float32)
        L.check(L.lib().hvx_llm_probe(e.h, 0, L.ptr(hid), L.ptr(lp), L.stream_ptr()))
        torch.cuda.synchronize()
        return hid, lp

    @torch.no_grad()
    def debug_step(self, head_k: int, ctx: int, fused: bool, n_layers: int = 0):
        """One decode step of slot 0 at context `ctx` over the current KV cache through the kernel-per-op path or the
        persistent fused kernel (hvx_llm_debug_step) -> (residual rows (head_k, hidden), logits (head_k, vocab) or None)."""
        d, dev = self.dims, self.engine.device
        h = torch.zeros(head_k, d.hidden, device=dev, dtype=torch.float32)
        full = n_layers <= 0 or n_layers >= d.layers
        lg = torch.zeros(head_k, d.speech_vocab, device=dev, dtype=torch.float32) if full else None
        L.check(L.lib().hvx_llm_debug_step(self.engine.h, int(head_k), int(ctx), int(bool(fused)), int(n_layers), L.ptr(h), L.ptr(lg)))
        torch.cuda.synchronize()
        return h, lg

    @torch.no_grad()
    def sample(self, logp: torch.Tensor, history: Sequence[int], min_len: int, u: torch.Tensor, sampling: Optional[Dict] = None):
        """sampling_ids (:151-166) for every row of logp (heads, vocab) against the same history snapshot."""
        dev = self.engine.device
        lp = logp.to(dev, torch.float32).contiguous()
        hist = torch.tensor(list(history), dtype=torch.int32, device=dev)
        uu = u.reshape(-1).to(dev, torch.float32).contiguous()
        ids = torch.zeros(lp.shape[0], dtype=torch.int32, device=dev)
        used = torch.zeros(1, dtype=torch.int32, device=dev)
        sp = self._sampler(sampling)
        L.check(L.lib().hvx_sample(self.engine.h, L.ptr(lp), int(lp.shape[0]), L.ptr(hist) if hist.numel() else None,
                                   int(hist.numel()), int(min_len), C.byref(sp), L.ptr(uu), int(uu.numel()), L.ptr(ids),
                                   L.ptr(used), L.stream_ptr()))
        return ids.cpu().tolist(), int(used.cpu())
