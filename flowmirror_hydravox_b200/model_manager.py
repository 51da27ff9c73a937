"""Host-side mirror of the reference's ``server/model_utils`` surface for the hot path.

``ModelManager`` keeps the reference's lifecycle (``load_models`` / ``load_pt`` / ``models[...]``,
server/model_utils/infer_speech_model.py:40-257) but its three ``models`` are the native stage objects
(``NativeLLM``, ``NativeFlow``, ``NativeHiFT``) backed by libhydravox_b200.so.  ``inference_zero_shot`` /
``inference_tts`` below take the *frontend output* (the ``model_input`` dict built by
cosyvoice/cli/frontend.py:157-184 — the tokenizer / ONNX speech tokenizer / CAM++ frontend is the row
SURVEY.md §8(f) marks "next") and run the same three stages in the same order as
infer_speech_model.py:549-592 / 631-670.  ``synthesize_batch`` is the one-call end-to-end path
(host buffers in, host waveform out) used for serving and by bench.py's ``e2e`` figure.

There is no CPU / PyTorch fallback: constructing the manager without a B200 or without the built
library raises (``_lib.HvxError``).
"""
from __future__ import annotations

import ctypes as C
import logging
import os
import time
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib as L
from . import dims as D
from .flow import NativeFlow
from .hift import NativeHiFT
from .llm import NativeLLM

logger = logging.getLogger(__name__)

_DROP = ("epoch", "step", "lr", "optimizer", "scheduler")      # non-tensor keys dropped by load_models (:69-94)


def _clean(sd: Dict) -> Dict[str, torch.Tensor]:
    return {k: v for k, v in sd.items() if isinstance(v, torch.Tensor) and k not in _DROP}


class ModelManager:
    def __init__(self, hd: D.HiftDims = D.HIFT_FULL, fd: D.FlowDims = D.FLOW_FULL, ld: D.LlmDims = D.LLM_FULL,
                 device: str = "cuda:0", max_ctx: int = 8192, max_seqs: int = 32, kv_f32: bool = False, seed: int = 0,
                 n_timesteps: int = 10, sine_seconds: float = 300.0, flow_precise: bool = False):
        """kv_f32 / flow_precise select the parity mode (fp32 KV cache, three-term split-fp16 flow GEMMs); the default is the
        serving mode (bf16 KV cache, fp16 x fp16 flow GEMMs = the reference's own serving precision)."""
        self.engine = L.Engine(hd=hd, fd=fd, ld=ld, max_ctx=max_ctx, max_seqs=max_seqs, device=device, kv_f32=kv_f32,
                               flow_precise=flow_precise)
        self.device = "cuda"
        self.configs = {"sample_rate": hd.sr}
        self.models = {"llm": NativeLLM(self.engine, seed=seed), "flow": NativeFlow(self.engine, n_timesteps=n_timesteps),
                       "hift": NativeHiFT(self.engine)}
        self.is_loaded = False
        self._sine_seconds = sine_seconds
        self._seed = seed
        self._pin: Dict[str, torch.Tensor] = {}
        self._loaded = set()

    # ---------------------------------------------------------------- lifecycle
    def load_state_dicts(self, llm_sd=None, flow_sd=None, hift_sd=None, sine_table: Optional[torch.Tensor] = None):
        if llm_sd is not None:
            self.models["llm"].load_state_dict(_clean(llm_sd))
        if flow_sd is not None:
            self.models["flow"].load_state_dict(_clean(flow_sd))
        if hift_sd is not None:
            self.models["hift"].load_state_dict(_clean(hift_sd))
        hift = self.models["hift"]
        if sine_table is not None:
            hift.set_sine_table(sine_table)
        elif hift.sine_table is None:
            # SineGen2.sine_waves (generator.py:226) is a construction-time torch.rand table that is not in the
            # state_dict; the engine draws its own from a seeded generator
            hd = self.engine.hd
            n = int(self._sine_seconds * hd.sr)
            g = torch.Generator().manual_seed(self._seed + 11)
            hift.set_sine_table(torch.rand(n, hd.harmonics, generator=g))
        for k, v in (("llm", llm_sd), ("flow", flow_sd), ("hift", hift_sd)):
            if v is not None:
                self._loaded.add(k)
        self.is_loaded = len(self._loaded) == 3
        return self

    def load_models(self, args) -> bool:
        """args.model_dir holds llm.pt / flow.pt / hift.pt (infer_speech_model.py:50-143)."""
        d = args.model_dir
        sds = {}
        for name in ("llm", "flow", "hift"):
            path = os.path.join(d, f"{name}.pt")
            sds[name] = torch.load(path, map_location="cpu", weights_only=True)
        self.load_state_dicts(sds["llm"], sds["flow"], sds["hift"])
        return True

    def load_pt(self, llm_pt: Optional[str] = None, flow_pt: Optional[str] = None) -> Dict[str, str]:
        """Hot swap of the llm / flow checkpoints (infer_speech_model.py:169-184).  The failure dict carries both
        'message' and 'error' (the router reads 'error')."""
        try:
            if llm_pt:
                self.models["llm"].load_state_dict(_clean(torch.load(llm_pt, map_location="cpu", weights_only=True)))
            if flow_pt:
                self.models["flow"].load_state_dict(_clean(torch.load(flow_pt, map_location="cpu", weights_only=True)))
            return {"status": "success", "message": "ok"}
        except Exception as ex:      # same contract as the reference: never raise out of load_pt
            return {"status": "error", "message": str(ex), "error": str(ex)}

    # ---------------------------------------------------------------- one-call end to end (host buffers)
    def _pinned(self, key: str, shape, dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        t = self._pin.get(key)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(max(n, 1), dtype=dtype).pin_memory()
            self._pin[key] = t
        return t[:n].view(*shape)

    @torch.no_grad()
    def synthesize_batch(self, requests: Sequence[Dict], head_k: Optional[int] = None, sampling: Optional[Dict] = None,
                         n_timesteps: Optional[int] = None, min_ratio: float = 2.0, max_ratio: float = 20.0,
                         u: Optional[torch.Tensor] = None, return_tokens: bool = False):
        """requests: dicts with CPU tensors text (n,), prompt_text (m,), prompt_speech (P,), prompt_feat (2P, mel) or
        None, embedding (spk_in,), optional speed / min_ratio / max_ratio.  Returns a list of CPU waveforms (1, n)
        (and token lists), plus per-stage device milliseconds in self.last_stage_ms."""
        if not self.is_loaded:
            raise ValueError("model not loaded")
        e, llm, flow, hift = self.engine, self.models["llm"], self.models["flow"], self.models["hift"]
        n = len(requests)
        head_k = int(llm.inference_head_num if head_k is None else head_k)
        steps = int(n_timesteps or flow.n_timesteps)
        frame = e.hd.frame_samples
        arr = (L.Request * n)()
        keep = []
        max_tok = 1
        for i, r in enumerate(requests):
            mn, mx = float(r.get("min_ratio", min_ratio)), float(r.get("max_ratio", max_ratio))
            n_new = int(r["text"].numel())
            max_tok = max(max_tok, int(n_new * mx))
        n_u = 4 * max_tok + 1024
        if u is None:
            u = torch.rand(n, n_u, generator=llm._gen)
        u = u.reshape(n, -1)
        for i, r in enumerate(requests):
            mn, mx = float(r.get("min_ratio", min_ratio)), float(r.get("max_ratio", max_ratio))
            text = torch.cat([r["prompt_text"].reshape(-1), r["text"].reshape(-1)]).to(torch.int32)
            ps = r["prompt_speech"].reshape(-1).to(torch.int32)
            t_text = self._pinned(f"text{i}", text.shape, torch.int32); t_text.copy_(text)
            t_ps = self._pinned(f"ps{i}", (max(int(ps.numel()), 1),), torch.int32)
            t_emb = self._pinned(f"emb{i}", (e.fd.spk_in,), torch.float32); t_emb.copy_(r["embedding"].reshape(-1).float())
            t_u = self._pinned(f"u{i}", (u.shape[1],), torch.float32); t_u.copy_(u[i])
            a = arr[i]
            a.text_ids_host, a.n_text_total, a.n_text_new = t_text.data_ptr(), int(text.numel()), int(r["text"].numel())
            a.n_prompt_speech = int(ps.numel())
            if ps.numel():
                t_ps[: ps.numel()].copy_(ps)
                pf = r["prompt_feat"].reshape(-1, e.fd.mel).float()
                assert pf.shape[0] == 2 * ps.numel(), "prompt_feat must hold 2 frames per prompt token (frontend.py:171-175)"
                t_pf = self._pinned(f"pf{i}", pf.shape, torch.float32); t_pf.copy_(pf)
                a.prompt_speech_host, a.prompt_feat_host = t_ps.data_ptr(), t_pf.data_ptr()
                keep.append(t_pf)
            a.embedding_host, a.u_host, a.n_u = t_emb.data_ptr(), t_u.data_ptr(), int(u.shape[1])
            a.min_ratio, a.max_ratio, a.speed = mn, mx, float(r.get("speed", 1.0))
            if not a.speed > 0:
                raise ValueError(f"Invalid speed: {a.speed}")
            keep += [t_text, t_ps, t_emb, t_u]
        min_speed = min(min(float(r.get("speed", 1.0)) for r in requests), 1.0)
        wav_stride = int(2 * max_tok / min_speed + 2) * frame
        wav = self._pinned("wav", (n, wav_stride), torch.float32)
        wav_len = self._pinned("wav_len", (n,), torch.int32)
        toks = self._pinned("toks", (n, max_tok), torch.int32)
        n_toks = self._pinned("n_toks", (n,), torch.int32)
        ms = self._pinned("ms", (4,), torch.float32)
        sp = llm._sampler(sampling)
        self.h2d_bytes = sum(int(a.n_text_total) * 4 + int(a.n_prompt_speech) * 4 * (1 + 2 * e.fd.mel) + e.fd.spk_in * 4 + int(a.n_u) * 4
                             for a in arr)
        L.check(L.lib().hvx_synthesize_host(e.h, arr, n, head_k, C.byref(sp), steps, L.ptr(flow.noise), L.ptr(hift.sine_table),
                                            C.c_int64(int(hift.sine_table.shape[0])), C.c_void_p(wav.data_ptr()), wav_stride, C.c_void_p(wav_len.data_ptr()),
                                            C.c_void_p(toks.data_ptr()), max_tok, C.c_void_p(n_toks.data_ptr()),
                                            C.c_void_p(ms.data_ptr()), L.stream_ptr()))
        self.last_stage_ms = dict(llm=float(ms[0]), flow=float(ms[1]), hift=float(ms[2]))
        lens = wav_len.tolist()
        self.d2h_bytes = sum(lens) * 4 + sum(int(arr[i].n_text_new * arr[i].max_ratio) for i in range(n)) * 4 + n * 4
        out = [wav[i, : lens[i]].clone().unsqueeze(0) for i in range(n)]
        if return_tokens:
            nt = n_toks.tolist()
            return out, [toks[i, : nt[i]].tolist() for i in range(n)]
        return out


def _req_from_model_input(mi: Dict, speed: float) -> Dict:
    z = torch.zeros(0, dtype=torch.int32)
    return dict(text=mi["text"].reshape(-1).cpu(), prompt_text=mi.get("prompt_text", z).reshape(-1).cpu(),
                prompt_speech=mi.get("llm_prompt_speech_token", z).reshape(-1).cpu(),
                prompt_feat=None if mi.get("prompt_speech_feat") is None else mi["prompt_speech_feat"].cpu(),
                embedding=mi["llm_embedding"].reshape(-1).cpu(), speed=speed)


def inference_zero_shot(model_manager: ModelManager, model_input: Dict, speed: float = 1.0) -> torch.Tensor:
    """infer_speech_model.py:523-610 from the frontend_zero_shot dict onward, stage by stage through the
    inner boundary (models['llm'|'flow'|'hift'].inference), returning tts_speech.cpu()."""
    if not model_manager.is_loaded:
        raise ValueError("model not loaded")
    try:
        mm, mi = model_manager, model_input
        start = time.time()
        toks = list(mm.models["llm"].inference(
            text=mi["text"], text_len=mi.get("text_len"), prompt_text=mi.get("prompt_text"),
            prompt_text_len=mi.get("prompt_text_len"), prompt_speech_token=mi.get("llm_prompt_speech_token"),
            prompt_speech_token_len=mi.get("llm_prompt_speech_token_len"), embedding=mi.get("llm_embedding")))
        llm_time = time.time() - start
        tps = len(toks) / llm_time if llm_time > 0 else 0
        token = torch.tensor(toks).unsqueeze(0)
        mel, _ = mm.models["flow"].inference(
            token=token, token_len=torch.tensor([token.shape[1]], dtype=torch.int32),
            prompt_token=mi.get("flow_prompt_speech_token"), prompt_token_len=mi.get("flow_prompt_speech_token_len"),
            prompt_feat=mi.get("prompt_speech_feat"), prompt_feat_len=mi.get("prompt_speech_feat_len"),
            embedding=mi["flow_embedding"], streaming=False, finalize=True)
        if speed <= 0:
            raise ValueError(f"Invalid speed: {speed}")
        if speed != 1.0:
            mel = speed_interp(mm, mel, max(1, int(mel.shape[2] / speed)))
        speech, _ = mm.models["hift"].inference(speech_feat=mel)
        total = time.time() - start
        logger.info("inference done, total %.2fs, TPS: %.2f, RTF: %.3f", total, tps, total / (speech.shape[-1] / mm.configs["sample_rate"]))
        return speech.cpu()
    except Exception as ex:
        raise ValueError(f"zero-shot inference failed: {ex}")


def inference_tts(model_manager: ModelManager, model_input: Dict, speed: float = 1.0) -> torch.Tensor:
    """infer_speech_model.py:612-689 from the frontend_sft dict onward (no prompt)."""
    return inference_zero_shot(model_manager, model_input, speed)


def speed_interp(model_manager: ModelManager, mel: torch.Tensor, t_out: int) -> torch.Tensor:
    """F.interpolate(mel, size=t_out, mode='linear') on the device (infer_speech_model.py:584-587)."""
    e = model_manager.engine
    m = mel.to(e.device, torch.float32).contiguous()
    out = torch.empty(1, m.shape[1], t_out, device=e.device, dtype=torch.float32)
    L.check(L.lib().hvx_speed_interp(e.h, L.ptr(m), int(m.shape[1]), int(m.shape[2]), int(t_out), L.ptr(out), L.stream_ptr()))
    return out
