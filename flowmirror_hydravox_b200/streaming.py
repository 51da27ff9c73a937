"""Streaming synthesis: AR decode overlapped with chunked flow + vocoder (BASELINE config 5).

Mirrors the upstream streaming orchestration the reference vendors in cosyvoice/cli/model.py:
``CosyVoice2Model.tts(stream=True)`` (:315-360) and ``CosyVoice3Model.token2wav`` (:405-430) —
  * the LLM runs in its own thread / CUDA stream and appends speech tokens;
  * every ``token_hop_len`` (25) tokens (+ ``pre_lookahead_len`` 3, + the prompt padding on the first chunk) the flow is
    re-run over *all tokens so far* with ``streaming=True, finalize=False`` and the new mel frames are appended to a cache;
  * the vocoder is re-run over the cached mel with ``finalize=False`` and only the new samples are emitted;
  * the last call uses ``finalize=True`` and — because ``CosyVoice2Model.tts`` does not pass ``stream`` to that last
    ``token2wav`` (:352-358) — ``streaming=False``: the final mel is computed under full attention, not the chunk mask.
``StreamingSynthesizerCV2`` is the other orchestration the reference vendors, ``CosyVoice2Model.token2wav`` (:279-313): the
vocoder (the ConvTranspose1d ``HiFTGenerator``) runs only on the NEW mel frames plus an 8-frame mel cache, its source signal is
continued from the previous chunk (``cache_source``), the first 3840 samples of every chunk are cross-faded with the previous
chunk's held-back tail under a Hamming window (``fade_in_out``, cosyvoice/utils/common.py:169-177) and every non-final chunk
holds its last 3840 samples back.
One request at a time per synthesizer (a lock serialises ``tts``); a consumer that abandons the generator (client disconnect)
cancels the decode (``hvx_llm_cancel``) and joins the LLM thread before the token buffers and sequence slot 0 can be reused.
The reference polls the token list every 100 ms (:334); here the consumer watches the device-side token counter that the
on-device sampler advances every step, so the first chunk starts as soon as its 28 tokens exist.
"""
from __future__ import annotations

import math
import threading
import time
from typing import Dict, Generator, Optional

import torch

from . import _lib as L


class StreamingSynthesizer:
    token_hop_len = 25            # cli/model.py:396 (must match the flow's static_chunk_size / token_mel_ratio)
    pre_lookahead_len = 3         # flow.pre_lookahead_len

    def __init__(self, model_manager, incremental: bool = False, max_frames: Optional[int] = None):
        """incremental: non-final chunks evaluate only their new frames through a flow streaming session (hvx_flow_stream_*: per
        Euler step and layer key / value caches under the chunk mask) instead of re-running the flow over all tokens so far."""
        self.mm = model_manager
        self.incremental, self.max_frames = incremental, max_frames
        self.dev = model_manager.engine.device
        self.side = torch.cuda.Stream(self.dev)           # flow + vocoder + counter polling; the LLM has its own stream
        self._cnt_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        self._bufs = {}
        self._lock = threading.Lock()                     # the engine has one LLM sequence slot 0 and one set of token buffers per shape

    def _poll(self, cnt: torch.Tensor) -> int:
        with torch.cuda.stream(self.side):
            self._cnt_host.copy_(cnt, non_blocking=True)
        self.side.synchronize()
        return int(self._cnt_host[0])

    @torch.no_grad()
    def tts(self, request: Dict, head_k: Optional[int] = None, sampling: Optional[Dict] = None, n_timesteps: int = 10,
            min_ratio: float = 2.0, max_ratio: float = 20.0, u: Optional[torch.Tensor] = None,
            debug: Optional[Dict] = None) -> Generator[Dict, None, None]:
        """request: dict with text, prompt_text, prompt_speech, prompt_feat, embedding (CPU or device tensors).
        Yields {'tts_speech': (1, n) CPU tensor} chunks exactly like CosyVoice2Model.tts(stream=True)."""
        with self._lock:
            yield from self._tts(request, head_k, sampling, n_timesteps, min_ratio, max_ratio, u, debug)

    def _vocode(self, state: Dict, mel: torch.Tensor, finalize: bool) -> torch.Tensor:
        """CosyVoice3Model.token2wav (cli/model.py:418-430): the causal vocoder re-runs over the cached mel, new samples are emitted"""
        hift = self.mm.models["hift"]
        state["mel"] = mel if "mel" not in state else torch.cat([state["mel"], mel], dim=2)
        wav, _ = hift.inference(speech_feat=state["mel"], finalize=finalize)
        off = state.get("speech_offset", 0)
        wav = wav[:, off:]
        state["speech_offset"] = off + wav.shape[1]
        return wav

    def _tts(self, request, head_k, sampling, n_timesteps, min_ratio, max_ratio, u, debug):
        mm, dev = self.mm, self.dev
        llm, flow = mm.models["llm"], mm.models["flow"]
        n_new = int(request["text"].numel())
        mn, mx = float(request.get("min_ratio", min_ratio)), float(request.get("max_ratio", max_ratio))
        max_out = int(n_new * mx) + 8
        if max_out not in self._bufs:                  # stable device pointers -> the engine reuses its decode graph
            self._bufs[max_out] = (torch.zeros(1, max_out, device=dev, dtype=torch.int32), torch.zeros(1, device=dev, dtype=torch.int32))
        out, cnt = self._bufs[max_out]
        out.zero_(); cnt.zero_()
        if u is None:
            u = torch.rand(1, 4 * max_out + 1024, generator=llm._gen)
        err = []

        def llm_job():
            try:
                torch.cuda.set_device(dev)
                llm.generate_batch([dict(text=request["text"], prompt_text=request["prompt_text"], prompt_speech=request["prompt_speech"])],
                                   head_k=head_k, u=u, sampling=sampling, min_ratio=mn, max_ratio=mx, out=out, cnt=cnt)
            except Exception as ex:         # surfaced by the consumer, like the worker turns exceptions into {"error": ...}
                err.append(ex)

        torch.cuda.synchronize(dev)
        th = threading.Thread(target=llm_job, daemon=True)
        t_start = time.perf_counter()
        th.start()
        P = int(request["prompt_speech"].numel())
        ptok = request["prompt_speech"].reshape(1, -1).to(dev, torch.int32) if P else None
        pfeat = request["prompt_feat"].reshape(1, -1, flow.dims.mel).to(dev, torch.float32) if P else None
        emb = request["embedding"].reshape(1, -1).to(dev, torch.float32)
        hop, la = self.token_hop_len, self.pre_lookahead_len
        prompt_pad = int(math.ceil(P / hop) * hop - P)
        token_offset = 0
        state: Dict = {}

        if self.incremental:
            with torch.cuda.stream(self.side):
                flow.stream_begin(n_timesteps, self.max_frames or min(flow.dims.noise_frames, 2 * (P + max_out) + 64))

        def token2wav(n_tok: int, finalize: bool):
            with torch.cuda.stream(self.side):
                if self.incremental and not finalize:
                    mel = flow.stream_append(out[:, :n_tok], emb, prompt_token=ptok, prompt_feat=pfeat)
                else:
                    # the reference's last token2wav call omits `stream` (cli/model.py:352-358): full attention on the final pass
                    mel, _ = flow.inference(token=out[:, :n_tok], embedding=emb, prompt_token=ptok, prompt_feat=pfeat,
                                            streaming=not finalize, finalize=finalize, n_timesteps=n_timesteps)
                    mel = mel[:, :, token_offset * 2:]
                wav = self._vocode(state, mel, finalize)
                res = wav.cpu()                       # D2H on the side stream, synchronises it
            if debug is not None:
                debug.setdefault("mel", []).append(mel.cpu())
                debug.setdefault("n_tok", []).append(n_tok)
            return res

        first = True
        try:
            while True:
                this_hop = hop + prompt_pad if token_offset == 0 else hop
                n = self._poll(cnt)
                alive = th.is_alive()
                if n - token_offset >= this_hop + la:
                    wav = token2wav(token_offset + this_hop + la, finalize=False)
                    token_offset += this_hop
                    if first and debug is not None:
                        debug["first_audio_ms"] = (time.perf_counter() - t_start) * 1e3
                    first = False
                    yield {"tts_speech": wav}
                    continue
                if not alive:
                    n = self._poll(cnt)
                    if n - token_offset < this_hop + la:
                        break
                    continue
                time.sleep(0.0002)
            th.join()
            if err:
                raise err[0]
            n = self._poll(cnt)
            if debug is not None:
                debug["tokens"] = out[0, :n].cpu().tolist()
                debug["mel_cache"] = lambda: state.get("mel")
            if n > 0:
                wav = token2wav(n, finalize=True)
                if first and debug is not None:
                    debug["first_audio_ms"] = (time.perf_counter() - t_start) * 1e3
                yield {"tts_speech": wav}
        finally:
            if self.incremental:
                flow.stream_end()
            if th.is_alive():                  # generator abandoned or failed mid-stream: stop the decode before anything is reused
                L.check(L.lib().hvx_llm_cancel(mm.engine.h))
                th.join()


class StreamingSynthesizerCV2(StreamingSynthesizer):
    """`CosyVoice2Model.tts(stream=True)` + `CosyVoice2Model.token2wav` (cosyvoice/cli/model.py:279-360): same chunk schedule, but the
    vocoder is the ConvTranspose1d HiFTGenerator run on [8 cached mel frames | new frames] with the source continued from the
    previous chunk and the chunk seams cross-faded (`fade_in_out`) — each chunk costs O(new frames), not O(frames so far).

    hift_t: a NativeHiFTTransposed whose weights are loaded.  noise_fn(n_samples) -> (n_samples, harmonics) pins the source
    module's Gaussian draw per vocoder call (tests); default: fresh noise per call, like the reference."""
    mel_cache_len = 8                                      # cli/model.py:249
    def __init__(self, model_manager, hift_t, noise_fn=None, incremental: bool = False, max_frames: Optional[int] = None):
        import numpy as np
        super().__init__(model_manager, incremental=incremental, max_frames=max_frames)
        self.hift_t, self.noise_fn = hift_t, noise_fn
        self.source_cache_len = self.mel_cache_len * hift_t.dims.frame_samples            # :250 (8 * 480)
        self.speech_window = torch.from_numpy(np.hamming(2 * self.source_cache_len)).to(self.dev)   # float64, :252

    def _vocode(self, state: Dict, mel: torch.Tensor, finalize: bool) -> torch.Tensor:
        ov = self.source_cache_len
        cache = state.get("hift")
        src_cache = None
        if cache is not None:
            mel = torch.cat([cache["mel"], mel], dim=2)
            src_cache = cache["source"]
        noise = None if self.noise_fn is None else self.noise_fn(mel.shape[2] * self.hift_t.dims.frame_samples)
        wav, src = self.hift_t.inference(speech_feat=mel, cache_source=src_cache, noise=noise)
        if cache is not None:                              # fade_in_out(tts_speech, cache speech, window) (:295, :308)
            L.check(L.lib().hvx_fade_in_out(self.mm.engine.h, L.ptr(wav), L.ptr(cache["speech"]), L.ptr(self.speech_window), ov, L.stream_ptr()))
        if finalize:
            return wav
        state["hift"] = {"mel": mel[:, :, -self.mel_cache_len:].contiguous(), "source": src[:, :, -ov:].contiguous(),
                         "speech": wav[:, -ov:].contiguous()}
        state["mel"] = mel
        return wav[:, :-ov]
