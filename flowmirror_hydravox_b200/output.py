"""The step right after the hot path: stitching of independently synthesised text segments
with random pauses and WAV / base64 encoding of the result.

Mirrors server/model_utils/infer_speech_model.py:419-450 (pause insertion of inference_tts_with_segmentation) and
:504-521 (audio_to_base64).  With last_prompt=False — what text_to_speech passes for long inputs (:782-800) — the
segments are independent utterances, so they go through ONE ModelManager.synthesize_batch call instead of a serial loop.
"""
from __future__ import annotations

import base64
import io
import random
import struct
from typing import Dict, List, Optional, Sequence

import torch


def stitch_segments(audio_segments: Sequence[torch.Tensor], sample_rate: int, rng: Optional[random.Random] = None) -> torch.Tensor:
    """cat(seg_0, silence_0, seg_1, ...) with 50-150 ms of silence after every segment but the last
    (infer_speech_model.py:419-441: pause_ms = random.uniform(50, 150); samples = int(pause_ms * sr / 1000))."""
    if not audio_segments:
        raise ValueError("no synthesised audio segments")
    rng = rng or random
    parts: List[torch.Tensor] = []
    for i, seg in enumerate(audio_segments):
        parts.append(seg)
        if i < len(audio_segments) - 1:
            pause_ms = rng.uniform(50, 150)
            shape = list(seg.shape)
            shape[-1] = int(pause_ms * sample_rate / 1000)
            parts.append(torch.zeros(shape, dtype=seg.dtype, device=seg.device))
    return torch.cat(parts, dim=-1)


def wav_bytes(audio: torch.Tensor, sample_rate: int) -> bytes:
    """RIFF/WAVE, IEEE float 32 — what torchaudio.save(buffer, float32 tensor, sr, format='wav') writes by default."""
    a = audio.detach().cpu().to(torch.float32)
    if a.dim() == 1:
        a = a[None]
    ch, n = a.shape
    data = a.t().contiguous().numpy().tobytes()
    fmt = struct.pack("<HHIIHH", 3, ch, sample_rate, sample_rate * ch * 4, ch * 4, 32)
    fact = struct.pack("<I", n)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"fact" + struct.pack("<I", 4) + fact + b"data" + struct.pack("<I", len(data)) + data
    return b"RIFF" + struct.pack("<I", len(body)) + body


def audio_to_base64(audio: torch.Tensor, sample_rate: int, format: str = "wav") -> str:
    """infer_speech_model.py:504-521."""
    if format != "wav":
        raise ValueError(f"audio to base64 failed: unsupported format {format}")
    return base64.b64encode(wav_bytes(audio, sample_rate)).decode("utf-8")


def synthesize_segments(model_manager, segment_requests: Sequence[Dict], rng: Optional[random.Random] = None, **kw) -> torch.Tensor:
    """inference_tts_with_segmentation(last_prompt=False) from the per-segment model inputs onward: every segment is an
    independent utterance -> one batched call, then the reference's pause stitching."""
    if len(segment_requests) == 1:
        return model_manager.synthesize_batch(list(segment_requests), **kw)[0]
    wavs = model_manager.synthesize_batch(list(segment_requests), **kw)
    return stitch_segments(wavs, model_manager.configs["sample_rate"], rng)


def synthesize_chained(model_manager, segment_requests: Sequence[Dict], reprompt, rng: Optional[random.Random] = None, **kw) -> torch.Tensor:
    """inference_tts_with_segmentation(last_prompt=True) (infer_speech_model.py:392-413) from the per-segment model inputs
    onward: segment 0 is synthesised as it is (speaker TTS); every later segment is a zero-shot request whose prompt is the
    text and the AUDIO of the segment before it, so the chain is serial inside a document (documents shard across GPUs:
    parallel.shard_documents).

    reprompt(request_i, prev_request, prev_wav) -> request_i with its prompt fields (prompt_text, prompt_speech, prompt_feat,
    embedding) rebuilt from the previous segment — the zero-shot frontend's job (cosyvoice/cli/frontend.py:157-184: speech
    tokenizer + CAM++ on the 16 kHz audio, 24 kHz mel of it; `frontend.NativeFrontendFeatures` provides the mel / fbank part
    on the GPU, the two ONNX networks are the caller's).  Then the reference's pause stitching (:419-441)."""
    if not segment_requests:
        raise ValueError("no text segments")
    wavs: List[torch.Tensor] = []
    prev_req, prev_wav = None, None
    for i, req in enumerate(segment_requests):
        try:
            cur = dict(req) if i == 0 else reprompt(dict(req), prev_req, prev_wav)
            wav = model_manager.synthesize_batch([cur], **kw)[0]
        except Exception as ex:                          # infer_speech_model.py:414-416
            raise ValueError(f"segment {i + 1} synthesis failed: {ex}")
        wavs.append(wav)
        prev_req, prev_wav = cur, wav
    if len(wavs) == 1:
        return wavs[0]
    return stitch_segments(wavs, model_manager.configs["sample_rate"], rng)


def synthesize_documents(model_manager, documents: Sequence[Sequence[Dict]], reprompt=None, last_prompt: bool = True,
                         rng: Optional[random.Random] = None, **kw) -> List[torch.Tensor]:
    """Several segmented documents on this GPU.  last_prompt=False: every segment of every document is independent -> ONE batched
    call for all of them.  last_prompt=True: the chains advance in lock-step — step k synthesises segment k of every document
    that still has one, as one batch — so the serial dependency costs max(len(doc)) batched calls, not sum(len(doc))."""
    sr = model_manager.configs["sample_rate"]
    if not last_prompt:
        flat = [r for doc in documents for r in doc]
        wavs = model_manager.synthesize_batch(flat, **kw) if flat else []
        out, p = [], 0
        for doc in documents:
            seg = wavs[p: p + len(doc)]; p += len(doc)
            out.append(seg[0] if len(seg) == 1 else stitch_segments(seg, sr, rng))
        return out
    if reprompt is None:
        raise ValueError("last_prompt=True needs the zero-shot frontend callable `reprompt`")
    segs: List[List[torch.Tensor]] = [[] for _ in documents]
    prev: List[Optional[Dict]] = [None] * len(documents)
    for k in range(max((len(d) for d in documents), default=0)):
        live = [i for i, d in enumerate(documents) if k < len(d)]
        cur = [dict(documents[i][k]) if k == 0 else reprompt(dict(documents[i][k]), prev[i], segs[i][-1]) for i in live]
        wavs = model_manager.synthesize_batch(cur, **kw)
        for i, r, w in zip(live, cur, wavs):
            segs[i].append(w); prev[i] = r
    return [s[0] if len(s) == 1 else stitch_segments(s, sr, rng) for s in segs]
