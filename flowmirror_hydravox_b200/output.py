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
