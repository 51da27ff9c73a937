"""NativeHiFT — drop-in for models['hift'] (CausalHiFTGenerator, cosyvoice/hifigan/generator.py:572-726)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .weights import pack_hift, pack_hift_t, pack_hifigan


class NativeHiFT:
    def __init__(self, engine: "L.Engine", sine_table: torch.Tensor | None = None):
        self.engine = engine
        self.dims = engine.hd
        self.sine_table = None
        if sine_table is not None:
            self.set_sine_table(sine_table)

    # ---- nn.Module-like surface used by ModelManager.load_models (infer_speech_model.py:92-118)
    def load_state_dict(self, sd, strict=True):
        self.engine.set_tensors(L.STAGE_HIFT, pack_hift(sd, self.dims))
        self.engine.finalize(L.STAGE_HIFT)
        return self

    def eval(self):
        return self

    def cuda(self):
        return self

    def to(self, *a, **k):
        return self

    def set_sine_table(self, table: torch.Tensor):
        """SineGen2.sine_waves rows (generator.py:226): (n_samples, harmonics) uniform[0,1)."""
        t = table.reshape(-1, self.dims.harmonics).to(self.engine.device, torch.float32).contiguous()
        self.sine_table = t

    @torch.no_grad()
    def inference(self, speech_feat: torch.Tensor, finalize: bool = True, f0: torch.Tensor | None = None,
                  return_f0: bool = False):
        """speech_feat (1, mel, T) fp32 -> (speech (1, n), source (1, 1, frame*T_f0))  [generator.py:713-726]."""
        assert speech_feat.dim() == 3 and speech_feat.shape[0] == 1, "reference asserts batch 1 (flow.py:387)"
        d, dev = self.dims, self.engine.device
        mel = speech_feat[0].to(dev, torch.float32).contiguous()
        T = mel.shape[1]
        frame = d.frame_samples
        Tf0 = T if finalize else T - 3
        Tx = T if finalize else T - 7
        n_out = Tx * frame if finalize else (Tx - 1) * frame
        if self.sine_table is None or self.sine_table.shape[0] < Tf0 * frame:
            raise L.HvxError("sine table missing or shorter than the utterance (set_sine_table)")
        wav = torch.empty(1, n_out, device=dev, dtype=torch.float32)
        src = torch.empty(1, 1, Tf0 * frame, device=dev, dtype=torch.float32)
        f0_out = torch.empty(Tf0, device=dev, dtype=torch.float32)
        f0_in = None if f0 is None else f0.reshape(-1).to(dev, torch.float32).contiguous()
        L.check(L.lib().hvx_hift_vocode(self.engine.h, L.ptr(mel), T, int(bool(finalize)), L.ptr(self.sine_table),
                                        C.c_int64(int(self.sine_table.shape[0])), L.ptr(f0_in), L.ptr(f0_out), L.ptr(wav), L.ptr(src), L.stream_ptr()))
        if return_f0:
            return wav, src, f0_out
        return wav, src


class NativeHiFTTransposed:
    """Drop-in for the non-causal `HiFTGenerator` (cosyvoice/hifigan/generator.py:378-569): ConvTranspose1d up-sampling,
    "same"-padded ResBlocks, ConvRNNF0Predictor on the GPU.  `inference(speech_feat, cache_source)` has the reference's
    signature (:557-569); the source module's Gaussian draw (:310) comes from `generator` (a torch.Generator on the engine's
    device, fresh noise per call like the reference) or from an explicit `noise` (n_samples, harmonics) — the pinned-RNG path."""

    def __init__(self, engine: "L.Engine", generator: torch.Generator | None = None):
        self.engine = engine
        self.dims = engine.hd
        self.generator = generator

    def load_state_dict(self, sd, strict=True):
        self.engine.set_tensors(L.STAGE_HIFT, pack_hift_t(sd, self.dims))
        self.engine.finalize(L.STAGE_HIFT)
        return self

    def eval(self):
        return self

    def cuda(self):
        return self

    def to(self, *a, **k):
        return self

    @torch.no_grad()
    def inference(self, speech_feat: torch.Tensor, cache_source: torch.Tensor | None = None, noise: torch.Tensor | None = None,
                  f0: torch.Tensor | None = None, return_f0: bool = False):
        """speech_feat (1, mel, T) fp32 -> (speech (1, frame*T), source (1, 1, frame*T))."""
        assert speech_feat.dim() == 3 and speech_feat.shape[0] == 1
        d, dev = self.dims, self.engine.device
        mel = speech_feat[0].to(dev, torch.float32).contiguous()
        T = mel.shape[1]
        n = T * d.frame_samples
        cache = None
        if cache_source is not None and cache_source.numel():
            cache = cache_source.reshape(-1)[:n].to(dev, torch.float32).contiguous()
        n_cache = 0 if cache is None else cache.numel()
        if noise is None and n_cache < n:
            noise = torch.randn(n, d.harmonics, device=dev, dtype=torch.float32, generator=self.generator)
        if noise is not None:
            noise = noise.reshape(-1, d.harmonics)[:n].to(dev, torch.float32).contiguous()
            if noise.shape[0] < n:
                raise L.HvxError("noise shorter than the utterance")
        wav = torch.empty(1, n, device=dev, dtype=torch.float32)
        src = torch.empty(1, 1, n, device=dev, dtype=torch.float32)
        f0_out = torch.empty(T, device=dev, dtype=torch.float32)
        f0_in = None if f0 is None else f0.reshape(-1).to(dev, torch.float32).contiguous()
        L.check(L.lib().hvx_hift_t_vocode(self.engine.h, L.ptr(mel), T, L.ptr(noise), L.ptr(cache), n_cache, L.ptr(f0_in),
                                          L.ptr(f0_out), L.ptr(wav), L.ptr(src), L.stream_ptr()))
        if return_f0:
            return wav, src, f0_out
        return wav, src

    @torch.no_grad()
    def decode(self, x: torch.Tensor, s: torch.Tensor):
        """HiFTGenerator.decode (:506-540): mel (1, mel, T) and an explicit source s (1, 1, frame*T) -> speech (1, frame*T)."""
        return self.inference(x, cache_source=s)[0]


class NativeHiFiGAN:
    """Drop-in for the classic HiFi-GAN `Generator` (matcha/hifigan/models.py:148-193): `forward(x)` / `__call__` take a mel
    (B, mel, T) fp32 and return the waveform (B, 1, prod(upsample_rates)*T) in (-1, 1), one hvx_hifigan_vocode per batch row.
    Build the engine with `hd=dims.HIFIGAN_V1` (or any HiftDims carrying this generator's rates / kernel sizes)."""

    def __init__(self, engine: "L.Engine"):
        self.engine = engine
        self.dims = engine.hd

    def load_state_dict(self, sd, strict=True):
        self.engine.set_tensors(L.STAGE_HIFT, pack_hifigan(sd, self.dims))
        return self

    def eval(self):
        return self

    def cuda(self):
        return self

    def to(self, *a, **k):
        return self

    def remove_weight_norm(self):          # models.py:195-203 — the packed weights are already folded
        return None

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert x.dim() == 3 and x.shape[1] == self.dims.mel
        dev = self.engine.device
        mel = x.to(dev, torch.float32).contiguous()
        B, _, T = mel.shape
        wav = torch.empty(B, 1, T * self.dims.frame_samples, device=dev, dtype=torch.float32)
        for b in range(B):
            L.check(L.lib().hvx_hifigan_vocode(self.engine.h, L.ptr(mel[b]), T, L.ptr(wav[b]), L.stream_ptr()))
        return wav

    __call__ = forward
