"""Zero-shot frontend features on the GPU: the 24 kHz prompt mel (`mel_spectrogram`,
matcha/utils/audio.py:42-82 — what cosyvoice/cli/frontend.py:117-122 calls `feat_extractor`) and the Kaldi fbank of the
speaker-embedding branch (cosyvoice/cli/frontend.py:108-112), both through hvx_frontend_fbank.

The host side builds, once per configuration and in float64, the linear map applied to a raw frame before the
non-linearity (window, DC removal, pre-emphasis, zero padding, DFT) and the filterbank; the device does
frames x basis -> magnitude/power -> filterbank -> log.  The speech tokenizer and CAM++ networks of the frontend are ONNX
files that are not part of the reference tree and are not rebuilt here."""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Tuple

import torch

from . import _lib as L


def dft_basis(n_fft: int, pre: torch.Tensor | None = None) -> torch.Tensor:
    """[frame_len][2*(n_fft//2+1)] float64: columns 0..n_bins-1 = cos(2 pi k n / n_fft), then -sin (the real and imaginary
    parts of rfft), left-multiplied by `pre` (frame_len x n_fft: the linear pre-processing of a frame, e.g. diag(window))."""
    n_bins = n_fft // 2 + 1
    n = torch.arange(n_fft, dtype=torch.int64)[:, None]
    k = torch.arange(n_bins, dtype=torch.int64)[None, :]
    ang = ((n * k) % n_fft).double() * (2.0 * math.pi / n_fft)          # reduce the angle in integers first
    F = torch.cat([torch.cos(ang), -torch.sin(ang)], dim=1)
    return F if pre is None else pre.double() @ F


def slaney_mel_basis(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> torch.Tensor:
    """librosa.filters.mel(htk=False, norm='slaney') — the filterbank matcha/utils/audio.py:53 asks librosa for (librosa is a
    pinned third-party dependency of the reference that is not installed here: restated from its published algorithm)."""
    def hz_to_mel(f):
        f = torch.as_tensor(f, dtype=torch.float64)
        f_sp, min_log_hz = 200.0 / 3, 1000.0
        min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0
        return torch.where(f >= min_log_hz, min_log_mel + torch.log(torch.clamp(f, min=1e-10) / min_log_hz) / logstep, f / f_sp)

    def mel_to_hz(m):
        f_sp, min_log_hz = 200.0 / 3, 1000.0
        min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0
        return torch.where(m >= min_log_mel, min_log_hz * torch.exp(logstep * (m - min_log_mel)), f_sp * m)

    fftfreqs = torch.linspace(0, sr / 2, 1 + n_fft // 2, dtype=torch.float64)
    mel_f = mel_to_hz(torch.linspace(float(hz_to_mel(fmin)), float(hz_to_mel(fmax)), n_mels + 2, dtype=torch.float64))
    fdiff = mel_f[1:] - mel_f[:-1]
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = torch.clamp(torch.minimum(lower, upper), min=0)
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (w * enorm[:, None]).float()


def kaldi_mel_banks(num_bins: int, padded: int, sr: float, low_freq: float = 20.0, high_freq: float = 0.0) -> torch.Tensor:
    """Kaldi mel banks without VTLN (torchaudio.compliance.kaldi.get_mel_banks + the zero Nyquist column fbank appends):
    (num_bins, padded//2 + 1).  Checked against torchaudio in tests/test_frontend_cpu.py."""
    nyq = 0.5 * sr
    if high_freq <= 0.0:
        high_freq += nyq
    mel = lambda f: 1127.0 * torch.log(1.0 + f / 700.0)
    n_fft_bins = padded // 2
    fft_bin_width = sr / padded
    mel_low, mel_high = 1127.0 * math.log(1.0 + low_freq / 700.0), 1127.0 * math.log(1.0 + high_freq / 700.0)
    delta = (mel_high - mel_low) / (num_bins + 1)
    b = torch.arange(num_bins, dtype=torch.float32)[:, None]
    left, center, right = mel_low + b * delta, mel_low + (b + 1.0) * delta, mel_low + (b + 2.0) * delta
    m = mel(fft_bin_width * torch.arange(n_fft_bins, dtype=torch.float32))[None, :]
    up, down = (m - left) / (center - left), (right - m) / (right - center)
    banks = torch.clamp(torch.minimum(up, down), min=0.0)
    return torch.nn.functional.pad(banks, (0, 1))


def kaldi_frame_map(frame_len: int, padded: int, preemph: float = 0.97, remove_dc: bool = True) -> torch.Tensor:
    """frame (1 x frame_len) -> windowed, zero-padded frame (1 x padded) as one matrix, float64: DC removal, pre-emphasis with
    the first sample replicated, povey window (torchaudio.compliance.kaldi._get_window, the published Kaldi order)."""
    N = frame_len
    M = torch.eye(N, dtype=torch.float64)
    if remove_dc:
        M = M @ (torch.eye(N, dtype=torch.float64) - torch.full((N, N), 1.0 / N, dtype=torch.float64))
    if preemph != 0.0:
        E = torch.eye(N, dtype=torch.float64)
        E[0, 0] -= preemph                                             # x'[0] = x[0] - c * x[0]
        E[torch.arange(N - 1), torch.arange(1, N)] = -preemph           # x'[i] = x[i] - c * x[i-1]
        M = M @ E
    w = torch.hann_window(N, periodic=False, dtype=torch.float64).pow(0.85)
    M = M * w[None, :]
    return torch.nn.functional.pad(M, (0, padded - N))


class _Fbank:
    def __init__(self, engine: "L.Engine", basis64: torch.Tensor, fb: torch.Tensor, frame_len: int, hop: int, pad: int, power: bool,
                 mag_eps: float, log_floor: float, subtract_mean: bool, channel_major: bool):
        self.engine = engine
        dev = engine.device
        self.basis = basis64.float().contiguous().to(dev)
        self.fb = fb.float().contiguous().to(dev)
        self.n_bins, self.n_mels = self.basis.shape[1] // 2, self.fb.shape[0]
        assert self.fb.shape[1] == self.n_bins and self.basis.shape[0] == frame_len
        self.frame_len, self.hop, self.pad = frame_len, hop, pad
        self.power, self.mag_eps, self.log_floor = int(power), float(mag_eps), float(log_floor)
        self.subtract_mean, self.channel_major = int(subtract_mean), int(channel_major)

    def n_frames(self, n_samples: int) -> int:
        return (n_samples + 2 * self.pad - self.frame_len) // self.hop + 1

    @torch.no_grad()
    def run(self, wav: torch.Tensor, drop_last: int = 0) -> torch.Tensor:
        dev = self.engine.device
        w = wav.reshape(-1).to(dev, torch.float32).contiguous()
        n = int(w.numel())
        nf = self.n_frames(n) - drop_last
        if nf < 1 or self.pad >= n:
            raise ValueError(f"waveform of {n} samples is too short for {self.frame_len}-sample frames")
        out = torch.empty((self.n_mels, nf) if self.channel_major else (nf, self.n_mels), device=dev, dtype=torch.float32)
        L.check(L.lib().hvx_frontend_fbank(self.engine.h, L.ptr(w), n, self.frame_len, self.hop, self.pad, L.ptr(self.basis), self.n_bins,
                                           L.ptr(self.fb), self.n_mels, self.power, C.c_float(self.mag_eps), C.c_float(self.log_floor),
                                           self.subtract_mean, self.channel_major, L.ptr(out), nf, L.stream_ptr()))
        return out


class MelSpectrogram:
    """`mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center=False)` of
    matcha/utils/audio.py:42-82 bound to its configuration (the `feat_extractor` partial of the reference's yaml:
    n_fft 1920, 80 mels, 24 kHz, hop 480, win 1920, fmin 0, fmax 8000): y (1, n) -> (1, num_mels, n // hop) fp32.
    `mel_basis` may be handed over (the reference's own librosa matrix) instead of the restated one."""

    def __init__(self, engine: "L.Engine", n_fft=1920, num_mels=80, sampling_rate=24000, hop_size=480, win_size=1920, fmin=0, fmax=8000,
                 center=False, mel_basis: torch.Tensor | None = None):
        if center or win_size > n_fft:
            raise ValueError("only the reference's call shape is built: center=False, win_size <= n_fft")
        win = torch.hann_window(win_size, dtype=torch.float64)         # periodic, as torch.hann_window(win_size) (:55)
        lpad = (n_fft - win_size) // 2                                 # torch.stft centres a short window inside n_fft
        pre = torch.zeros(n_fft, n_fft, dtype=torch.float64)
        pre[torch.arange(lpad, lpad + win_size), torch.arange(lpad, lpad + win_size)] = win
        fb = slaney_mel_basis(sampling_rate, n_fft, num_mels, fmin, fmax) if mel_basis is None else mel_basis
        self.impl = _Fbank(engine, dft_basis(n_fft, pre), fb, n_fft, hop_size, int((n_fft - hop_size) / 2), power=False, mag_eps=1e-9,
                           log_floor=1e-5, subtract_mean=False, channel_major=True)

    def __call__(self, y: torch.Tensor) -> torch.Tensor:
        if y.dim() != 2 or y.shape[0] != 1:
            raise ValueError("expected a (1, n_samples) waveform, as load_wav returns it")
        return self.impl.run(y)[None]


class KaldiFbank:
    """`kaldi.fbank(speech, num_mel_bins=80, dither=0, sample_frequency=16000)` followed by `feat - feat.mean(dim=0)`
    (cosyvoice/cli/frontend.py:108-112): speech (1, n) -> (n_frames, num_mel_bins) fp32.  Kaldi defaults: 25 ms frames,
    10 ms shift, snip_edges, DC removal, pre-emphasis 0.97, povey window, power spectrum, low_freq 20 Hz."""

    def __init__(self, engine: "L.Engine", num_mel_bins=80, sample_frequency=16000, frame_length_ms=25.0, frame_shift_ms=10.0,
                 subtract_mean=True):
        flen, hop = int(sample_frequency * frame_length_ms * 0.001), int(sample_frequency * frame_shift_ms * 0.001)
        padded = 1 << (flen - 1).bit_length()
        basis = kaldi_frame_map(flen, padded) @ dft_basis(padded)
        fb = kaldi_mel_banks(num_mel_bins, padded, float(sample_frequency))
        self.impl = _Fbank(engine, basis, fb, flen, hop, 0, power=True, mag_eps=0.0, log_floor=torch.finfo(torch.float32).eps,
                           subtract_mean=subtract_mean, channel_major=False)

    def __call__(self, speech: torch.Tensor) -> torch.Tensor:
        return self.impl.run(speech)


class WhisperLogMel:
    """`whisper.log_mel_spectrogram(speech, n_mels=128)` as cosyvoice/cli/frontend.py:95 calls it (the input of the speech
    tokenizer): 16 kHz, n_fft 400, hop 160, centred reflect-padded hann STFT, power spectrum without the last frame, slaney mel
    filterbank, log10 with a 1e-10 floor, clamp to 8 below the maximum, (x + 4) / 4.  speech (1, n) -> (1, n_mels, n // 160).
    whisper is a pinned dependency of the reference that is not installed here: restated from its published algorithm
    (oracle/frontend_ref.py: whisper_log_mel), parity unpinned; `filters` may be handed over (whisper's mel_filters.npz)."""

    def __init__(self, engine: "L.Engine", n_mels: int = 128, filters: torch.Tensor | None = None):
        win = torch.hann_window(400, dtype=torch.float64)
        fb = slaney_mel_basis(16000, 400, n_mels, 0.0, 8000.0) if filters is None else filters
        self.impl = _Fbank(engine, dft_basis(400, torch.diag(win)), fb, 400, 160, 200, power=True, mag_eps=0.0, log_floor=1e-10,
                           subtract_mean=False, channel_major=True)

    def __call__(self, speech: torch.Tensor) -> torch.Tensor:
        out = self.impl.run(speech, drop_last=1)                       # stft[..., :-1]
        L.check(L.lib().hvx_frontend_whisper_post(self.impl.engine.h, L.ptr(out), int(out.numel()), C.c_float(1.0 / math.log(10.0)),
                                                  C.c_float(8.0), C.c_float(4.0), C.c_float(4.0), L.stream_ptr()))
        return out[None] if speech.dim() == 2 else out


class NativeFrontendFeatures:
    """The two feature extractors of `CosyVoiceFrontEnd` that feed the hot path, with the reference's method names and return
    shapes (cosyvoice/cli/frontend.py:104-122), already-resampled waveforms in (load_wav / resampling stay on the host side):
      _extract_speech_feat(speech_24k (1, n))  -> (speech_feat (1, n // 480, 80) on the device, speech_feat_len int32 (1,))
      _extract_spk_fbank(speech_16k (1, n))    -> mean-normalised fbank (n_frames, 80): the input of the CAM++ session."""

    def __init__(self, engine: "L.Engine", mel_basis: torch.Tensor | None = None):
        self.engine = engine
        self.feat_extractor = MelSpectrogram(engine, mel_basis=mel_basis)
        self.fbank = KaldiFbank(engine)

    def _extract_speech_feat(self, speech: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        feat = self.feat_extractor(speech).squeeze(dim=0).transpose(0, 1).unsqueeze(dim=0)
        return feat, torch.tensor([feat.shape[1]], dtype=torch.int32, device=feat.device)

    def _extract_spk_fbank(self, speech: torch.Tensor) -> torch.Tensor:
        return self.fbank(speech)

    @staticmethod
    def align_prompt(speech_feat: torch.Tensor, speech_token: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """frontend_zero_shot's `force speech_feat % speech_token = 2` (cosyvoice/cli/frontend.py:170-174)"""
        token_len = min(int(speech_feat.shape[1] / 2), speech_token.shape[1])
        return speech_feat[:, : 2 * token_len], speech_token[:, :token_len]
