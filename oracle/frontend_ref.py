"""TEST INFRASTRUCTURE ONLY — CPU restatements of the two filterbank front ends of the zero-shot frontend.

  * mel_spectrogram: matcha/utils/audio.py:42-82 (reflect pad, torch.stft with a periodic hann window, sqrt(|X|^2 + 1e-9),
    mel matmul, log(clamp(., 1e-5))).  The mel matrix is an argument: the reference obtains it from librosa (pinned in its
    requirements, not installed here; restated in flowmirror_hydravox_b200.frontend.slaney_mel_basis — unpinned).  Pinned
    against the reference function itself by oracle/make_golden.py (tests/golden/frontend.pt).
  * kaldi_fbank: torchaudio.compliance.kaldi.fbank as cosyvoice/cli/frontend.py:108-112 calls it (num_mel_bins=80, dither=0,
    16 kHz, Kaldi defaults) plus the mean subtraction — restated op by op; torchaudio is importable wherever the tests run,
    so tests/test_frontend_cpu.py pins this restatement against the library directly.
"""
from __future__ import annotations

import math

import torch


def mel_spectrogram(y, mel_basis, n_fft=1920, hop_size=480, win_size=1920):
    pad = int((n_fft - hop_size) / 2)
    y = torch.nn.functional.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    spec = torch.view_as_real(torch.stft(y, n_fft, hop_length=hop_size, win_length=win_size, window=torch.hann_window(win_size),
                                         center=False, pad_mode="reflect", normalized=False, onesided=True, return_complex=True))
    spec = torch.sqrt(spec.pow(2).sum(-1) + 1e-9)
    return torch.log(torch.clamp(torch.matmul(mel_basis, spec), min=1e-5))


def kaldi_fbank(speech, num_mel_bins=80, sample_frequency=16000.0, frame_length=25.0, frame_shift=10.0, preemph=0.97,
                low_freq=20.0, subtract_mean=True):
    """speech (1, n) -> (m, num_mel_bins)"""
    wav = speech[0].float()
    flen, hop = int(sample_frequency * frame_length * 0.001), int(sample_frequency * frame_shift * 0.001)
    padded = 1 << (flen - 1).bit_length()
    m = 1 + (wav.numel() - flen) // hop                                     # snip_edges
    fr = wav.unfold(0, flen, hop)[:m].clone()
    fr = fr - fr.mean(dim=1, keepdim=True)                                  # remove_dc_offset
    prev = torch.cat([fr[:, :1], fr[:, :-1]], dim=1)                        # replicate-padded shift
    fr = fr - preemph * prev
    fr = fr * torch.hann_window(flen, periodic=False).pow(0.85)             # povey
    fr = torch.nn.functional.pad(fr, (0, padded - flen))
    spec = torch.fft.rfft(fr).abs().pow(2.0)
    nyq = 0.5 * sample_frequency
    mel = lambda f: 1127.0 * torch.log(1.0 + f / 700.0)
    mlo, mhi = 1127.0 * math.log(1.0 + low_freq / 700.0), 1127.0 * math.log(1.0 + nyq / 700.0)
    delta = (mhi - mlo) / (num_mel_bins + 1)
    b = torch.arange(num_mel_bins, dtype=torch.float32)[:, None]
    left, center, right = mlo + b * delta, mlo + (b + 1.0) * delta, mlo + (b + 2.0) * delta
    mf = mel((sample_frequency / padded) * torch.arange(padded // 2, dtype=torch.float32))[None, :]
    banks = torch.clamp(torch.minimum((mf - left) / (center - left), (right - mf) / (right - center)), min=0.0)
    banks = torch.nn.functional.pad(banks, (0, 1))
    feat = torch.clamp(spec @ banks.T, min=torch.finfo(torch.float32).eps).log()
    return feat - feat.mean(dim=0, keepdim=True) if subtract_mean else feat


def whisper_log_mel(audio, filters):
    """whisper.log_mel_spectrogram (whisper/audio.py; called at cosyvoice/cli/frontend.py:95) restated from the published
    algorithm — the package is not installed here, parity unpinned.  audio (n,) or (1, n); filters (n_mels, 201)."""
    stft = torch.stft(audio, 400, 160, window=torch.hann_window(400), return_complex=True)
    magnitudes = stft[..., :-1].abs() ** 2
    log_spec = torch.clamp(filters @ magnitudes, min=1e-10).log10()
    log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    return (log_spec + 4.0) / 4.0
