"""CPU checks of the frontend feature path: the oracle against the reference fixture and against torchaudio, and
the host-side folding (window / DC removal / pre-emphasis / DFT as one basis) against the oracle in float64 — everything but the
CUDA kernels themselves (tests/test_frontend_gpu.py)."""
import torch

from flowmirror_hydravox_b200 import frontend as F
from oracle import frontend_ref


def test_oracle_matches_reference_fixture(golden):
    g = golden("frontend")
    basis = F.slaney_mel_basis(24000, 1920, 80, 0, 8000)
    mel = frontend_ref.mel_spectrogram(g["y24"], basis)
    assert mel.shape == g["mel"].shape == (1, 80, g["y24"].shape[1] // 480)
    assert (mel - g["mel"]).abs().max() < 1e-5
    fb = frontend_ref.kaldi_fbank(g["s16"])
    assert fb.shape == g["fbank"].shape and (fb - g["fbank"]).abs().max() < 1e-5


def test_kaldi_restatement_matches_torchaudio():
    import torchaudio.compliance.kaldi as kaldi
    s = torch.randn(1, 16000 + 123, generator=torch.Generator().manual_seed(3)) * 0.2
    ref = kaldi.fbank(s, num_mel_bins=80, dither=0, sample_frequency=16000)
    ours = frontend_ref.kaldi_fbank(s, subtract_mean=False)
    assert ours.shape == ref.shape and (ours - ref).abs().max() < 1e-4
    banks, _ = kaldi.get_mel_banks(80, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    mine = F.kaldi_mel_banks(80, 512, 16000.0)
    assert mine.shape == (80, 257) and torch.equal(mine[:, :256], banks) and float(mine[:, 256].abs().max()) == 0.0


def test_slaney_mel_basis_properties():
    fb = F.slaney_mel_basis(24000, 1920, 80, 0, 8000)
    assert fb.shape == (80, 961) and float(fb.min()) >= 0.0
    peak = fb.argmax(dim=1)
    assert bool((peak[1:] > peak[:-1]).all())                       # centres strictly increasing
    assert float(fb[:, int(8000 / 12.5) + 2:].abs().max()) == 0.0   # nothing above fmax (bin width 12.5 Hz)
    # slaney normalisation: each triangle integrates to ~1 over frequency in Hz
    area = fb.sum(dim=1) * 12.5
    assert float((area - 1.0).abs().max()) < 0.1


def _folded(wav, basis64, fb, frame_len, hop, pad, power, mag_eps, floor):
    w = wav.double()
    if pad:
        w = torch.nn.functional.pad(w[None, None], (pad, pad), mode="reflect")[0, 0]
    fr = w.unfold(0, frame_len, hop)
    spec = fr @ basis64
    nb = basis64.shape[1] // 2
    p = spec[:, :nb] ** 2 + spec[:, nb:] ** 2
    mag = p if power else torch.sqrt(p + mag_eps)
    return torch.log(torch.clamp(mag @ fb.double().T, min=floor))


def test_folded_basis_equals_oracle(golden):
    g = golden("frontend")
    win = torch.hann_window(1920, dtype=torch.float64)
    basis = F.dft_basis(1920, torch.diag(win))
    fb = F.slaney_mel_basis(24000, 1920, 80, 0, 8000)
    mel = _folded(g["y24"][0], basis, fb, 1920, 480, 720, False, 1e-9, 1e-5).T[None]
    assert (mel.float() - g["mel"]).abs().max() < 2e-4
    kb = F.kaldi_frame_map(400, 512) @ F.dft_basis(512)
    out = _folded(g["s16"][0], kb, F.kaldi_mel_banks(80, 512, 16000.0), 400, 160, 0, True, 0.0, torch.finfo(torch.float32).eps)
    out = out - out.mean(dim=0, keepdim=True)
    assert out.shape == g["fbank"].shape and (out.float() - g["fbank"]).abs().max() < 2e-3


def test_whisper_folded_basis_equals_oracle(golden):
    """centred reflect-padded frames x (hann . DFT) -> power -> slaney mel -> log10 / clamp / affine == the torch.stft restatement"""
    import math
    g = golden("frontend")
    audio = g["s16"][0]
    fb = F.slaney_mel_basis(16000, 400, 128, 0.0, 8000.0)
    ref = frontend_ref.whisper_log_mel(audio, fb)
    x = _folded(audio, F.dft_basis(400, torch.diag(torch.hann_window(400, dtype=torch.float64))), fb, 400, 160, 200, True, 0.0, 1e-10)
    x = (x[:-1] / math.log(10.0)).T
    x = (torch.maximum(x, x.max() - 8.0) + 4.0) / 4.0
    assert x.shape == ref.shape == (128, audio.numel() // 160)
    assert (x.float() - ref).abs().max() < 1e-4
