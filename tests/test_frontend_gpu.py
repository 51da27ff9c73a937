"""GPU parity of hvx_frontend_fbank: the 24 kHz prompt mel against the reference's mel_spectrogram fixture, the
Kaldi fbank against torchaudio, plus ragged lengths against the oracle."""
import pytest
import torch

from flowmirror_hydravox_b200 import frontend as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fe():
    from flowmirror_hydravox_b200 import _lib as L
    e = L.Engine()
    yield e, F.MelSpectrogram(e), F.KaldiFbank(e)
    e.close()


def test_mel_matches_reference_fixture(fe, golden):
    e, mel, _ = fe
    g = golden("frontend")
    out = mel(g["y24"]).cpu()
    assert out.shape == g["mel"].shape
    err = (out - g["mel"]).abs()
    print(f"[frontend mel] max-abs {err.max():.3e} mean-abs {err.mean():.3e}")
    assert err.max().item() < 1e-3                       # north_star: 1e-3 on mel frames (this mel is the flow's prompt_feat)


def test_fbank_matches_torchaudio_fixture(fe, golden):
    e, _, fbank = fe
    g = golden("frontend")
    out = fbank(g["s16"]).cpu()
    assert out.shape == g["fbank"].shape
    err = (out - g["fbank"]).abs()
    print(f"[frontend fbank] max-abs {err.max():.3e} mean-abs {err.mean():.3e}")
    assert err.max().item() < 2e-3 and err.mean().item() < 1e-4


@pytest.mark.parametrize("n", [960, 2399, 24000 + 7, 5 * 24000])
def test_mel_lengths_vs_oracle(fe, n):
    from oracle import frontend_ref
    e, mel, _ = fe
    y = (torch.rand(1, n, generator=torch.Generator().manual_seed(n)) * 2 - 1) * 0.5
    ref = frontend_ref.mel_spectrogram(y, F.slaney_mel_basis(24000, 1920, 80, 0, 8000))
    out = mel(y).cpu()
    assert out.shape == ref.shape == (1, 80, n // 480)
    assert (out - ref).abs().max().item() < 1e-3


@pytest.mark.parametrize("n", [400, 559, 16000, 30 * 16000])
def test_fbank_lengths_vs_oracle(fe, n):
    from oracle import frontend_ref
    e, _, fbank = fe
    s = torch.randn(1, n, generator=torch.Generator().manual_seed(n)) * 0.1
    ref = frontend_ref.kaldi_fbank(s)
    out = fbank(s).cpu()
    assert out.shape == ref.shape == (1 + (n - 400) // 160, 80)
    assert (out - ref).abs().max().item() < 2e-3


def test_frontend_rejects_short_input(fe):
    e, mel, fbank = fe
    with pytest.raises(ValueError):
        fbank(torch.zeros(1, 399))
    with pytest.raises(ValueError):
        mel(torch.zeros(1, 100))


def test_frontend_feature_objects_mirror_reference_shapes(fe, golden):
    """_extract_speech_feat returns (1, T, 80) + length like cosyvoice/cli/frontend.py:117-122; align_prompt trims to 2 frames
    per prompt token (:170-174)"""
    e, _, _ = fe
    g = golden("frontend")
    ff = F.NativeFrontendFeatures(e)
    feat, n = ff._extract_speech_feat(g["y24"])
    assert feat.shape == (1, g["mel"].shape[2], 80) and int(n[0]) == feat.shape[1]
    assert (feat.cpu() - g["mel"].transpose(1, 2)).abs().max().item() < 1e-3
    tok = torch.arange(45)[None]
    f2, t2 = ff.align_prompt(feat, tok)
    assert f2.shape[1] == 2 * t2.shape[1] == 90
    fb = ff._extract_spk_fbank(g["s16"])
    assert fb.shape == g["fbank"].shape and abs(float(fb.mean())) < 1e-4


@pytest.mark.parametrize("n", [400, 16000 + 33, 30 * 16000])
def test_whisper_log_mel_vs_oracle(fe, n):
    """hvx_frontend_fbank + hvx_frontend_whisper_post vs the restated whisper.log_mel_spectrogram (unpinned: whisper is not installed)"""
    from oracle import frontend_ref
    e, _, _ = fe
    wl = F.WhisperLogMel(e)
    s = torch.randn(1, n, generator=torch.Generator().manual_seed(n)) * 0.1
    s[:, : n // 3] *= 1e-4                                            # a quiet stretch: exercises the max-8 clamp
    ref = frontend_ref.whisper_log_mel(s, F.slaney_mel_basis(16000, 400, 128, 0.0, 8000.0))
    out = wl(s).cpu()
    assert out.shape == ref.shape == (1, 128, n // 160)
    assert (out - ref).abs().max().item() < 1e-3, (out - ref).abs().max().item()
    assert abs(float(out.min()) - float(ref.min())) < 1e-4            # the clamp floor itself
