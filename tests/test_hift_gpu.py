"""HiFT stage on the B200 through the C-ABI vs the CPU oracle and the reference fixtures."""
import pytest
import torch

from flowmirror_hydravox_b200 import dims as D, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engines():
    from flowmirror_hydravox_b200 import _lib as L
    from flowmirror_hydravox_b200.hift import NativeHiFT
    out = {}
    for name, hd in (("tiny", D.HIFT_TINY), ("full", D.HIFT_FULL)):
        e = L.Engine(hd=hd)
        h = NativeHiFT(e)
        h.load_state_dict(synth.hift_state_dict(hd, 0))
        out[name] = (e, h, hd)
    yield out
    for e, _, _ in out.values():
        e.close()


@pytest.mark.parametrize("name", ["tiny", "full"])
def test_hift_matches_reference_fixture(engines, golden, name):
    e, h, hd = engines[name]
    g = golden(f"hift_{name}")
    h.set_sine_table(synth.hift_sine_table(hd, g["T"]))
    # F0 predictor alone (fp32 convs): relative to the reference's CPU result
    wav, src, f0 = h.inference(g["mel"], return_f0=True)
    rel = ((f0.cpu() - g["f0"][0]).abs() / (g["f0"][0].abs() + 1)).max().item()
    assert rel < 2e-4, rel
    # waveform with the F0 track pinned (see oracle/hift_ref.inference docstring)
    wav, src = h.inference(g["mel"], f0=g["f0"])
    assert wav.shape == g["wav"].shape
    err = (wav.cpu() - g["wav"]).abs()
    rms = err.pow(2).mean().sqrt().item()
    assert (src.cpu() - g["src"]).abs().max().item() < 1e-5
    assert rms < 1e-4 and err.max().item() < 1e-3, (rms, err.max().item())     # north_star: <=1e-4 RMS on waveform
    # streaming branch (finalize=False)
    wav_s, _ = h.inference(g["mel"], finalize=False, f0=g["f0_stream"])
    assert wav_s.shape == g["wav_stream"].shape
    assert (wav_s.cpu() - g["wav_stream"]).pow(2).mean().sqrt().item() < 1e-4


@pytest.mark.parametrize("T", [1, 2, 9, 300])
def test_hift_matches_oracle_lengths(engines, T):
    from oracle import hift_ref
    e, h, hd = engines["tiny"]
    sd = synth.hift_state_dict(hd, 0)
    table = synth.hift_sine_table(hd, T)
    h.set_sine_table(table)
    g = torch.Generator().manual_seed(T)
    mel = torch.rand(1, hd.mel, T, generator=g) * 6 - 6
    w = hift_ref.fold_weight_norm(sd)
    f0 = hift_ref.f0_predict(w, mel)
    ref, _ = hift_ref.inference(sd, mel, table, hd, f0=f0)
    wav, _ = h.inference(mel, f0=f0)
    assert wav.shape == ref.shape == (1, T * hd.frame_samples)
    assert (wav.cpu() - ref).pow(2).mean().sqrt().item() < 1e-4


def test_hift_full_size_properties(engines):
    """BASELINE config-2 size (2048 frames = 40.96 s): causal-vocoder property of the reference's own
    self-check (generator.py:729-746) — a streamed prefix equals the offline result."""
    e, h, hd = engines["full"]
    T = 2048
    h.set_sine_table(synth.hift_sine_table(hd, T))
    g = torch.Generator().manual_seed(5)
    mel = torch.rand(1, hd.mel, T, generator=g) * 6 - 6
    wav, src, f0 = h.inference(mel, return_f0=True)
    assert wav.shape == (1, T * 480) and torch.isfinite(wav).all() and wav.abs().max() <= 0.99
    Tc = 520
    f0c = torch.cat([f0[: Tc - 3]])
    wav_c, _ = h.inference(mel[:, :, :Tc], finalize=False, f0=f0c)
    n = wav_c.shape[1]
    assert n == (Tc - 8) * 480
    assert (wav_c - wav[:, :n]).abs().max().item() < 2e-3
