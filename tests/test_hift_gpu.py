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


@pytest.mark.parametrize("name,T", [("tiny", 37), ("tiny", 1), ("full", 64), ("full", 3)])
@pytest.mark.parametrize("finalize", [True, False])
def test_hift_tensor_core_decode_equals_fp32_decode(engines, monkeypatch, name, T, finalize):
    """The decode stack as split-fp16 implicit GEMMs on tcgen05 (default) against the fp32 CUDA-core convolutions
    (HVX_HIFT_FP32=1) on the same F0 track: two implementations of the same arithmetic to ~22 mantissa bits per operand."""
    e, h, hd = engines[name]
    if not finalize and T < 9:
        pytest.skip("streaming branch needs > 8 frames")
    h.set_sine_table(synth.hift_sine_table(hd, T))
    g = torch.Generator().manual_seed(100 + T)
    mel = torch.rand(1, hd.mel, T, generator=g) * 6 - 6
    monkeypatch.setenv("HVX_HIFT_FP32", "1")
    ref, _, f0 = h.inference(mel, finalize=finalize, return_f0=True)
    monkeypatch.setenv("HVX_HIFT_FP32", "0")
    wav, _ = h.inference(mel, finalize=finalize, f0=f0)
    d = (wav - ref)
    rms, mx = d.pow(2).mean().sqrt().item(), d.abs().max().item()
    print(f"[hift tc vs fp32 {name} T={T} finalize={finalize}] rms {rms:.3e} max-abs {mx:.3e} (wav rms {ref.pow(2).mean().sqrt():.3f})")
    assert wav.shape == ref.shape and torch.isfinite(wav).all()
    assert rms < 2e-5 and mx < 5e-4


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


# ---------------------------------------------------------------- a12': non-causal ConvTranspose1d HiFTGenerator
@pytest.fixture(scope="module")
def engines_t():
    from flowmirror_hydravox_b200 import _lib as L
    from flowmirror_hydravox_b200.hift import NativeHiFTTransposed
    out = {}
    for name, hd in (("tiny", D.HIFT_TINY), ("full", D.HIFT_FULL)):
        e = L.Engine(hd=hd)
        h = NativeHiFTTransposed(e)
        h.load_state_dict(synth.hift_t_state_dict(hd, 0))
        out[name] = (e, h, hd)
    yield out
    for e, _, _ in out.values():
        e.close()


@pytest.mark.parametrize("name", ["tiny", "full"])
def test_hift_transposed_matches_reference_fixture(engines_t, golden, name):
    e, h, hd = engines_t[name]
    g = golden(f"hift_t_{name}")
    # decode(mel, s) with an explicit source: the deterministic part of HiFTGenerator (generator.py:506-540)
    wav = h.decode(g["mel"], g["s"])
    assert wav.shape == g["wav"].shape
    err = (wav.cpu() - g["wav"]).abs()
    assert err.pow(2).mean().sqrt().item() < 1e-4 and err.max().item() < 1e-3, (err.pow(2).mean().sqrt().item(), err.max().item())
    # F0 predictor (ConvRNNF0Predictor) on the GPU vs the reference's fp32 result
    _, _, f0 = h.inference(g["mel"], noise=g["noise"], return_f0=True)
    rel = ((f0.cpu() - g["f0"][0]).abs() / (g["f0"][0].abs() + 1)).max().item()
    assert rel < 2e-4, rel
    # whole inference, RNG pinned (noise explicit), F0 pinned: source module + decode
    wav_i, src = h.inference(g["mel"], noise=g["noise"], f0=g["f0"])
    assert (src.cpu() - g["src_inf"]).abs().max().item() < 5e-4      # phase ~1e4 rad in fp32: 1 ulp = 1e-3 rad x 0.1 amplitude
    e_i = (wav_i.cpu() - g["wav_inf"]).abs()
    assert e_i.pow(2).mean().sqrt().item() < 1e-4, e_i.pow(2).mean().sqrt().item()
    # cache_source overwrites the head of the source (:566-567)
    cache = g["s"][:, :, : 3 * hd.frame_samples]
    wav_c, src_c = h.inference(g["mel"], cache_source=cache, noise=g["noise"], f0=g["f0"])
    assert torch.equal(src_c.cpu()[:, :, : cache.shape[2]], cache)
    assert (wav_c.cpu() - g["wav_cache"]).pow(2).mean().sqrt().item() < 1e-4


@pytest.mark.parametrize("T", [1, 2, 7, 150])
def test_hift_transposed_matches_oracle_lengths(engines_t, T):
    from oracle import hift_ref
    e, h, hd = engines_t["tiny"]
    sd = synth.hift_t_state_dict(hd, 0)
    g = torch.Generator().manual_seed(T)
    mel = torch.rand(1, hd.mel, T, generator=g) * 6 - 6
    noise = torch.randn(T * hd.frame_samples, hd.harmonics, generator=g)
    w = hift_ref.fold_weight_norm(sd)
    f0 = hift_ref.f0_predict_nc(w, mel)
    ref, s_ref = hift_ref.inference_transposed(sd, mel, noise, hd, f0=f0)
    wav = h.decode(mel, s_ref)                         # source pinned: decode parity at any length
    assert wav.shape == ref.shape == (1, T * hd.frame_samples)
    assert (wav.cpu() - ref).pow(2).mean().sqrt().item() < 1e-4
    _, src = h.inference(mel, noise=noise, f0=f0)
    tol = 5e-4 if T <= 7 else 2e-2                      # fp32 phase grows with the clip: 150 frames reach ~1e5 rad
    assert (src.cpu() - s_ref).abs().max().item() < tol


def test_hift_transposed_fresh_noise_is_stochastic_like_reference(engines_t):
    """Without pinned noise every call draws fresh Gaussian noise, as SineGen2 does (generator.py:310)."""
    e, h, hd = engines_t["tiny"]
    mel = torch.rand(1, hd.mel, 12, generator=torch.Generator().manual_seed(0)) * 6 - 6
    a, _ = h.inference(mel)
    b, _ = h.inference(mel)
    assert a.shape == (1, 12 * hd.frame_samples) and torch.isfinite(a).all() and not torch.equal(a, b)


# ---------------------------------------------------------------- a12' (second variant): classic HiFi-GAN Generator
@pytest.fixture(scope="module")
def engines_gan():
    from flowmirror_hydravox_b200 import _lib as L
    from flowmirror_hydravox_b200.hift import NativeHiFiGAN
    out = {}
    for name, hd in (("tiny", D.HIFIGAN_TINY), ("v1", D.HIFIGAN_V1)):
        e = L.Engine(hd=hd)
        h = NativeHiFiGAN(e)
        h.load_state_dict(synth.hifigan_state_dict(hd, 0))
        out[name] = (e, h, hd)
    yield out
    for e, _, _ in out.values():
        e.close()


@pytest.mark.parametrize("name", ["tiny", "v1"])
def test_hifigan_matches_reference_fixture(engines_gan, golden, name):
    """hvx_hifigan_vocode vs the unmodified matcha.hifigan.models.Generator (tests/golden/hifigan_*.pt), batch of 2."""
    e, h, hd = engines_gan[name]
    g = golden(f"hifigan_{name}")
    wav = h(g["mel"])
    assert wav.shape == g["wav"].shape
    err = (wav.cpu() - g["wav"]).abs()
    assert err.pow(2).mean().sqrt().item() < 1e-5 and err.max().item() < 1e-4, (err.pow(2).mean().sqrt().item(), err.max().item())


@pytest.mark.parametrize("T", [1, 2, 5, 300])
def test_hifigan_matches_oracle_lengths(engines_gan, T):
    from oracle import hifigan_ref
    e, h, hd = engines_gan["tiny"]
    sd = synth.hifigan_state_dict(hd, 0)
    mel = torch.rand(1, hd.mel, T, generator=torch.Generator().manual_seed(T)) * 6 - 6
    ref = hifigan_ref.generator(sd, mel, hd)
    wav = h(mel)
    assert wav.shape == ref.shape == (1, 1, T * hd.frame_samples)
    assert (wav.cpu() - ref).abs().max().item() < 1e-4


def test_hifigan_rejects_hift_weights(engines_gan):
    """an engine whose HiFT stage holds the 18-channel ISTFT head must not be run as a classic HiFi-GAN"""
    from flowmirror_hydravox_b200 import _lib as L
    from flowmirror_hydravox_b200.hift import NativeHiFT, NativeHiFiGAN
    e = L.Engine(hd=D.HIFT_TINY)
    try:
        NativeHiFT(e).load_state_dict(synth.hift_state_dict(D.HIFT_TINY, 0))
        with pytest.raises(L.HvxError):
            NativeHiFiGAN(e)(torch.zeros(1, D.HIFT_TINY.mel, 4))
    finally:
        e.close()
