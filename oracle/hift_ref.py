"""TEST INFRASTRUCTURE ONLY — CPU restatement (fp32 torch) of the HiFT vocoder stage.

Follows, function by function:
  cosyvoice/hifigan/generator.py:572-726  CausalHiFTGenerator (decode :672-711, inference :713-726)
  cosyvoice/hifigan/generator.py:192-375  SineGen2 / SourceModuleHnNSF (eval + causal branch)
  cosyvoice/hifigan/generator.py:491-505  _stft / _istft
  cosyvoice/hifigan/generator.py:46-117   ResBlock
  cosyvoice/hifigan/f0_predictor.py:62-103 CausalConvRNNF0Predictor
  cosyvoice/transformer/convolution.py:150-258 CausalConv1d / DownSample / Upsample
  cosyvoice/transformer/activation.py:73-84 Snake
Pinned against the real reference modules by oracle/make_golden.py (container) and
against tests/golden/hift_*.pt everywhere (tests/test_oracle_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


def fold_weight_norm(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """w = g * v / ||v||_2 over (in, k) per output channel (torch weight_norm, dim=0)."""
    out = {}
    for k, v in sd.items():
        if k.endswith("parametrizations.weight.original1"):
            base = k[: -len("parametrizations.weight.original1")]
            g = sd[base + "parametrizations.weight.original0"].float()
            vv = v.float()
            norm = vv.reshape(vv.shape[0], -1).norm(2, 1).reshape(-1, 1, 1)
            out[base + "weight"] = vv * (g / norm)          # same op order as torch._weight_norm
        elif k.endswith("parametrizations.weight.original0"):
            continue
        else:
            out[k] = v.float() if v.is_floating_point() else v
    return out


def _cconv(x, w, b, dilation=1, side="left"):
    k = w.shape[-1]
    pad = (k - 1) * dilation if k % 2 == 1 else int((k * dilation - dilation) / 2) * 2 + 1
    x = F.pad(x, (pad, 0) if side == "left" else (0, pad))
    return F.conv1d(x, w, b, dilation=dilation)


def _cconv_ctx(x, ctx, w, b):
    """right-causal conv with explicit look-ahead context (finalize=False)."""
    return F.conv1d(torch.cat([x, ctx], dim=2), w, b)


def snake(x, alpha):
    a = alpha.view(1, -1, 1)
    return x + (1.0 / (a + 1e-9)) * torch.sin(x * a) ** 2


def f0_predict(w, mel, finalize=True):
    p = "f0_predictor."
    if finalize:
        x = _cconv(mel, w[p + "condnet.0.weight"], w[p + "condnet.0.bias"], side="right")
    else:
        x = _cconv_ctx(mel[:, :, :-3], mel[:, :, -3:], w[p + "condnet.0.weight"], w[p + "condnet.0.bias"])
    x = F.elu(x)
    for i in (2, 4, 6, 8):
        x = F.elu(_cconv(x, w[p + f"condnet.{i}.weight"], w[p + f"condnet.{i}.bias"]))
    return torch.abs(F.linear(x.transpose(1, 2), w[p + "classifier.weight"], w[p + "classifier.bias"]).squeeze(-1))


def source(w, f0, sine_table, dims):
    """Closed form of SineGen2+SourceModuleHnNSF in eval/causal mode.

    f0: (1, T) frame-rate.  sine_table: (L>=T*frame, H) uniform[0,1) rows (SineGen2.sine_waves).
    Phase accumulates at frame rate in float64 (torch CPU cumsum accumulates float in double)
    and is held constant over each frame's samples; rand_ini has no effect.
    """
    H, up = dims.harmonics, dims.frame_samples
    T = f0.shape[1]
    h = torch.arange(1, H + 1, dtype=torch.float32).view(1, 1, H)
    fn = f0.view(1, T, 1) * h
    rad = (fn / dims.sr) % 1
    c = torch.cumsum(rad.double(), dim=1).float()
    phase = (c * 2 * math.pi) * up
    sines = torch.sin(phase) * 0.1                               # (1,T,H)
    uv = (f0 > 10).float().view(1, T, 1)
    namp = uv * 0.003 + (1 - uv) * 0.1 / 3
    sines = sines.repeat_interleave(up, dim=1)
    uv_s = uv.repeat_interleave(up, dim=1)
    namp_s = namp.repeat_interleave(up, dim=1)
    sw = sines * uv_s + namp_s * sine_table[: T * up].view(1, T * up, H)
    s = torch.tanh(F.linear(sw, w["m_source.l_linear.weight"], w["m_source.l_linear.bias"]))
    return s.transpose(1, 2)                                     # (1,1,T*up)


def _resblock(w, pfx, x, dil):
    for i, d in enumerate(dil):
        xt = snake(x, w[f"{pfx}.activations1.{i}.alpha"])
        xt = _cconv(xt, w[f"{pfx}.convs1.{i}.weight"], w[f"{pfx}.convs1.{i}.bias"], dilation=d)
        xt = snake(xt, w[f"{pfx}.activations2.{i}.alpha"])
        xt = _cconv(xt, w[f"{pfx}.convs2.{i}.weight"], w[f"{pfx}.convs2.{i}.bias"])
        x = xt + x
    return x


def decode(w, mel, s, dims, finalize=True):
    n_fft, hop = dims.n_fft, dims.hop
    win = torch.hann_window(n_fft, periodic=True, dtype=torch.float32)
    spec = torch.stft(s.squeeze(1), n_fft, hop, n_fft, window=win, return_complex=True)
    sr, si = spec.real, spec.imag
    up_prod = 1
    for u in dims.ups:
        up_prod *= u
    if finalize:
        x = _cconv(mel, w["conv_pre.weight"], w["conv_pre.bias"], side="right")
    else:
        x = _cconv_ctx(mel[:, :, :-4], mel[:, :, -4:], w["conv_pre.weight"], w["conv_pre.bias"])
        sr, si = sr[:, :, : -up_prod * 4], si[:, :, : -up_prod * 4]
    s_stft = torch.cat([sr, si], dim=1)
    nk = len(dims.rb_k)
    for i, u in enumerate(dims.ups):
        x = F.leaky_relu(x, 0.1)
        x = x.repeat_interleave(u, dim=2)
        wk = w[f"ups.{i}.weight"]
        x = F.conv1d(F.pad(x, (wk.shape[-1] - 1, 0)), wk, w[f"ups.{i}.bias"])
        if i == len(dims.ups) - 1:
            x = F.pad(x, (1, 0), mode="reflect")
        wd = w[f"source_downs.{i}.weight"]
        if wd.shape[-1] == 1:
            sd_ = F.conv1d(s_stft, wd, w[f"source_downs.{i}.bias"])
        else:
            stride = wd.shape[-1] // 2
            sd_ = F.conv1d(F.pad(s_stft, (stride - 1, 0)), wd, w[f"source_downs.{i}.bias"], stride=stride)
        sd_ = _resblock(w, f"source_resblocks.{i}", sd_, dims.rb_d)
        x = x + sd_
        xs = None
        for j in range(nk):
            r = _resblock(w, f"resblocks.{i * nk + j}", x, dims.rb_d)
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.leaky_relu(x)
    x = _cconv(x, w["conv_post.weight"], w["conv_post.bias"])
    nb = n_fft // 2 + 1
    mag = torch.clip(torch.exp(x[:, :nb]), max=1e2)
    ph = torch.sin(x[:, nb:])
    y = torch.istft(torch.complex(mag * torch.cos(ph), mag * torch.sin(ph)), n_fft, hop, n_fft, window=win)
    if not finalize:
        y = y[:, : -up_prod * hop]
    return torch.clamp(y, -0.99, 0.99)


@torch.no_grad()
def inference(sd, mel, sine_table, dims, finalize=True, folded=False, f0=None):
    """mel (1, 80, T) fp32 -> (wav (1, frame*T'), source (1,1,frame*T_f0)).

    `f0` pins the F0 track: the harmonic phase is 2*pi*frame*cumsum(f0*h/sr), so a 1e-6 relative
    difference in f0 (any two fp32 conv implementations) moves the phase of harmonic 9 by ~1e-3 rad
    per second of audio.  Waveform parity is therefore stated with the F0 track pinned, and F0
    parity is stated separately ("HiFT parity")."""
    w = sd if folded else fold_weight_norm(sd)
    if f0 is None:
        f0 = f0_predict(w, mel, finalize)
    s = source(w, f0, sine_table, dims)
    wav = decode(w, mel if finalize else mel[:, :, :-3], s, dims, finalize)
    return wav, s


# --------------------------------------------------------------------------- a12': non-causal HiFTGenerator
# (cosyvoice/hifigan/generator.py:378-569: weight-normed ConvTranspose1d upsampling, symmetric-padded ResBlocks,
#  ConvRNNF0Predictor f0_predictor.py:9-55).  Its source module (SineGen2 with causal=False, :233-317) draws a random
#  initial phase and fresh Gaussian noise on every call: the noise is an explicit input here (RNG pinned the way the LLM
#  sampler's uniform stream is), the initial phase provably has no effect (see source_nc).
def f0_predict_nc(w, mel):
    p = "f0_predictor."
    x = mel
    for i in (0, 2, 4, 6, 8):
        x = F.elu(F.conv1d(x, w[p + f"condnet.{i}.weight"], w[p + f"condnet.{i}.bias"], padding=1))
    return torch.abs(F.linear(x.transpose(1, 2), w[p + "classifier.weight"], w[p + "classifier.bias"]).squeeze(-1))


def _resblock_nc(w, pfx, x, dil):
    for i, d in enumerate(dil):
        k = w[f"{pfx}.convs1.{i}.weight"].shape[-1]
        xt = snake(x, w[f"{pfx}.activations1.{i}.alpha"])
        xt = F.conv1d(xt, w[f"{pfx}.convs1.{i}.weight"], w[f"{pfx}.convs1.{i}.bias"], dilation=d, padding=int((k * d - d) / 2))
        xt = snake(xt, w[f"{pfx}.activations2.{i}.alpha"])
        xt = F.conv1d(xt, w[f"{pfx}.convs2.{i}.weight"], w[f"{pfx}.convs2.{i}.bias"], padding=int((k - 1) / 2))
        x = xt + x
    return x


@torch.no_grad()
def decode_transposed(sd, mel, s, dims, folded=False):
    """HiFTGenerator.decode (generator.py:506-540): mel (1,80,T), source s (1,1,frame*T) -> wav (1, frame*T)."""
    w = sd if folded else fold_weight_norm(sd)
    n_fft, hop = dims.n_fft, dims.hop
    win = torch.hann_window(n_fft, periodic=True, dtype=torch.float32)
    spec = torch.stft(s.squeeze(1), n_fft, hop, n_fft, window=win, return_complex=True)
    s_stft = torch.cat([spec.real, spec.imag], dim=1)
    x = F.conv1d(mel, w["conv_pre.weight"], w["conv_pre.bias"], padding=3)
    nk = len(dims.rb_k)
    for i, u in enumerate(dims.ups):
        x = F.leaky_relu(x, 0.1)
        k = w[f"ups.{i}.weight"].shape[-1]
        x = F.conv_transpose1d(x, w[f"ups.{i}.weight"], w[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        if i == len(dims.ups) - 1:
            x = F.pad(x, (1, 0), mode="reflect")
        wd = w[f"source_downs.{i}.weight"]
        if wd.shape[-1] == 1:
            si = F.conv1d(s_stft, wd, w[f"source_downs.{i}.bias"])
        else:
            st = wd.shape[-1] // 2
            si = F.conv1d(s_stft, wd, w[f"source_downs.{i}.bias"], stride=st, padding=st // 2)
        si = _resblock_nc(w, f"source_resblocks.{i}", si, dims.rb_d)
        x = x + si
        xs = None
        for j in range(nk):
            r = _resblock_nc(w, f"resblocks.{i * nk + j}", x, dims.rb_d)
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.leaky_relu(x)
    x = F.conv1d(x, w["conv_post.weight"], w["conv_post.bias"], padding=3)
    nb = n_fft // 2 + 1
    mag = torch.clip(torch.exp(x[:, :nb]), max=1e2)
    ph = torch.sin(x[:, nb:])
    y = torch.istft(torch.complex(mag * torch.cos(ph), mag * torch.sin(ph)), n_fft, hop, n_fft, window=win)
    return torch.clamp(y, -0.99, 0.99)


def source_nc(w, f0, noise, dims):
    """SineGen2(causal=False) + SourceModuleHnNSF (generator.py:233-317, 358-375) with the Gaussian draw explicit.

    f0 (1, T) frame rate; noise (T*frame, H) = the values of `torch.randn_like(sine_waves)` (:310).  Restated facts:
    * f0 is nearest-upsampled x frame (:553), so the linear 1/frame down-sampling of `rad` (:253-255) reads sample
      positions frame*i + frame/2 - 0.5, i.e. two samples of the same frame: rad_frames[i] = (f0_i*h/sr) % 1 exactly, and
      `rand_ini`, added to sample 0 only (:246-250), never reaches the output;
    * phase = cumsum(rad_frames)*2*pi (:257), then `* frame` and LINEAR x frame up-sampling (:258-259, align_corners=False).
    """
    H, up = dims.harmonics, dims.frame_samples
    T = f0.shape[1]
    h = torch.arange(1, H + 1, dtype=torch.float32).view(1, 1, H)
    rad = ((f0.view(1, T, 1) * h) / dims.sr) % 1
    phase = torch.cumsum(rad.double(), dim=1).float() * 2 * math.pi
    phase = F.interpolate(phase.transpose(1, 2) * up, scale_factor=up, mode="linear").transpose(1, 2)   # (1, T*up, H)
    sines = torch.sin(phase) * 0.1
    uv = (f0 > 10).float().view(1, T, 1).repeat_interleave(up, dim=1)
    namp = uv * 0.003 + (1 - uv) * 0.1 / 3
    sw = sines * uv + namp * noise.view(1, T * up, H)
    s = torch.tanh(F.linear(sw, w["m_source.l_linear.weight"], w["m_source.l_linear.bias"]))
    return s.transpose(1, 2)


def draw_source_noise(n_samples, H, generator=None):
    """The reference's RNG consumption in SineGen2.forward (causal=False): `torch.rand(1, H)` for rand_ini (:248), then
    `randn_like` of a (1, n, H) tensor whose memory layout is (1, H, n) transposed (:259,310) — element order matters."""
    torch.rand(1, H, generator=generator)
    return torch.empty(1, H, n_samples).transpose(1, 2).normal_(generator=generator)[0]


@torch.no_grad()
def inference_transposed(sd, mel, noise, dims, f0=None, cache_source=None, folded=False):
    """HiFTGenerator.inference (generator.py:557-569): mel (1,80,T) -> (wav (1, frame*T), source (1,1,frame*T))."""
    w = sd if folded else fold_weight_norm(sd)
    if f0 is None:
        f0 = f0_predict_nc(w, mel)
    s = source_nc(w, f0, noise, dims)
    if cache_source is not None and cache_source.shape[2] != 0:
        s = s.clone()
        s[:, :, : cache_source.shape[2]] = cache_source
    return decode_transposed(w, mel, s, dims, folded=True), s
