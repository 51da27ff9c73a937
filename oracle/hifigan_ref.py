"""TEST INFRASTRUCTURE ONLY — CPU restatement (fp32 torch functional ops) of the classic HiFi-GAN generator the reference
vendors under matcha/hifigan (SURVEY.md §8 a12'): `Generator.forward` models.py:176-193, `ResBlock1.forward` :86-93,
`get_padding` xutils.py:51-52, weight_norm folding (torch.nn.utils.weight_norm: w = g * v / ||v||, norm over all dims but 0).
Pinned against the reference module itself by oracle/make_golden.py (fixtures tests/golden/hifigan_*.pt)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1                                  # models.py:11


def fold(sd: Dict[str, torch.Tensor], base: str) -> torch.Tensor:
    if base + ".weight" in sd:
        return sd[base + ".weight"].float()
    g, v = sd[base + ".weight_g"].float(), sd[base + ".weight_v"].float()
    return v * (g / v.reshape(v.shape[0], -1).norm(2, 1).reshape(-1, 1, 1))


def _same(k, d=1):
    return int((k * d - d) / 2)                    # xutils.py:51-52


def resblock1(sd, pfx, x, k, dils):
    for t, d in enumerate(dils):                   # models.py:86-93
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, fold(sd, f"{pfx}.convs1.{t}"), sd[f"{pfx}.convs1.{t}.bias"].float(), dilation=d, padding=_same(k, d))
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, fold(sd, f"{pfx}.convs2.{t}"), sd[f"{pfx}.convs2.{t}.bias"].float(), padding=_same(k, 1))
        x = xt + x
    return x


def generator(sd: Dict[str, torch.Tensor], mel: torch.Tensor, dims) -> torch.Tensor:
    """mel (B, mel, T) -> (B, 1, prod(ups)*T)   (models.py:176-193)"""
    x = F.conv1d(mel.float(), fold(sd, "conv_pre"), sd["conv_pre.bias"].float(), padding=3)
    nk = len(dims.rb_k)
    for i, (u, k) in enumerate(zip(dims.ups, dims.up_k)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, fold(sd, f"ups.{i}"), sd[f"ups.{i}.bias"].float(), stride=u, padding=(k - u) // 2)
        xs = None
        for j, rk in enumerate(dims.rb_k):
            y = resblock1(sd, f"resblocks.{i * nk + j}", x, rk, dims.rb_d)
            xs = y if xs is None else xs + y
        x = xs / nk
    x = F.leaky_relu(x)                            # default slope 0.01 (:189)
    x = F.conv1d(x, fold(sd, "conv_post"), sd["conv_post.bias"].float(), padding=3)
    return torch.tanh(x)
