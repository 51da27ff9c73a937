"""Checkpoint ingest: reference state_dict keys -> the engine's packed device tensors.

Mirrors what `load_state_dict` does in ModelManager.load_models / load_pt
(server/model_utils/infer_speech_model.py:69-94,169-184), plus the one-time re-layout each kernel
wants (weight-norm folding, [Cin][K][Cout] conv packing, fused QKV, bf16 casts, row padding).
"""
from __future__ import annotations

from typing import Dict

import torch

from . import dims as D


def _fold_wn(sd, base):
    """w = v * (g / ||v||) over (in, k) per out channel — torch._weight_norm op order."""
    if base + ".weight" in sd:
        return sd[base + ".weight"].float()
    g = sd[base + ".parametrizations.weight.original0"].float()
    v = sd[base + ".parametrizations.weight.original1"].float()
    n = v.reshape(v.shape[0], -1).norm(2, 1).reshape(-1, 1, 1)
    return v * (g / n)


def _conv(sd, base, out, name):
    w = _fold_wn(sd, base)                               # (Cout, Cin, K)
    out[name + ".w"] = w.permute(1, 2, 0).contiguous()   # [Cin][K][Cout]
    out[name + ".b"] = sd[base + ".bias"].float().contiguous()


def pack_hift(sd: Dict[str, torch.Tensor], d: D.HiftDims) -> Dict[str, torch.Tensor]:
    o: Dict[str, torch.Tensor] = {}
    for i, idx in enumerate((0, 2, 4, 6, 8)):
        _conv(sd, f"f0_predictor.condnet.{idx}", o, f"f0.c{i}")
    o["f0.cls.w"] = sd["f0_predictor.classifier.weight"].float().reshape(-1).contiguous()
    o["f0.cls.b"] = sd["f0_predictor.classifier.bias"].float().reshape(-1).contiguous()
    o["src.lin.w"] = sd["m_source.l_linear.weight"].float().reshape(-1).contiguous()
    o["src.lin.b"] = sd["m_source.l_linear.bias"].float().reshape(-1).contiguous()
    _conv(sd, "conv_pre", o, "conv_pre")
    _conv(sd, "conv_post", o, "conv_post")

    def rb(src, dst):
        for j in range(len(d.rb_d)):
            _conv(sd, f"{src}.convs1.{j}", o, f"{dst}.c1.{j}")
            _conv(sd, f"{src}.convs2.{j}", o, f"{dst}.c2.{j}")
            o[f"{dst}.a1.{j}"] = sd[f"{src}.activations1.{j}.alpha"].float().contiguous()
            o[f"{dst}.a2.{j}"] = sd[f"{src}.activations2.{j}.alpha"].float().contiguous()

    for i in range(len(d.ups)):
        _conv(sd, f"ups.{i}", o, f"ups.{i}")
        _conv(sd, f"source_downs.{i}", o, f"sdown.{i}")
        rb(f"source_resblocks.{i}", f"srb.{i}")
        for j in range(len(d.rb_k)):
            n = i * len(d.rb_k) + j
            rb(f"resblocks.{n}", f"rb.{n}")
    return o
