"""Synthetic (seeded, random-init) checkpoints and inputs for the HydraVox hot path.

The real ``llm.pt / flow.pt / hift.pt`` are downloaded from ModelScope at install time
(reference README.md:73) and are not available offline, so bench.py and the parity tests
use state_dicts with *exactly the reference's parameter names and shapes*
(checked against the real reference modules by oracle/make_golden.py) filled from a
seeded ``torch.Generator``.  Scales are chosen so activations stay O(1) through each stage
(no saturation of the vocoder's exp/clamp, no softmax collapse), which keeps parity
comparisons meaningful.
"""
from __future__ import annotations

from typing import Dict

import torch

from .dims import FlowDims, HiftDims, LlmDims


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _randn(g, *shape, std=1.0):
    return torch.randn(*shape, generator=g) * std


# --------------------------------------------------------------------------- HiFT
def hift_state_dict(d: HiftDims, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = _gen(seed)
    sd: Dict[str, torch.Tensor] = {}

    def wn_conv(name, cout, cin, k, gain=1.0):
        v = _randn(g, cout, cin, k)
        gmag = v.reshape(cout, -1).norm(dim=1).reshape(cout, 1, 1) * (gain / (cin * k) ** 0.5)
        sd[name + ".bias"] = _randn(g, cout, std=0.05)
        sd[name + ".parametrizations.weight.original0"] = gmag
        sd[name + ".parametrizations.weight.original1"] = v

    def plain_conv(name, cout, cin, k, gain=1.0):
        sd[name + ".weight"] = _randn(g, cout, cin, k, std=gain / (cin * k) ** 0.5)
        sd[name + ".bias"] = _randn(g, cout, std=0.05)

    def resblock(pfx, ch, k):
        for c in ("convs1", "convs2"):
            for i in range(len(d.rb_d)):
                wn_conv(f"{pfx}.{c}.{i}", ch, ch, k, gain=0.6)
        for a in ("activations1", "activations2"):
            for i in range(len(d.rb_d)):
                sd[f"{pfx}.{a}.{i}.alpha"] = 0.5 + torch.rand(ch, generator=g)

    sd["m_source.l_linear.weight"] = _randn(g, 1, d.harmonics, std=1.0)
    sd["m_source.l_linear.bias"] = _randn(g, 1, std=0.1)
    wn_conv("conv_pre", d.base, d.mel, 5, gain=0.5)
    n_up = len(d.ups)
    for i in range(n_up):
        wn_conv(f"ups.{i}", d.base >> (i + 1), d.base >> i, d.up_k[i])
    cum = [1]
    for u in list(d.ups)[::-1][:-1]:
        cum.append(cum[-1] * u)
    for i, u in enumerate(cum[::-1]):
        plain_conv(f"source_downs.{i}", d.base >> (i + 1), d.n_fft + 2, 1 if u == 1 else 2 * u, gain=2.0)
        resblock(f"source_resblocks.{i}", d.base >> (i + 1), d.src_k[i])
    for i in range(n_up):
        for j, k in enumerate(d.rb_k):
            resblock(f"resblocks.{i * len(d.rb_k) + j}", d.base >> (i + 1), k)
    wn_conv("conv_post", d.n_fft + 2, d.base >> n_up, 7, gain=0.25)
    p = "f0_predictor."
    wn_conv(p + "condnet.0", d.f0_ch, d.mel, 4, gain=0.5)
    for i in (2, 4, 6, 8):
        wn_conv(p + f"condnet.{i}", d.f0_ch, d.f0_ch, 3, gain=1.4)
    # classifier scaled so |f0| straddles the voiced threshold (10 Hz) and reaches a few 100 Hz
    sd[p + "classifier.weight"] = _randn(g, 1, d.f0_ch, std=400.0 / d.f0_ch ** 0.5)
    sd[p + "classifier.bias"] = torch.tensor([5.0])
    return sd


def hift_t_state_dict(d: HiftDims, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Checkpoint of the non-causal HiFTGenerator variant (generator.py:378-569, SURVEY 8 a12'): same parameter names as the
    causal one; `ups.{i}` are weight-normed ConvTranspose1d weights (Cin, Cout, k), `source_downs` have kernel 2u / stride u,
    the F0 predictor is ConvRNNF0Predictor (k3, pad 1)."""
    sd = hift_state_dict(d, seed)
    g = _gen(seed + 7)
    for i, (u, k) in enumerate(zip(d.ups, d.up_k)):
        cin, cout = d.base >> i, d.base >> (i + 1)
        v = _randn(g, cin, cout, k)
        sd[f"ups.{i}.parametrizations.weight.original1"] = v
        sd[f"ups.{i}.parametrizations.weight.original0"] = v.reshape(cin, -1).norm(dim=1).reshape(cin, 1, 1) * (1.0 / (cin * k / u) ** 0.5)
        sd[f"ups.{i}.bias"] = _randn(g, cout, std=0.05)
    v = _randn(g, d.base, d.mel, 7)
    sd["conv_pre.parametrizations.weight.original1"] = v
    sd["conv_pre.parametrizations.weight.original0"] = v.reshape(d.base, -1).norm(dim=1).reshape(d.base, 1, 1) * (0.5 / (d.mel * 7) ** 0.5)
    v = _randn(g, d.f0_ch, d.mel, 3)
    sd["f0_predictor.condnet.0.parametrizations.weight.original1"] = v
    sd["f0_predictor.condnet.0.parametrizations.weight.original0"] = v.reshape(d.f0_ch, -1).norm(dim=1).reshape(d.f0_ch, 1, 1) * (0.5 / (d.mel * 3) ** 0.5)
    return sd


def hifigan_state_dict(d: HiftDims, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Checkpoint of the classic HiFi-GAN `Generator` (matcha/hifigan/models.py:148-193) with torch.nn.utils.weight_norm keys
    (`weight_g`, `weight_v`): Conv1d norms over (in, k) per output channel, ConvTranspose1d norms per input channel (dim 0)."""
    g = _gen(seed + 21)
    sd: Dict[str, torch.Tensor] = {}

    def wn(name, shape, fan, gain):
        v = _randn(g, *shape)
        sd[name + ".weight_v"] = v
        sd[name + ".weight_g"] = v.reshape(shape[0], -1).norm(dim=1).reshape(shape[0], 1, 1) * (gain / fan ** 0.5)

    wn("conv_pre", (d.base, d.mel, 7), d.mel * 7, 1.0)
    sd["conv_pre.bias"] = _randn(g, d.base, std=0.05)
    for i, (u, k) in enumerate(zip(d.ups, d.up_k)):
        cin, cout = d.base >> i, d.base >> (i + 1)
        wn(f"ups.{i}", (cin, cout, k), cin * k / u, 1.0)
        sd[f"ups.{i}.bias"] = _randn(g, cout, std=0.05)
        for j, rk in enumerate(d.rb_k):
            n = i * len(d.rb_k) + j
            for c in ("convs1", "convs2"):
                for t in range(len(d.rb_d)):
                    wn(f"resblocks.{n}.{c}.{t}", (cout, cout, rk), cout * rk, 0.8)
                    sd[f"resblocks.{n}.{c}.{t}.bias"] = _randn(g, cout, std=0.05)
    ch = d.base >> len(d.ups)
    wn("conv_post", (1, ch, 7), ch * 7, 0.3)
    sd["conv_post.bias"] = _randn(g, 1, std=0.05)
    return sd


def hift_sine_table(d: HiftDims, n_frames: int, seed: int = 11) -> torch.Tensor:
    """Stand-in for SineGen2.sine_waves (generator.py:226): uniform[0,1) rows, (n_frames*frame, H)."""
    return torch.rand(n_frames * d.frame_samples, d.harmonics, generator=_gen(seed))


# --------------------------------------------------------------------------- U-Net estimator (a7')
def unet_state_dict(d, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Checkpoint of CausalConditionalDecoder (cosyvoice/flow/decoder.py:294-400) at dims.UnetDims: same keys and shapes as
    the reference module's state_dict (pinned by load_state_dict(strict=True) in oracle/make_golden.py)."""
    g = _gen(seed + 31)
    sd: Dict[str, torch.Tensor] = {}
    C, inner, tdim = d.ch, d.heads * d.head_dim, 4 * d.ch

    def lin(name, out, inp, gain=1.0, bias=True):
        sd[name + ".weight"] = _randn(g, out, inp, std=gain / inp ** 0.5)
        if bias:
            sd[name + ".bias"] = _randn(g, out, std=0.05)

    def conv(name, out, inp, k, gain=1.0):
        sd[name + ".weight"] = _randn(g, out, inp, k, std=gain / (inp * k) ** 0.5)
        sd[name + ".bias"] = _randn(g, out, std=0.05)

    def ln(name, n):
        sd[name + ".weight"] = 1.0 + 0.1 * _randn(g, n)
        sd[name + ".bias"] = _randn(g, n, std=0.05)

    def resnet(p, cin):
        lin(p + ".mlp.1", C, tdim)
        for b, ci in (("block1", cin), ("block2", C)):
            conv(f"{p}.{b}.block.0", C, ci, 3, gain=1.4)
            ln(f"{p}.{b}.block.2", C)
        conv(p + ".res_conv", C, cin, 1)

    def tfm(p):
        ln(p + ".norm1", C)
        for n in "qkv":
            lin(f"{p}.attn1.to_{n}", inner, C, gain=1.5 if n != "v" else 1.0, bias=False)
        lin(p + ".attn1.to_out.0", C, inner, gain=0.5)
        ln(p + ".norm3", C)
        lin(p + ".ff.net.0.proj", d.ff_mult * C, C)
        lin(p + ".ff.net.2", C, d.ff_mult * C, gain=0.5)

    lin("time_mlp.linear_1", tdim, d.in_ch)
    lin("time_mlp.linear_2", tdim, tdim)
    stages = [("down_blocks.0", d.in_ch)] + [(f"mid_blocks.{i}", C) for i in range(d.n_mid)] + [("up_blocks.0", 2 * C)]
    for p, cin in stages:
        resnet(p + ".0", cin)
        for j in range(d.n_blocks):
            tfm(f"{p}.1.{j}")
        if not p.startswith("mid"):
            conv(p + ".2", C, C, 3)
    conv("final_block.block.0", C, C, 3, gain=1.4)
    ln("final_block.block.2", C)
    conv("final_proj", d.mel, C, 1)
    return sd


def unet_nc_state_dict(d, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Checkpoint of the non-causal multi-level ConditionalDecoder (cosyvoice/flow/decoder.py:88-205) at dims.UnetNcDims: same keys
    and shapes as the reference module's state_dict (pinned by load_state_dict(strict=True) in oracle/make_golden.py)."""
    g = _gen(seed + 37)
    sd: Dict[str, torch.Tensor] = {}
    C, inner, tdim, L = d.ch, d.heads * d.head_dim, 4 * d.ch, d.levels

    def lin(name, out, inp, gain=1.0, bias=True):
        sd[name + ".weight"] = _randn(g, out, inp, std=gain / inp ** 0.5)
        if bias:
            sd[name + ".bias"] = _randn(g, out, std=0.05)

    def conv(name, out, inp, k, gain=1.0, transpose=False):
        shape = (inp, out, k) if transpose else (out, inp, k)
        sd[name + ".weight"] = _randn(g, *shape, std=gain / (inp * k) ** 0.5)
        sd[name + ".bias"] = _randn(g, out, std=0.05)

    def norm(name, n):
        sd[name + ".weight"] = 1.0 + 0.1 * _randn(g, n)
        sd[name + ".bias"] = _randn(g, n, std=0.05)

    def resnet(p, cin):
        lin(p + ".mlp.1", C, tdim)
        for b, ci in (("block1", cin), ("block2", C)):
            conv(f"{p}.{b}.block.0", C, ci, 3, gain=1.4)
            norm(f"{p}.{b}.block.1", C)                       # GroupNorm(8, C)
        conv(p + ".res_conv", C, cin, 1)

    def tfm(p):
        norm(p + ".norm1", C)
        for n in "qkv":
            lin(f"{p}.attn1.to_{n}", inner, C, gain=1.5 if n != "v" else 1.0, bias=False)
        lin(p + ".attn1.to_out.0", C, inner, gain=0.5)
        norm(p + ".norm3", C)
        lin(p + ".ff.net.0.proj", d.ff_mult * C, C)
        lin(p + ".ff.net.2", C, d.ff_mult * C, gain=0.5)

    lin("time_mlp.linear_1", tdim, d.in_ch)
    lin("time_mlp.linear_2", tdim, tdim)
    for i in range(L):
        p = f"down_blocks.{i}"
        resnet(p + ".0", d.in_ch if i == 0 else C)
        for j in range(d.n_blocks):
            tfm(f"{p}.1.{j}")
        conv(p + (".2" if i == L - 1 else ".2.conv"), C, C, 3)            # Conv1d(k3, pad 1) on the last level, else Downsample1D
    for i in range(d.n_mid):
        resnet(f"mid_blocks.{i}.0", C)
        for j in range(d.n_blocks):
            tfm(f"mid_blocks.{i}.1.{j}")
    for i in range(L):
        p = f"up_blocks.{i}"
        resnet(p + ".0", 2 * C)
        for j in range(d.n_blocks):
            tfm(f"{p}.1.{j}")
        if i == L - 1:
            conv(p + ".2", C, C, 3)
        else:
            conv(p + ".2.conv", C, C, 4, transpose=True)                 # Upsample1D: ConvTranspose1d(C, C, 4, 2, 1)
    conv("final_block.block.0", C, C, 3, gain=1.4)
    norm("final_block.block.1", C)
    conv("final_proj", d.mel, C, 1)
    return sd


# --------------------------------------------------------------------------- flow
def flow_state_dict(d: FlowDims, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = _gen(seed)
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, out, inp, gain=1.0, bias=True):
        sd[name + ".weight"] = _randn(g, out, inp, std=gain / inp ** 0.5)
        if bias:
            sd[name + ".bias"] = _randn(g, out, std=0.05)

    sd["input_embedding.weight"] = _randn(g, d.vocab, d.mel)
    lin("spk_embed_affine_layer", d.mel, d.spk_in)
    sd["pre_lookahead_layer.conv1.weight"] = _randn(g, d.pla_ch, d.mel, 4, std=1.0 / (d.mel * 4) ** 0.5)
    sd["pre_lookahead_layer.conv1.bias"] = _randn(g, d.pla_ch, std=0.05)
    sd["pre_lookahead_layer.conv2.weight"] = _randn(g, d.mel, d.pla_ch, 3, std=1.0 / (d.pla_ch * 3) ** 0.5)
    sd["pre_lookahead_layer.conv2.bias"] = _randn(g, d.mel, std=0.05)
    p = "decoder.estimator."
    lin(p + "time_embed.time_mlp.0", d.dim, 256)
    lin(p + "time_embed.time_mlp.2", d.dim, d.dim)
    lin(p + "input_embed.proj", d.dim, 4 * d.mel)
    cg = d.dim // d.pos_groups
    for c in ("conv1", "conv2"):
        sd[p + f"input_embed.conv_pos_embed.{c}.0.weight"] = _randn(g, d.dim, cg, d.pos_k, std=1.0 / (cg * d.pos_k) ** 0.5)
        sd[p + f"input_embed.conv_pos_embed.{c}.0.bias"] = _randn(g, d.dim, std=0.05)
    sd[p + "rotary_embed.inv_freq"] = 1.0 / (10000.0 ** (torch.arange(0, d.dim_head, 2).float() / d.dim_head))
    inner = d.heads * d.dim_head
    for i in range(d.depth):
        bp = p + f"transformer_blocks.{i}."
        lin(bp + "attn_norm.linear", 6 * d.dim, d.dim, gain=0.5)
        lin(bp + "attn.to_q", inner, d.dim)
        lin(bp + "attn.to_k", inner, d.dim)
        lin(bp + "attn.to_v", inner, d.dim)
        lin(bp + "attn.to_out.0", d.dim, inner)
        lin(bp + "ff.ff.0.0", d.dim * d.ff_mult, d.dim)
        lin(bp + "ff.ff.2", d.dim, d.dim * d.ff_mult)
    lin(p + "norm_out.linear", 2 * d.dim, d.dim, gain=0.5)
    lin(p + "proj_out", d.mel, d.dim)
    return sd


def flow_noise(d: FlowDims) -> torch.Tensor:
    """CausalConditionalCFM.rand_noise (flow_matching.py:200-201): randn(1,80,15000) right after
    set_all_random_seed(0).  torch.manual_seed(0) + randn reproduces it bit-for-bit on CPU
    (first values -1.1258, -1.1524, -0.2506)."""
    g = _gen(0)
    return torch.randn(1, d.mel, 50 * 300, generator=g)[:, :, : d.noise_frames].contiguous()


# --------------------------------------------------------------------------- LLM
def llm_state_dict(d: LlmDims, seed: int = 0, dtype=torch.float32, eos_scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """eos_scale scales the llm_decoder rows of the stop range (ids >= speech_token_size).  Random weights put a stop
    id on top of the nucleus every few dozen steps, which under the fixed-length protocol (min_len == max_len)
    exhausts the reference's 100 EOS retries; bench.py and the long-run tests use eos_scale=0 (stop logits == 0)."""
    g = _gen(seed)
    sd: Dict[str, torch.Tensor] = {}

    def lin(name, out, inp, gain=1.0, bias=False):
        sd[name + ".weight"] = _randn(g, out, inp, std=gain / inp ** 0.5).to(dtype)
        if bias:
            sd[name + ".bias"] = _randn(g, out, std=0.1).to(dtype)

    def norm(name):
        sd[name + ".weight"] = (1.0 + 0.1 * _randn(g, d.hidden)).to(dtype)

    sd["llm.model.model.embed_tokens.weight"] = _randn(g, d.text_vocab, d.hidden).to(dtype)
    for l in range(d.layers):
        p = f"llm.model.model.layers.{l}."
        lin(p + "self_attn.q_proj", d.q_heads * d.head_dim, d.hidden, bias=True)
        lin(p + "self_attn.k_proj", d.kv_heads * d.head_dim, d.hidden, bias=True)
        lin(p + "self_attn.v_proj", d.kv_heads * d.head_dim, d.hidden, bias=True)
        lin(p + "self_attn.o_proj", d.hidden, d.q_heads * d.head_dim, gain=0.5)
        lin(p + "mlp.gate_proj", d.inter, d.hidden)
        lin(p + "mlp.up_proj", d.inter, d.hidden)
        lin(p + "mlp.down_proj", d.hidden, d.inter, gain=0.5)
        norm(p + "input_layernorm")
        norm(p + "post_attention_layernorm")
    norm("llm.model.model.norm")
    sd["llm.model.lm_head.weight"] = sd["llm.model.model.embed_tokens.weight"]
    lin("llm_decoder", d.speech_vocab, d.hidden, gain=3.0)
    if eos_scale != 1.0:
        sd["llm_decoder.weight"][d.speech_token_size:] *= eos_scale
    mh = d.mtp_attn_heads * (d.hidden // d.mtp_attn_heads)
    for j in range(d.mtp_heads):
        p = f"mtp_block.{j}."
        lin(p + "self_attn.q_proj", mh, d.hidden, bias=True)
        lin(p + "self_attn.k_proj", mh, d.hidden, bias=True)
        lin(p + "self_attn.v_proj", mh, d.hidden, bias=True)
        lin(p + "self_attn.o_proj", d.hidden, mh, gain=0.5)
        lin(p + "mlp.gate_proj", d.mtp_inter, d.hidden)
        lin(p + "mlp.up_proj", d.mtp_inter, d.hidden)
        lin(p + "mlp.down_proj", d.hidden, d.mtp_inter, gain=0.5)
        norm(p + "input_layernorm")
        norm(p + "post_attention_layernorm")
    sd["speech_embedding.weight"] = _randn(g, d.speech_vocab, d.hidden).to(dtype)
    return sd


# --------------------------------------------------------------------------- inputs
def utterance(ld: LlmDims, fd: FlowDims, n_text: int, seed: int = 1986, zero_shot: bool = True,
              prompt_tokens: int = 125, prompt_text: int = 16):
    """One synthetic request: text ids, prompt, speaker embedding (seed 1986 as in SURVEY §8d)."""
    g = _gen(seed)
    u = dict(
        text=torch.randint(0, ld.text_vocab, (n_text,), generator=g, dtype=torch.int32),
        embedding=torch.rand(fd.spk_in, generator=g),
    )
    if zero_shot:
        u["prompt_text"] = torch.randint(0, ld.text_vocab, (prompt_text,), generator=g, dtype=torch.int32)
        u["prompt_speech"] = torch.randint(0, min(ld.speech_token_size, fd.vocab), (prompt_tokens,), generator=g,
                                           dtype=torch.int32)
        u["prompt_feat"] = (torch.rand(2 * prompt_tokens, fd.mel, generator=g) * -6.0)
    else:
        u["prompt_text"] = torch.zeros(0, dtype=torch.int32)
        u["prompt_speech"] = torch.zeros(0, dtype=torch.int32)
        u["prompt_feat"] = None
    return u
