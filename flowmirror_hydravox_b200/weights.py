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
    if base + ".weight_g" in sd:                         # torch.nn.utils.weight_norm (the classic HiFi-GAN checkpoints)
        g, v = sd[base + ".weight_g"].float(), sd[base + ".weight_v"].float()
    else:
        g = sd[base + ".parametrizations.weight.original0"].float()
        v = sd[base + ".parametrizations.weight.original1"].float()
    n = v.reshape(v.shape[0], -1).norm(2, 1).reshape(-1, 1, 1)
    return v * (g / n)


def _conv(sd, base, out, name, tc: bool = False):
    w = _fold_wn(sd, base)                               # (Cout, Cin, K)
    out[name + ".w"] = w.permute(1, 2, 0).contiguous()   # [Cin][K][Cout]
    out[name + ".b"] = sd[base + ".bias"].float().contiguous()
    if tc:
        # tensor-core operand of the implicit-GEMM convolution (csrc/hift.cu: tc_conv): [Cout][2*K*Cp] fp16 = [hi | lo],
        # column = tap*Cp + ci, input channels zero-padded to whole 64-wide k-blocks
        cout, cin, k = w.shape
        cp = (cin + 63) // 64 * 64
        wp = torch.zeros(cout, k, cp)
        wp[:, :, :cin] = w.permute(0, 2, 1)
        out[name + ".w16"] = _split16(wp.reshape(cout, k * cp))


def _conv_transpose(sd, base, out, name):
    """Weight-normed ConvTranspose1d (Cin, Cout, K), norm over (Cout, K) per input channel (dim 0).  Stored tap-reversed as
    [Cin][K][Cout]: conv_transpose(x, w, stride u) == conv(zero_stuff(x, u), flip_k(w)) — what conv1d_tile_kernel runs."""
    w = _fold_wn(sd, base)                               # (Cin, Cout, K)
    out[name + ".w"] = w.flip(2).permute(0, 2, 1).contiguous()
    out[name + ".b"] = sd[base + ".bias"].float().contiguous()


def pack_hift_t(sd: Dict[str, torch.Tensor], d: D.HiftDims) -> Dict[str, torch.Tensor]:
    """HiFTGenerator (generator.py:378-569) state_dict -> the same packed tensor names as the causal vocoder."""
    return pack_hift(sd, d, transposed=True)


def pack_hifigan(sd: Dict[str, torch.Tensor], d: D.HiftDims) -> Dict[str, torch.Tensor]:
    """Classic HiFi-GAN `Generator` (matcha/hifigan/models.py:148-193) state_dict (weight-normed or with the norm removed,
    :195-203) -> packed tensors for hvx_hifigan_vocode."""
    o: Dict[str, torch.Tensor] = {}
    _conv(sd, "conv_pre", o, "conv_pre")
    _conv(sd, "conv_post", o, "conv_post")
    for i in range(len(d.ups)):
        _conv_transpose(sd, f"ups.{i}", o, f"ups.{i}")
        for j in range(len(d.rb_k)):
            n = i * len(d.rb_k) + j
            for t in range(len(d.rb_d)):
                _conv(sd, f"resblocks.{n}.convs1.{t}", o, f"rb.{n}.c1.{t}")
                _conv(sd, f"resblocks.{n}.convs2.{t}", o, f"rb.{n}.c2.{t}")
    return o


def pack_hift(sd: Dict[str, torch.Tensor], d: D.HiftDims, transposed: bool = False) -> Dict[str, torch.Tensor]:
    o: Dict[str, torch.Tensor] = {}
    for i, idx in enumerate((0, 2, 4, 6, 8)):
        _conv(sd, f"f0_predictor.condnet.{idx}", o, f"f0.c{i}")
    o["f0.cls.w"] = sd["f0_predictor.classifier.weight"].float().reshape(-1).contiguous()
    o["f0.cls.b"] = sd["f0_predictor.classifier.bias"].float().reshape(-1).contiguous()
    o["src.lin.w"] = sd["m_source.l_linear.weight"].float().reshape(-1).contiguous()
    o["src.lin.b"] = sd["m_source.l_linear.bias"].float().reshape(-1).contiguous()
    tc = not transposed                 # the causal vocoder's decode stack runs on tensor cores (split-fp16 implicit GEMMs)
    _conv(sd, "conv_pre", o, "conv_pre", tc)
    _conv(sd, "conv_post", o, "conv_post", tc)

    def rb(src, dst):
        for j in range(len(d.rb_d)):
            _conv(sd, f"{src}.convs1.{j}", o, f"{dst}.c1.{j}", tc)
            _conv(sd, f"{src}.convs2.{j}", o, f"{dst}.c2.{j}", tc)
            for a in ("1", "2"):
                al = sd[f"{src}.activations{a}.{j}.alpha"].float().contiguous()
                o[f"{dst}.a{a}.{j}"] = al
                o[f"{dst}.ia{a}.{j}"] = (1.0 / (al + 1e-9)).contiguous()       # Snake: x + 1/(alpha + 1e-9) * sin(alpha x)^2 (activation.py:79-84)

    for i in range(len(d.ups)):
        if transposed:
            _conv_transpose(sd, f"ups.{i}", o, f"ups.{i}")
        else:
            _conv(sd, f"ups.{i}", o, f"ups.{i}", tc)
        _conv(sd, f"source_downs.{i}", o, f"sdown.{i}")
        rb(f"source_resblocks.{i}", f"srb.{i}")
        for j in range(len(d.rb_k)):
            n = i * len(d.rb_k) + j
            rb(f"resblocks.{n}", f"rb.{n}")
    return o


def _split16(w: torch.Tensor) -> torch.Tensor:
    """[N][K] fp32 -> [N][2K] fp16 = [hi | lo] with lo = fp16(w - hi): ~22 mantissa bits for the parity-mode GEMMs."""
    w = w.to(torch.float32)
    hi = w.to(torch.float16)
    lo = (w - hi.float()).to(torch.float16)
    return torch.cat([hi, lo], dim=1).contiguous()


def pack_flow(sd: Dict[str, torch.Tensor], d: D.FlowDims, precise: bool = False) -> Dict[str, torch.Tensor]:
    """CausalMaskedDiffWithDiT state_dict (flow.py:296-365, DiT/dit.py:104-143) -> engine tensors.
    GEMM operands are fp16 (the reference serves this stage with .half(), infer_speech_model.py:105-117),
    biases / tiny layers fp32.  Conv weights become implicit-GEMM B operands: column = tap*C + ci."""
    h, f = torch.float16, torch.float32
    o: Dict[str, torch.Tensor] = {}
    o["spk.w"] = sd["spk_embed_affine_layer.weight"].to(f).contiguous()
    o["spk.b"] = sd["spk_embed_affine_layer.bias"].to(f).contiguous()
    o["emb"] = sd["input_embedding.weight"].to(f).contiguous()
    cp = (d.mel + 63) // 64 * 64
    w1 = sd["pre_lookahead_layer.conv1.weight"].to(f)                      # (pla, mel, 4)
    w1p = torch.zeros(w1.shape[0], w1.shape[2], cp)
    w1p[:, :, : d.mel] = w1.permute(0, 2, 1)
    o["pla1.w"] = w1p.reshape(w1.shape[0], -1).to(h).contiguous()
    o["pla1.b"] = sd["pre_lookahead_layer.conv1.bias"].to(f).contiguous()
    w2 = sd["pre_lookahead_layer.conv2.weight"].to(f)                      # (mel, pla, 3)
    o["pla2.w"] = w2.permute(0, 2, 1).reshape(w2.shape[0], -1).to(h).contiguous()
    o["pla2.b"] = sd["pre_lookahead_layer.conv2.bias"].to(f).contiguous()
    p = "decoder.estimator."
    o["tm0.w"] = sd[p + "time_embed.time_mlp.0.weight"].to(f).contiguous()
    o["tm0.b"] = sd[p + "time_embed.time_mlp.0.bias"].to(f).contiguous()
    o["tm2.w"] = sd[p + "time_embed.time_mlp.2.weight"].to(f).contiguous()
    o["tm2.b"] = sd[p + "time_embed.time_mlp.2.bias"].to(f).contiguous()
    o["in.w"] = sd[p + "input_embed.proj.weight"].to(h).contiguous()
    o["in.b"] = sd[p + "input_embed.proj.bias"].to(f).contiguous()
    for i, c in enumerate(("conv1", "conv2"), 1):
        w = sd[p + f"input_embed.conv_pos_embed.{c}.0.weight"].to(f)        # (dim, dim/groups, k)
        o[f"pos{i}.w"] = w.permute(0, 2, 1).reshape(w.shape[0], -1).to(h).contiguous()
        o[f"pos{i}.b"] = sd[p + f"input_embed.conv_pos_embed.{c}.0.bias"].to(f).contiguous()
    o["rope.inv_freq"] = sd[p + "rotary_embed.inv_freq"].to(f).contiguous()
    mods_w, mods_b = [], []
    for i in range(d.depth):
        bp = p + f"transformer_blocks.{i}."
        o[f"blk{i}.qkv.w"] = torch.cat([sd[bp + f"attn.to_{n}.weight"] for n in "qkv"], 0).to(h).contiguous()
        o[f"blk{i}.qkv.b"] = torch.cat([sd[bp + f"attn.to_{n}.bias"] for n in "qkv"], 0).to(f).contiguous()
        o[f"blk{i}.out.w"] = sd[bp + "attn.to_out.0.weight"].to(h).contiguous()
        o[f"blk{i}.out.b"] = sd[bp + "attn.to_out.0.bias"].to(f).contiguous()
        o[f"blk{i}.ff1.w"] = sd[bp + "ff.ff.0.0.weight"].to(h).contiguous()
        o[f"blk{i}.ff1.b"] = sd[bp + "ff.ff.0.0.bias"].to(f).contiguous()
        o[f"blk{i}.ff2.w"] = sd[bp + "ff.ff.2.weight"].to(h).contiguous()
        o[f"blk{i}.ff2.b"] = sd[bp + "ff.ff.2.bias"].to(f).contiguous()
        mods_w.append(sd[bp + "attn_norm.linear.weight"])
        mods_b.append(sd[bp + "attn_norm.linear.bias"])
    mods_w.append(sd[p + "norm_out.linear.weight"])
    mods_b.append(sd[p + "norm_out.linear.bias"])
    o["mod.w"] = torch.cat(mods_w, 0).to(h).contiguous()
    o["mod.b"] = torch.cat(mods_b, 0).to(f).contiguous()
    o["proj.w"] = sd[p + "proj_out.weight"].to(h).contiguous()
    o["proj.b"] = sd[p + "proj_out.bias"].to(f).contiguous()
    if precise:                                   # parity mode: every GEMM operand gets split-fp16 weights from the fp32 originals
        o["pla1.w"] = _split16(w1p.reshape(w1.shape[0], -1))
        o["pla2.w"] = _split16(w2.permute(0, 2, 1).reshape(w2.shape[0], -1))
        for i, c in enumerate(("conv1", "conv2"), 1):
            w = sd[p + f"input_embed.conv_pos_embed.{c}.0.weight"].to(f)
            o[f"pos{i}.w"] = _split16(w.permute(0, 2, 1).reshape(w.shape[0], -1))
        o["in.w"] = _split16(sd[p + "input_embed.proj.weight"])
        o["proj.w"] = _split16(sd[p + "proj_out.weight"])
        o["mod.w"] = _split16(torch.cat(mods_w, 0))
        for i in range(d.depth):
            bp = p + f"transformer_blocks.{i}."
            o[f"blk{i}.qkv.w"] = _split16(torch.cat([sd[bp + f"attn.to_{n}.weight"] for n in "qkv"], 0))
            o[f"blk{i}.out.w"] = _split16(sd[bp + "attn.to_out.0.weight"])
            o[f"blk{i}.ff1.w"] = _split16(sd[bp + "ff.ff.0.0.weight"])
            o[f"blk{i}.ff2.w"] = _split16(sd[bp + "ff.ff.2.weight"])
    return o


def pack_unet(sd: Dict[str, torch.Tensor], d: D.UnetDims, precise: bool = False) -> Dict[str, torch.Tensor]:
    """CausalConditionalDecoder state_dict (cosyvoice/flow/decoder.py:294-400, channels == (C,)) -> engine tensors of the
    HVX_STAGE_UNET stage.  GEMM operands fp16 ([hi | lo] in parity mode); causal k3 convs become implicit-GEMM B operands with
    column = tap*Cin + ci; q/k/v fused; every resnet's time-embedding Linear stacked into one fp32 matrix."""
    import math
    f = torch.float32
    cvt = _split16 if precise else (lambda w: w.to(torch.float16).contiguous())
    o: Dict[str, torch.Tensor] = {}
    half = d.in_ch // 2
    o["time.freqs"] = torch.exp(torch.arange(half).float() * -(math.log(10000) / (half - 1))).contiguous()   # SinusoidalPosEmb
    for i, n in ((1, "linear_1"), (2, "linear_2")):
        o[f"time.l{i}.w"] = sd[f"time_mlp.{n}.weight"].to(f).contiguous()
        o[f"time.l{i}.b"] = sd[f"time_mlp.{n}.bias"].to(f).contiguous()

    def conv(key):
        w = sd[key + ".weight"].to(f)                                    # (Cout, Cin, k)
        return cvt(w.permute(0, 2, 1).reshape(w.shape[0], -1)), sd[key + ".bias"].to(f).contiguous()

    stages = ["down_blocks.0"] + [f"mid_blocks.{i}" for i in range(d.n_mid)] + ["up_blocks.0"]
    rw, rb = [], []
    for i, p in enumerate(stages):
        r = p + ".0"
        rw.append(sd[r + ".mlp.1.weight"].to(f)); rb.append(sd[r + ".mlp.1.bias"].to(f))
        for n, blk in ((1, "block1"), (2, "block2")):
            o[f"res{i}.c{n}.w"], o[f"res{i}.c{n}.b"] = conv(f"{r}.{blk}.block.0")
            o[f"res{i}.ln{n}.g"] = sd[f"{r}.{blk}.block.2.weight"].to(f).contiguous()
            o[f"res{i}.ln{n}.b"] = sd[f"{r}.{blk}.block.2.bias"].to(f).contiguous()
        o[f"res{i}.rc.w"], o[f"res{i}.rc.b"] = conv(r + ".res_conv")
        for j in range(d.n_blocks):
            t, q = f"{p}.1.{j}", f"tfm{i * d.n_blocks + j}"
            o[q + ".n1.g"], o[q + ".n1.b"] = sd[t + ".norm1.weight"].to(f).contiguous(), sd[t + ".norm1.bias"].to(f).contiguous()
            o[q + ".n3.g"], o[q + ".n3.b"] = sd[t + ".norm3.weight"].to(f).contiguous(), sd[t + ".norm3.bias"].to(f).contiguous()
            o[q + ".qkv.w"] = cvt(torch.cat([sd[f"{t}.attn1.to_{n}.weight"].to(f) for n in "qkv"], 0))
            o[q + ".out.w"], o[q + ".out.b"] = cvt(sd[t + ".attn1.to_out.0.weight"].to(f)), sd[t + ".attn1.to_out.0.bias"].to(f).contiguous()
            o[q + ".ff1.w"], o[q + ".ff1.b"] = cvt(sd[t + ".ff.net.0.proj.weight"].to(f)), sd[t + ".ff.net.0.proj.bias"].to(f).contiguous()
            o[q + ".ff2.w"], o[q + ".ff2.b"] = cvt(sd[t + ".ff.net.2.weight"].to(f)), sd[t + ".ff.net.2.bias"].to(f).contiguous()
    o["rmlp.w"], o["rmlp.b"] = torch.cat(rw, 0).contiguous(), torch.cat(rb, 0).contiguous()
    o["down.w"], o["down.b"] = conv("down_blocks.0.2")
    o["up.w"], o["up.b"] = conv("up_blocks.0.2")
    o["fin.w"], o["fin.b"] = conv("final_block.block.0")
    o["fin.g"], o["fin.bt"] = sd["final_block.block.2.weight"].to(f).contiguous(), sd["final_block.block.2.bias"].to(f).contiguous()
    o["proj.w"], o["proj.b"] = conv("final_proj")
    return o


def pack_unet_nc(sd: Dict[str, torch.Tensor], d, precise: bool = False) -> Dict[str, torch.Tensor]:
    """Non-causal multi-level ConditionalDecoder state_dict (cosyvoice/flow/decoder.py:88-205; dims.UnetNcDims) -> engine tensors of
    the HVX_STAGE_UNET stage.  As pack_unet, with: GroupNorm affine parameters under the ln* names; Downsample1D (Conv1d k3 stride 2
    pad 1) repacked as a 2-tap convolution over frame pairs — operand row t = (x[2t], x[2t+1]), y[t] = [0 | w0] row[t-1] +
    [w1 | w2] row[t]; Upsample1D (ConvTranspose1d(C, C, 4, 2, 1), weight (Cin, Cout, 4)) repacked as a k3 pad-1 convolution with 2C
    outputs — output row m = (y[2m], y[2m+1]), y[2m] = w3 x[m-1] + w1 x[m], y[2m+1] = w2 x[m] + w0 x[m+1]."""
    import math
    f = torch.float32
    cvt = _split16 if precise else (lambda w: w.to(torch.float16).contiguous())
    o: Dict[str, torch.Tensor] = {}
    L, C = d.levels, d.ch
    half = d.in_ch // 2
    o["time.freqs"] = torch.exp(torch.arange(half).float() * -(math.log(10000) / (half - 1))).contiguous()
    for i, n in ((1, "linear_1"), (2, "linear_2")):
        o[f"time.l{i}.w"] = sd[f"time_mlp.{n}.weight"].to(f).contiguous()
        o[f"time.l{i}.b"] = sd[f"time_mlp.{n}.bias"].to(f).contiguous()

    def conv(key):
        w = sd[key + ".weight"].to(f)                                    # (Cout, Cin, k) -> column = tap*Cin + ci
        return cvt(w.permute(0, 2, 1).reshape(w.shape[0], -1)), sd[key + ".bias"].to(f).contiguous()

    stages = [f"down_blocks.{i}" for i in range(L)] + [f"mid_blocks.{i}" for i in range(d.n_mid)] + [f"up_blocks.{i}" for i in range(L)]
    rw, rb = [], []
    for i, p in enumerate(stages):
        r = p + ".0"
        rw.append(sd[r + ".mlp.1.weight"].to(f)); rb.append(sd[r + ".mlp.1.bias"].to(f))
        for n, blk in ((1, "block1"), (2, "block2")):
            o[f"res{i}.c{n}.w"], o[f"res{i}.c{n}.b"] = conv(f"{r}.{blk}.block.0")
            o[f"res{i}.ln{n}.g"] = sd[f"{r}.{blk}.block.1.weight"].to(f).contiguous()          # GroupNorm(8, C)
            o[f"res{i}.ln{n}.b"] = sd[f"{r}.{blk}.block.1.bias"].to(f).contiguous()
        o[f"res{i}.rc.w"], o[f"res{i}.rc.b"] = conv(r + ".res_conv")
        for j in range(d.n_blocks):
            t, q = f"{p}.1.{j}", f"tfm{i * d.n_blocks + j}"
            o[q + ".n1.g"], o[q + ".n1.b"] = sd[t + ".norm1.weight"].to(f).contiguous(), sd[t + ".norm1.bias"].to(f).contiguous()
            o[q + ".n3.g"], o[q + ".n3.b"] = sd[t + ".norm3.weight"].to(f).contiguous(), sd[t + ".norm3.bias"].to(f).contiguous()
            o[q + ".qkv.w"] = cvt(torch.cat([sd[f"{t}.attn1.to_{n}.weight"].to(f) for n in "qkv"], 0))
            o[q + ".out.w"], o[q + ".out.b"] = cvt(sd[t + ".attn1.to_out.0.weight"].to(f)), sd[t + ".attn1.to_out.0.bias"].to(f).contiguous()
            o[q + ".ff1.w"], o[q + ".ff1.b"] = cvt(sd[t + ".ff.net.0.proj.weight"].to(f)), sd[t + ".ff.net.0.proj.bias"].to(f).contiguous()
            o[q + ".ff2.w"], o[q + ".ff2.b"] = cvt(sd[t + ".ff.net.2.weight"].to(f)), sd[t + ".ff.net.2.bias"].to(f).contiguous()
    o["rmlp.w"], o["rmlp.b"] = torch.cat(rw, 0).contiguous(), torch.cat(rb, 0).contiguous()
    for l in range(L):
        if l == L - 1:
            o[f"down{l}.w"], o[f"down{l}.b"] = conv(f"down_blocks.{l}.2")
            o[f"up{l}.w"], o[f"up{l}.b"] = conv(f"up_blocks.{l}.2")
            continue
        w = sd[f"down_blocks.{l}.2.conv.weight"].to(f)                    # (C, C, 3), stride 2
        z = torch.zeros_like(w[:, :, 0])
        o[f"down{l}.w"] = cvt(torch.cat([z, w[:, :, 0], w[:, :, 1], w[:, :, 2]], dim=1))       # (C, [tap0: 0 | w0], [tap1: w1 | w2])
        o[f"down{l}.b"] = sd[f"down_blocks.{l}.2.conv.bias"].to(f).contiguous()
        wt = sd[f"up_blocks.{l}.2.conv.weight"].to(f)                     # (Cin, Cout, 4)
        k = [wt[:, :, i].t() for i in range(4)]                           # (Cout, Cin) per tap
        z = torch.zeros_like(k[0])
        even = torch.cat([k[3], k[1], z], dim=1)                          # y[2m]   = w3 x[m-1] + w1 x[m]
        odd = torch.cat([z, k[2], k[0]], dim=1)                           # y[2m+1] = w2 x[m]   + w0 x[m+1]
        o[f"up{l}.w"] = cvt(torch.cat([even, odd], dim=0))
        b = sd[f"up_blocks.{l}.2.conv.bias"].to(f)
        o[f"up{l}.b"] = torch.cat([b, b]).contiguous()
    o["fin.w"], o["fin.b"] = conv("final_block.block.0")
    o["fin.g"], o["fin.bt"] = sd["final_block.block.1.weight"].to(f).contiguous(), sd["final_block.block.1.bias"].to(f).contiguous()
    o["proj.w"], o["proj.b"] = conv("final_proj")
    return o


def _rope_pair_perm(n_heads: int, head_dim: int = 64) -> torch.Tensor:
    """Row order that puts the HF half-split RoPE pair (d, d + head_dim/2) on adjacent rows (2i, 2i+1)."""
    half = head_dim // 2
    one = torch.tensor([i // 2 + half * (i % 2) for i in range(head_dim)])
    return torch.cat([one + h * head_dim for h in range(n_heads)])


def pack_llm(sd: Dict[str, torch.Tensor], d: D.LlmDims) -> Dict[str, torch.Tensor]:
    """CosyVoice3LM state_dict (llm_multi_head_v3.py:622-689; HF Qwen2 names under llm.model.*) -> engine tensors.
    bf16 weights (the reference serves the LLM with .to(bfloat16), infer_speech_model.py:102), fp32 norms/biases.
    q/k rows are permuted per head for adjacent RoPE pairs, gate/up rows interleaved for fused SwiGLU, the MTP heads'
    dead q/k projections are dropped and their tensors stacked over heads."""
    b, f = torch.bfloat16, torch.float32
    o: Dict[str, torch.Tensor] = {}
    o["embed"] = sd["llm.model.model.embed_tokens.weight"].to(b).contiguous()
    o["speech_emb"] = sd["speech_embedding.weight"].to(b).contiguous()
    o["dec.w"] = sd["llm_decoder.weight"].to(b).contiguous()
    o["norm"] = sd["llm.model.model.norm.weight"].to(f).contiguous()
    pq, pk = _rope_pair_perm(d.q_heads, d.head_dim), _rope_pair_perm(d.kv_heads, d.head_dim)

    def gu(p):
        g, u = sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"]
        return torch.stack([g, u], dim=1).reshape(2 * g.shape[0], g.shape[1]).to(b).contiguous()

    for l in range(d.layers):
        p = f"llm.model.model.layers.{l}."
        o[f"L{l}.qkv.w"] = torch.cat([sd[p + "self_attn.q_proj.weight"][pq], sd[p + "self_attn.k_proj.weight"][pk],
                                      sd[p + "self_attn.v_proj.weight"]], 0).to(b).contiguous()
        o[f"L{l}.qkv.b"] = torch.cat([sd[p + "self_attn.q_proj.bias"][pq], sd[p + "self_attn.k_proj.bias"][pk],
                                      sd[p + "self_attn.v_proj.bias"]], 0).to(f).contiguous()
        o[f"L{l}.o.w"] = sd[p + "self_attn.o_proj.weight"].to(b).contiguous()
        o[f"L{l}.gu.w"] = gu(p)
        o[f"L{l}.down.w"] = sd[p + "mlp.down_proj.weight"].to(b).contiguous()
        o[f"L{l}.ln1"] = sd[p + "input_layernorm.weight"].to(f).contiguous()
        o[f"L{l}.ln2"] = sd[p + "post_attention_layernorm.weight"].to(f).contiguous()
    hs = [f"mtp_block.{j}." for j in range(d.mtp_heads)]
    o["mtp.v.w"] = torch.stack([sd[p + "self_attn.v_proj.weight"] for p in hs]).to(b).contiguous()
    o["mtp.v.b"] = torch.stack([sd[p + "self_attn.v_proj.bias"] for p in hs]).to(f).contiguous()
    o["mtp.o.w"] = torch.stack([sd[p + "self_attn.o_proj.weight"] for p in hs]).to(b).contiguous()
    o["mtp.gu.w"] = torch.stack([gu(p) for p in hs]).contiguous()
    o["mtp.down.w"] = torch.stack([sd[p + "mlp.down_proj.weight"] for p in hs]).to(b).contiguous()
    o["mtp.ln1"] = torch.stack([sd[p + "input_layernorm.weight"] for p in hs]).to(f).contiguous()
    o["mtp.ln2"] = torch.stack([sd[p + "post_attention_layernorm.weight"] for p in hs]).to(f).contiguous()
    return o
