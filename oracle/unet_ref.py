"""TEST INFRASTRUCTURE ONLY — CPU restatement (fp32 torch functional ops on a state_dict) of the U-Net flow-matching
estimator `CausalConditionalDecoder.forward` (cosyvoice/flow/decoder.py:405-494; SURVEY.md §8 a7'), for channels == (C,)
(one down block, n_mid mid blocks, one up block — the only shape the CosyVoice2-generation configs use; no resampling).

Follows: CausalConv1d :36-62, CausalBlock1D :65-79, ResnetBlock1D matcha/models/components/decoder.py:46-61,
SinusoidalPosEmb :14-28, TimestepEmbedding :73-113, BasicTransformerBlock matcha/models/components/transformer.py:246-316,
FeedForward :83-134, masks cosyvoice/utils/mask.py:127-236 (add_optional_chunk_mask / subsequent_chunk_mask) and
mask_to_bias cosyvoice/utils/common.py.  Third-party arithmetic not in the tree: diffusers==0.29.0 `Attention`
(AttnProcessor2_0: q/k/v projections without bias, scaled_dot_product_attention with an additive mask, out projection with
bias) and `GELU` (Linear + exact-erf gelu) — restated from the published algorithm; parity for those two is anchored on the
reference's call sites (transformer.py:196-204,110-126) and otherwise unpinned (diffusers is not installed here; the fixture
generator oracle/refshim.py supplies the same restatement to the reference module).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


def time_embedding(sd, t, in_ch):
    """SinusoidalPosEmb(in_ch)(t) -> TimestepEmbedding (decoder.py:424-425)"""
    half = in_ch // 2
    emb = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half).float() * -emb)
    emb = 1000 * t.float().reshape(-1, 1) * emb[None]
    emb = torch.cat((emb.sin(), emb.cos()), dim=-1)
    h = F.silu(F.linear(emb, sd["time_mlp.linear_1.weight"], sd["time_mlp.linear_1.bias"]))
    return F.linear(h, sd["time_mlp.linear_2.weight"], sd["time_mlp.linear_2.bias"])


def _cconv(x, w, b):
    return F.conv1d(F.pad(x, (w.shape[2] - 1, 0)), w, b)


def _block(sd, p, x, mask):
    """CausalBlock1D: conv k3 -> LayerNorm over channels -> Mish, masked in and out"""
    h = _cconv(x * mask, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"])
    h = F.layer_norm(h.transpose(1, 2), (h.shape[1],), sd[p + ".block.2.weight"], sd[p + ".block.2.bias"]).transpose(1, 2)
    return F.mish(h) * mask


def _resnet(sd, p, x, mask, temb):
    h = _block(sd, p + ".block1", x, mask)
    h = h + F.linear(F.mish(temb), sd[p + ".mlp.1.weight"], sd[p + ".mlp.1.bias"]).unsqueeze(-1)
    h = _block(sd, p + ".block2", h, mask)
    return h + F.conv1d(x * mask, sd[p + ".res_conv.weight"], sd[p + ".res_conv.bias"])


def _tfm(sd, p, x, bias, heads):
    """BasicTransformerBlock without cross attention: x (B, T, C), bias (B, T, T) additive"""
    B, T, C = x.shape
    n = F.layer_norm(x, (C,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])
    q = F.linear(n, sd[p + ".attn1.to_q.weight"]).view(B, T, heads, -1).transpose(1, 2)
    k = F.linear(n, sd[p + ".attn1.to_k.weight"]).view(B, T, heads, -1).transpose(1, 2)
    v = F.linear(n, sd[p + ".attn1.to_v.weight"]).view(B, T, heads, -1).transpose(1, 2)
    a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(q.shape[-1]) + bias[:, None], dim=-1) @ v
    a = a.transpose(1, 2).reshape(B, T, -1)
    x = x + F.linear(a, sd[p + ".attn1.to_out.0.weight"], sd[p + ".attn1.to_out.0.bias"])
    n = F.layer_norm(x, (C,), sd[p + ".norm3.weight"], sd[p + ".norm3.bias"])
    f = F.gelu(F.linear(n, sd[p + ".ff.net.0.proj.weight"], sd[p + ".ff.net.0.proj.bias"]))
    return x + F.linear(f, sd[p + ".ff.net.2.weight"], sd[p + ".ff.net.2.bias"])


def attn_bias(mask, streaming, chunk):
    """(B, 1, T) 0/1 -> additive (B, T, T): padding mask, AND the block-causal chunk mask when streaming (mask.py:127-236),
    as mask_to_bias makes it: (1 - m) * -1e10"""
    B, _, T = mask.shape
    m = mask.bool().expand(B, T, T)
    if streaming:
        i = torch.arange(T)
        m = m & (i[None, :] < ((i[:, None] // chunk) + 1) * chunk)[None]
    return (1.0 - m.float()) * -1.0e10


def estimator(sd: Dict[str, torch.Tensor], x, mask, mu, t, spks, cond, dims, streaming=False):
    """x, mu, cond (B, mel, T); mask (B, 1, T); t (B,); spks (B, mel) -> (B, mel, T)   (decoder.py:405-494)"""
    sd = {k: v.float() for k, v in sd.items()}
    x, mask, mu, spks, cond = x.float(), mask.float(), mu.float(), spks.float(), cond.float()
    temb = time_embedding(sd, t, dims.in_ch)
    h = torch.cat([x, mu, spks.unsqueeze(-1).expand(-1, -1, x.shape[-1]), cond], dim=1)
    bias = attn_bias(mask, streaming, dims.chunk)

    def stage(p, h):
        h = _resnet(sd, p + ".0", h, mask, temb).transpose(1, 2)
        for j in range(dims.n_blocks):
            h = _tfm(sd, f"{p}.1.{j}", h, bias, dims.heads)
        return h.transpose(1, 2)

    h = stage("down_blocks.0", h)
    skip = h
    h = _cconv(h * mask, sd["down_blocks.0.2.weight"], sd["down_blocks.0.2.bias"])
    for i in range(dims.n_mid):
        h = stage(f"mid_blocks.{i}", h)
    h = stage("up_blocks.0", torch.cat([h[:, :, : skip.shape[-1]], skip], dim=1))
    h = _cconv(h * mask, sd["up_blocks.0.2.weight"], sd["up_blocks.0.2.bias"])
    h = _block(sd, "final_block", h, mask)
    return F.conv1d(h * mask, sd["final_proj.weight"], sd["final_proj.bias"]) * mask


def cfm_solve(sd, mu, spks, cond, noise, n_timesteps, dims, cfg_rate=0.7, temperature=1.0, streaming=False):
    """CausalConditionalCFM.forward + solve_euler (cosyvoice/flow/flow_matching.py:203-228,71-124) over `estimator`:
    mu, cond (1, mel, T), spks (1, mel), noise = rand_noise (1, mel, >=T) -> (1, mel, T)"""
    T = mu.shape[2]
    x = noise[:, :, :T].float() * temperature
    t_span = torch.linspace(0, 1, n_timesteps + 1)
    t_span = 1 - torch.cos(t_span * 0.5 * torch.pi)
    t, dt = t_span[0].unsqueeze(0), t_span[1] - t_span[0]
    mask = torch.ones(2, 1, T)
    zeros = torch.zeros_like(mu)
    for step in range(1, n_timesteps + 1):
        d = estimator(sd, torch.cat([x, x]), mask, torch.cat([mu, zeros]), torch.cat([t, t]), torch.cat([spks, torch.zeros_like(spks)]),
                      torch.cat([cond, zeros]), dims, streaming=streaming)
        d = (1.0 + cfg_rate) * d[:1] - cfg_rate * d[1:]
        x = x + dt * d
        t = t + dt
        if step < n_timesteps:
            dt = t_span[step + 1] - t
    return x


# ------------------------------------------------------------------------------------------------------------------
# The non-causal multi-level ConditionalDecoder (cosyvoice/flow/decoder.py:88-291): Block1D = Conv1d(k3, pad 1) -> GroupNorm(8)
# -> Mish (matcha/models/components/decoder.py:31-43), ResnetBlock1D :46-61, Downsample1D :64-70 (Conv1d k3 stride 2 pad 1),
# Upsample1D :116-158 (ConvTranspose1d(4, 2, 1)), skip concatenation and the padding-mask path (mask[:, :, ::2] per level).
def _block_nc(sd, p, x, mask, groups):
    h = F.conv1d(x * mask, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"], padding=1)
    h = F.group_norm(h, groups, sd[p + ".block.1.weight"], sd[p + ".block.1.bias"])
    return F.mish(h) * mask


def _resnet_nc(sd, p, x, mask, temb, groups):
    h = _block_nc(sd, p + ".block1", x, mask, groups)
    h = h + F.linear(F.mish(temb), sd[p + ".mlp.1.weight"], sd[p + ".mlp.1.bias"]).unsqueeze(-1)
    h = _block_nc(sd, p + ".block2", h, mask, groups)
    return h + F.conv1d(x * mask, sd[p + ".res_conv.weight"], sd[p + ".res_conv.bias"])


def estimator_nc(sd: Dict[str, torch.Tensor], x, mask, mu, t, spks, cond, dims):
    """x, mu, cond (B, mel, T); mask (B, 1, T) 0/1; t (B,); spks (B, mel) -> (B, mel, T)   (decoder.py:210-291)"""
    sd = {k: v.float() for k, v in sd.items()}
    x, mask, mu, spks, cond = x.float(), mask.float(), mu.float(), spks.float(), cond.float()
    temb = time_embedding(sd, t, dims.in_ch)
    h = torch.cat([x, mu, spks.unsqueeze(-1).expand(-1, -1, x.shape[-1]), cond], dim=1)
    L, G = dims.levels, dims.groups

    def stage(p, h, m):
        h = _resnet_nc(sd, p + ".0", h, m, temb, G).transpose(1, 2)
        bias = attn_bias(m, False, 0)
        for j in range(dims.n_blocks):
            h = _tfm(sd, f"{p}.1.{j}", h, bias, dims.heads)
        return h.transpose(1, 2)

    hiddens, masks = [], [mask]
    for i in range(L):
        m = masks[-1]
        h = stage(f"down_blocks.{i}", h, m)
        hiddens.append(h)
        if i == L - 1:
            h = F.conv1d(h * m, sd[f"down_blocks.{i}.2.weight"], sd[f"down_blocks.{i}.2.bias"], padding=1)
        else:
            h = F.conv1d(h * m, sd[f"down_blocks.{i}.2.conv.weight"], sd[f"down_blocks.{i}.2.conv.bias"], stride=2, padding=1)
        masks.append(m[:, :, ::2])
    masks = masks[:-1]
    m_mid = masks[-1]
    for i in range(dims.n_mid):
        h = stage(f"mid_blocks.{i}", h, m_mid)
    for i in range(L):
        m = masks.pop()
        skip = hiddens.pop()
        h = stage(f"up_blocks.{i}", torch.cat([h[:, :, : skip.shape[-1]], skip], dim=1), m)
        if i == L - 1:
            h = F.conv1d(h * m, sd[f"up_blocks.{i}.2.weight"], sd[f"up_blocks.{i}.2.bias"], padding=1)
        else:
            h = F.conv_transpose1d(h * m, sd[f"up_blocks.{i}.2.conv.weight"], sd[f"up_blocks.{i}.2.conv.bias"], stride=2, padding=1)
    h = _block_nc(sd, "final_block", h, m, G)
    return F.conv1d(h * m, sd["final_proj.weight"], sd["final_proj.bias"]) * mask


def cfm_forward_nc(sd, mu, z, spks, cond, n_timesteps, dims, cfg_rate=0.7, prompt_len=0, cache=None):
    """ConditionalCFM.forward + solve_euler (cosyvoice/flow/flow_matching.py:36-69,71-124) over estimator_nc.  z = the
    `torch.randn_like(mu) * temperature` draw of :52 (passed in: RNG pinned); cache (1, mel, n, 2) overwrites the head of z and mu
    (:53-57) and the new cache = [prompt | last 34 frames] of both (:58-60).  Returns (mel (1, mel, T), cache)."""
    mu, z = mu.float().clone(), z.float().clone()
    if cache is not None and cache.shape[2] != 0:
        cs = cache.shape[2]
        z[:, :, :cs] = cache[:, :, :, 0]
        mu[:, :, :cs] = cache[:, :, :, 1]
    new_cache = torch.stack([torch.cat([z[:, :, :prompt_len], z[:, :, -34:]], dim=2),
                             torch.cat([mu[:, :, :prompt_len], mu[:, :, -34:]], dim=2)], dim=-1)
    T = mu.shape[2]
    x = z
    t_span = torch.linspace(0, 1, n_timesteps + 1)
    t_span = 1 - torch.cos(t_span * 0.5 * torch.pi)
    t, dt = t_span[0].unsqueeze(0), t_span[1] - t_span[0]
    mask = torch.ones(2, 1, T)
    zeros = torch.zeros_like(mu)
    for step in range(1, n_timesteps + 1):
        d = estimator_nc(sd, torch.cat([x, x]), mask, torch.cat([mu, zeros]), torch.cat([t, t]),
                         torch.cat([spks, torch.zeros_like(spks)]), torch.cat([cond, zeros]), dims)
        d = (1.0 + cfg_rate) * d[:1] - cfg_rate * d[1:]
        x = x + dt * d
        t = t + dt
        if step < n_timesteps:
            dt = t_span[step + 1] - t
    return x, new_cache
