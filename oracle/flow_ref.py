"""TEST INFRASTRUCTURE ONLY — CPU restatement (fp32 torch) of the token->mel flow stage.

Follows, function by function:
  cosyvoice/flow/flow.py:367-430                 CausalMaskedDiffWithDiT.inference (pre-net, cond, slicing)
  cosyvoice/transformer/upsample_encoder.py:82-103 PreLookaheadLayer.forward
  cosyvoice/flow/flow_matching.py:71-124,203-228 solve_euler / CausalConditionalCFM.forward (CFG, cosine t)
  cosyvoice/flow/DiT/dit.py:145-176              DiT.forward (+ InputEmbedding :76-98)
  cosyvoice/flow/DiT/modules.py:71-83,115-144,230-282,349-407,500-530,606-616
  cosyvoice/utils/mask.py:127-158                subsequent_chunk_mask (streaming mask)
Third-party arithmetic restated: x_transformers==2.12.2 RotaryEmbedding/apply_rotary_pos_emb
(interleaved pairs, first 64 channels of the un-split 1024-wide q/k, fp32) — not vendored in the
reference; its published algorithm is restated here and anchored on call sites DiT/dit.py:158,
DiT/modules.py:367-373.

The oracle computes in fp32 with an fp32 ODE state (the reference keeps everything in fp16/bf16);
`state_dtype`/`model_dtype` knobs emulate the reference's low-precision carries when needed.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def prenet(sd, token, embedding, dims, prompt_token=None, prompt_feat=None, finalize=True):
    """token (1,N) int, embedding (1,spk_in), prompt_token (1,P), prompt_feat (1,2P,mel)
    -> mu (1,mel,T), spks (1,mel), cond (1,mel,T), mel_len1."""
    emb = F.normalize(embedding.float(), dim=1)
    spks = F.linear(emb, sd["spk_embed_affine_layer.weight"], sd["spk_embed_affine_layer.bias"])
    if prompt_token is not None:
        token = torch.cat([prompt_token, token], dim=1)
    tok = F.embedding(torch.clamp(token, min=0), sd["input_embedding.weight"])         # (1,L,mel)
    x = tok.transpose(1, 2)
    if finalize:
        y = F.pad(x, (0, 3))
        res = tok
    else:
        y = x                                            # last 3 tokens are the look-ahead context
        res = tok[:, :-3]
    y = F.leaky_relu(F.conv1d(y, sd["pre_lookahead_layer.conv1.weight"], sd["pre_lookahead_layer.conv1.bias"]))
    y = F.conv1d(F.pad(y, (2, 0)), sd["pre_lookahead_layer.conv2.weight"], sd["pre_lookahead_layer.conv2.bias"])
    h = y.transpose(1, 2) + res
    h = h.repeat_interleave(2, dim=1)                    # token_mel_ratio
    T = h.shape[1]
    mel_len1 = 0 if prompt_feat is None else prompt_feat.shape[1]
    cond = torch.zeros(1, T, dims.mel)
    if prompt_feat is not None:
        cond[:, :mel_len1] = prompt_feat.float()
    return h.transpose(1, 2).contiguous(), spks, cond.transpose(1, 2).contiguous(), mel_len1


def time_embed(sd, t, p="decoder.estimator.time_embed."):
    half = 128
    e = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half).float() * -e)
    a = 1000.0 * t.float().unsqueeze(1) * freqs.unsqueeze(0)
    th = torch.cat((a.sin(), a.cos()), dim=-1)
    th = F.linear(th, sd[p + "time_mlp.0.weight"], sd[p + "time_mlp.0.bias"])
    return F.linear(F.silu(th), sd[p + "time_mlp.2.weight"], sd[p + "time_mlp.2.bias"])


def rotary_first_head(q, T, rot_dim=64):
    """x_transformers partial rotary on channels [0, rot_dim) of (B,T,D), interleaved pairs."""
    inv = 1.0 / (10000.0 ** (torch.arange(0, rot_dim, 2).float() / rot_dim))
    ang = torch.arange(T).float()[:, None] * inv[None, :]             # (T, rot/2)
    cos, sin = ang.cos(), ang.sin()
    x = q[..., :rot_dim].reshape(*q.shape[:-1], rot_dim // 2, 2)
    x1, x2 = x[..., 0], x[..., 1]
    r1 = x1 * cos - x2 * sin
    r2 = x2 * cos + x1 * sin
    rot = torch.stack((r1, r2), dim=-1).reshape(*q.shape[:-1], rot_dim)
    return torch.cat((rot, q[..., rot_dim:]), dim=-1)


def attn_mask(T, lens, streaming, chunk):
    """(B,T,T) bool: key j visible to query i."""
    B = len(lens)
    ar = torch.arange(T)
    key_ok = ar[None, :] < torch.as_tensor(lens)[:, None]                  # (B,T)
    m = key_ok[:, None, :].expand(B, T, T)
    if streaming:
        blk = (ar // chunk + 1) * chunk
        m = m & (ar[None, :] < blk[:, None])[None]
    return m


def dit_forward(sd, x, mu, t, spks, cond, dims, lens=None, streaming=False):
    """x, mu, cond: (B,mel,T); t: (B,); spks: (B,mel) -> (B,mel,T)   [DiT/dit.py:145-176]"""
    p = "decoder.estimator."
    B, _, T = x.shape
    lens = [T] * B if lens is None else lens
    temb = time_embed(sd, t)
    xin = torch.cat([x.transpose(1, 2), cond.transpose(1, 2), mu.transpose(1, 2),
                     spks[:, None, :].expand(B, T, spks.shape[-1])], dim=-1)
    h = F.linear(xin, sd[p + "input_embed.proj.weight"], sd[p + "input_embed.proj.bias"])
    k = dims.pos_k
    c = h.transpose(1, 2)
    for name in ("conv1", "conv2"):
        c = F.pad(c, (k - 1, 0))
        c = F.mish(F.conv1d(c, sd[p + f"input_embed.conv_pos_embed.{name}.0.weight"],
                            sd[p + f"input_embed.conv_pos_embed.{name}.0.bias"], groups=dims.pos_groups))
    h = c.transpose(1, 2) + h
    mask = attn_mask(T, lens, streaming, dims.chunk)[:, None]              # (B,1,T,T)
    row_ok = (torch.arange(T)[None, :] < torch.as_tensor(lens)[:, None])[..., None]
    H, dh = dims.heads, dims.dim_head
    st = F.silu(temb)
    for i in range(dims.depth):
        bp = p + f"transformer_blocks.{i}."
        mod = F.linear(st, sd[bp + "attn_norm.linear.weight"], sd[bp + "attn_norm.linear.bias"])
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = mod.chunk(6, dim=1)
        n = F.layer_norm(h, (dims.dim,), eps=1e-6) * (1 + sc_a[:, None]) + sh_a[:, None]
        q = F.linear(n, sd[bp + "attn.to_q.weight"], sd[bp + "attn.to_q.bias"])
        kk = F.linear(n, sd[bp + "attn.to_k.weight"], sd[bp + "attn.to_k.bias"])
        v = F.linear(n, sd[bp + "attn.to_v.weight"], sd[bp + "attn.to_v.bias"])
        q, kk = rotary_first_head(q, T), rotary_first_head(kk, T)
        q = q.view(B, T, H, dh).transpose(1, 2)
        kk = kk.view(B, T, H, dh).transpose(1, 2)
        v = v.view(B, T, H, dh).transpose(1, 2)
        s = (q @ kk.transpose(-1, -2)) * (dh ** -0.5)
        s = s.masked_fill(~mask, float("-inf"))
        a = torch.softmax(s, dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, T, H * dh)
        a = F.linear(a, sd[bp + "attn.to_out.0.weight"], sd[bp + "attn.to_out.0.bias"])
        a = a.masked_fill(~row_ok, 0.0)
        h = h + g_a[:, None] * a
        n = F.layer_norm(h, (dims.dim,), eps=1e-6) * (1 + sc_m[:, None]) + sh_m[:, None]
        f = F.gelu(F.linear(n, sd[bp + "ff.ff.0.0.weight"], sd[bp + "ff.ff.0.0.bias"]), approximate="tanh")
        f = F.linear(f, sd[bp + "ff.ff.2.weight"], sd[bp + "ff.ff.2.bias"])
        h = h + g_m[:, None] * f
    mod = F.linear(st, sd[p + "norm_out.linear.weight"], sd[p + "norm_out.linear.bias"])
    sc, sh = mod.chunk(2, dim=1)
    h = F.layer_norm(h, (dims.dim,), eps=1e-6) * (1 + sc)[:, None] + sh[:, None]
    return F.linear(h, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"]).transpose(1, 2)


def t_schedule(n_timesteps, dtype=torch.float32):
    ts = torch.linspace(0, 1, n_timesteps + 1, dtype=dtype)
    return 1 - torch.cos(ts * 0.5 * torch.pi)


def solve_euler(sd, mu, spks, cond, noise, dims, n_timesteps=10, streaming=False, return_all=False):
    """CFG Euler solve (flow_matching.py:71-124): row0 conditional, row1 mu=spks=cond=0."""
    T = mu.shape[2]
    x = noise[:, :, :T].float().clone()
    ts = t_schedule(n_timesteps)
    t, dt = ts[0], ts[1] - ts[0]
    sol = []
    for step in range(1, n_timesteps + 1):
        xin = torch.cat([x, x], dim=0)
        muin = torch.cat([mu, torch.zeros_like(mu)], dim=0)
        spin = torch.cat([spks, torch.zeros_like(spks)], dim=0)
        cin = torch.cat([cond, torch.zeros_like(cond)], dim=0)
        tin = t.reshape(1).repeat(2)
        d = dit_forward(sd, xin, muin, tin, spin, cin, dims, streaming=streaming)
        v = (1.0 + dims.cfg_rate) * d[:1] - dims.cfg_rate * d[1:]
        x = x + dt * v
        t = t + dt
        sol.append(x)
        if step < n_timesteps:
            dt = ts[step + 1] - t
    return sol if return_all else x


@torch.no_grad()
def inference(sd, token, embedding, noise, dims, n_timesteps=10, prompt_token=None, prompt_feat=None,
              streaming=False, finalize=True):
    sd = {k: v.float() for k, v in sd.items()}
    mu, spks, cond, l1 = prenet(sd, token, embedding, dims, prompt_token, prompt_feat, finalize)
    x = solve_euler(sd, mu, spks, cond, noise, dims, n_timesteps, streaming)
    return x[:, :, l1:]
