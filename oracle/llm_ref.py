"""TEST INFRASTRUCTURE ONLY — CPU restatement (fp32 torch) of the multi-head AR speech-token decode.

Follows, function by function:
  cosyvoice/llm/llm_multi_head_v3.py:925-960  CosyVoice3LM.inference (prompt assembly, min/max len)
  cosyvoice/llm/llm_multi_head_v3.py:861-922  inference_wrapper else-branch (multi-head loop)
  cosyvoice/llm/llm_multi_head_v3.py:151-166  sampling_ids (EOS retry, <=100 trials)
  cosyvoice/llm/llm_multi_head_v3.py:248-260  Qwen2Encoder.forward_one_step
  cosyvoice/utils/common.py:138-166           ras_sampling / nucleus_sampling / random_sampling
Third-party arithmetic restated (not vendored in /root/reference): Hugging Face
transformers==4.40.1 (requirements.txt:42) Qwen2 — RMSNorm (fp32, eps 1e-6), q/k/v bias, half-split
RoPE (theta from config), GQA by repeat_kv, SwiGLU MLP, final norm; call sites
llm_multi_head_v3.py:24-26,240-258,658-666,887.

Two deliberate restatements:
  * the base model is KV-cached (K new rows per step) — identical to the reference's full-prefix
    recompute under causal attention;
  * an MTP head on one token with no cache is h1 = h + Wo(Wv n1(h) + bv); out = h1 + mlp(n2(h1))
    (softmax over one key == 1, RoPE at position 0 == identity; Wq/Wk are dead).
RNG: torch.multinomial is replaced by inverse-CDF sampling over an explicit uniform stream `u`
consumed in call order (one draw per multinomial call).  The reference is pinned to this oracle by
monkey-patching Tensor.multinomial to pop from the same stream (oracle/make_golden.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


class UStream:
    def __init__(self, u: torch.Tensor):
        self.u = u.flatten().tolist()
        self.pos = 0

    def next(self) -> float:
        v = self.u[self.pos]
        self.pos += 1
        return v


def multinomial_u(weights: torch.Tensor, u: float) -> int:
    """Inverse-CDF draw over unnormalised non-negative weights: first i with cumsum_i > u*total."""
    c = torch.cumsum(weights.float(), dim=0)
    target = torch.tensor(u, dtype=torch.float32) * c[-1]
    i = int(torch.searchsorted(c, target, right=True))
    return min(i, weights.numel() - 1)


def nucleus_sampling(logp, us: UStream, top_p=0.8, top_k=25):
    p = logp.float().softmax(dim=0)
    sv, si = p.sort(descending=True, stable=True)
    cum, n = 0.0, 0
    cum = torch.zeros((), dtype=torch.float32)
    while n < len(sv) and float(cum) < top_p and n < top_k:
        cum = cum + sv[n]
        n += 1
    return int(si[multinomial_u(sv[:n], us.next())])


def random_sampling(logp, us: UStream):
    return multinomial_u(logp.float().softmax(dim=0), us.next())


def ras_sampling(logp, decoded: List[int], us: UStream, top_p=0.8, top_k=25, win_size=10, tau_r=0.1):
    tid = nucleus_sampling(logp, us, top_p, top_k)
    window = decoded[-win_size:] if win_size != 0 else decoded
    rep = sum(1 for t in window if t == tid)
    if rep >= win_size * tau_r:
        tid = random_sampling(logp, us)
    return tid


def sampling_ids(logp, decoded, us, speech_token_size, ignore_eos, sp):
    trials = 0
    while True:
        tid = ras_sampling(logp, decoded, us, **sp)
        if (not ignore_eos) or tid < speech_token_size:
            return tid
        trials += 1
        if trials > 100:
            raise RuntimeError("sampling reaches max_trials 100 and still get eos when ignore_eos is True")


# --------------------------------------------------------------------------- Qwen2 math
def rmsnorm(x, w, eps=1e-6):
    v = x.float()
    v = v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + eps)
    return w.float() * v


def rope_half(x, pos, theta):
    """x (n, heads, 64); pos (n,). HF half-split convention."""
    d = x.shape[-1]
    inv = 1.0 / (theta ** (torch.arange(0, d, 2).float() / d))
    ang = pos.float()[:, None] * inv[None, :]
    cos = torch.cat([ang.cos(), ang.cos()], -1)[:, None, :]
    sin = torch.cat([ang.sin(), ang.sin()], -1)[:, None, :]
    x1, x2 = x[..., : d // 2], x[..., d // 2:]
    return x * cos + torch.cat([-x2, x1], -1) * sin


class LlmOracle:
    def __init__(self, sd: Dict[str, torch.Tensor], dims, kv_dtype=None):
        """kv_dtype=torch.bfloat16 rounds K (after RoPE) and V as they enter the cache, mirroring an engine that
        keeps its KV cache in bf16 (the reference's own GPU path keeps every activation in bf16)."""
        self.sd = {k: v.float() for k, v in sd.items()}
        self.d = dims
        self.kv_dtype = kv_dtype
        self.reset()

    def reset(self):
        self.k = [None] * self.d.layers
        self.v = [None] * self.d.layers
        self.n = 0

    def forward_rows(self, x: torch.Tensor) -> torch.Tensor:
        """x (n, hidden) new rows at positions self.n.. ; returns final-normed hidden (n, hidden)."""
        d, sd = self.d, self.sd
        n = x.shape[0]
        pos = torch.arange(self.n, self.n + n)
        h = x.float()
        for l in range(d.layers):
            p = f"llm.model.model.layers.{l}."
            a = rmsnorm(h, sd[p + "input_layernorm.weight"], d.eps)
            q = F.linear(a, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"]).view(n, d.q_heads, d.head_dim)
            k = F.linear(a, sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.k_proj.bias"]).view(n, d.kv_heads, d.head_dim)
            v = F.linear(a, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"]).view(n, d.kv_heads, d.head_dim)
            q, k = rope_half(q, pos, d.rope_theta), rope_half(k, pos, d.rope_theta)
            if self.kv_dtype is not None:
                k, v = k.to(self.kv_dtype).float(), v.to(self.kv_dtype).float()
            self.k[l] = k if self.k[l] is None else torch.cat([self.k[l], k], 0)
            self.v[l] = v if self.v[l] is None else torch.cat([self.v[l], v], 0)
            K, V = self.k[l], self.v[l]                                   # (ctx, kvh, 64)
            rep = d.q_heads // d.kv_heads
            Kr, Vr = K.repeat_interleave(rep, 1), V.repeat_interleave(rep, 1)
            s = torch.einsum("nhd,chd->hnc", q, Kr) * (d.head_dim ** -0.5)
            ctx = K.shape[0]
            causal = torch.arange(ctx)[None, :] <= pos[:, None]           # (n, ctx)
            s = s.masked_fill(~causal[None], float("-inf"))
            o = torch.einsum("hnc,chd->nhd", s.softmax(-1), Vr).reshape(n, d.q_heads * d.head_dim)
            h = h + F.linear(o, sd[p + "self_attn.o_proj.weight"])
            a = rmsnorm(h, sd[p + "post_attention_layernorm.weight"], d.eps)
            m = F.silu(F.linear(a, sd[p + "mlp.gate_proj.weight"])) * F.linear(a, sd[p + "mlp.up_proj.weight"])
            h = h + F.linear(m, sd[p + "mlp.down_proj.weight"])
        self.n += n
        return rmsnorm(h, sd["llm.model.model.norm.weight"], d.eps)

    def mtp_head(self, j: int, h: torch.Tensor) -> torch.Tensor:
        """h (hidden,) final-normed last hidden -> head output (hidden,)."""
        sd, d = self.sd, self.d
        p = f"mtp_block.{j}."
        a = rmsnorm(h, sd[p + "input_layernorm.weight"], d.eps)
        v = F.linear(a, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"])
        h1 = h + F.linear(v, sd[p + "self_attn.o_proj.weight"])
        a = rmsnorm(h1, sd[p + "post_attention_layernorm.weight"], d.eps)
        m = F.silu(F.linear(a, sd[p + "mlp.gate_proj.weight"])) * F.linear(a, sd[p + "mlp.up_proj.weight"])
        return h1 + F.linear(m, sd[p + "mlp.down_proj.weight"])

    def head_logp(self, j, h):
        return F.linear(self.mtp_head(j, h), self.sd["llm_decoder.weight"]).log_softmax(-1)

    def prompt_embeds(self, text, prompt_text, prompt_speech):
        sd, d = self.sd, self.d
        sos, task = d.speech_token_size, d.speech_token_size + 2
        t = torch.cat([prompt_text, text]).long()
        parts = [sd["speech_embedding.weight"][sos][None], sd["llm.model.model.embed_tokens.weight"][t],
                 sd["speech_embedding.weight"][task][None]]
        if prompt_speech.numel():
            parts.append(sd["speech_embedding.weight"][prompt_speech.long()])
        return torch.cat(parts, 0)


@torch.no_grad()
def inference(sd, dims, text, prompt_text, prompt_speech, u, head_k=1, sp=None,
              min_ratio=2.0, max_ratio=20.0, return_logp=False, kv_dtype=None):
    """1-D int tensors in; returns list of emitted speech tokens (and per-step head log-probs)."""
    sp = sp or dict(top_p=0.8, top_k=25, win_size=10, tau_r=0.1)
    m = LlmOracle(sd, dims, kv_dtype)
    head_k = max(1, min(int(head_k), dims.mtp_heads))          # llm_multi_head_v3.py:866-868
    us = u if isinstance(u, UStream) else UStream(u)
    x = m.prompt_embeds(text, prompt_text, prompt_speech)
    min_len, max_len = int(text.numel() * min_ratio), int(text.numel() * max_ratio)
    out: List[int] = []
    logps = []
    stop_from = dims.speech_token_size
    while len(out) < max_len:
        hid = m.forward_rows(x)
        last = hid[-1]
        lp = [m.head_logp(j, last) for j in range(head_k)]
        logps.append(torch.stack(lp))
        snap = list(out)
        ids = [sampling_ids(lp[j], snap, us, dims.speech_token_size, (len(snap) + j) < min_len, sp)
               for j in range(head_k)]
        group, stop = [], False
        for t in ids:
            if t >= stop_from:
                stop = True
                break
            out.append(t)
            group.append(t)
            if len(out) >= max_len:
                stop = True
                break
        if stop or not group:
            break
        x = m.sd["speech_embedding.weight"][torch.tensor(group)]
    return (out, logps) if return_logp else out
