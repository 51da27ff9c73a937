"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

Import the *unmodified* reference modules from ``/root/reference`` inside the
build container (the GPU box has no ``/root/reference``; nothing run there may
import this file).  Used by ``oracle/make_golden.py`` to pin the restated oracle
(``oracle/hift_ref.py``, ``oracle/flow_ref.py``, ``oracle/llm_ref.py``) against
the reference's own code and to mint the committed fixtures in ``tests/golden``.

What is shimmed:
  * import-only stubs for third-party modules that the reference imports at
    module scope but never calls on the hot path (lightning, hydra, diffusers,
    conformer, gdown, wget, matplotlib, omegaconf);
  * a numerical stand-in for ``x_transformers.x_transformers.RotaryEmbedding``
    / ``apply_rotary_pos_emb`` (x_transformers==2.12.2 is pinned by the
    reference's requirements.txt:49 but is not installed here) following the
    published 2.x semantics: interleaved frequencies, rotate_half on
    interleaved pairs, partial rotary on the first rot_dim channels, fp32 math;
  * a compat subclass of HF ``Qwen2DecoderLayer`` so the MTP heads can be
    called as ``layer(hidden)[0]`` under transformers>=4.48
    (llm_multi_head_v3.py:887 was written against 4.40.1).
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import os
import sys
import types

import torch
from torch import nn

REF_ROOT = os.environ.get("HVX_REFERENCE_ROOT", "/root/reference")

_STUB_ROOTS = (
    "lightning", "hydra", "gdown", "wget", "matplotlib", "conformer", "diffusers",
    "omegaconf", "x_transformers", "hyperpyyaml", "librosa",
)


class _Anything:
    """Attribute sink: any attribute/call/subscript returns another sink."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]          # behaves as an identity decorator
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __getitem__(self, k):
        return _Anything()

    def __mro_entries__(self, bases):
        return (object,)


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name in ("ConformerBlock",):
            return type(name, (nn.Module,), {"__init__": lambda self, *a, **k: nn.Module.__init__(self)})
        if name == "DictConfig":
            return _AttrDict
        return _Anything()


class _AttrDict(dict):
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


# ---------------------------------------------------------------------------
# x_transformers 2.x rotary stand-in (see module docstring)
# ---------------------------------------------------------------------------
class RotaryEmbedding(nn.Module):
    def __init__(self, dim, theta=10000.0):
        super().__init__()
        inv_freq = 1.0 / (theta ** (torch.arange(0, dim, 2).float() / dim))
        self.register_buffer("inv_freq", inv_freq)

    def forward_from_seq_len(self, seq_len):
        t = torch.arange(seq_len, device=self.inv_freq.device)
        return self.forward(t)

    @torch.autocast("cuda", enabled=False)
    def forward(self, t):
        if t.ndim == 1:
            t = t[None, :]
        freqs = torch.einsum("b i , j -> b i j", t.type_as(self.inv_freq), self.inv_freq)
        freqs = torch.stack((freqs, freqs), dim=-1).reshape(*freqs.shape[:-1], -1)  # interleaved
        return freqs, 1.0


def _rotate_half(x):
    x = x.reshape(*x.shape[:-1], -1, 2)
    x1, x2 = x.unbind(dim=-1)
    x = torch.stack((-x2, x1), dim=-1)
    return x.reshape(*x.shape[:-2], -1)


@torch.autocast("cuda", enabled=False)
def apply_rotary_pos_emb(t, freqs, scale=1.0):
    rot_dim, seq_len, orig_dtype = freqs.shape[-1], t.shape[-2], t.dtype
    freqs = freqs[:, -seq_len:, :]
    scale = scale[:, -seq_len:, :] if isinstance(scale, torch.Tensor) else scale
    if t.ndim == 4 and freqs.ndim == 3:
        freqs = freqs[:, None]
    t, t_unrotated = t[..., :rot_dim], t[..., rot_dim:]
    t = (t * freqs.cos() * scale) + (_rotate_half(t) * freqs.sin() * scale)
    out = torch.cat((t, t_unrotated), dim=-1)
    return out.type(orig_dtype)


_installed = False


# ---------------------------------------------------------------------------
# diffusers 0.29 stand-ins for the U-Net estimator's transformer blocks (matcha/models/components/transformer.py:5-14):
# published semantics of `Attention` with AttnProcessor2_0 (self-attention, no norms), `GELU`, `LoRACompatibleLinear`.
# ---------------------------------------------------------------------------
class DiffusersGELU(nn.Module):
    def __init__(self, dim_in, dim_out, approximate="none", bias=True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)
        self.approximate = approximate

    def forward(self, hidden_states):
        return torch.nn.functional.gelu(self.proj(hidden_states), approximate=self.approximate)


class DiffusersAttention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, out_bias=True, **_):
        super().__init__()
        inner = dim_head * heads
        self.heads = heads
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(kv_dim, inner, bias=bias)
        self.to_v = nn.Linear(kv_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=out_bias), nn.Dropout(dropout)])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **_):
        B, T, _c = hidden_states.shape
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q, k, v = self.to_q(hidden_states), self.to_k(ctx), self.to_v(ctx)
        hd = q.shape[-1] // self.heads
        q, k, v = (z.view(B, -1, self.heads, hd).transpose(1, 2) for z in (q, k, v))
        if attention_mask is not None:                     # prepare_attention_mask: (B, T, S) -> per head
            attention_mask = attention_mask.repeat_interleave(self.heads, dim=0).view(B, self.heads, -1, attention_mask.shape[-1])
        o = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(B, -1, self.heads * hd).to(q.dtype)
        return self.to_out[1](self.to_out[0](o))


def _install_diffusers():
    att = importlib.import_module("diffusers.models.attention")
    att.GELU = DiffusersGELU
    importlib.import_module("diffusers.models.attention_processor").Attention = DiffusersAttention
    importlib.import_module("diffusers.models.lora").LoRACompatibleLinear = nn.Linear
    importlib.import_module("diffusers.utils.torch_utils").maybe_allow_in_graph = lambda cls: cls
    importlib.import_module("diffusers.models.activations").get_activation = lambda name: {"silu": nn.SiLU, "swish": nn.SiLU, "mish": nn.Mish, "gelu": nn.GELU}[name]()


def install():
    """Make ``cosyvoice.*`` / ``matcha.*`` importable from the reference tree."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference tree not found at {REF_ROOT} (refshim is container-only)")
    # resolve the real third-party stacks BEFORE the stub finder can answer their
    # optional-dependency probes (transformers checks find_spec("matplotlib") etc.)
    import transformers  # noqa: F401
    from transformers import Qwen2ForCausalLM  # noqa: F401
    from transformers.models.qwen2 import modeling_qwen2  # noqa: F401
    import torchaudio  # noqa: F401
    # never stub a module that really exists
    real = tuple(r for r in _STUB_ROOTS if _really_exists(r))
    finder = _StubFinder()
    finder_roots = tuple(r for r in _STUB_ROOTS if r not in real)
    globals()["_STUB_ROOTS"] = finder_roots
    sys.meta_path.append(finder)
    sys.path[:0] = [os.path.join(REF_ROOT, "server", "model_utils"), REF_ROOT]
    xt = importlib.import_module("x_transformers.x_transformers")
    xt.RotaryEmbedding = RotaryEmbedding
    xt.apply_rotary_pos_emb = apply_rotary_pos_emb
    if "diffusers" in finder_roots:
        _install_diffusers()
    if "librosa" in finder_roots:          # matcha/utils/audio.py:4 — librosa.filters.mel, restated (slaney scale and norm)
        from flowmirror_hydravox_b200.frontend import slaney_mel_basis
        importlib.import_module("librosa.filters").mel = (
            lambda sr, n_fft, n_mels, fmin, fmax: slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax).numpy())
    _installed = True


def add_stub_roots(*roots):
    """Import-only stand-ins for further third-party modules (e.g. the frontend's onnxruntime / whisper / inflect / wetext when the
    reference's server-side host module is imported for a call-sequence test); a module that really exists is never stubbed."""
    install()
    globals()["_STUB_ROOTS"] = tuple(_STUB_ROOTS) + tuple(r for r in roots if r not in _STUB_ROOTS and not _really_exists(r))


def _really_exists(root):
    for p in sys.path:
        if os.path.isdir(os.path.join(p, root)) or os.path.isfile(os.path.join(p, root + ".py")):
            return True
    return False


# ---------------------------------------------------------------------------
# constructors for the three reference stages at arbitrary (reduced) dims
# ---------------------------------------------------------------------------
def build_hift(cfg, seed=0):
    """cfg: oracle.dims.HiftDims. Returns the reference CausalHiFTGenerator (eval)."""
    install()
    from cosyvoice.hifigan.generator import CausalHiFTGenerator
    from cosyvoice.hifigan.f0_predictor import CausalConvRNNF0Predictor
    torch.manual_seed(seed)
    m = CausalHiFTGenerator(
        in_channels=cfg.mel, base_channels=cfg.base, nb_harmonics=cfg.harmonics - 1,
        sampling_rate=cfg.sr, nsf_alpha=0.1, nsf_sigma=0.003, nsf_voiced_threshold=10,
        upsample_rates=list(cfg.ups), upsample_kernel_sizes=list(cfg.up_k),
        istft_params={"n_fft": cfg.n_fft, "hop_len": cfg.hop},
        resblock_kernel_sizes=list(cfg.rb_k), resblock_dilation_sizes=[list(cfg.rb_d)] * len(cfg.rb_k),
        source_resblock_kernel_sizes=list(cfg.src_k),
        source_resblock_dilation_sizes=[list(cfg.rb_d)] * len(cfg.src_k),
        lrelu_slope=0.1, audio_limit=0.99, conv_pre_look_right=4,
        f0_predictor=CausalConvRNNF0Predictor(1, cfg.mel, cfg.f0_ch))
    # random weights: make snake alphas and biases non-trivial so parity tests bite
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if n.endswith(".alpha"):
                p.copy_(0.5 + torch.rand(p.shape, generator=g))
            elif n.endswith(".bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    return m.eval()


def build_hift_t(cfg, seed=0):
    """The reference's non-causal HiFTGenerator (ConvTranspose1d upsampling, generator.py:378-569) at cfg dims (eval)."""
    install()
    from cosyvoice.hifigan.generator import HiFTGenerator
    from cosyvoice.hifigan.f0_predictor import ConvRNNF0Predictor
    torch.manual_seed(seed)
    m = HiFTGenerator(
        in_channels=cfg.mel, base_channels=cfg.base, nb_harmonics=cfg.harmonics - 1,
        sampling_rate=cfg.sr, nsf_alpha=0.1, nsf_sigma=0.003, nsf_voiced_threshold=10,
        upsample_rates=list(cfg.ups), upsample_kernel_sizes=list(cfg.up_k),
        istft_params={"n_fft": cfg.n_fft, "hop_len": cfg.hop},
        resblock_kernel_sizes=list(cfg.rb_k), resblock_dilation_sizes=[list(cfg.rb_d)] * len(cfg.rb_k),
        source_resblock_kernel_sizes=list(cfg.src_k),
        source_resblock_dilation_sizes=[list(cfg.rb_d)] * len(cfg.src_k),
        lrelu_slope=0.1, audio_limit=0.99, f0_predictor=ConvRNNF0Predictor(1, cfg.mel, cfg.f0_ch))
    return m.eval()


def build_hifigan(cfg):
    """The reference's classic HiFi-GAN `Generator` (matcha/hifigan/models.py:148-193) at cfg dims (eval, weight-normed)."""
    install()
    from matcha.hifigan.models import Generator
    h = _AttrDict(resblock="1", upsample_rates=list(cfg.ups), upsample_kernel_sizes=list(cfg.up_k),
                  upsample_initial_channel=cfg.base, resblock_kernel_sizes=list(cfg.rb_k),
                  resblock_dilation_sizes=[list(cfg.rb_d)] * len(cfg.rb_k))
    return Generator(h).eval()


def build_unet(cfg):
    """The reference's CausalConditionalDecoder (cosyvoice/flow/decoder.py:294-494) at cfg dims (oracle.dims.UnetDims), eval."""
    install()
    from cosyvoice.flow.decoder import CausalConditionalDecoder
    m = CausalConditionalDecoder(in_channels=cfg.in_ch, out_channels=cfg.mel, channels=[cfg.ch], dropout=0.0,
                                 attention_head_dim=cfg.head_dim, n_blocks=cfg.n_blocks, num_mid_blocks=cfg.n_mid,
                                 num_heads=cfg.heads, act_fn="gelu", static_chunk_size=cfg.chunk, num_decoding_left_chunks=-1)
    return m.eval()


def build_unet_nc(cfg):
    """The reference's non-causal multi-level ConditionalDecoder (cosyvoice/flow/decoder.py:88-291) at cfg dims (dims.UnetNcDims)."""
    install()
    from cosyvoice.flow.decoder import ConditionalDecoder
    m = ConditionalDecoder(in_channels=cfg.in_ch, out_channels=cfg.mel, channels=list(cfg.channels), dropout=0.0,
                           attention_head_dim=cfg.head_dim, n_blocks=cfg.n_blocks, num_mid_blocks=cfg.n_mid,
                           num_heads=cfg.heads, act_fn="gelu")
    return m.eval()


def build_unet_nc_cfm(cfg):
    """The non-causal ConditionalCFM (flow_matching.py:22-69) over the reference ConditionalDecoder, cfm_params as the CosyVoice yaml."""
    install()
    from cosyvoice.flow.flow_matching import ConditionalCFM
    return ConditionalCFM(
        in_channels=3 * cfg.mel, n_spks=1, spk_emb_dim=cfg.mel,
        cfm_params=_AttrDict(sigma_min=1e-6, solver="euler", t_scheduler="cosine", training_cfg_rate=0.2,
                             inference_cfg_rate=0.7, reg_loss_type="l1"),
        estimator=build_unet_nc(cfg)).eval()


def build_unet_cfm(cfg):
    """CausalConditionalCFM (flow_matching.py:197-228) over the reference U-Net estimator, cfm_params as the CosyVoice2 yaml."""
    install()
    from cosyvoice.flow.flow_matching import CausalConditionalCFM
    return CausalConditionalCFM(
        in_channels=3 * cfg.mel, n_spks=1, spk_emb_dim=cfg.mel,
        cfm_params=_AttrDict(sigma_min=1e-6, solver="euler", t_scheduler="cosine", training_cfg_rate=0.2,
                             inference_cfg_rate=0.7, reg_loss_type="l1"),
        estimator=build_unet(cfg)).eval()


def build_flow(cfg, seed=0, dtype=torch.float32):
    install()
    from cosyvoice.flow.flow import CausalMaskedDiffWithDiT
    from cosyvoice.flow.flow_matching import CausalConditionalCFM
    from cosyvoice.flow.DiT.dit import DiT
    from cosyvoice.transformer.upsample_encoder import PreLookaheadLayer
    torch.manual_seed(seed)
    est = DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, dim_head=cfg.dim_head, ff_mult=cfg.ff_mult,
              mel_dim=cfg.mel, mu_dim=cfg.mel, spk_dim=cfg.mel, out_channels=cfg.mel,
              static_chunk_size=cfg.chunk, num_decoding_left_chunks=-1)
    cfm = CausalConditionalCFM(
        in_channels=3 * cfg.mel, n_spks=1, spk_emb_dim=cfg.mel,
        cfm_params=_AttrDict(sigma_min=1e-6, solver="euler", t_scheduler="cosine", training_cfg_rate=0.2,
                             inference_cfg_rate=0.7, reg_loss_type="l1"),
        estimator=est)
    m = CausalMaskedDiffWithDiT(
        input_size=cfg.mel, output_size=cfg.mel, spk_embed_dim=cfg.spk_in, vocab_size=cfg.vocab,
        input_frame_rate=25, token_mel_ratio=2, pre_lookahead_len=3,
        pre_lookahead_layer=PreLookaheadLayer(cfg.mel, cfg.pla_ch, 3), decoder=cfm)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.ndim >= 2 and "embedding" not in n:
                p.copy_(torch.randn(p.shape, generator=g) * (0.7 / (p[0].numel() ** 0.5)))
            elif n.endswith(".bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    m.bf16 = dtype == torch.bfloat16
    m.fp16 = dtype == torch.float16
    return m.to(dtype).eval()


def build_llm(cfg, seed=0, dtype=torch.float32):
    install()
    import cosyvoice.llm.llm_multi_head_v3 as v3
    from transformers import Qwen2ForCausalLM
    from transformers.models.qwen2.configuration_qwen2 import Qwen2Config
    from transformers.models.qwen2 import modeling_qwen2 as mq

    class _CompatLayer(mq.Qwen2DecoderLayer):
        """transformers>=4.48 compat: layer(h) -> (h',) with position 0.. rotary."""

        def __init__(self, config, layer_idx):
            config._attn_implementation = "eager"
            super().__init__(config, layer_idx)
            self._rot = mq.Qwen2RotaryEmbedding(config)

        def forward(self, hidden_states, **kw):
            L = hidden_states.shape[1]
            pos = torch.arange(L, device=hidden_states.device)[None]
            pe = self._rot(hidden_states, pos)
            out = super().forward(hidden_states, position_embeddings=pe, attention_mask=None)
            return out if isinstance(out, tuple) else (out,)

    v3.Qwen2DecoderLayer = _CompatLayer
    _orig_cfg = v3.Qwen2Config

    def _mtp_cfg(**kw):
        kw.setdefault("intermediate_size", cfg.mtp_inter)
        return _orig_cfg(**kw)

    v3.Qwen2Config = _mtp_cfg if cfg.mtp_inter != 22016 else _orig_cfg
    torch.manual_seed(seed)
    enc = v3.Qwen2Encoder.__new__(v3.Qwen2Encoder)
    nn.Module.__init__(enc)
    qc = Qwen2Config(hidden_size=cfg.hidden, num_hidden_layers=cfg.layers, num_attention_heads=cfg.q_heads,
                     num_key_value_heads=cfg.kv_heads, intermediate_size=cfg.inter, vocab_size=cfg.text_vocab,
                     rope_theta=cfg.rope_theta, tie_word_embeddings=True, rms_norm_eps=1e-6,
                     max_position_embeddings=32768)
    qc._attn_implementation = "eager"
    enc.model = Qwen2ForCausalLM(qc)
    m = v3.CosyVoice3LM(llm_input_size=cfg.hidden, llm_output_size=cfg.hidden, speech_token_size=cfg.speech_vocab - 200,
                        llm=enc, sampling=None, head_num=cfg.mtp_heads, inference_head_num=cfg.mtp_heads,
                        mtp_head_num=cfg.mtp_attn_heads)
    v3.Qwen2Config = _orig_cfg
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "norm" in n:
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif p.ndim >= 2:
                p.copy_(torch.randn(p.shape, generator=g) * (1.0 / (p.shape[-1] ** 0.5)))
            elif n.endswith(".bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
    return m.to(dtype).eval()
