"""Model dimensions of the HydraVox hot path (SURVEY.md §8 header).

``hydravox.yaml`` is not in the reference tree; the FULL presets are the
CosyVoice3-derived dims pinned by scripts/post_process/
add_mtp_weights_to_cosyvoice3lm_ckpt.py:130-154, llm_multi_head_v3.py:643-652,
flow.py:436-451 and generator.py:739-746.  TINY presets keep every structural
property (GQA, partial rotary, grouped conv, 3 upsample stages, MTP heads) at
sizes whose weights fit in a committed fixture.
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import Tuple


@dataclass(frozen=True)
class HiftDims:
    mel: int = 80
    base: int = 512
    f0_ch: int = 512
    harmonics: int = 9            # nb_harmonics(8) + fundamental
    sr: int = 24000
    ups: Tuple[int, ...] = (8, 5, 3)
    up_k: Tuple[int, ...] = (16, 11, 7)
    n_fft: int = 16
    hop: int = 4
    rb_k: Tuple[int, ...] = (3, 7, 11)
    rb_d: Tuple[int, ...] = (1, 3, 5)
    src_k: Tuple[int, ...] = (7, 7, 11)

    @property
    def frame_samples(self) -> int:      # 480 at full dims
        p = self.hop
        for u in self.ups:
            p *= u
        return p


@dataclass(frozen=True)
class FlowDims:
    mel: int = 80
    spk_in: int = 192
    vocab: int = 6561
    pla_ch: int = 1024
    dim: int = 1024
    depth: int = 22
    heads: int = 16
    dim_head: int = 64
    ff_mult: int = 2
    chunk: int = 50               # static_chunk_size (frames) for the streaming mask
    pos_k: int = 31
    pos_groups: int = 16
    cfg_rate: float = 0.7
    noise_frames: int = 15000


@dataclass(frozen=True)
class UnetDims:
    """CausalConditionalDecoder (cosyvoice/flow/decoder.py:294-400) as the CosyVoice2-generation configs instantiate it:
    in_channels 320 = x|mu|spks|cond, channels [256] (one down / one up block, no resampling), 4 transformer blocks per
    stage, 12 mid blocks, 8 heads x 64, GELU feed-forward x4, 50-frame streaming chunks."""
    mel: int = 80
    ch: int = 256
    n_blocks: int = 4
    n_mid: int = 12
    heads: int = 8
    head_dim: int = 64
    ff_mult: int = 4
    chunk: int = 50

    @property
    def in_ch(self) -> int:
        return 4 * self.mel

    @property
    def n_res(self) -> int:               # resnet blocks: down + mid + up
        return self.n_mid + 2


@dataclass(frozen=True)
class UnetNcDims:
    """The non-causal multi-level `ConditionalDecoder` (cosyvoice/flow/decoder.py:88-291; blocks
    matcha/models/components/decoder.py:31-158) as the CosyVoice-generation configs instantiate it: in_channels 320, channels
    [256, 256] (two down / two up levels: Downsample1D = Conv1d(k3, stride 2) and Upsample1D = ConvTranspose1d(4, 2, 1) between
    them, plain Conv1d(k3) on the last), GroupNorm(8) Block1Ds, 4 transformer blocks per stage, 12 mid blocks, 8 heads x 64."""
    mel: int = 80
    channels: tuple = (256, 256)
    n_blocks: int = 4
    n_mid: int = 12
    heads: int = 8
    head_dim: int = 64
    ff_mult: int = 4
    groups: int = 8
    chunk: int = 0

    @property
    def in_ch(self) -> int:
        return 4 * self.mel

    @property
    def ch(self) -> int:
        assert len(set(self.channels)) == 1, "the engine builds equal-width levels (the reference configs use [256, 256])"
        return self.channels[0]

    @property
    def levels(self) -> int:
        return len(self.channels)

    @property
    def n_res(self) -> int:               # resnet blocks: down levels + mid + up levels
        return self.n_mid + 2 * len(self.channels)


@dataclass(frozen=True)
class LlmDims:
    hidden: int = 896
    layers: int = 24
    q_heads: int = 14
    kv_heads: int = 2
    head_dim: int = 64
    inter: int = 4864
    text_vocab: int = 151936
    speech_vocab: int = 6761      # speech_token_size(6561) + 200
    rope_theta: float = 1e6
    mtp_heads: int = 5
    mtp_attn_heads: int = 14
    mtp_inter: int = 22016
    eps: float = 1e-6

    @property
    def speech_token_size(self) -> int:
        return self.speech_vocab - 200


HIFT_FULL = HiftDims()
FLOW_FULL = FlowDims()
LLM_FULL = LlmDims()
UNET_FULL = UnetDims()
UNET_TINY = UnetDims(mel=16, ch=128, n_blocks=2, n_mid=2, heads=2, chunk=6)
UNET_NC_FULL = UnetNcDims()
UNET_NC_TINY = UnetNcDims(mel=16, channels=(128, 128), n_blocks=1, n_mid=1, heads=2)
UNET_NC_SMALL = UnetNcDims(channels=(128, 128), n_blocks=1, n_mid=2, heads=2)   # mel stays 80: solve_euler hard-codes it (flow_matching.py:94-99)
UNET_SMALL = UnetDims(ch=128, n_blocks=1, n_mid=1, heads=2, chunk=10)     # mel stays 80: solve_euler hard-codes it (flow_matching.py:94-99)

HIFT_TINY = HiftDims(base=64, f0_ch=64)
# classic HiFi-GAN v1 generator (matcha/hifigan/config.py:1-28, models.py:148-193): 22.05 kHz, hop 256 = 8*8*2*2, no ISTFT head
# (hop=1 makes frame_samples the product of the rates); f0_ch / harmonics / n_fft / src_k are unused by this variant
HIFIGAN_V1 = HiftDims(sr=22050, ups=(8, 8, 2, 2), up_k=(16, 16, 4, 4), hop=1)
HIFIGAN_TINY = HiftDims(sr=22050, base=64, ups=(4, 2, 2), up_k=(8, 4, 4), hop=1, rb_k=(3, 7))
FLOW_TINY = FlowDims(vocab=512, pla_ch=128, depth=2, noise_frames=600)   # dim stays 1024: the reference hard-codes 16 conv groups
LLM_TINY = LlmDims(hidden=128, layers=2, q_heads=2, kv_heads=1, inter=256, text_vocab=512,
                   speech_vocab=456, mtp_heads=3, mtp_attn_heads=2, mtp_inter=384)


def to_dict(d):
    return asdict(d)
