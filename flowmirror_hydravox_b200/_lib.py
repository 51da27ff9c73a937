"""ctypes binding of libhydravox_b200.so (include/hydravox_b200.h).  No CPU fallback: importing the
compute path without the built library, or creating an engine without a B200, raises."""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import dims as D

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HVX_LIB_PATH") or os.path.join(_HERE, "libhydravox_b200.so")      # HVX_LIB_PATH: A/B runs against another build

STAGE_LLM, STAGE_FLOW, STAGE_HIFT, STAGE_UNET = 0, 1, 2, 3
_DT = {torch.float32: 0, torch.bfloat16: 1, torch.int32: 2, torch.float16: 3}


class HvxError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [
        ("hift_mel", C.c_int), ("hift_base", C.c_int), ("hift_f0_ch", C.c_int), ("hift_harmonics", C.c_int),
        ("hift_sr", C.c_int), ("hift_n_ups", C.c_int), ("hift_ups", C.c_int * 4), ("hift_up_k", C.c_int * 4),
        ("hift_n_fft", C.c_int), ("hift_hop", C.c_int), ("hift_n_rb", C.c_int), ("hift_rb_k", C.c_int * 4),
        ("hift_n_dil", C.c_int), ("hift_rb_d", C.c_int * 4), ("hift_src_k", C.c_int * 4),
        ("flow_mel", C.c_int), ("flow_spk_in", C.c_int), ("flow_vocab", C.c_int), ("flow_pla_ch", C.c_int),
        ("flow_dim", C.c_int), ("flow_depth", C.c_int), ("flow_heads", C.c_int), ("flow_dim_head", C.c_int),
        ("flow_ff_mult", C.c_int), ("flow_chunk", C.c_int), ("flow_pos_k", C.c_int), ("flow_pos_groups", C.c_int),
        ("flow_noise_frames", C.c_int), ("flow_cfg_rate", C.c_float), ("flow_precise", C.c_int),
        ("llm_hidden", C.c_int), ("llm_layers", C.c_int), ("llm_q_heads", C.c_int), ("llm_kv_heads", C.c_int),
        ("llm_head_dim", C.c_int), ("llm_inter", C.c_int), ("llm_text_vocab", C.c_int), ("llm_speech_vocab", C.c_int),
        ("llm_mtp_heads", C.c_int), ("llm_mtp_inter", C.c_int), ("llm_max_ctx", C.c_int), ("llm_max_seqs", C.c_int),
        ("llm_rope_theta", C.c_float), ("llm_eps", C.c_float), ("llm_kv_f32", C.c_int),
        ("unet_mel", C.c_int), ("unet_ch", C.c_int), ("unet_n_blocks", C.c_int), ("unet_n_mid", C.c_int),
        ("unet_heads", C.c_int), ("unet_ff_mult", C.c_int), ("unet_chunk", C.c_int),
        ("unet_noncausal", C.c_int), ("unet_levels", C.c_int), ("unet_groups", C.c_int),
    ]


class Sampler(C.Structure):
    _fields_ = [("top_p", C.c_double), ("tau_r", C.c_double), ("top_k", C.c_int), ("win_size", C.c_int)]


class Request(C.Structure):
    _fields_ = [
        ("text_ids_host", C.c_void_p), ("n_text_total", C.c_int), ("n_text_new", C.c_int),
        ("prompt_speech_host", C.c_void_p), ("n_prompt_speech", C.c_int),
        ("prompt_feat_host", C.c_void_p), ("embedding_host", C.c_void_p),
        ("u_host", C.c_void_p), ("n_u", C.c_int),
        ("min_ratio", C.c_float), ("max_ratio", C.c_float), ("speed", C.c_double),
    ]


def make_config(hd: D.HiftDims, fd: D.FlowDims, ld: D.LlmDims, max_ctx: int = 8192, max_seqs: int = 32,
                kv_f32: bool = False, flow_precise: bool = False, ud: D.UnetDims | None = None) -> Config:
    c = Config()
    c.hift_mel, c.hift_base, c.hift_f0_ch, c.hift_harmonics, c.hift_sr = hd.mel, hd.base, hd.f0_ch, hd.harmonics, hd.sr
    c.hift_n_ups = len(hd.ups)
    for i, (u, k) in enumerate(zip(hd.ups, hd.up_k)):
        c.hift_ups[i], c.hift_up_k[i] = u, k
    c.hift_n_fft, c.hift_hop = hd.n_fft, hd.hop
    c.hift_n_rb, c.hift_n_dil = len(hd.rb_k), len(hd.rb_d)
    for i, k in enumerate(hd.rb_k):
        c.hift_rb_k[i] = k
    for i, k in enumerate(hd.rb_d):
        c.hift_rb_d[i] = k
    for i, k in enumerate(hd.src_k):
        c.hift_src_k[i] = k
    c.flow_mel, c.flow_spk_in, c.flow_vocab, c.flow_pla_ch = fd.mel, fd.spk_in, fd.vocab, fd.pla_ch
    c.flow_dim, c.flow_depth, c.flow_heads, c.flow_dim_head = fd.dim, fd.depth, fd.heads, fd.dim_head
    c.flow_ff_mult, c.flow_chunk, c.flow_pos_k, c.flow_pos_groups = fd.ff_mult, fd.chunk, fd.pos_k, fd.pos_groups
    c.flow_noise_frames, c.flow_cfg_rate = fd.noise_frames, fd.cfg_rate
    c.flow_precise = int(bool(flow_precise))
    c.llm_hidden, c.llm_layers, c.llm_q_heads, c.llm_kv_heads = ld.hidden, ld.layers, ld.q_heads, ld.kv_heads
    c.llm_head_dim, c.llm_inter, c.llm_text_vocab, c.llm_speech_vocab = ld.head_dim, ld.inter, ld.text_vocab, ld.speech_vocab
    c.llm_mtp_heads, c.llm_mtp_inter, c.llm_max_ctx, c.llm_max_seqs = ld.mtp_heads, ld.mtp_inter, max_ctx, max_seqs
    c.llm_rope_theta, c.llm_eps = ld.rope_theta, ld.eps
    c.llm_kv_f32 = int(bool(kv_f32))
    if ud is not None:
        c.unet_mel, c.unet_ch, c.unet_n_blocks, c.unet_n_mid = ud.mel, ud.ch, ud.n_blocks, ud.n_mid
        c.unet_heads, c.unet_ff_mult, c.unet_chunk = ud.heads, ud.ff_mult, ud.chunk
        if hasattr(ud, "channels"):                       # dims.UnetNcDims: the non-causal multi-level ConditionalDecoder
            c.unet_noncausal, c.unet_levels, c.unet_groups = 1, ud.levels, ud.groups
    return c


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HvxError(f"{LIB_PATH} is missing: run `python -m flowmirror_hydravox_b200.build` "
                           "(there is no CPU/PyTorch fallback for the hot path)")
        _lib = C.CDLL(LIB_PATH)
        _lib.hvx_last_error.restype = C.c_char_p
        _lib.hvx_kernel_launches.restype = C.c_int64
    return _lib


def check(rc: int):
    if rc != 0:
        raise HvxError(f"hvx error {rc}: {lib().hvx_last_error().decode()}")


def ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Engine:
    """Owns one hvx_engine (one per process/GPU, like one reference worker per GPU)."""

    def __init__(self, hd=D.HIFT_FULL, fd=D.FLOW_FULL, ld=D.LLM_FULL, max_ctx=8192, max_seqs=32, device="cuda:0",
                 kv_f32=False, flow_precise=False, ud=None):
        if not torch.cuda.is_available():
            raise HvxError("no CUDA device: the HydraVox B200 engine has no CPU fallback")
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)          # make sure the primary context exists
        self.hd, self.fd, self.ld, self.ud = hd, fd, ld, ud
        self.cfg = make_config(hd, fd, ld, max_ctx, max_seqs, kv_f32, flow_precise, ud)
        self.flow_precise = bool(flow_precise)
        self.h = C.c_void_p()
        check(lib().hvx_create(C.byref(self.h), C.byref(self.cfg)))
        self._keep = {0: {}, 1: {}, 2: {}, 3: {}}          # tensors borrowed by the engine
        self._pending = {0: None, 1: None, 2: None, 3: None}   # previous tensors of a stage between set_tensors and finalize

    def _register(self, stage: int, tensors: dict):
        for name, t in tensors.items():
            shape = (C.c_int64 * t.ndim)(*t.shape)
            check(lib().hvx_set_tensor(self.h, stage, name.encode(), ptr(t), _DT[t.dtype], shape, t.ndim))

    def set_tensors(self, stage: int, tensors: dict):
        """Stage a checkpoint: upload the packed tensors and register them with the engine.  The tensors the engine was
        serving from are kept alive until `finalize` succeeds; if it fails they are registered again, so a bad checkpoint
        (hot swap through `load_pt`, infer_speech_model.py:169-184) never leaves the engine pointing at freed memory — the
        reference's failed `load_state_dict` leaves the old weights in place too."""
        new = {name: t.to(self.device).contiguous() for name, t in tensors.items()}
        if self._pending[stage] is None:
            self._pending[stage] = dict(self._keep[stage])          # what to fall back to
        self._keep[stage] = {**self._keep[stage], **new}
        self._register(stage, new)

    def finalize(self, stage: int):
        old, self._pending[stage] = self._pending[stage], None
        rc = lib().hvx_finalize(self.h, stage)
        if rc != 0:
            msg = lib().hvx_last_error().decode()
            if old:                                                   # roll back: re-register the previous tensors, rebuild the stage
                torch.cuda.synchronize(self.device)
                self._keep[stage] = old
                self._register(stage, old)
                lib().hvx_finalize(self.h, stage)
            raise HvxError(f"hvx error {rc}: {msg}")

    def launches(self) -> int:
        return int(lib().hvx_kernel_launches(self.h))

    PROF_CLASSES = ("gemm", "attention", "hift_conv", "llm_step", "layernorm", "llm_prefill", "_6", "_7")

    def profile(self, on: bool):
        """bracket every kernel-class launch with CUDA events on its own stream (hvx_profile_enable)"""
        check(lib().hvx_profile_enable(self.h, int(bool(on))))

    def profile_collect(self) -> dict:
        """synchronise and return {class: (ms, work, call sites)} accumulated since the last collect (hvx_profile_collect)"""
        ms, work, n = (C.c_double * 8)(), (C.c_double * 8)(), (C.c_int64 * 8)()
        check(lib().hvx_profile_collect(self.h, ms, work, n))
        return {k: (ms[i], work[i], int(n[i])) for i, k in enumerate(self.PROF_CLASSES) if not k.startswith("_")}

    def close(self):
        if self.h:
            lib().hvx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
