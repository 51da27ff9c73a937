"""The drop-in boundary against the reference's OWN host code.

The reference's `server/model_utils/infer_speech_model.py` is imported unmodified (oracle/refshim.py supplies import-only stand-ins
for the third-party modules of its frontend) and its `inference_zero_shot`, `inference_tts` and `ModelManager.load_pt` run against
stage objects auto-specced from NativeLLM / NativeFlow / NativeHiFT: `unittest.mock.create_autospec` enforces the native classes'
real signatures, so a keyword the reference passes that a native object would not accept — or a method the reference calls that
a native object does not have — fails here, on the CPU, without an engine.  The arithmetic behind the same calls is covered on
the GPU (tests/test_e2e_gpu.py::test_model_input_surface, tests/test_c1_gpu.py).  Needs /root/reference (build container only)."""
import inspect
import os
from functools import partial
from unittest import mock

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "server", "model_utils")), reason="reference tree not present")


@pytest.fixture(scope="module")
def ism():
    from oracle import refshim
    refshim.add_stub_roots("onnxruntime", "whisper", "inflect", "ttsfrd", "wetext", "pyworld", "soundfile")
    import server.model_utils.infer_speech_model as m
    return m


def _native_specs(n_tok=12):
    from flowmirror_hydravox_b200.flow import NativeFlow
    from flowmirror_hydravox_b200.hift import NativeHiFT
    from flowmirror_hydravox_b200.llm import NativeLLM
    llm = mock.create_autospec(NativeLLM, instance=True)
    flow = mock.create_autospec(NativeFlow, instance=True)
    hift = mock.create_autospec(NativeHiFT, instance=True)
    llm.inference.side_effect = lambda *a, **k: iter(range(100, 100 + n_tok))
    flow.inference.side_effect = lambda *a, **k: (torch.zeros(1, 80, 2 * k["token"].shape[1]), None)
    hift.inference.side_effect = lambda *a, **k: (torch.zeros(1, 480 * k["speech_feat"].shape[2]), torch.zeros(1, 1, 480 * k["speech_feat"].shape[2]))
    for m_ in (llm, flow, hift):                       # `.eval().cuda().to(dtype)` / `.half()` chains of load_models / load_pt
        for name in ("eval", "cuda", "to", "half"):    # (the reference never calls hift.half(): infer_speech_model.py:100-118)
            if hasattr(m_, name):
                getattr(m_, name).return_value = m_
    return llm, flow, hift


class _Frontend:
    """what CosyVoiceFrontEnd hands the stage chain (cosyvoice/cli/frontend.py:157-184)"""

    def text_normalize(self, text, split=True, text_frontend=True):
        return [text] if split else text

    def frontend_zero_shot(self, tts_text, prompt_text, prompt_audio, resample_rate, zero_shot_spk_id=""):
        i32 = torch.int32
        return {"text": torch.arange(7, dtype=i32)[None], "text_len": torch.tensor([7], dtype=i32),
                "prompt_text": torch.arange(3, dtype=i32)[None], "prompt_text_len": torch.tensor([3], dtype=i32),
                "llm_prompt_speech_token": torch.arange(5, dtype=i32)[None], "llm_prompt_speech_token_len": torch.tensor([5], dtype=i32),
                "flow_prompt_speech_token": torch.arange(5, dtype=i32)[None], "flow_prompt_speech_token_len": torch.tensor([5], dtype=i32),
                "prompt_speech_feat": torch.zeros(1, 10, 80), "prompt_speech_feat_len": torch.tensor([10], dtype=i32),
                "llm_embedding": torch.zeros(1, 192), "flow_embedding": torch.zeros(1, 192)}

    def frontend_sft(self, tts_text, spk_id):
        return {"text": torch.arange(7, dtype=torch.int32)[None], "text_len": torch.tensor([7], dtype=torch.int32),
                "llm_embedding": torch.zeros(192), "flow_embedding": torch.zeros(192)}


def _manager(ism, specs):
    mm = ism.ModelManager()
    mm.models = dict(zip(("llm", "flow", "hift"), specs))
    mm.frontend, mm.configs, mm.device, mm.is_loaded = _Frontend(), {"sample_rate": 24000}, "cpu", True
    return mm


def test_reference_zero_shot_chain_accepts_native_objects(ism):
    specs = _native_specs(12)
    mm = _manager(ism, specs)
    wav = ism.inference_zero_shot(mm, "hello world", "a prompt", torch.zeros(1, 16000), 16000, speed=1.0)
    assert wav.shape == (1, 480 * 24)
    llm, flow, hift = specs
    kw = llm.inference.call_args.kwargs            # infer_speech_model.py:549-557
    assert set(kw) == {"text", "text_len", "prompt_text", "prompt_text_len", "prompt_speech_token", "prompt_speech_token_len", "embedding"}
    kw = flow.inference.call_args.kwargs           # :570-580
    assert set(kw) == {"token", "token_len", "prompt_token", "prompt_token_len", "prompt_feat", "prompt_feat_len", "embedding",
                       "streaming", "finalize"}
    assert kw["token"].tolist() == [list(range(100, 112))] and kw["streaming"] is False and kw["finalize"] is True
    assert set(hift.inference.call_args.kwargs) == {"speech_feat"}      # :590-592
    # speed != 1 resamples the mel between the flow and the vocoder (:584-587)
    wav = ism.inference_zero_shot(mm, "hello world", "a prompt", torch.zeros(1, 16000), 16000, speed=2.0)
    assert wav.shape == (1, 480 * 12)


def test_reference_tts_chain_accepts_native_objects(ism):
    specs = _native_specs(9)
    mm = _manager(ism, specs)
    wav = ism.inference_tts(mm, "hello", "spk0", speed=1.0)
    assert wav.shape == (1, 480 * 18)
    kw = specs[0].inference.call_args.kwargs       # :631-639: empty 1-D prompt text, prompt_speech_token=None
    assert kw["prompt_speech_token"] is None and kw["prompt_text"].numel() == 0
    kw = specs[1].inference.call_args.kwargs       # :652-658: no prompt at all
    assert set(kw) == {"token", "token_len", "embedding", "streaming", "finalize"}
    with pytest.raises(ValueError):                # :660-661 through the reference's own error wrapping
        ism.inference_tts(mm, "hello", "spk0", speed=0.0)


def test_reference_load_pt_accepts_native_objects(ism, tmp_path):
    specs = _native_specs()
    mm = _manager(ism, specs)
    torch.save({"w": torch.zeros(1)}, tmp_path / "llm.pt")
    torch.save({"w": torch.ones(1)}, tmp_path / "flow.pt")
    out = mm.load_pt(str(tmp_path / "llm.pt"), str(tmp_path / "flow.pt"))          # :169-184
    assert out["status"] == "success", out
    llm, flow, _ = specs
    assert llm.load_state_dict.call_count == 1 and flow.load_state_dict.call_count == 1
    llm.to.assert_called_once_with(torch.bfloat16)
    flow.half.assert_called_once_with()
    assert llm.bf16 is True and flow.fp16 is True
    # a failing re-pack surfaces as the reference's {"status": "error"} (the engine keeps its previous weights: _lib.Engine two-phase swap)
    llm.load_state_dict.side_effect = RuntimeError("numel mismatch")
    assert mm.load_pt(str(tmp_path / "llm.pt"), str(tmp_path / "flow.pt"))["status"] == "error"


def test_native_llm_honours_the_workers_per_request_attributes():
    """server/worker.py:57-65 overwrites `models['llm'].sampling` (a functools.partial whose keywords carry top_p / top_k / win_size /
    tau_r) and `.inference_head_num` before every request; the native object reads both."""
    from flowmirror_hydravox_b200.llm import NativeLLM
    src = inspect.getsource(NativeLLM)
    assert "inference_head_num" in src and ".keywords" in src
    sig = inspect.signature(NativeLLM.inference)
    ref_kw = ["text", "text_len", "prompt_text", "prompt_text_len", "prompt_speech_token", "prompt_speech_token_len", "embedding",
              "sampling", "max_token_text_ratio", "min_token_text_ratio", "uuid"]                 # llm_multi_head_v3.py:926-939
    assert [p for p in sig.parameters if p != "self"] == ref_kw
    assert sig.parameters["max_token_text_ratio"].default == 20 and sig.parameters["min_token_text_ratio"].default == 2
    p = partial(lambda **k: None, top_p=0.9, top_k=10, win_size=24, tau_r=0.2)
    assert set(p.keywords) == {"top_p", "top_k", "win_size", "tau_r"}
