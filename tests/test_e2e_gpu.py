"""End to end through the reference-facing surface: ModelManager (server/model_utils mirror) -> C-ABI.

tiny dims for all three stages so the CPU oracle chain (llm_ref -> flow_ref -> hift_ref) finishes in seconds."""
import pytest
import torch
import torch.nn.functional as F

from flowmirror_hydravox_b200 import dims as D, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mm():
    from flowmirror_hydravox_b200.model_manager import ModelManager
    m = ModelManager(hd=D.HIFT_TINY, fd=D.FLOW_TINY, ld=D.LLM_TINY, max_ctx=512, max_seqs=4, n_timesteps=5, sine_seconds=20.0)
    sds = (synth.llm_state_dict(D.LLM_TINY, 0, eos_scale=0.0), synth.flow_state_dict(D.FLOW_TINY, 0), synth.hift_state_dict(D.HIFT_TINY, 0))
    m.load_state_dicts(*sds)
    m.sds = sds
    yield m
    m.engine.close()


def _requests():
    g = torch.Generator().manual_seed(1986)
    ld, fd = D.LLM_TINY, D.FLOW_TINY
    reqs = []
    for n_text, P in ((6, 5), (4, 0), (9, 12)):
        r = dict(text=torch.randint(0, ld.text_vocab, (n_text,), generator=g, dtype=torch.int32),
                 prompt_text=torch.randint(0, ld.text_vocab, (3 if P else 0,), generator=g, dtype=torch.int32),
                 prompt_speech=torch.randint(0, min(ld.speech_token_size, fd.vocab), (P,), generator=g, dtype=torch.int32),
                 prompt_feat=(torch.rand(2 * P, fd.mel, generator=g) * -6.0) if P else None,
                 embedding=torch.rand(fd.spk_in, generator=g))
        reqs.append(r)
    return reqs


def test_pipeline_matches_oracle_chain(mm):
    from oracle import flow_ref, hift_ref, llm_ref
    reqs = _requests()
    sp = dict(top_p=0.9, top_k=10, win_size=24, tau_r=0.2)
    u = torch.rand(len(reqs), 1024, generator=torch.Generator().manual_seed(5))
    wavs, toks = mm.synthesize_batch(reqs, head_k=2, sampling=sp, n_timesteps=5, min_ratio=4, max_ratio=4, u=u, return_tokens=True)
    llm_sd, flow_sd, hift_sd = mm.sds
    llm_sdb = {k: (v.to(torch.bfloat16).float() if v.ndim >= 2 else v.float()) for k, v in llm_sd.items()}
    noise = mm.models["flow"].noise.cpu()[None]
    table = mm.models["hift"].sine_table.cpu()
    for i, r in enumerate(reqs):
        ref_tok = llm_ref.inference(llm_sdb, D.LLM_TINY, r["text"], r["prompt_text"], r["prompt_speech"], u[i], head_k=2, sp=sp,
                                    min_ratio=4, max_ratio=4, kv_dtype=torch.bfloat16)
        assert toks[i] == ref_tok and len(ref_tok) == 4 * r["text"].numel()
        P = r["prompt_speech"].numel()
        mel = flow_ref.inference(flow_sd, torch.tensor(ref_tok)[None], r["embedding"][None], noise, D.FLOW_TINY, 5,
                                 r["prompt_speech"][None].long() if P else None, r["prompt_feat"][None] if P else None)
        eng_mel, _ = mm.models["flow"].inference(token=torch.tensor(ref_tok)[None], embedding=r["embedding"][None],
                                                 prompt_token=r["prompt_speech"][None] if P else None,
                                                 prompt_feat=r["prompt_feat"][None] if P else None, n_timesteps=5)
        e_mel = (eng_mel.cpu() - mel).abs().max().item()
        # vocoder on the engine's own mel with the oracle's F0 pinned (hift_ref.inference docstring)
        w = hift_ref.fold_weight_norm(hift_sd)
        f0 = hift_ref.f0_predict(w, eng_mel.cpu())
        ref_wav, _ = hift_ref.inference(hift_sd, eng_mel.cpu(), table, D.HIFT_TINY, f0=f0)
        eng_wav, _ = mm.models["hift"].inference(eng_mel, f0=f0)
        rms = (eng_wav.cpu() - ref_wav).pow(2).mean().sqrt().item()
        print(f"[e2e {i}] tokens {len(ref_tok)} equal; mel max-abs {e_mel:.2e}; wav rms (F0 pinned) {rms:.2e}")
        assert e_mel < 1e-2 and rms < 1e-4
        assert wavs[i].shape == (1, 2 * len(ref_tok) * D.HIFT_TINY.frame_samples) and torch.isfinite(wavs[i]).all()
        # the one-call path and the stage-by-stage path are the same kernels on the same inputs
        stage_wav, _ = mm.models["hift"].inference(eng_mel)
        assert (stage_wav.cpu() - wavs[i]).abs().max().item() < 1e-5


def test_model_input_surface(mm):
    """inference_zero_shot / inference_tts on the frontend's model_input dict (frontend.py:157-184)."""
    from flowmirror_hydravox_b200.model_manager import inference_tts, inference_zero_shot
    from functools import partial
    r = _requests()[0]
    mm.models["llm"].sampling = partial(lambda **kw: None, top_p=0.9, top_k=10, win_size=24, tau_r=0.2)   # worker.py:57-63
    mm.models["llm"].inference_head_num = 2
    mi = {"text": r["text"][None], "text_len": torch.tensor([r["text"].numel()]), "prompt_text": r["prompt_text"][None],
          "prompt_text_len": torch.tensor([3]), "llm_prompt_speech_token": r["prompt_speech"][None],
          "llm_prompt_speech_token_len": torch.tensor([5]), "flow_prompt_speech_token": r["prompt_speech"][None],
          "flow_prompt_speech_token_len": torch.tensor([5]), "prompt_speech_feat": r["prompt_feat"][None],
          "prompt_speech_feat_len": torch.tensor([10]), "llm_embedding": r["embedding"][None], "flow_embedding": r["embedding"][None]}
    wav = inference_zero_shot(mm, mi, speed=1.0)
    assert wav.device.type == "cpu" and wav.dim() == 2 and wav.shape[1] % D.HIFT_TINY.frame_samples == 0 and wav.abs().max() <= 0.99
    wav2 = inference_tts(mm, {"text": r["text"][None], "text_len": torch.tensor([6]), "llm_embedding": r["embedding"][None],
                              "flow_embedding": r["embedding"][None]}, speed=1.25)
    assert wav2.shape[1] % D.HIFT_TINY.frame_samples == 0
    with pytest.raises(ValueError):
        inference_tts(mm, mi, speed=0.0)


@pytest.mark.parametrize("T,speed", [(17, 0.5), (200, 1.3), (64, 2.0), (1, 0.7)])
def test_speed_interp_matches_torch(mm, T, speed):
    from flowmirror_hydravox_b200.model_manager import speed_interp
    mel = torch.randn(1, 80, T, generator=torch.Generator().manual_seed(T))
    t_out = max(1, int(T / speed))
    ref = F.interpolate(mel, size=t_out, mode="linear")
    out = speed_interp(mm, mel, t_out).cpu()
    assert out.shape == ref.shape and (out - ref).abs().max().item() < 2e-6


def test_load_pt_contract(mm, tmp_path):
    p = tmp_path / "flow.pt"
    sd = dict(mm.sds[1]); sd["epoch"] = 3
    torch.save(sd, p)
    assert mm.load_pt(None, str(p))["status"] == "success"
    bad = mm.load_pt(str(tmp_path / "missing.pt"), None)
    assert bad["status"] == "error" and "error" in bad and "message" in bad


def test_segmented_synthesis_batches_independent_segments(mm):
    """inference_tts_with_segmentation(last_prompt=False): the segments are independent utterances, so one batched call +
    the reference's pause stitching must equal per-segment synthesis stitched the same way."""
    import random
    from flowmirror_hydravox_b200 import output
    reqs = _requests()
    u = torch.rand(len(reqs), 1024, generator=torch.Generator().manual_seed(9))
    kw = dict(head_k=2, sampling=dict(top_p=0.9, top_k=10, win_size=24, tau_r=0.2), n_timesteps=5, min_ratio=4, max_ratio=4)
    merged = output.synthesize_segments(mm, reqs, rng=random.Random(1), u=u, **kw)
    singles = [mm.synthesize_batch([r], u=u[i:i + 1], **kw)[0] for i, r in enumerate(reqs)]
    ref = output.stitch_segments(singles, 24000, random.Random(1))
    assert merged.shape == ref.shape and (merged - ref).abs().max().item() < 1e-5
    assert len(output.audio_to_base64(merged, 24000)) > 100
