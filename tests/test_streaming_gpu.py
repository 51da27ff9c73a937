"""Streaming orchestration (cli/model.py:315-430) on the B200: chunk schedule, causal consistency and
parity of every streamed mel chunk with the oracle's flow (streaming mask, finalize=False) on the same tokens."""
import pytest
import torch

from flowmirror_hydravox_b200 import dims as D, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mm():
    from flowmirror_hydravox_b200.model_manager import ModelManager
    m = ModelManager(hd=D.HIFT_TINY, fd=D.FLOW_TINY, ld=D.LLM_TINY, max_ctx=1024, max_seqs=2, n_timesteps=4, sine_seconds=30.0)
    sds = (synth.llm_state_dict(D.LLM_TINY, 0, eos_scale=0.0), synth.flow_state_dict(D.FLOW_TINY, 0), synth.hift_state_dict(D.HIFT_TINY, 0))
    m.load_state_dicts(*sds)
    m.sds = sds
    yield m
    m.engine.close()


def test_streaming_matches_oracle_and_offline(mm):
    from flowmirror_hydravox_b200.streaming import StreamingSynthesizer
    from oracle import flow_ref
    r = synth.utterance(D.LLM_TINY, D.FLOW_TINY, 12, seed=1986, prompt_tokens=7, prompt_text=3)
    u = torch.rand(1, 2048, generator=torch.Generator().manual_seed(2))
    sp = dict(top_p=0.9, top_k=10, win_size=24, tau_r=0.2)
    dbg = {}
    s = StreamingSynthesizer(mm)
    chunks = [c["tts_speech"] for c in s.tts(r, head_k=2, sampling=sp, n_timesteps=4, min_ratio=8, max_ratio=8, u=u, debug=dbg)]
    toks = dbg["tokens"]
    assert len(toks) == 96
    # chunk schedule of CosyVoice2Model.tts: first hop = 25 + prompt pad (18) tokens, then 25, all + 3 look-ahead; then the rest
    assert dbg["n_tok"] == [46, 71, 96, 96] or dbg["n_tok"] == [46, 71, 96]
    frame = D.HIFT_TINY.frame_samples
    wav = torch.cat(chunks, dim=1)
    assert wav.shape[1] == 2 * 96 * frame and torch.isfinite(wav).all()
    # every streamed mel chunk == oracle flow with the streaming mask on the same token prefix
    flow_sd = mm.sds[1]
    noise = mm.models["flow"].noise.cpu()[None]
    off = 0
    for mel, n_tok in zip(dbg["mel"], dbg["n_tok"]):
        fin = n_tok == 96 and off + mel.shape[2] // 2 >= 96
        # the final pass runs under full attention (cli/model.py:352-358 does not pass `stream` to the last token2wav)
        ref = flow_ref.inference(flow_sd, torch.tensor(toks[:n_tok])[None], r["embedding"][None], noise, D.FLOW_TINY, 4,
                                 r["prompt_speech"][None].long(), r["prompt_feat"][None], streaming=not fin, finalize=fin)
        ref = ref[:, :, 2 * off:]
        assert ref.shape == mel.shape
        assert (ref - mel).abs().max().item() < 1e-2
        off += mel.shape[2] // 2
    assert off == 96
    # causal vocoder: the streamed waveform equals one offline pass over the accumulated mel (generator.py:729-746)
    full, _ = mm.models["hift"].inference(speech_feat=dbg["mel_cache"]())
    assert (full.cpu() - wav).abs().max().item() < 5e-3
    assert dbg["first_audio_ms"] > 0


def test_abandoned_stream_does_not_leak_into_next_request(mm):
    """A consumer that stops iterating (client disconnect) cancels the decode; the next request on the same synthesizer
    gets exactly the tokens it gets on a fresh one."""
    from flowmirror_hydravox_b200.streaming import StreamingSynthesizer
    r1 = synth.utterance(D.LLM_TINY, D.FLOW_TINY, 24, seed=5, prompt_tokens=7, prompt_text=3)
    r2 = synth.utterance(D.LLM_TINY, D.FLOW_TINY, 12, seed=6, prompt_tokens=7, prompt_text=3)
    u = torch.rand(1, 4096, generator=torch.Generator().manual_seed(3))
    sp = dict(top_p=0.9, top_k=10, win_size=24, tau_r=0.2)
    s = StreamingSynthesizer(mm)
    ref = {}
    list(s.tts(r2, head_k=2, sampling=sp, n_timesteps=2, min_ratio=8, max_ratio=8, u=u, debug=ref))
    gen = s.tts(r1, head_k=2, sampling=sp, n_timesteps=2, min_ratio=8, max_ratio=8, u=u)
    next(gen)                      # first chunk only
    gen.close()                    # abandon: the finally block cancels + joins the LLM thread
    dbg = {}
    wav = torch.cat([c["tts_speech"] for c in s.tts(r2, head_k=2, sampling=sp, n_timesteps=2, min_ratio=8, max_ratio=8, u=u, debug=dbg)], 1)
    assert dbg["tokens"] == ref["tokens"] and len(dbg["tokens"]) == 96
    assert torch.isfinite(wav).all()


def test_streaming_cv2_fade_in_out_matches_reference_orchestration(mm):
    """CosyVoice2Model.token2wav (cli/model.py:279-313): 8-frame mel cache, source continuation, fade_in_out cross-fade and the
    3840-sample hold-back.  The streamed chunks must equal the restated orchestration (oracle/stream_ref.py) driven over the same
    per-call mel slices with the same vocoder and the same pinned source noise; hvx_fade_in_out is bit-exact vs the reference blend."""
    from flowmirror_hydravox_b200 import _lib as L
    from flowmirror_hydravox_b200.hift import NativeHiFTTransposed
    from flowmirror_hydravox_b200.streaming import StreamingSynthesizerCV2
    from oracle import stream_ref
    hd = D.HIFT_TINY
    e = L.Engine(hd=hd)
    try:
        ht = NativeHiFTTransposed(e)
        ht.load_state_dict(synth.hift_t_state_dict(hd, 0))
        noise_all = torch.randn(400 * hd.frame_samples, hd.harmonics, generator=torch.Generator().manual_seed(11)).cuda()
        noise_fn = lambda n: noise_all[:n]
        r = synth.utterance(D.LLM_TINY, D.FLOW_TINY, 12, seed=1986, prompt_tokens=7, prompt_text=3)
        u = torch.rand(1, 2048, generator=torch.Generator().manual_seed(2))
        sp = dict(top_p=0.9, top_k=10, win_size=24, tau_r=0.2)
        dbg = {}
        s = StreamingSynthesizerCV2(mm, ht, noise_fn=noise_fn)
        chunks = [c["tts_speech"] for c in s.tts(r, head_k=2, sampling=sp, n_timesteps=4, min_ratio=8, max_ratio=8, u=u, debug=dbg)]
        assert len(dbg["tokens"]) == 96 and len(chunks) == len(dbg["mel"]) >= 3
        frame, ov = hd.frame_samples, 8 * hd.frame_samples
        # every non-final chunk holds `ov` samples back, the final one releases them: the total is the whole utterance
        assert sum(c.shape[1] for c in chunks) == 2 * 96 * frame
        assert chunks[0].shape[1] == dbg["mel"][0].shape[2] * frame - ov

        def vocoder(mel, cache_source):
            w, src = ht.inference(mel, cache_source=cache_source, noise=noise_fn(mel.shape[2] * frame))
            return w.cpu(), src.cpu()
        ref = stream_ref.token2wav_cv2(vocoder, dbg["mel"], mel_cache_len=8, frame_samples=frame)
        for i, (a, b) in enumerate(zip(chunks, ref)):
            assert a.shape == b.shape, (i, a.shape, b.shape)
            assert torch.equal(a, b), (i, (a - b).abs().max().item())
    finally:
        e.close()


def test_streaming_incremental_equals_recompute(mm):
    """StreamingSynthesizer(incremental=True): the same chunks as the recompute-all orchestration (the flow of every non-final chunk
    runs through a key / value cache session, the final chunk is the reference's full-attention pass either way)"""
    from flowmirror_hydravox_b200.streaming import StreamingSynthesizer
    r = synth.utterance(D.LLM_TINY, D.FLOW_TINY, 12, seed=1986, prompt_tokens=7, prompt_text=3)
    u = torch.rand(1, 2048, generator=torch.Generator().manual_seed(2))
    sp = dict(top_p=0.9, top_k=10, win_size=24, tau_r=0.2)
    res = {}
    for inc in (False, True):
        dbg = {}
        s = StreamingSynthesizer(mm, incremental=inc)
        chunks = [c["tts_speech"] for c in s.tts(r, head_k=2, sampling=sp, n_timesteps=4, min_ratio=8, max_ratio=8, u=u, debug=dbg)]
        res[inc] = (chunks, dbg)
    (c0, d0), (c1, d1) = res[False], res[True]
    assert d0["tokens"] == d1["tokens"] and d0["n_tok"] == d1["n_tok"] and len(c0) == len(c1)
    for a, b in zip(d0["mel"], d1["mel"]):
        assert a.shape == b.shape and (a - b).abs().max().item() < 2e-3
    for a, b in zip(c0, c1):
        assert a.shape == b.shape and (a - b).abs().max().item() < 5e-3
