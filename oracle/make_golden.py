"""TEST INFRASTRUCTURE ONLY — mint tests/golden/*.pt from the UNMODIFIED reference modules.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
For every stage and for TINY and FULL dims it
  1. builds the reference module (oracle/refshim.py) and loads the synthetic checkpoint
     (flowmirror_hydravox_b200/synth.py) with strict=True  -> pins parameter names/shapes;
  2. runs the reference's own inference on seeded inputs;
  3. runs the restated oracle on the same inputs and asserts agreement;
  4. stores inputs + reference outputs (+ a checkpoint checksum) as a small fixture.
The fixtures travel to the GPU box; the reference does not.
"""
from __future__ import annotations

import os
import sys
from functools import partial

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from flowmirror_hydravox_b200 import dims as D, synth  # noqa: E402
from oracle import flow_ref, frontend_ref, hift_ref, hifigan_ref, llm_ref, refshim, unet_ref  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))


class _TorchF32Proxy:
    """flow.py:378-385 casts to bf16/fp16 unconditionally; evaluate the reference in fp32 by
    presenting torch.bfloat16/float16 as float32 to that module only (source untouched)."""

    def __getattr__(self, n):
        return torch.float32 if n in ("bfloat16", "float16") else getattr(torch, n)


def golden_hift(name, dims, T, seed):
    m = refshim.build_hift(dims)
    sd = synth.hift_state_dict(dims, seed)
    m.load_state_dict(sd, strict=True)
    table = synth.hift_sine_table(dims, T)
    m.m_source.l_sin_gen.sine_waves = table[None]
    g = torch.Generator().manual_seed(seed + 100)
    mel = torch.rand(1, dims.mel, T, generator=g) * 6.0 - 6.0
    wav, src = m.inference(speech_feat=mel, finalize=True)
    f0 = m.f0_predictor(mel)
    w = hift_ref.fold_weight_norm(sd)
    f0_o = hift_ref.f0_predict(w, mel)
    ef0 = ((f0 - f0_o).abs() / (f0.abs() + 1.0)).max().item()
    wav_o, src_o = hift_ref.inference(sd, mel, table, dims, f0=f0)          # F0 pinned
    e = (wav - wav_o).abs().max().item()
    wav_free, _ = hift_ref.inference(sd, mel, table, dims)                  # F0 from the oracle itself
    e_free = (wav - wav_free).abs().max().item()
    wav_s, _ = m.inference(speech_feat=mel, finalize=False)
    f0_s = m.f0_predictor(mel, finalize=False)
    wav_so, _ = hift_ref.inference(sd, mel, table, dims, finalize=False, f0=f0_s)
    es = (wav_s - wav_so).abs().max().item()
    voiced = (f0 > 10).float().mean().item()
    print(f"[hift:{name}] T={T} ref-vs-oracle: f0 rel {ef0:.2e}; wav max-abs (F0 pinned) {e:.2e} (stream {es:.2e}), "
          f"(F0 free) {e_free:.2e}; rms {wav.pow(2).mean().sqrt():.3f} sat {(wav.abs() >= 0.99).float().mean():.3f} "
          f"voiced {voiced:.2f} f0 max {f0.max():.0f}")
    assert ef0 < 1e-4 and e < 5e-6 and es < 5e-6
    torch.save(dict(dims=name, seed=seed, T=T, mel=mel, wav=wav, src=src, f0=f0, wav_stream=wav_s, f0_stream=f0_s,
                    sd_checksum=checksum(sd)), os.path.join(OUT, f"hift_{name}.pt"))


def c2_hift_mel(dims, T, seed):
    """the mel of the config-2 size HiFT fixture (regenerated from the seed by the test: a CPU generator is bit-reproducible)"""
    g = torch.Generator().manual_seed(seed + 100)
    return torch.rand(1, dims.mel, T, generator=g) * 6.0 - 6.0


def golden_hift_c2(dims, T, seed):
    """BASELINE config-2 size (1024 speech tokens -> 2048 mel frames = 40.96 s): the reference module's waveform and its
    CPU-fp32 F0 track.  Only wav + f0 are stored (3.9 MB); mel and the sine table are regenerated from their seeds."""
    m = refshim.build_hift(dims)
    sd = synth.hift_state_dict(dims, seed)
    m.load_state_dict(sd, strict=True)
    table = synth.hift_sine_table(dims, T)
    m.m_source.l_sin_gen.sine_waves = table[None]
    mel = c2_hift_mel(dims, T, seed)
    wav, _ = m.inference(speech_feat=mel, finalize=True)
    f0 = m.f0_predictor(mel)
    wav_o, _ = hift_ref.inference(sd, mel, table, dims, f0=f0)
    e = (wav - wav_o).abs().max().item()
    rms = (wav - wav_o).pow(2).mean().sqrt().item()
    wav_free, _ = hift_ref.inference(sd, mel, table, dims)
    rms_free = (wav - wav_free).pow(2).mean().sqrt().item()
    print(f"[hift:c2] T={T} ref-vs-oracle wav max-abs {e:.2e} rms {rms:.2e} (F0 pinned); rms {rms_free:.2e} (oracle's own F0); "
          f"voiced {(f0 > 10).float().mean():.2f}")
    assert e < 2e-5
    torch.save(dict(dims="c2", seed=seed, T=T, wav=wav, f0=f0, oracle_free_f0_rms=rms_free, sd_checksum=checksum(sd)),
               os.path.join(OUT, "hift_c2.pt"))


def golden_llm_c2(dims, seed, n_text=128, n_ptext=16, P=125, K=2, ratio=8, depths=(0, 1, 64, 255, 511)):
    """BASELINE config-2 size: 16 + 128 text tokens, 125 prompt speech tokens, inference_head_num=2, fixed-length protocol
    (min = max ratio 8 -> 1024 speech tokens) through the UNMODIFIED CosyVoice3LM.inference (no KV cache: ~300 TFLOP
    on the CPU).  Weights are the synthetic checkpoint rounded to bf16 (what the engine holds; the reference serves in bf16,
    infer_speech_model.py:102) evaluated in fp32, stop logits zeroed (synth.llm_state_dict eos_scale=0) as in bench.py.
    Stored: the 1024 token ids, and the teacher-forced head log-probs of the reference modules at several depths."""
    refshim.install()
    from cosyvoice.utils.common import ras_sampling
    sp = dict(top_p=0.9, top_k=10, win_size=24, tau_r=0.2)
    m = refshim.build_llm(dims)
    sd = {k: v.to(torch.bfloat16).float() for k, v in synth.llm_state_dict(dims, seed, eos_scale=0.0).items()}
    m.load_state_dict(sd, strict=True)
    u0 = synth.utterance(dims, D.FLOW_FULL, n_text, seed=1986, prompt_tokens=P, prompt_text=n_ptext)
    text, ptext, pspeech = u0["text"].long()[None], u0["prompt_text"].long()[None], u0["prompt_speech"].long()[None]
    u = torch.rand(8192, generator=torch.Generator().manual_seed(seed + 7))
    m.sampling = partial(ras_sampling, **sp)
    m.inference_head_num = K
    us = llm_ref.UStream(u)
    orig = torch.Tensor.multinomial
    torch.Tensor.multinomial = lambda self, n, replacement=False: torch.tensor([llm_ref.multinomial_u(self, us.next())])
    import time
    t0 = time.time()
    try:
        ref = list(m.inference(text=text, text_len=torch.tensor([n_text]), prompt_text=ptext, prompt_text_len=torch.tensor([n_ptext]),
                               prompt_speech_token=pspeech, prompt_speech_token_len=torch.tensor([P]), embedding=None,
                               min_token_text_ratio=ratio, max_token_text_ratio=ratio))
    finally:
        torch.Tensor.multinomial = orig
    t_ref = time.time() - t0
    ora, logps = llm_ref.inference(sd, dims, text[0], ptext[0], pspeech[0], u, head_k=K, sp=sp, min_ratio=ratio, max_ratio=ratio,
                                   return_logp=True)
    n_same = next((i for i, (a, b) in enumerate(zip(ref, ora)) if a != b), min(len(ref), len(ora)))
    print(f"[llm:c2] reference {len(ref)} tokens in {t_ref:.0f} s (u used {us.pos}); oracle {len(ora)}, common prefix {n_same}")
    assert len(ref) == n_text * ratio
    # teacher-forced head log-probs from the reference modules at several decode depths (step s: 2*s tokens already emitted)
    o = llm_ref.LlmOracle(sd, dims)
    emb = sd["speech_embedding.weight"]
    lp_ref, e_max = {}, 0.0
    for s_ in depths:
        lm_in = torch.cat([o.prompt_embeds(text[0], ptext[0], pspeech[0]), emb[torch.tensor(ref[: K * s_], dtype=torch.long)]], 0)[None]
        y, _ = m.llm.forward_one_step(lm_in, masks=torch.tril(torch.ones(1, lm_in.shape[1], lm_in.shape[1])).bool())
        last = y[:, -1:, :]
        lp = torch.stack([m.llm_decoder(m.mtp_block[j](last)[0][:, -1]).log_softmax(-1)[0] for j in range(K)])
        lp_ref[s_] = lp
        if n_same >= K * s_:
            e_max = max(e_max, (lp - logps[s_][:K]).abs().max().item())
    print(f"[llm:c2] teacher-forced head log-probs at steps {list(depths)}: reference-vs-oracle max-abs {e_max:.2e}")
    assert e_max < 5e-4
    torch.save(dict(dims="c2", seed=seed, n_text=n_text, n_ptext=n_ptext, P=P, K=K, ratio=ratio, sp=sp, text=text[0], prompt_text=ptext[0],
                    prompt_speech=pspeech[0], u=u, tokens=ref, oracle_tokens=ora, oracle_common_prefix=n_same, depths=list(depths),
                    head_logp={int(k): v for k, v in lp_ref.items()}, u_used=us.pos, sd_checksum=checksum(sd)),
               os.path.join(OUT, "llm_c2.pt"))


def golden_e2e_c1(seed=0, n_text=32, K=1, ratio=8, n_steps=10):
    """BASELINE configs[0] ("C1", SURVEY 8d: one 32-text-token utterance, inference_head_num=1, 10 CFM Euler steps, no prompt ->
    256 speech tokens, 512 mel frames, 245 760 samples = 10.24 s) through the three UNMODIFIED reference modules at full dims,
    chained the way `inference_tts` chains them (infer_speech_model.py:629-668: empty prompt text, no prompt speech tokens, the
    flow called with token / embedding only, the vocoder on the flow's mel).  Each stage's restatement is checked on the way."""
    ld, fd, hd = D.LLM_FULL, D.FLOW_FULL, D.HIFT_FULL
    refshim.install()
    from cosyvoice.utils.common import ras_sampling
    import cosyvoice.flow.flow as flowmod
    import time
    sp = dict(top_p=0.9, top_k=10, win_size=24, tau_r=0.2)         # server tts defaults (router.py:22-44)
    u0 = synth.utterance(ld, fd, n_text, seed=1986, zero_shot=False)
    text = u0["text"].long()[None]
    u = torch.rand(4096, generator=torch.Generator().manual_seed(seed + 7))
    # ---- stage 1: CosyVoice3LM.inference (no KV cache), weights = the bf16-rounded synthetic checkpoint evaluated in fp32
    m = refshim.build_llm(ld)
    sd_l = {k: v.to(torch.bfloat16).float() for k, v in synth.llm_state_dict(ld, seed, eos_scale=0.0).items()}
    m.load_state_dict(sd_l, strict=True)
    m.sampling = partial(ras_sampling, **sp)
    m.inference_head_num = K
    us = llm_ref.UStream(u)
    orig = torch.Tensor.multinomial
    torch.Tensor.multinomial = lambda self, n, replacement=False: torch.tensor([llm_ref.multinomial_u(self, us.next())])
    t0 = time.time()
    try:
        toks = list(m.inference(text=text, text_len=torch.tensor([n_text]), prompt_text=torch.zeros(1, 0, dtype=torch.long),
                                prompt_text_len=torch.tensor([0]), prompt_speech_token=None, prompt_speech_token_len=torch.tensor([0]),
                                embedding=None, min_token_text_ratio=ratio, max_token_text_ratio=ratio))
    finally:
        torch.Tensor.multinomial = orig
    ora = llm_ref.inference(sd_l, ld, text[0], torch.zeros(0, dtype=torch.long), torch.zeros(0, dtype=torch.long), u, head_k=K, sp=sp,
                            min_ratio=ratio, max_ratio=ratio)
    print(f"[e2e:c1] llm: reference {len(toks)} tokens in {time.time() - t0:.0f} s (u used {us.pos}); oracle identical: {ora == toks}")
    assert len(toks) == n_text * ratio and ora == toks
    del m
    # ---- stage 2: CausalMaskedDiffWithDiT.inference in fp32 (n_timesteps=10 is the reference's own hard-coded value, flow.py:425)
    assert n_steps == 10
    flowmod.torch = _TorchF32Proxy()
    f = refshim.build_flow(fd)
    sd_f = synth.flow_state_dict(fd, seed)
    f.load_state_dict(sd_f, strict=True)
    f.bf16 = True
    noise = synth.flow_noise(fd)
    assert torch.equal(f.decoder.rand_noise[:, :, : noise.shape[2]], noise), "rand_noise recipe drifted"
    tok = torch.tensor(toks)[None]
    emb = u0["embedding"][None]
    mel, _ = f.inference(token=tok, token_len=torch.tensor([tok.shape[1]]), embedding=emb, streaming=False, finalize=True)
    flowmod.torch = torch
    mel_o = flow_ref.inference(sd_f, tok, emb, noise, fd, n_steps)
    e_mel = (mel - mel_o).abs().max().item()
    print(f"[e2e:c1] flow: mel {tuple(mel.shape)} ref-vs-oracle max-abs {e_mel:.2e} mean|mel| {mel.abs().mean():.3f}")
    assert mel.shape == (1, fd.mel, 2 * len(toks)) and e_mel < 2e-4
    del f
    # ---- stage 3: CausalHiFTGenerator.inference on the reference's mel (F0 predictor on the CPU in fp32, as the reference runs it)
    h = refshim.build_hift(hd)
    sd_h = synth.hift_state_dict(hd, seed)
    h.load_state_dict(sd_h, strict=True)
    table = synth.hift_sine_table(hd, mel.shape[2])
    h.m_source.l_sin_gen.sine_waves = table[None]
    wav, _ = h.inference(speech_feat=mel)
    f0 = h.f0_predictor(mel)
    wav_o, _ = hift_ref.inference(sd_h, mel, table, hd, f0=f0)
    e_wav = (wav - wav_o).abs().max().item()
    wav_free, _ = hift_ref.inference(sd_h, mel, table, hd)
    rms_free = (wav - wav_free).pow(2).mean().sqrt().item()
    print(f"[e2e:c1] hift: {wav.shape[1]} samples, ref-vs-oracle max-abs {e_wav:.2e} (F0 pinned), rms {rms_free:.2e} (oracle's own F0); "
          f"wav rms {wav.pow(2).mean().sqrt():.3f}")
    assert wav.shape == (1, mel.shape[2] * hd.frame_samples) and e_wav < 2e-5
    torch.save(dict(dims="c1", seed=seed, n_text=n_text, K=K, ratio=ratio, n_steps=n_steps, sp=sp, text=u0["text"], embedding=u0["embedding"],
                    u=u, u_used=us.pos, tokens=toks, mel=mel, wav=wav, f0=f0, oracle_free_f0_rms=rms_free,
                    sd_checksum=dict(llm=checksum(sd_l), flow=checksum(sd_f), hift=checksum(sd_h))),
               os.path.join(OUT, "e2e_c1.pt"))


def golden_hift_t(name, dims, T, seed):
    """a12': the non-causal ConvTranspose1d HiFTGenerator.  Its source module is stochastic (fresh noise + random initial
    phase per call): the fixture pins decode(mel, s) for an explicit source s, the F0 predictor, and the whole
    `inference` with the global RNG seeded (the oracle replays the same draws through hift_ref.draw_source_noise)."""
    m = refshim.build_hift_t(dims)
    sd = synth.hift_t_state_dict(dims, seed)
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed + 200)
    mel = torch.rand(1, dims.mel, T, generator=g) * 6.0 - 6.0
    s = torch.tanh(torch.randn(1, 1, T * dims.frame_samples, generator=g) * 0.3)
    wav = m.decode(x=mel, s=s)
    f0 = m.f0_predictor(mel)
    wav_o = hift_ref.decode_transposed(sd, mel, s, dims)
    f0_o = hift_ref.f0_predict_nc(hift_ref.fold_weight_norm(sd), mel)
    e, ef = (wav - wav_o).abs().max().item(), ((f0 - f0_o).abs() / (f0.abs() + 1)).max().item()
    print(f"[hift_t:{name}] T={T} ref-vs-oracle wav max-abs {e:.2e} f0 rel {ef:.2e}; rms {wav.pow(2).mean().sqrt():.3f} sat {(wav.abs() >= 0.99).float().mean():.3f}")
    assert e < 5e-6 and ef < 1e-4
    torch.manual_seed(seed + 300)
    wav_i, s_i = m.inference(speech_feat=mel)
    torch.manual_seed(seed + 300)
    noise = hift_ref.draw_source_noise(T * dims.frame_samples, dims.harmonics)
    wav_io, s_io = hift_ref.inference_transposed(sd, mel, noise, dims, f0=f0)
    es, ei = (s_i - s_io).abs().max().item(), (wav_i - wav_io).abs().max().item()
    cache = s[:, :, : 3 * dims.frame_samples]
    torch.manual_seed(seed + 300)
    wav_c, s_c = m.inference(speech_feat=mel, cache_source=cache)
    wav_co, _ = hift_ref.inference_transposed(sd, mel, noise, dims, f0=f0, cache_source=cache)
    ec = (wav_c - wav_co).abs().max().item()
    print(f"[hift_t:{name}] inference (RNG pinned): source max-abs {es:.2e} wav max-abs {ei:.2e} (cache_source {ec:.2e}); voiced {(f0 > 10).float().mean():.2f}")
    assert es < 2e-4 and ei < 2e-3 and ec < 2e-3
    torch.save(dict(dims=name, seed=seed, T=T, mel=mel, s=s, wav=wav, f0=f0, noise=noise, wav_inf=wav_i, src_inf=s_i,
                    wav_cache=wav_c, sd_checksum=checksum(sd)), os.path.join(OUT, f"hift_t_{name}.pt"))


def golden_hifigan(name, dims, T, seed):
    """a12' (second variant): the classic HiFi-GAN Generator vendored under matcha/hifigan."""
    m = refshim.build_hifigan(dims)
    sd = synth.hifigan_state_dict(dims, seed)
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed + 400)
    mel = torch.rand(2, dims.mel, T, generator=g) * 6.0 - 6.0
    with torch.no_grad():
        wav = m(mel)
    wav_o = hifigan_ref.generator(sd, mel, dims)
    e = (wav - wav_o).abs().max().item()
    print(f"[hifigan:{name}] T={T} ref-vs-oracle wav max-abs {e:.2e}; out {tuple(wav.shape)} rms {wav.pow(2).mean().sqrt():.3f} |max| {wav.abs().max():.3f}")
    assert e < 5e-6 and wav.shape[-1] == T * dims.frame_samples
    torch.save(dict(dims=name, seed=seed, T=T, mel=mel, wav=wav, sd_checksum=checksum(sd)), os.path.join(OUT, f"hifigan_{name}.pt"))


def golden_unet(name, dims, T, seed):
    """a7': the U-Net estimator CausalConditionalDecoder at the forward_estimator seam (CFG batch of 2, all-true mask),
    offline and with the streaming chunk mask."""
    m = refshim.build_unet(dims)
    sd = synth.unet_state_dict(dims, seed)
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed + 500)
    x = torch.randn(2, dims.mel, T, generator=g)
    mu = torch.randn(2, dims.mel, T, generator=g)
    cond = torch.randn(2, dims.mel, T, generator=g) * 2.0 - 3.0
    spks = torch.randn(2, dims.mel, generator=g)
    mu[1], cond[1], spks[1] = 0, 0, 0                                  # the CFG row (flow_matching.py:100-112)
    x[1] = x[0]
    t = torch.tensor([0.3141, 0.3141])
    mask = torch.ones(2, 1, T)
    out = {}
    for key, streaming in (("full", False), ("stream", True)):
        y = m(x, mask, mu, t, spks, cond, streaming=streaming)
        y_o = unet_ref.estimator(sd, x, mask, mu, t, spks, cond, dims, streaming=streaming)
        e = (y - y_o).abs().max().item()
        print(f"[unet:{name}] T={T} {key}: ref-vs-oracle max-abs {e:.2e}; |out| mean {y.abs().mean():.3f} max {y.abs().max():.2f}")
        assert e < 2e-4 * max(1.0, y.abs().max().item())
        out["y_" + key] = y
    # ragged batch: the second row is padded (mask 0 from frame T-5 on) -- the reference's padding-mask path
    mask2 = mask.clone(); mask2[1, :, T - 5:] = 0
    y = m(x, mask2, mu, t, spks, cond, streaming=False)
    y_o = unet_ref.estimator(sd, x, mask2, mu, t, spks, cond, dims)
    assert (y - y_o).abs().max().item() < 2e-4 * max(1.0, y.abs().max().item())
    print(f"[unet:{name}] stream-vs-full max-abs {(out['y_full'] - out['y_stream']).abs().max():.2e}")
    torch.save(dict(dims=name, seed=seed, T=T, x=x, mu=mu, cond=cond, spks=spks, t=t, sd_checksum=checksum(sd), **out),
               os.path.join(OUT, f"unet_{name}.pt"))


def golden_unet_nc(name, dims, T, seed):
    """a7' (first variant named in SURVEY 8): the non-causal multi-level ConditionalDecoder at the forward_estimator seam, CFG
    batch of 2: all-true mask (odd and even T: the stride-2 level and the ConvTranspose1d length handling), and a padded
    batch (row 1 valid up to T - 7) through the reference's padding-mask path."""
    m = refshim.build_unet_nc(dims)
    sd = synth.unet_nc_state_dict(dims, seed)
    m.load_state_dict(sd, strict=True)
    out = {}
    for key, Tk, pad in (("full", T, 0), ("odd", T - 1, 0), ("masked", T, 7)):
        g = torch.Generator().manual_seed(seed + 800 + Tk + pad)
        x = torch.randn(2, dims.mel, Tk, generator=g)
        mu = torch.randn(2, dims.mel, Tk, generator=g)
        cond = torch.randn(2, dims.mel, Tk, generator=g) * 2.0 - 3.0
        spks = torch.randn(2, dims.mel, generator=g)
        t = torch.tensor([0.3141, 0.3141])
        mask = torch.ones(2, 1, Tk)
        if pad:
            mask[1, :, Tk - pad:] = 0
        else:
            mu[1], cond[1], spks[1] = 0, 0, 0
            x[1] = x[0]
        y = m(x, mask, mu, t, spks, cond)
        y_o = unet_ref.estimator_nc(sd, x, mask, mu, t, spks, cond, dims)
        e = (y - y_o).abs().max().item()
        print(f"[unet_nc:{name}] T={Tk} {key}: ref-vs-oracle max-abs {e:.2e}; |out| mean {y.abs().mean():.3f} max {y.abs().max():.2f}")
        assert e < 2e-4 * max(1.0, y.abs().max().item())
        out[key] = dict(x=x, mu=mu, cond=cond, spks=spks, t=t, mask=mask, y=y)
    torch.save(dict(dims=name, seed=seed, T=T, sd_checksum=checksum(sd), **out), os.path.join(OUT, f"unet_nc_{name}.pt"))


def golden_unet_nc_cfm(name, dims, T, n_steps, seed):
    """ConditionalCFM.forward (flow_matching.py:36-69) over the non-causal estimator: two chained calls, the second one fed the
    first one's cache (prompt + 34-frame overlap of z and mu).  z is the reference's own torch.randn_like draw under a pinned seed."""
    cfm = refshim.build_unet_nc_cfm(dims)
    sd = synth.unet_nc_state_dict(dims, seed)
    cfm.estimator.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed + 900)
    P = T // 3
    out, cache, cache_o = {}, torch.zeros(1, dims.mel, 0, 2), None
    for call in range(2):
        mu = torch.randn(1, dims.mel, T, generator=g)
        cond = torch.zeros(1, dims.mel, T)
        cond[:, :, :P] = torch.rand(1, dims.mel, P, generator=g) * 6.0 - 6.0
        spks = torch.randn(1, dims.mel, generator=g)
        torch.manual_seed(seed + 910 + call)
        z = torch.randn_like(mu)
        torch.manual_seed(seed + 910 + call)
        cache_in = cache.clone()
        mel, cache = cfm(mu.clone(), torch.ones(1, 1, T), n_steps, temperature=1.0, spks=spks, cond=cond, prompt_len=P, cache=cache)
        mel_o, cache_o = unet_ref.cfm_forward_nc(sd, mu, z, spks, cond, n_steps, dims, prompt_len=P, cache=cache_in)
        e = (mel - mel_o).abs().max().item()
        print(f"[unet_nc_cfm:{name}] call {call} T={T} steps={n_steps}: ref-vs-oracle mel max-abs {e:.2e}, cache equal {torch.equal(cache, cache_o)}")
        assert e < 5e-4 and torch.equal(cache, cache_o)
        out[f"call{call}"] = dict(mu=mu, cond=cond, spks=spks, z=z, cache_in=cache_in, mel=mel, cache=cache)
    torch.save(dict(dims=name, seed=seed, T=T, n_steps=n_steps, prompt_len=P, sd_checksum=checksum(sd), **out),
               os.path.join(OUT, f"unet_nc_cfm_{name}.pt"))


def golden_unet_cfm(name, dims, T, n_steps, seed):
    """the whole Euler solve over the U-Net estimator: CausalConditionalCFM.forward (flow_matching.py:203-228)"""
    cfm = refshim.build_unet_cfm(dims)
    sd = synth.unet_state_dict(dims, seed)
    cfm.estimator.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed + 600)
    mu = torch.randn(1, dims.mel, T, generator=g)
    cond = torch.zeros(1, dims.mel, T)
    cond[:, :, : T // 3] = torch.rand(1, dims.mel, T // 3, generator=g) * 6.0 - 6.0      # prompt mel in the head, zeros after (flow.py:415-419)
    spks = torch.randn(1, dims.mel, generator=g)
    out = {}
    for key, streaming in (("full", False), ("stream", True)):
        y, _ = cfm(mu, torch.ones(1, 1, T), n_steps, spks=spks, cond=cond, streaming=streaming)
        y_o = unet_ref.cfm_solve(sd, mu, spks, cond, cfm.rand_noise, n_steps, dims, streaming=streaming)
        e = (y - y_o).abs().max().item()
        print(f"[unet_cfm:{name}] T={T} steps={n_steps} {key}: ref-vs-oracle max-abs {e:.2e}; |mel| mean {y.abs().mean():.3f} max {y.abs().max():.2f}")
        assert e < 5e-4 * max(1.0, y.abs().max().item())
        out["mel_" + key] = y.clone()
    assert torch.equal(cfm.rand_noise, __import__("flowmirror_hydravox_b200.flow", fromlist=["rand_noise"]).rand_noise(dims.mel, 15000))
    torch.save(dict(dims=name, seed=seed, T=T, n_steps=n_steps, mu=mu, cond=cond, spks=spks, sd_checksum=checksum(sd), **out),
               os.path.join(OUT, f"unet_cfm_{name}.pt"))


def golden_frontend(seed):
    """f1: the reference's own mel_spectrogram (matcha/utils/audio.py:42-82) on a synthetic 24 kHz prompt, with the yaml's
    feat_extractor arguments; torchaudio's kaldi.fbank (the call of cosyvoice/cli/frontend.py:108-112) on a 16 kHz one."""
    refshim.install()
    from matcha.utils.audio import mel_spectrogram
    import torchaudio.compliance.kaldi as kaldi
    from flowmirror_hydravox_b200.frontend import slaney_mel_basis
    g = torch.Generator().manual_seed(seed + 700)
    t = torch.arange(24000 * 2 + 331) / 24000.0
    y = (0.4 * torch.sin(2 * torch.pi * 220.0 * t) + 0.2 * torch.sin(2 * torch.pi * 1760.0 * t * (1 + 0.1 * t))
         + 0.05 * torch.randn(t.shape, generator=g))[None].clamp(-1, 1)
    mel = mel_spectrogram(y, n_fft=1920, num_mels=80, sampling_rate=24000, hop_size=480, win_size=1920, fmin=0, fmax=8000, center=False)
    basis = slaney_mel_basis(24000, 1920, 80, 0, 8000)
    mel_o = frontend_ref.mel_spectrogram(y, basis)
    e = (mel - mel_o).abs().max().item()
    t16 = torch.arange(16000 * 3 + 77) / 16000.0
    s16 = (0.3 * torch.sin(2 * torch.pi * 150.0 * t16) + 0.1 * torch.randn(t16.shape, generator=g))[None]
    fb = kaldi.fbank(s16, num_mel_bins=80, dither=0, sample_frequency=16000)
    fb = fb - fb.mean(dim=0, keepdim=True)
    fb_o = frontend_ref.kaldi_fbank(s16)
    e2 = (fb - fb_o).abs().max().item()
    print(f"[frontend] mel {tuple(mel.shape)} ref-vs-oracle max-abs {e:.2e} (range {mel.min():.2f}..{mel.max():.2f}); "
          f"kaldi fbank {tuple(fb.shape)} torchaudio-vs-oracle max-abs {e2:.2e}")
    assert e < 1e-4 and e2 < 1e-3 and mel.shape == (1, 80, y.shape[1] // 480)
    torch.save(dict(seed=seed, y24=y, mel=mel, s16=s16, fbank=fb), os.path.join(OUT, "frontend.pt"))


def golden_flow(name, dims, N, P, n_steps, seed, modes=("full", "stream", "chunk"), with_est=True, with_bf16=True):
    """modes/with_est/with_bf16 trim the work for the BASELINE config-2 size fixture (`c2`: 1024 + 125 tokens -> 2298 frames,
    25 Euler steps: one fp32 solve of the reference is several CPU-minutes)."""
    refshim.install()
    import cosyvoice.flow.flow as flowmod
    flowmod.torch = _TorchF32Proxy()
    m = refshim.build_flow(dims)
    sd = synth.flow_state_dict(dims, seed)
    m.load_state_dict(sd, strict=True)
    m.bf16 = True
    noise = synth.flow_noise(dims)
    assert torch.equal(m.decoder.rand_noise[:, :, : noise.shape[2]], noise), "rand_noise recipe drifted"
    g = torch.Generator().manual_seed(seed + 100)
    tok = torch.randint(0, dims.vocab, (1, N), generator=g)
    ptok = torch.randint(0, dims.vocab, (1, P), generator=g)
    pfeat = torch.rand(1, 2 * P, dims.mel, generator=g) * -6.0
    emb = torch.rand(1, dims.spk_in, generator=g)
    orig = m.decoder.forward
    m.decoder.forward = lambda **kw: orig(**{**kw, "n_timesteps": n_steps})
    kw = dict(token=tok, token_len=torch.tensor([N]), embedding=emb, prompt_token=ptok,
              prompt_token_len=torch.tensor([P]), prompt_feat=pfeat, prompt_feat_len=torch.tensor([2 * P]))
    out = {}
    for key, streaming, finalize in (("full", False, True), ("stream", True, True), ("chunk", True, False)):
        if key not in modes:
            continue
        ref, _ = m.inference(finalize=finalize, streaming=streaming, **kw)
        ora = flow_ref.inference(sd, tok, emb, noise, dims, n_steps, ptok, pfeat, streaming=streaming, finalize=finalize)
        e = (ref - ora).abs().max().item()
        print(f"[flow:{name}] {key} mel {tuple(ref.shape)} ref-vs-oracle max-abs {e:.2e} mean|mel| {ref.abs().mean():.3f}")
        assert e < 2e-4
        out["mel_" + key] = ref
    extra = {}
    if with_est:
        # one estimator call (the TRT seam, flow_matching.py:126-153) for kernel-level parity
        T = 2 * (N + P)
        xg = torch.randn(2, dims.mel, T, generator=g)
        mug = torch.randn(2, dims.mel, T, generator=g)
        cg = torch.randn(2, dims.mel, T, generator=g)
        sg = torch.randn(2, dims.mel, generator=g)
        tg = torch.tensor([0.3, 0.3])
        est = m.decoder.estimator(xg, torch.ones(2, 1, T), mug, tg, sg, cg, streaming=False)
        est_o = flow_ref.dit_forward(sd, xg, mug, tg, sg, cg, dims)
        assert (est - est_o).abs().max().item() < 1e-4
        extra.update(est_in=dict(x=xg, mu=mug, cond=cg, spks=sg, t=tg), est_out=est)
    flowmod.torch = torch
    if with_bf16:
        # what the reference's own low-precision path loses against fp32 (parity budget)
        mb = refshim.build_flow(dims, dtype=torch.bfloat16)
        mb.load_state_dict({k: v.to(torch.bfloat16) for k, v in sd.items()}, strict=True)
        mb.bf16 = True
        origb = mb.decoder.forward
        mb.decoder.forward = lambda **k2: origb(**{**k2, "n_timesteps": n_steps})
        refb, _ = mb.inference(finalize=True, streaming=False, **kw)
        dev = (refb - out["mel_full"]).abs()
        print(f"[flow:{name}] reference bf16 path vs fp32: max-abs {dev.max():.3e} mean-abs {dev.mean():.3e}")
        extra.update(ref_bf16_maxabs=float(dev.max()), ref_bf16_meanabs=float(dev.mean()))
    torch.save(dict(dims=name, seed=seed, N=N, P=P, n_steps=n_steps, token=tok, prompt_token=ptok, prompt_feat=pfeat,
                    embedding=emb, sd_checksum=checksum(sd), **extra, **out), os.path.join(OUT, f"flow_{name}.pt"))


def golden_llm(name, dims, n_text, n_ptext, P, cases, seed):
    refshim.install()
    from cosyvoice.utils.common import ras_sampling
    m = refshim.build_llm(dims)
    sd = synth.llm_state_dict(dims, seed)
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed + 100)
    text = torch.randint(0, dims.text_vocab, (1, n_text), generator=g)
    ptext = torch.randint(0, dims.text_vocab, (1, n_ptext), generator=g)
    pspeech = torch.randint(0, dims.speech_token_size, (1, P), generator=g)
    u = torch.rand(8192, generator=g)
    rec = []
    for K, ratio, sp in cases:
        m.sampling = partial(ras_sampling, **sp)
        m.inference_head_num = K
        us = llm_ref.UStream(u)
        orig = torch.Tensor.multinomial
        torch.Tensor.multinomial = lambda self, n, replacement=False: torch.tensor(
            [llm_ref.multinomial_u(self, us.next())])
        try:
            ref = list(m.inference(text=text, text_len=torch.tensor([n_text]), prompt_text=ptext,
                                   prompt_text_len=torch.tensor([n_ptext]), prompt_speech_token=pspeech,
                                   prompt_speech_token_len=torch.tensor([P]), embedding=None,
                                   min_token_text_ratio=ratio[0], max_token_text_ratio=ratio[1]))
        finally:
            torch.Tensor.multinomial = orig
        ora, logps = llm_ref.inference(sd, dims, text[0], ptext[0], pspeech[0], u, head_k=K, sp=sp,
                                       min_ratio=ratio[0], max_ratio=ratio[1], return_logp=True)
        print(f"[llm:{name}] K={K} ratio={ratio} sp={sp} ref {len(ref)} tok, oracle equal: {ref == ora}, u used {us.pos}")
        assert ref == ora
        rec.append(dict(K=K, ratio=ratio, sp=sp, tokens=ref, u_used=us.pos, logp0=logps[0]))
    # first-step head log-probs straight from the reference modules (teacher-forced logits parity)
    lm_in = llm_ref.LlmOracle(sd, dims).prompt_embeds(text[0], ptext[0], pspeech[0])[None]
    y, _ = m.llm.forward_one_step(lm_in, masks=torch.tril(torch.ones(1, lm_in.shape[1], lm_in.shape[1])).bool())
    last = y[:, -1:, :]
    ref_lp = torch.stack([m.llm_decoder(m.mtp_block[j](last)[0][:, -1]).log_softmax(-1)[0] for j in range(dims.mtp_heads)])
    o = llm_ref.LlmOracle(sd, dims)
    hid = o.forward_rows(lm_in[0])
    ora_lp = torch.stack([o.head_logp(j, hid[-1]) for j in range(dims.mtp_heads)])
    e = (ref_lp - ora_lp).abs().max().item()
    print(f"[llm:{name}] head log-prob ref-vs-oracle max-abs {e:.2e}")
    assert e < 2e-4
    torch.save(dict(dims=name, seed=seed, text=text[0], prompt_text=ptext[0], prompt_speech=pspeech[0], u=u,
                    cases=rec, head_logp=ref_lp, last_hidden=y[0, -1], sd_checksum=checksum(sd)),
               os.path.join(OUT, f"llm_{name}.pt"))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    if sys.argv[1:] == ["mid"]:                                     # the 400-frame full-dim flow fixture (tests/golden/flow_mid.pt)
        with torch.no_grad():
            golden_flow("mid", D.FLOW_FULL, 150, 50, 10, 1)
        return
    if sys.argv[1:2] == ["c2"]:                                     # BASELINE config-2 size fixtures (tens of CPU-minutes)
        which = sys.argv[2:] or ["hift", "flow", "llm"]
        with torch.no_grad():
            if "hift" in which:
                golden_hift_c2(D.HIFT_FULL, 2048, 0)
            if "flow" in which:
                golden_flow("c2", D.FLOW_FULL, 1024, 125, 25, 0, modes=("full",), with_est=False, with_bf16=False)
            if "llm" in which:
                golden_llm_c2(D.LLM_FULL, 0)
        return
    if sys.argv[1:] == ["c1"]:                                      # BASELINE configs[0]: the three reference modules chained
        with torch.no_grad():
            golden_e2e_c1()
        return
    if sys.argv[1:] == ["frontend"]:
        with torch.no_grad():
            golden_frontend(0)
        return
    if sys.argv[1:] == ["unet_nc"]:
        with torch.no_grad():
            golden_unet_nc("tiny", D.UNET_NC_TINY, 38, 0)
            golden_unet_nc("full", D.UNET_NC_FULL, 130, 0)
            golden_unet_nc_cfm("small", D.UNET_NC_SMALL, 75, 6, 0)
        return
    if sys.argv[1:] == ["unet"]:
        with torch.no_grad():
            golden_unet("tiny", D.UNET_TINY, 37, 0)
            golden_unet("full", D.UNET_FULL, 130, 0)
            golden_unet_cfm("small", D.UNET_SMALL, 57, 10, 0)
            golden_unet_cfm("full", D.UNET_FULL, 96, 5, 0)
        return
    if sys.argv[1:] == ["hifigan"]:
        with torch.no_grad():
            golden_hifigan("tiny", D.HIFIGAN_TINY, 37, 0)
            golden_hifigan("v1", D.HIFIGAN_V1, 24, 0)
        return
    if sys.argv[1:] == ["hift_t"]:                                  # regenerate only the a12' fixtures
        with torch.no_grad():
            golden_hift_t("tiny", D.HIFT_TINY, 29, 0)
            golden_hift_t("full", D.HIFT_FULL, 20, 0)
        return
    with torch.no_grad():
        golden_hift("tiny", D.HIFT_TINY, 37, 0)
        golden_hift("full", D.HIFT_FULL, 24, 0)
        golden_hift_t("tiny", D.HIFT_TINY, 29, 0)
        golden_hift_t("full", D.HIFT_FULL, 20, 0)
        golden_hifigan("tiny", D.HIFIGAN_TINY, 37, 0)
        golden_hifigan("v1", D.HIFIGAN_V1, 24, 0)
        golden_unet("tiny", D.UNET_TINY, 37, 0)
        golden_unet("full", D.UNET_FULL, 130, 0)
        golden_unet_cfm("small", D.UNET_SMALL, 57, 10, 0)
        golden_unet_cfm("full", D.UNET_FULL, 96, 5, 0)
        golden_unet_nc("tiny", D.UNET_NC_TINY, 38, 0)
        golden_unet_nc("full", D.UNET_NC_FULL, 130, 0)
        golden_unet_nc_cfm("small", D.UNET_NC_SMALL, 75, 6, 0)
        golden_frontend(0)
        golden_flow("tiny", D.FLOW_TINY, 21, 10, 10, 0)
        golden_flow("full", D.FLOW_FULL, 24, 8, 4, 0)
        golden_flow("mid", D.FLOW_FULL, 150, 50, 10, 1)
        sp1 = dict(top_p=0.9, top_k=10, win_size=24, tau_r=0.2)     # server tts defaults (router.py:22-44)
        sp2 = dict(top_p=0.8, top_k=25, win_size=10, tau_r=0.1)     # ras_sampling defaults (common.py:138)
        sp3 = dict(top_p=0.9, top_k=10, win_size=0, tau_r=0.2)      # win_size=0 edge: always random sampling
        golden_llm("tiny", D.LLM_TINY, 12, 4, 6,
                   [(1, (2, 20), sp1), (3, (8, 8), sp1), (2, (2, 6), sp2), (5, (2, 3), sp3)], 0)
        golden_llm("full", D.LLM_FULL, 6, 2, 3, [(2, (4, 4), sp1)], 0)


if __name__ == "__main__":
    main()
