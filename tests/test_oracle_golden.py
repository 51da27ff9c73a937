"""Oracle (oracle/*.py, CPU fp32) against the fixtures minted from the UNMODIFIED reference
modules by oracle/make_golden.py.  These run without a GPU and without /root/reference."""
import pytest
import torch

from flowmirror_hydravox_b200 import dims as D, synth
from oracle import flow_ref, hift_ref, llm_ref


def _checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))


@pytest.mark.parametrize("name,dims", [("tiny", D.HIFT_TINY), ("full", D.HIFT_FULL)])
def test_hift_oracle_matches_reference(golden, name, dims):
    g = golden(f"hift_{name}")
    sd = synth.hift_state_dict(dims, g["seed"])
    assert abs(_checksum(sd) - g["sd_checksum"]) < 1e-6 * g["sd_checksum"]
    table = synth.hift_sine_table(dims, g["T"])
    w = hift_ref.fold_weight_norm(sd)
    f0 = hift_ref.f0_predict(w, g["mel"])
    assert ((f0 - g["f0"]).abs() / (g["f0"].abs() + 1)).max() < 1e-4
    wav, src = hift_ref.inference(sd, g["mel"], table, dims, f0=g["f0"])
    assert wav.shape == g["wav"].shape == (1, g["T"] * dims.frame_samples)
    assert (wav - g["wav"]).abs().max() < 5e-6            # fp32, F0 pinned (see hift_ref.inference)
    assert (src - g["src"]).abs().max() < 5e-6
    wav_s, _ = hift_ref.inference(sd, g["mel"], table, dims, finalize=False, f0=g["f0_stream"])
    assert (wav_s - g["wav_stream"]).abs().max() < 5e-6
    # causal-vocoder self-check of the reference (generator.py:729-746): streamed prefix == offline prefix
    n = wav_s.shape[1]
    assert (wav_s - g["wav"][:, :n]).abs().max() < 2e-3   # F0 differs in the look-ahead frames only


@pytest.mark.parametrize("name,dims", [("tiny", D.HIFT_TINY), ("full", D.HIFT_FULL)])
def test_hift_transposed_oracle_matches_reference(golden, name, dims):
    """a12': the non-causal ConvTranspose1d HiFTGenerator (generator.py:378-569), fixtures from the unmodified module."""
    g = golden(f"hift_t_{name}")
    sd = synth.hift_t_state_dict(dims, g["seed"])
    assert abs(_checksum(sd) - g["sd_checksum"]) < 1e-6 * g["sd_checksum"]
    w = hift_ref.fold_weight_norm(sd)
    f0 = hift_ref.f0_predict_nc(w, g["mel"])
    assert ((f0 - g["f0"]).abs() / (g["f0"].abs() + 1)).max() < 1e-4
    wav = hift_ref.decode_transposed(sd, g["mel"], g["s"], dims)
    assert wav.shape == g["wav"].shape == (1, g["T"] * dims.frame_samples)
    assert (wav - g["wav"]).abs().max() < 5e-6
    # whole inference with the RNG pinned: source bit-for-bit up to fp32 rounding, waveform through it
    wav_i, src = hift_ref.inference_transposed(sd, g["mel"], g["noise"], dims, f0=g["f0"])
    assert (src - g["src_inf"]).abs().max() < 1e-5
    assert (wav_i - g["wav_inf"]).abs().max() < 5e-6
    wav_c, _ = hift_ref.inference_transposed(sd, g["mel"], g["noise"], dims, f0=g["f0"], cache_source=g["s"][:, :, : 3 * dims.frame_samples])
    assert (wav_c - g["wav_cache"]).abs().max() < 5e-6


@pytest.mark.parametrize("name,dims", [("tiny", D.HIFIGAN_TINY), ("v1", D.HIFIGAN_V1)])
def test_hifigan_oracle_matches_reference(golden, name, dims):
    """a12' (second variant): classic HiFi-GAN Generator (matcha/hifigan/models.py:148-193), fixtures from the unmodified module."""
    from oracle import hifigan_ref
    g = golden(f"hifigan_{name}")
    sd = synth.hifigan_state_dict(dims, g["seed"])
    assert abs(_checksum(sd) - g["sd_checksum"]) < 1e-6 * g["sd_checksum"]
    wav = hifigan_ref.generator(sd, g["mel"], dims)
    assert wav.shape == g["wav"].shape == (2, 1, g["T"] * dims.frame_samples)
    assert (wav - g["wav"]).abs().max() < 5e-6
    # the packed (folded, tap-reversed) weights are the same linear maps: conv of the zero-stuffed signal == conv_transpose
    from flowmirror_hydravox_b200.weights import pack_hifigan
    pk = pack_hifigan(sd, dims)
    x = torch.randn(1, dims.base, 9, generator=torch.Generator().manual_seed(1))
    u, k = dims.ups[0], dims.up_k[0]
    ref = torch.nn.functional.conv_transpose1d(x, hifigan_ref.fold(sd, "ups.0"), sd["ups.0.bias"], stride=u, padding=(k - u) // 2)
    z = torch.zeros(1, dims.base, (9 - 1) * u + 1)
    z[:, :, ::u] = x
    pl = k - 1 - (k - u) // 2
    z = torch.nn.functional.pad(z, (pl, 9 * u + k - 1 - pl - z.shape[2]))
    mine = torch.nn.functional.conv1d(z, pk["ups.0.w"].permute(2, 0, 1), pk["ups.0.b"])
    assert mine.shape == ref.shape and (mine - ref).abs().max() < 1e-5


def test_hift_source_noise_replays_reference_rng(golden):
    """draw_source_noise consumes the global generator like SineGen2.forward did when the fixture was minted
    (rand(1,H), then randn_like of a (1,n,H) view of (1,H,n) memory)."""
    g = golden("hift_t_tiny")
    torch.manual_seed(g["seed"] + 300)
    n = hift_ref.draw_source_noise(g["T"] * D.HIFT_TINY.frame_samples, D.HIFT_TINY.harmonics)
    assert n.shape == g["noise"].shape and torch.equal(n, g["noise"])


@pytest.mark.parametrize("name,dims", [("tiny", D.UNET_TINY), ("full", D.UNET_FULL)])
def test_unet_oracle_matches_reference(golden, name, dims):
    """a7': the U-Net estimator CausalConditionalDecoder (cosyvoice/flow/decoder.py:405-494), fixtures from the reference module."""
    from oracle import unet_ref
    g = golden(f"unet_{name}")
    sd = synth.unet_state_dict(dims, g["seed"])
    assert abs(_checksum(sd) - g["sd_checksum"]) < 1e-6 * g["sd_checksum"]
    mask = torch.ones(2, 1, g["T"])
    for key, streaming in (("full", False), ("stream", True)):
        y = unet_ref.estimator(sd, g["x"], mask, g["mu"], g["t"], g["spks"], g["cond"], dims, streaming=streaming)
        assert y.shape == g["y_" + key].shape == (2, dims.mel, g["T"])
        assert (y - g["y_" + key]).abs().max() < 5e-5
    # packing keeps every parameter: conv B operands are the same linear maps
    from flowmirror_hydravox_b200.weights import pack_unet
    pk = pack_unet(sd, dims)
    w = sd["mid_blocks.0.0.block1.block.0.weight"]
    assert torch.equal(pk["res1.c1.w"].float().reshape(dims.ch, 3, dims.ch), w.permute(0, 2, 1).half().float())
    pk2 = pack_unet(sd, dims, precise=True)
    hi, lo = pk2["tfm0.qkv.w"].float().chunk(2, dim=1)
    full = torch.cat([sd[f"down_blocks.0.1.0.attn1.to_{n}.weight"] for n in "qkv"], 0)
    assert (hi + lo - full).abs().max() < 1e-6


@pytest.mark.parametrize("name,dims", [("tiny", D.UNET_NC_TINY), ("full", D.UNET_NC_FULL)])
def test_unet_nc_oracle_matches_reference(golden, name, dims):
    """a7': the non-causal multi-level ConditionalDecoder (cosyvoice/flow/decoder.py:88-291): GroupNorm blocks, the stride-2
    level, ConvTranspose1d(4,2,1), skip concatenation and the padding-mask path; fixtures from the reference module."""
    from oracle import unet_ref
    g = golden(f"unet_nc_{name}")
    sd = synth.unet_nc_state_dict(dims, g["seed"])
    assert abs(_checksum(sd) - g["sd_checksum"]) < 1e-6 * g["sd_checksum"]
    for key in ("full", "odd", "masked"):
        c = g[key]
        y = unet_ref.estimator_nc(sd, c["x"], c["mask"], c["mu"], c["t"], c["spks"], c["cond"], dims)
        assert y.shape == c["y"].shape
        assert (y - c["y"]).abs().max() < 2e-4 * max(1.0, c["y"].abs().max().item())
    assert (g["masked"]["y"][1, :, -7:] == 0).all()


def test_unet_cfm_oracle_matches_reference(golden):
    """CausalConditionalCFM.forward over the U-Net estimator (flow_matching.py:203-228): the oracle's Euler solve vs the fixture"""
    from oracle import unet_ref
    from flowmirror_hydravox_b200.flow import rand_noise
    dims = D.UNET_SMALL
    g = golden("unet_cfm_small")
    sd = synth.unet_state_dict(dims, g["seed"])
    assert abs(_checksum(sd) - g["sd_checksum"]) < 1e-6 * g["sd_checksum"]
    noise = rand_noise(dims.mel, 15000)
    for key, streaming in (("full", False), ("stream", True)):
        y = unet_ref.cfm_solve(sd, g["mu"], g["spks"], g["cond"], noise, g["n_steps"], dims, streaming=streaming)
        assert y.shape == g["mel_" + key].shape == (1, dims.mel, g["T"])
        assert (y - g["mel_" + key]).abs().max() < 5e-5


@pytest.mark.parametrize("name,dims", [("tiny", D.FLOW_TINY), ("full", D.FLOW_FULL)])
def test_flow_oracle_matches_reference(golden, name, dims):
    g = golden(f"flow_{name}")
    sd = synth.flow_state_dict(dims, g["seed"])
    assert abs(_checksum(sd) - g["sd_checksum"]) < 1e-6 * g["sd_checksum"]
    noise = synth.flow_noise(dims)
    assert abs(noise[0, 0, 0].item() + 1.1258) < 1e-4      # flow_matching.py:200-201 table
    for key, streaming, finalize in (("full", False, True), ("stream", True, True), ("chunk", True, False)):
        mel = flow_ref.inference(sd, g["token"], g["embedding"], noise, dims, g["n_steps"], g["prompt_token"],
                                 g["prompt_feat"], streaming=streaming, finalize=finalize)
        assert mel.shape == g["mel_" + key].shape
        assert (mel - g["mel_" + key]).abs().max() < 2e-4
    e = g["est_in"]
    est = flow_ref.dit_forward({k: v.float() for k, v in sd.items()}, e["x"], e["mu"], e["t"], e["spks"], e["cond"], dims)
    assert (est - g["est_out"]).abs().max() < 1e-4


def test_flow_t_schedule_and_mask():
    ts = flow_ref.t_schedule(10)
    assert ts[0] == 0 and abs(ts[-1].item() - 1.0) < 1e-6 and (ts[1:] > ts[:-1]).all()
    m = flow_ref.attn_mask(4, [4], True, 2)[0]
    assert m.int().tolist() == [[1, 1, 0, 0], [1, 1, 0, 0], [1, 1, 1, 1], [1, 1, 1, 1]]   # mask.py:141-146
    m = flow_ref.attn_mask(4, [3], False, 2)[0]
    assert m.int().tolist() == [[1, 1, 1, 0]] * 4


@pytest.mark.parametrize("name,dims", [("tiny", D.LLM_TINY), pytest.param("full", D.LLM_FULL, marks=pytest.mark.slow)])
def test_llm_oracle_matches_reference(golden, name, dims):
    g = golden(f"llm_{name}")
    sd = synth.llm_state_dict(dims, g["seed"])
    assert abs(_checksum(sd) - g["sd_checksum"]) < 1e-6 * g["sd_checksum"]
    for c in g["cases"]:
        us = llm_ref.UStream(g["u"])
        toks = llm_ref.inference(sd, dims, g["text"], g["prompt_text"], g["prompt_speech"], us, head_k=c["K"],
                                 sp=c["sp"], min_ratio=c["ratio"][0], max_ratio=c["ratio"][1])
        assert toks == c["tokens"]                          # bit-exact token ids on the pinned u-stream
        assert us.pos == c["u_used"]
    o = llm_ref.LlmOracle(sd, dims)
    hid = o.forward_rows(o.prompt_embeds(g["text"], g["prompt_text"], g["prompt_speech"]))
    assert (hid[-1] - g["last_hidden"]).abs().max() < 1e-4
    lp = torch.stack([o.head_logp(j, hid[-1]) for j in range(dims.mtp_heads)])
    assert (lp - g["head_logp"]).abs().max() < 2e-4


def test_sampler_edges():
    # kept set = longest prefix with cum_before < top_p and count < top_k (common.py:146-156)
    logp = torch.log(torch.tensor([0.5, 0.3, 0.1, 0.06, 0.04]))
    us = llm_ref.UStream(torch.tensor([0.0, 0.7]))
    assert llm_ref.nucleus_sampling(logp, us, top_p=0.8, top_k=25) == 0
    assert llm_ref.nucleus_sampling(logp, us, top_p=0.8, top_k=25) == 1     # kept {0,1}: 0.7*0.8=0.56 > 0.5
    us = llm_ref.UStream(torch.tensor([0.99]))
    assert llm_ref.nucleus_sampling(logp, us, top_p=0.8, top_k=25) == 1     # crossing element (0.3) is kept, 0.1 is not
    us = llm_ref.UStream(torch.tensor([0.99]))
    assert llm_ref.nucleus_sampling(logp, us, top_p=0.99, top_k=1) == 0     # top_k bound
    # repetition-aware fallback draws a second u over the full distribution (common.py:139-143)
    us = llm_ref.UStream(torch.tensor([0.0, 0.999]))
    assert llm_ref.ras_sampling(logp, [0, 0, 0], us, top_p=0.8, top_k=25, win_size=3, tau_r=0.1) == 4
    assert us.pos == 2
    # stable sort: equal probabilities keep index order
    logp = torch.log(torch.tensor([0.25, 0.25, 0.25, 0.25]))
    us = llm_ref.UStream(torch.tensor([0.30]))
    assert llm_ref.nucleus_sampling(logp, us, top_p=0.6, top_k=25) == 0   # kept {0,1,2}: .3*.75=.225<.25


@pytest.mark.slow
def test_e2e_c1_oracle_chain_matches_reference_chain(golden):
    """BASELINE configs[0] at full dims: the CPU restatements chained (llm_ref -> flow_ref -> hift_ref) against the fixture the three
    unmodified reference modules produced when chained like inference_tts (oracle/make_golden.py c1): 256 token ids identical,
    mel within 2e-4, waveform within 2e-5 with the reference's F0 track."""
    import os
    if not os.path.exists(os.path.join(os.path.dirname(__file__), "golden", "e2e_c1.pt")):
        pytest.skip("e2e_c1 fixture not minted")
    g = golden("e2e_c1")
    ld, fd, hd = D.LLM_FULL, D.FLOW_FULL, D.HIFT_FULL
    sd_l = {k: v.to(torch.bfloat16).float() for k, v in synth.llm_state_dict(ld, g["seed"], eos_scale=0.0).items()}
    assert abs(_checksum(sd_l) - g["sd_checksum"]["llm"]) < 1e-6 * g["sd_checksum"]["llm"]
    us = llm_ref.UStream(g["u"])
    empty = torch.zeros(0, dtype=torch.long)
    toks = llm_ref.inference(sd_l, ld, g["text"], empty, empty, us, head_k=g["K"], sp=g["sp"], min_ratio=g["ratio"], max_ratio=g["ratio"])
    assert toks == g["tokens"] and us.pos == g["u_used"]
    del sd_l
    mel = flow_ref.inference(synth.flow_state_dict(fd, g["seed"]), torch.tensor(toks)[None], g["embedding"][None], synth.flow_noise(fd), fd,
                             g["n_steps"])
    assert mel.shape == g["mel"].shape and (mel - g["mel"]).abs().max() < 2e-4
    table = synth.hift_sine_table(hd, mel.shape[2])
    wav, _ = hift_ref.inference(synth.hift_state_dict(hd, g["seed"]), g["mel"], table, hd, f0=g["f0"])
    assert wav.shape == g["wav"].shape and (wav - g["wav"]).abs().max() < 2e-5
