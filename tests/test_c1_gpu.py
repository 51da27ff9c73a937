"""BASELINE configs[0] ("C1"): ONE 32-text-token utterance, inference_head_num=1, 10 CFM Euler steps, no prompt
(the `inference_tts` call shape, infer_speech_model.py:629-668) -> 256 speech tokens, 512 mel frames, 245 760 samples — the
whole chain at FULL model dims against the fixture minted by chaining the three UNMODIFIED reference modules on the CPU
(oracle/make_golden.py c1: CosyVoice3LM.inference -> CausalMaskedDiffWithDiT.inference -> CausalHiFTGenerator.inference).

north_star: token ids identical, <= 1e-3 max-abs on mel frames, <= 1e-4 RMS on waveform samples with the RNG pinned; the engine
runs in the mode bench.py's headline uses (fp32 KV cache, three-term split-fp16 flow GEMMs)."""
import os

import pytest
import torch

from flowmirror_hydravox_b200 import dims as D, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_e2e_c1_full_dims_matches_reference_chain(golden):
    if not os.path.exists(os.path.join(GOLD, "e2e_c1.pt")):
        pytest.skip("e2e_c1 fixture not minted")
    from flowmirror_hydravox_b200.model_manager import ModelManager
    g = golden("e2e_c1")
    ld, fd, hd = D.LLM_FULL, D.FLOW_FULL, D.HIFT_FULL
    ref_tok, K, ratio, steps = g["tokens"], g["K"], g["ratio"], g["n_steps"]
    assert (len(ref_tok), K, steps) == (256, 1, 10)
    T = 2 * len(ref_tok)
    mm = ModelManager(hd=hd, fd=fd, ld=ld, max_ctx=512, max_seqs=1, n_timesteps=steps, kv_f32=True, flow_precise=True)
    mm.load_state_dicts(synth.llm_state_dict(ld, g["seed"], dtype=torch.bfloat16, eos_scale=0.0), synth.flow_state_dict(fd, g["seed"]),
                        synth.hift_state_dict(hd, g["seed"]), sine_table=synth.hift_sine_table(hd, T))
    req = dict(text=g["text"], prompt_text=torch.zeros(0, dtype=torch.int32), prompt_speech=torch.zeros(0, dtype=torch.int32),
               prompt_feat=None, embedding=g["embedding"])
    # 1) the public one-call path on host buffers: tokens -> mel -> waveform, nothing pinned
    wavs, toks = mm.synthesize_batch([req], head_k=K, sampling=g["sp"], n_timesteps=steps, min_ratio=ratio, max_ratio=ratio,
                                     u=g["u"][None], return_tokens=True)
    n_same = next((i for i, (a, b) in enumerate(zip(toks[0], ref_tok)) if a != b), min(len(toks[0]), len(ref_tok)))
    print(f"[c1] engine {len(toks[0])} tokens, reference {len(ref_tok)}, identical prefix {n_same}")
    assert toks[0] == ref_tok
    assert wavs[0].shape == g["wav"].shape and torch.isfinite(wavs[0]).all()
    # 2) the flow on those tokens vs the reference module's mel
    tok = torch.tensor(ref_tok)[None]
    mel, _ = mm.models["flow"].inference(token=tok, embedding=g["embedding"][None], n_timesteps=steps)
    e_mel = (mel.cpu() - g["mel"]).abs()
    print(f"[c1] mel {tuple(mel.shape)} max-abs {e_mel.max():.3e} mean-abs {e_mel.mean():.3e}")
    assert mel.shape == g["mel"].shape and e_mel.max().item() < 1e-3
    # 3) the vocoder on the REFERENCE mel with the reference's CPU F0 track pinned (hift_ref.inference docstring), and on the
    #    engine's own mel with the same F0: the whole mel -> waveform leg
    wav_p, _ = mm.models["hift"].inference(g["mel"], f0=g["f0"])
    rms_p = (wav_p.cpu() - g["wav"]).pow(2).mean().sqrt().item()
    wav_c, _ = mm.models["hift"].inference(mel, f0=g["f0"])
    rms_c = (wav_c.cpu() - g["wav"]).pow(2).mean().sqrt().item()
    # 4) nothing pinned: the one-call waveform (GPU F0 predictor on the engine's mel) — reported next to the CPU restatement's own
    #    free-running number, bounded loosely (the harmonic phase integrates F0 over the utterance)
    rms_free = (wavs[0] - g["wav"]).pow(2).mean().sqrt().item()
    print(f"[c1] wav {wav_p.shape[1]} samples: rms {rms_p:.3e} (reference mel, F0 pinned), {rms_c:.3e} (engine mel, F0 pinned), "
          f"{rms_free:.3e} (one call, nothing pinned; CPU oracle with its own F0: {g['oracle_free_f0_rms']:.3e})")
    assert rms_p < 1e-4
    assert rms_c < 1e-4 and rms_free < 5e-2          # measured on B200: 2.4e-6 / 2.4e-6 / 6.5e-3
    mm.engine.close()
