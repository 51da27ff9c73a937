"""Host-side logic that needs no GPU: checkpoint packing (the layouts the kernels rely on), roofline arithmetic,
and the no-CPU-fallback contract of the reference-facing objects."""
import pytest
import torch

from flowmirror_hydravox_b200 import dims as D, synth, weights
from oracle import hift_ref, llm_ref


def test_rope_pair_permutation_keeps_q_dot_k():
    """pack_llm permutes q/k rows so the HF half-split RoPE pair (d, d+32) sits on adjacent features; rotating adjacent
    pairs of the permuted vectors must give the same q.k as the reference rotation of the original ones."""
    g = torch.Generator().manual_seed(0)
    perm = weights._rope_pair_perm(1, 64)
    q, k = torch.randn(1, 1, 64, generator=g), torch.randn(1, 1, 64, generator=g)
    pq, pk = torch.tensor([17]), torch.tensor([5])
    ref = (llm_ref.rope_half(q, pq, 1e6) * llm_ref.rope_half(k, pk, 1e6)).sum()

    def rot_pairs(x, pos):
        inv = 1.0 / (1e6 ** (torch.arange(0, 64, 2).float() / 64))
        a = pos.float() * inv
        x = x.reshape(32, 2)
        return torch.stack([x[:, 0] * a.cos() - x[:, 1] * a.sin(), x[:, 1] * a.cos() + x[:, 0] * a.sin()], 1).reshape(64)
    got = (rot_pairs(q[0, 0][perm], pq) * rot_pairs(k[0, 0][perm], pk)).sum()
    assert abs(ref - got) < 1e-4


def test_pack_llm_layouts():
    d = D.LLM_TINY
    sd = synth.llm_state_dict(d, 0)
    o = weights.pack_llm(sd, d)
    qd, kd = d.q_heads * d.head_dim, d.kv_heads * d.head_dim
    assert o["L0.qkv.w"].shape == (qd + 2 * kd, d.hidden) and o["L0.qkv.w"].dtype == torch.bfloat16
    assert o["L0.qkv.b"].dtype == torch.float32
    # v rows are not permuted; q row 1 is original row 32 of head 0
    assert torch.equal(o["L0.qkv.w"][qd + kd:], sd["llm.model.model.layers.0.self_attn.v_proj.weight"].bfloat16())
    assert torch.equal(o["L0.qkv.w"][1], sd["llm.model.model.layers.0.self_attn.q_proj.weight"][32].bfloat16())
    # gate/up interleave: row 2n = gate_n, row 2n+1 = up_n
    assert torch.equal(o["L1.gu.w"][6], sd["llm.model.model.layers.1.mlp.gate_proj.weight"][3].bfloat16())
    assert torch.equal(o["L1.gu.w"][7], sd["llm.model.model.layers.1.mlp.up_proj.weight"][3].bfloat16())
    # MTP heads stacked, dead q/k projections dropped
    assert o["mtp.gu.w"].shape == (d.mtp_heads, 2 * d.mtp_inter, d.hidden)
    assert not any("q_proj" in k or ".q." in k for k in o if k.startswith("mtp"))


def test_pack_flow_conv_as_gemm_operands():
    d = D.FLOW_TINY
    sd = synth.flow_state_dict(d, 0)
    o = weights.pack_flow(sd, d)
    w = sd["decoder.estimator.input_embed.conv_pos_embed.conv1.0.weight"]          # (dim, 64, k)
    assert o["pos1.w"].shape == (d.dim, d.pos_k * 64)
    assert o["pos1.w"][5, 7 * 64 + 9] == w[5, 9, 7].half()                           # column = tap*64 + ci
    w1 = sd["pre_lookahead_layer.conv1.weight"]                                      # (pla, mel, 4)
    assert o["pla1.w"].shape == (d.pla_ch, 4 * 128)
    assert o["pla1.w"][3, 2 * 128 + 11] == w1[3, 11, 2].half() and o["pla1.w"][3, 2 * 128 + 100] == 0   # padded to 128
    assert o["mod.w"].shape == (d.depth * 6 * d.dim + 2 * d.dim, d.dim)
    assert o["blk0.qkv.w"].shape == (3 * d.heads * d.dim_head, d.dim)


def test_pack_hift_folds_weight_norm_like_the_oracle():
    d = D.HIFT_TINY
    sd = synth.hift_state_dict(d, 0)
    o = weights.pack_hift(sd, d)
    w = hift_ref.fold_weight_norm(sd)
    assert torch.allclose(o["conv_pre.w"], w["conv_pre.weight"].permute(1, 2, 0), atol=0, rtol=0)     # [Cin][K][Cout]
    assert torch.allclose(o["rb.2.c1.1.w"], w["resblocks.2.convs1.1.weight"].permute(1, 2, 0), atol=0, rtol=0)


def test_roofline_arithmetic_matches_survey():
    import bench
    # 971 MB of live bf16 weights per decode step at K=2 (+ KV traffic)
    b = bench.llm_step_bytes(D.LLM_FULL, 2, 0, 0)
    assert abs(b - 970.98e6) < 0.5e6
    kv = bench.llm_step_bytes(D.LLM_FULL, 2, 1000, 1) - b
    assert abs(kv - 24 * 2 * 128 * 2 * 1002) < 1
    # 0.756*T + 1.80e-4*T^2 GFLOP per NFE
    f = bench.flow_nfe_flops(D.FLOW_FULL, 2298) / 1e9
    assert abs(f - (0.756 * 2298 + 1.80e-4 * 2298 ** 2)) / f < 0.01


def test_no_cpu_fallback():
    from flowmirror_hydravox_b200 import _lib as L
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from flowmirror_hydravox_b200.model_manager import ModelManager
    with pytest.raises(L.HvxError):
        ModelManager(hd=D.HIFT_TINY, fd=D.FLOW_TINY, ld=D.LLM_TINY)


def test_stitching_and_wav_encoding():
    """pause insertion of inference_tts_with_segmentation (infer_speech_model.py:419-441) and audio_to_base64 (:504-521)."""
    import base64, io, random
    from scipy.io import wavfile
    from flowmirror_hydravox_b200 import output
    segs = [torch.full((1, 100), 0.1), torch.full((1, 50), -0.2), torch.full((1, 70), 0.3)]
    out = output.stitch_segments(segs, 24000, random.Random(3))
    r = random.Random(3)
    pauses = [int(r.uniform(50, 150) * 24000 / 1000) for _ in range(2)]
    assert out.shape == (1, 220 + sum(pauses)) and all(1200 <= p <= 3600 for p in pauses)
    assert torch.equal(out[:, :100], segs[0]) and out[:, 100:100 + pauses[0]].abs().max() == 0
    assert torch.equal(out[:, -70:], segs[2])
    b64 = output.audio_to_base64(out, 24000)
    sr, data = wavfile.read(io.BytesIO(base64.b64decode(b64)))
    assert sr == 24000 and data.dtype.kind == "f" and data.shape[0] == out.shape[1]
    assert torch.equal(torch.from_numpy(data.copy()).reshape(1, -1), out)
    with pytest.raises(ValueError):
        output.stitch_segments([], 24000)


def test_stream_oracle_fade_in_out_matches_reference_function():
    """oracle/stream_ref.fade_in_out == cosyvoice/utils/common.py:169-177 (run when /root/reference is present), and the blend is
    evaluated in float64 (float32 tensor x float64 numpy window)"""
    import os
    import numpy as np
    import torch
    from oracle import stream_ref
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(1, 9000, generator=g), torch.randn(1, 3840, generator=g)
    w = np.hamming(2 * 3840)
    out = stream_ref.fade_in_out(a, b, w)
    want = (a[:, :3840].double() * torch.from_numpy(w[:3840]) + b.double() * torch.from_numpy(w[3840:])).float()
    assert torch.equal(out[:, :3840], want) and torch.equal(out[:, 3840:], a[:, 3840:])
    if os.path.isdir("/root/reference"):
        from oracle import refshim
        refshim.install()
        from cosyvoice.utils.common import fade_in_out as ref_fade
        assert torch.equal(ref_fade(a.clone(), b, w), out)


def test_pack_unet_nc_resampling_convs_are_exact_repacks():
    """weights.pack_unet_nc turns Downsample1D (Conv1d k3, stride 2, pad 1) into a 2-tap convolution over frame pairs and Upsample1D
    (ConvTranspose1d(C, C, 4, 2, 1)) into a k3 convolution with 2C outputs (matcha/models/components/decoder.py:64-70,116-158): both
    are the same linear maps, checked here against torch on the CPU (up to the fp16 rounding of the packed weights)."""
    import torch
    import torch.nn.functional as F
    from flowmirror_hydravox_b200 import dims as D, synth
    from flowmirror_hydravox_b200.weights import pack_unet_nc
    d = D.UNET_NC_TINY
    sd = synth.unet_nc_state_dict(d, 0)
    pk = pack_unet_nc(sd, d)
    C = d.ch
    g = torch.Generator().manual_seed(3)
    for T in (11, 12):                                   # odd and even lengths: ceil(T / 2) output frames
        x = torch.randn(1, C, T, generator=g)
        w, b = sd["down_blocks.0.2.conv.weight"].half().float(), sd["down_blocks.0.2.conv.bias"]
        ref = F.conv1d(x, w, b, stride=2, padding=1)
        To = (T + 1) // 2
        pairs = F.pad(x, (0, 2 * To - T)).transpose(1, 2).reshape(To, 2 * C)            # row t = (x[2t], x[2t+1])
        rows = torch.cat([F.pad(pairs, (0, 0, 1, 0))[:-1], pairs], 1)                   # [row t-1 | row t]
        y = (rows @ pk["down0.w"].float().t() + pk["down0.b"]).t()[None]
        assert y.shape == ref.shape and (y - ref).abs().max() < 1e-5
        wt, bt = sd["up_blocks.0.2.conv.weight"].half().float(), sd["up_blocks.0.2.conv.bias"]
        ref = F.conv_transpose1d(x, wt, bt, stride=2, padding=1)
        xr = x[0].t()
        rows = torch.cat([F.pad(xr, (0, 0, 1, 0))[:-1], xr, F.pad(xr, (0, 0, 0, 1))[1:]], 1)   # [x[m-1] | x[m] | x[m+1]]
        y = (rows @ pk["up0.w"].float().t() + pk["up0.b"]).reshape(-1, C).t()[None]      # row m = (y[2m], y[2m+1])
        assert y.shape == ref.shape and (y - ref).abs().max() < 1e-5
    # parity mode packs [hi | lo]: hi + lo reproduces the fp32 weights
    pk2 = pack_unet_nc(sd, d, precise=True)
    hi, lo = pk2["res0.c1.w"].float().chunk(2, dim=1)
    w = sd["down_blocks.0.0.block1.block.0.weight"]
    assert (hi + lo - w.permute(0, 2, 1).reshape(w.shape[0], -1)).abs().max() < 1e-6


def test_product_path_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package (or the C sources) may import, call or load it — the engine has no
    CPU fallback.  Checked on the syntax tree of every module of the package and by a fresh interpreter's sys.modules after import."""
    import ast
    import glob
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "flowmirror_hydravox_b200")
    for path in glob.glob(os.path.join(pkg, "*.py")):
        for node in ast.walk(ast.parse(open(path).read())):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n == "oracle" or n.startswith("oracle.") for n in names), path
    code = ("import sys; sys.path.insert(0, %r); import flowmirror_hydravox_b200 as p; "
            "from flowmirror_hydravox_b200 import _lib, llm, flow, hift, model_manager, streaming, parallel, output, frontend, synth, dims; "
            "print(sorted(m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')))" % root)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-1500:]
    assert out.stdout.strip().splitlines()[-1] == "[]"
    import re
    for path in glob.glob(os.path.join(pkg, "csrc", "*")):          # the C sources cannot reach Python or another library at run time
        src = re.sub(r"//[^\n]*|/\*.*?\*/", "", open(path, errors="ignore").read(), flags=re.S)
        assert not re.search(r"\b(dlopen|popen|system|execv\w*|Py_\w+)\s*\(", src), path
