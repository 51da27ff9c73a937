"""tcgen05/TMA GEMM and attention kernels vs torch fp32 math on the same bf16 inputs."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from flowmirror_hydravox_b200 import _lib as L
    e = L.Engine()
    yield e
    e.close()


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 256), (300, 1024, 1024), (4596, 3072, 1024),
                                   (25, 2048, 1024), (517, 80, 1024), (1000, 1024, 2048), (64, 6144, 320),
                                   (4596, 1024, 1024), (4596, 2048, 1024), (4596, 1024, 2048), (20000, 1024, 1024)])
@pytest.mark.parametrize("out_f32,act", [(0, 0), (1, 0), (0, 1), (2, 0), (3, 1)])     # bit 1 of out_f32: fp16 operands
def test_gemm(eng, M, N, K, out_f32, act):
    from flowmirror_hydravox_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    dt = torch.float16 if out_f32 & 2 else torch.bfloat16
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(dt)
    B = (torch.randn(N, K, device="cuda", generator=g) * (1.0 / K ** 0.5)).to(dt)
    bias = torch.randn(N, device="cuda", generator=g)
    Cout = torch.zeros(M, N, device="cuda", dtype=torch.float32 if out_f32 & 1 else dt)
    L.check(L.lib().hvx_gemm_bf16(eng.h, L.ptr(A), L.ptr(B), L.ptr(bias), L.ptr(Cout), M, N, K, out_f32, act, L.stream_ptr()))
    torch.cuda.synchronize()
    ref = A.float() @ B.float().T + bias
    if act == 1:
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    tol = 2e-3 if out_f32 & 1 else (4e-3 if out_f32 & 2 else 2e-2)
    err = (Cout.float() - ref).abs().max().item()
    assert err < tol, err


def _split16(x):
    hi = x.to(torch.float16)
    return torch.cat([hi, (x - hi.float()).to(torch.float16)], dim=1).contiguous()


@pytest.mark.parametrize("M,N,K", [(4596, 1024, 1024), (4596, 3072, 1024), (4596, 1024, 2048), (16200, 2048, 1024), (300, 1024, 1024),
                                   (2298, 80, 1024), (130, 256, 256), (4596, 1024, 64)])
@pytest.mark.parametrize("pair", [True, False])
def test_gemm_three_term_split_precision(eng, monkeypatch, M, N, K, pair):
    """A_hi W_hi + A_lo W_hi + A_hi W_lo on split-fp16 operands (the flow's parity mode; tile-sharing stages, and the CTA-pair
    cta_group::2 kernel for the big shapes unless HVX_NO_PAIR): fp32-level agreement with the fp64 product."""
    from flowmirror_hydravox_b200 import _lib as L
    if not pair:
        monkeypatch.setenv("HVX_NO_PAIR", "1")
    else:
        monkeypatch.delenv("HVX_NO_PAIR", raising=False)
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g) * 0.5
    B = torch.randn(N, K, device="cuda", generator=g) * (1.0 / K ** 0.5)
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32)
    A2, B2 = _split16(A), _split16(B)                       # keep the operands alive for the duration of the call
    L.check(L.lib().hvx_gemm_bf16(eng.h, L.ptr(A2), L.ptr(B2), L.ptr(bias), L.ptr(out), M, N, K, 1 | 2 | 4, 0, L.stream_ptr()))
    torch.cuda.synchronize()
    ref = (A.double() @ B.double().T + bias.double()).float()
    err = (out - ref).abs().max().item()
    print(f"[gemm3 {M}x{N}x{K} pair={pair}] max-abs {err:.3e}")
    assert err < 5e-5, err


@pytest.mark.parametrize("B,T,H,chunk", [(1, 128, 1, 0), (2, 200, 2, 0), (2, 1000, 16, 0), (2, 333, 4, 50), (1, 2298, 16, 0)])
def test_attention(eng, B, T, H, chunk):
    from flowmirror_hydravox_b200 import _lib as L
    g = torch.Generator(device="cuda").manual_seed(T)
    q = torch.randn(B, T, H, 64, device="cuda", generator=g).bfloat16()
    k = torch.randn(B, T, H, 64, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, T, H, 64, device="cuda", generator=g).bfloat16()
    qk = torch.cat([q.reshape(B * T, H * 64), k.reshape(B * T, H * 64)], dim=1).contiguous()
    Tp = (T + 7) // 8 * 8
    vt = torch.zeros(B * H * 64, Tp, device="cuda", dtype=torch.bfloat16)
    vt[:, :T] = v.permute(0, 2, 3, 1).reshape(B * H * 64, T)
    out = torch.zeros(B * T, H * 64, device="cuda", dtype=torch.bfloat16)
    L.check(L.lib().hvx_attention_bf16(eng.h, L.ptr(qk), L.ptr(vt), Tp, L.ptr(out), B, T, H, chunk, L.stream_ptr()))
    torch.cuda.synchronize()
    qf, kf, vf = (x.float().permute(0, 2, 1, 3) for x in (q, k, v))
    s = qf @ kf.transpose(-1, -2) * 0.125
    if chunk:
        ar = torch.arange(T, device="cuda")
        s = s.masked_fill(~(ar[None, :] < ((ar // chunk + 1) * chunk)[:, None]), float("-inf"))
    ref = (s.softmax(-1) @ vf).permute(0, 2, 1, 3).reshape(B * T, H * 64)
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2, err
