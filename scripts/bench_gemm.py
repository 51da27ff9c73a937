import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmirror_hydravox_b200 import _lib as L
e = L.Engine()
def run(M, N, K, out_f32=2, act=0, reps=20):
    A = (torch.randn(M, K, device="cuda") * 0.5).half(); B = (torch.randn(N, K, device="cuda") / K ** 0.5).half()
    bias = torch.randn(N, device="cuda"); C = torch.zeros(M, N, device="cuda", dtype=torch.float32 if out_f32 & 1 else torch.float16)
    f = lambda: L.check(L.lib().hvx_gemm_bf16(e.h, L.ptr(A), L.ptr(B), L.ptr(bias), L.ptr(C), M, N, K, out_f32, act, L.stream_ptr()))
    for _ in range(3): f()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0.record()
    for _ in range(reps): f()
    t1.record(); torch.cuda.synchronize()
    us = t0.elapsed_time(t1) * 1e3 / reps
    print(f"M={M} N={N} K={K} out_f32={out_f32} act={act}: {us:8.1f} us  {2.0*M*N*K/us/1e6:7.1f} TFLOP/s")
    ref = torch.matmul(A, B.T)
    t0.record()
    for _ in range(reps): torch.matmul(A, B.T)
    t1.record(); torch.cuda.synchronize()
    us = t0.elapsed_time(t1) * 1e3 / reps
    print(f"   cuBLAS fp16: {us:8.1f} us  {2.0*M*N*K/us/1e6:7.1f} TFLOP/s")
for K in (256, 1024, 4096):
    run(4596, 3072, K)
run(4596, 1024, 1024); run(4596, 2048, 1024, act=1); run(4596, 1024, 2048, out_f32=3)
run(128, 1024, 1024); run(128, 128, 4096)
