"""Device-side timeline of one CTA of the tile GEMM at the 128-row decode shapes (HVX_GEMM_TIMELINE=1 prints it):
python scripts/gemm_timeline.py   — qkv (N=1152), gate/up (N=9728), o-proj (N=896) with K'=1792, and M = 32 / 128 rows."""
import os, sys, ctypes as C, torch
os.environ["HVX_GEMM_TIMELINE"] = "1"
os.environ["HVX_NO_PERSIST"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmirror_hydravox_b200 import _lib as L
e = L.Engine()
g = torch.Generator().manual_seed(0)
for M, N, K in ((128, 1152, 1792), (32, 1152, 1792), (128, 9728, 1792), (128, 896, 1792), (128, 896, 9728), (128, 1152, 448)):
    A = (torch.randn(M, K, generator=g) * 0.1).to(torch.bfloat16).cuda()
    B = (torch.randn(N, K, generator=g) * 0.1).to(torch.bfloat16).cuda()
    out = torch.empty(M, N, dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for it in range(2):
        flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        L.check(L.lib().hvx_gemm_bf16(e.h, L.ptr(A), L.ptr(B), None, L.ptr(out), M, N, K, 1, 0, L.stream_ptr()))
        b.record(); torch.cuda.synchronize()
        print(f"M={M} N={N} K={K} run {it}: {a.elapsed_time(b) * 1e3:.1f} us (incl. timeline sync)", flush=True)
    ref = A.float() @ B.float().t()
    print("   max-abs err", (out - ref).abs().max().item(), flush=True)
    # kernel duration by CUDA events without the timeline dump (5 back-to-back launches, L2 flushed before each)
    os.environ.pop("HVX_GEMM_TIMELINE", None)

# plain event timing in a fresh process state is not possible (the env switch is read once); use torch.profiler on the same calls
from torch.profiler import profile, ProfilerActivity
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for M, N, K in ((128, 1152, 1792), (32, 1152, 1792), (128, 9728, 1792), (128, 896, 1792)):
        A = (torch.randn(M, K, generator=g) * 0.1).to(torch.bfloat16).cuda()
        B = (torch.randn(N, K, generator=g) * 0.1).to(torch.bfloat16).cuda()
        out = torch.empty(M, N, dtype=torch.float32, device="cuda")
        for it in range(3):
            flush.zero_()
            L.check(L.lib().hvx_gemm_bf16(e.h, L.ptr(A), L.ptr(B), None, L.ptr(out), M, N, K, 1, 0, L.stream_ptr()))
        torch.cuda.synchronize()
for ev in prof.events():
    if "gemm" in ev.name:
        print(f"CUPTI {ev.name[:60]} {ev.device_time:.1f} us", flush=True)
