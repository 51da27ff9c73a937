"""U-Net estimator (a7') timing on one B200: ms per NFE and tensor throughput at the bench utterance length.
python scripts/bench_unet.py [T]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmirror_hydravox_b200 import dims as D, synth, _lib as L
from flowmirror_hydravox_b200.flow import NativeUNetEstimator

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2298
ud = D.UNET_FULL
C, inner, ff = ud.ch, ud.heads * ud.head_dim, ud.ch * ud.ff_mult
res = lambda cin: 3 * cin * C + 3 * C * C + cin * C
mac = res(ud.in_ch) + ud.n_mid * res(C) + res(2 * C) + ud.n_res * ud.n_blocks * (C * 3 * inner + inner * C + 2 * C * ff) + 3 * 3 * C * C + C * ud.mel
flops = 2 * (2 * T) * mac + 2 * ud.n_res * ud.n_blocks * 4 * T * T * inner       # CFG batch of 2
print(f"T={T}: {mac / 1e6:.1f} M MAC per frame, {flops / 1e12:.3f} TFLOP per NFE")
g = torch.Generator().manual_seed(0)
x, mu, cond = (torch.randn(2, ud.mel, T, generator=g).cuda() for _ in range(3))
spks, t = torch.randn(2, ud.mel, generator=g).cuda(), torch.tensor([0.5, 0.5]).cuda()
for precise in (False, True):
    e = L.Engine(ud=ud, flow_precise=precise)
    m = NativeUNetEstimator(e).load_state_dict(synth.unet_state_dict(ud, 0))
    for _ in range(3):
        m(x, None, mu, t, spks, cond)
    torch.cuda.synchronize()
    n0 = e.launches()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        m(x, None, mu, t, spks, cond)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"precise={int(precise)}: {ms:.3f} ms per NFE, {flops / ms / 1e9:.1f} TFLOP/s, {(e.launches() - n0) // 10} launches per NFE")
    e.close()

# whole Euler solve (hvx_cfm_solve_unet): CausalConditionalCFM.forward at n_timesteps = 10 (the reference's default) and 25
from flowmirror_hydravox_b200.flow import NativeUNetCFM
e = L.Engine(ud=ud)
cfm = NativeUNetCFM(e); cfm.load_state_dict(synth.unet_state_dict(ud, 0))
mu1, cond1, spk1 = mu[:1].contiguous(), cond[:1].contiguous(), spks[:1].contiguous()
for steps in (10, 25):
    for _ in range(2):
        cfm(mu1, None, steps, spks=spk1, cond=cond1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        cfm(mu1, None, steps, spks=spk1, cond=cond1)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"solve, {steps} Euler steps, T={T} ({T / 50:.1f} s of audio): {ms:.1f} ms, {flops * steps / ms / 1e9:.1f} TFLOP/s")
e.close()
