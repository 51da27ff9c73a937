"""Per-hop cost of the streaming flow, full dims: recompute-all (flow.inference over all tokens so far, the reference's schedule)
vs the incremental session (hvx_flow_stream_append).  python scripts/bench_stream_flow.py [n_tokens] [n_timesteps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmirror_hydravox_b200 import dims as D, synth, _lib as L
from flowmirror_hydravox_b200.flow import NativeFlow
n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
fd = D.FLOW_FULL
hop, la, P = 25, 3, 125
for precise in (False, True):
    e = L.Engine(fd=fd, flow_precise=precise)
    f = NativeFlow(e, n_timesteps=steps); f.load_state_dict(synth.flow_state_dict(fd, 0))
    g = torch.Generator().manual_seed(1)
    tok = torch.randint(0, fd.vocab, (1, n_total), generator=g).cuda()
    ptok = torch.randint(0, fd.vocab, (1, P), generator=g).cuda()
    pfeat = (torch.rand(1, 2 * P, fd.mel, generator=g) * 6 - 6).cuda()
    emb = torch.randn(1, fd.spk_in, generator=g).cuda()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for mode in ("recompute", "incremental"):
        for rep in range(2):
            if mode == "incremental":
                f.stream_begin(max_frames=2 * (P + n_total))
            torch.cuda.synchronize()
            off, t_hops, worst = 0, [], 0.0
            while n_total - off >= hop + la:
                n_tok = off + hop + la
                a, b = ev(), ev()
                a.record()
                if mode == "incremental":
                    mel = f.stream_append(tok[:, :n_tok], emb, prompt_token=ptok, prompt_feat=pfeat)
                else:
                    mel, _ = f.inference(token=tok[:, :n_tok], embedding=emb, prompt_token=ptok, prompt_feat=pfeat, streaming=True, finalize=False)
                b.record(); torch.cuda.synchronize()
                t_hops.append(a.elapsed_time(b))
                off += hop
            if mode == "incremental":
                f.stream_end()
        n = len(t_hops)
        print(f"[{'parity' if precise else 'serving'} mode, {steps} Euler steps, {n} hops to {2 * (P + off)} frames] {mode}: total {sum(t_hops):.0f} ms, "
              f"first hop {t_hops[0]:.1f} ms, hop at ~1000 frames {t_hops[min(n - 1, 14)]:.1f} ms, last hop {t_hops[-1]:.1f} ms", flush=True)
    e.close()
