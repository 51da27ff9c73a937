"""torchrun entry: utterance-sharded synthesis over the GPUs of the box (parallel.synthesize_sharded, NCCL scatter/gather).
rank 0 checks that every waveform equals the one the same request produces on a single GPU."""
import os, sys, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowmirror_hydravox_b200 import dims as D, synth, parallel
from flowmirror_hydravox_b200.model_manager import ModelManager

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
mm = ModelManager(hd=D.HIFT_TINY, fd=D.FLOW_TINY, ld=D.LLM_TINY, device=f"cuda:{local}", max_ctx=512, max_seqs=8, n_timesteps=4, sine_seconds=20.0)
mm.load_state_dicts(synth.llm_state_dict(D.LLM_TINY, 0, eos_scale=0.0), synth.flow_state_dict(D.FLOW_TINY, 0), synth.hift_state_dict(D.HIFT_TINY, 0))
sp = dict(top_p=0.9, top_k=10, win_size=24, tau_r=0.2)
N = 7
reqs = None
if rank == 0:
    reqs = []
    for i in range(N):
        r = synth.utterance(D.LLM_TINY, D.FLOW_TINY, 4 + i, seed=100 + i, prompt_tokens=3 + i, prompt_text=2)
        r["min_ratio"] = r["max_ratio"] = 4.0
        r["u"] = torch.rand(1024, generator=torch.Generator().manual_seed(i))
        reqs.append(r)

def synth_fn(mine):
    if not mine:
        return []
    u = torch.stack([r["u"] for r in mine])
    return mm.synthesize_batch(mine, head_k=2, sampling=sp, n_timesteps=4, u=u)

out = parallel.synthesize_sharded(synth_fn, reqs)
if rank == 0:
    ref = [mm.synthesize_batch([r], head_k=2, sampling=sp, n_timesteps=4, u=r["u"][None])[0] for r in reqs]
    ok = all(a.shape == b.shape and (a - b).abs().max().item() < 1e-5 for a, b in zip(out, ref))
    print("SHARDED_OK" if ok else "SHARDED_MISMATCH", [tuple(a.shape) for a in out])
dist.barrier()
dist.destroy_process_group()
