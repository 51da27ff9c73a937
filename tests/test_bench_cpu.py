"""bench.py's reference arm (the reference algorithm's CPU path = the oracle port on the host cores) runs without a GPU: its JSON line
keeps the driver's contract, and under torchrun only rank 0 works."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], capture_output=True, text=True,
                          env=env, timeout=900)


def test_reference_arm_line_keeps_the_contract():
    r = _run({}, "--steps", "1", "--warmup", "0", "--cpu-tokens", "16", "--cfm-steps", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "speech_tokens_per_s" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "16 speech tokens" in cb["sample"]
    assert "configs[2]" in d["config"]["workload"] and "inference_head_num=4" in d["config"]["workload"]      # the native arm's workload
    assert set(cb["stage_s"]) == {"llm", "flow", "hift"}


def test_reference_arm_other_ranks_exit_without_work():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_native_arm_fails_loudly_without_a_gpu():
    """no CPU fallback: the native arm of the bench exits non-zero with an error and prints no JSON line when there is no CUDA device"""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--no-extras", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert "Error" in r.stderr
