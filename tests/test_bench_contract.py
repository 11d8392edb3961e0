"""CPU: the parts of bench.py's output contract that can be checked without a GPU -- the reference arm's JSON line (one real CPU
create_proof at a small k) and the device arm's refusal to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True, cwd=ROOT,
                          env=e, timeout=600)


def test_reference_arm_line():
    r = run_bench("--impl", "reference", "--k", "9", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly one JSON line"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "proofs/s" and d["higher_is_better"] is True
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    assert d["value"] > 0 and abs(d["value"] - 1000.0 / d["ms_per_step"]) < 1e-9 * d["value"] + 1e-12
    assert d["e2e"] == {"value": d["value"], "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "create_proof" in cb["sample"]
    assert d["proof_verified"] is True and len(d["proof_sha256"]) == 64 and d["proof_bytes"] > 0 and d["keygen_s"] > 0
    assert d["vs_baseline"] is None and "workload" in d["config"] and "shape" in d["config"]
    assert "synthetic" in d["data"] and "NOT the reference circuit" in d["data"]


def test_reference_arm_under_torchrun_only_rank0_prints():
    r = run_bench("--impl", "reference", "--k", "9", "--steps", "1", "--warmup", "0", "--gpus", "2",
                  env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29533"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_device_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = run_bench("--steps", "1", "--warmup", "1", "--k", "9")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) or "no CPU fallback" in (r.stderr + r.stdout)
