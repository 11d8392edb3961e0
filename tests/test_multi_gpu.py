"""GPU, >= 2 devices: one proof computed jointly by a group of GPUs (csrc/group.cu) equals the single-GPU proof and the CPU
prover's bytes -- with the peer-to-peer exchange (stores into the peers' windows from k_msm_final's epilogue + flags) and with
the NCCL all-gathers (ZK_GROUP_NO_P2P=1).  Skipped on a one-GPU box; run with `gpurun --gpus 2`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


def _run(world, shape, k, expect_p2p, extra_env=None, port=29641):
    env = dict(os.environ, **(extra_env or {}))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "workers", "group_proof_worker.py"), shape, str(k), "1" if expect_p2p else "0"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ok=True" in r.stdout, r.stdout[-2000:]
    return r.stdout


@pytest.mark.parametrize("shape,k", [("toy", 6), ("qe3", 11)])
def test_group_of_two_p2p(shape, k):
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = _run(2, shape, k, expect_p2p=True)
    assert "p2p=True" in out


def test_group_of_two_nccl_only():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = _run(2, "qe3", 11, expect_p2p=False, extra_env={"ZK_GROUP_NO_P2P": "1"}, port=29642)
    assert "p2p=False" in out


def test_group_of_all_gpus():
    n = _gpus()
    if n < 4:
        pytest.skip("needs >= 4 GPUs")
    world = 8 if n >= 8 else 4
    _run(world, "qe3", 12, expect_p2p=True, port=29643)
