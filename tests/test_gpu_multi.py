"""Multi-GPU parity (needs >= 2 visible GPUs; skipped on a single-GPU box): one rank per GPU under
torchrun, RCB partition, NVLink peer-memory exchanges inside the Krylov loop, fields against the
single-domain oracle (tools/mgpu_check.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("args", [["--peer", "--precond", "ilu0"], ["--peer", "--fused", "--precond", "jacobi"],
                                  ["--kind", "tri", "--nx", "30", "--ny", "26"],
                                  ["--peer", "--precond", "amg", "--nx", "64", "--ny", "48"],
                                  ["--precond", "amg", "--kind", "tri", "--nx", "30", "--ny", "26", "--amg-scope", "local"],
                                  ["--peer", "--strip", "--nx", "40", "--ny", "64", "--precond", "jacobi"]])
def test_two_rank_parity(args):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "mgpu_check.py")] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(" OK") == 2


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")
@pytest.mark.parametrize("args", [["--peer"], ["--peer", "--fuse-rows", "3000"], ["--fuse-rows", "3000", "--precision", "single"]])
def test_two_rank_cycle_equals_transcription(args):
    """the distributed device V-cycle (peer-memory or NCCL ghost refreshes, fused replicated tail) = the scipy
    transcription of the same distributed hierarchy"""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29537", os.path.join(ROOT, "tools", "mgpu_cycle_check.py")] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and " OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
