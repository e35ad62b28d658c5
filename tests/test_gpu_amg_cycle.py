"""One application of the device V-cycle against the scipy transcription of the SAME hierarchy (the library's host
setup read back level by level): separate launches per level, and the fused small-level kernel (k_amg_tail), in
double (1e-10) and single precision.  Converged fields cannot tell a correct preconditioner from a merely
convergent one; this can."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("args", [["--fuse-rows", "0"], ["--fuse-rows", "10000"], ["--fuse-rows", "10000", "--precision", "single"],
                                  ["--fuse-rows", "200000", "--nx", "700", "--ny", "600", "--coarsest", "1000"]])
def test_device_cycle_equals_transcription(args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "mgpu_cycle_check.py")] + args, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and " OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    if args[1] != "0":
        assert "launches per cycle" in r.stdout
