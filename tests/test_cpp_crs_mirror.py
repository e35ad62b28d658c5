"""include/phase/CrsEquation.h (the host fall-through of Seam 2) against the reference's own
CrsEquation compiled in place (oracle/_ref): identical arrays, bit for bit, on random scripts."""
import os
import subprocess

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(O.ref_lib() is None, reason="oracle/_ref not built and /root/reference absent")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = tmp_path_factory.mktemp("crs") / "crs_mirror_check"
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "crs_mirror_check.cpp"), "-o", str(out)])
    return str(out)


@pytest.mark.parametrize("seed", range(5))
def test_mirror_matches_reference(exe, seed):
    rng = np.random.default_rng(100 + seed)
    n = 13
    script, ref = [], {}
    for i, nnz in enumerate((3, 5, 2)):
        script.append("new %d %d %d" % (i, n, nnz))
        ref[i] = O.RefCrs(n, nnz)
    for _ in range(160):
        i = int(rng.integers(3)); r = int(rng.integers(n)); c = int(rng.integers(n))
        v = float(rng.choice([0.0, 1.5, -2.0, float(rng.standard_normal())]))
        k = rng.random()
        if k < 0.6:
            script.append("add %d %d %d %r" % (i, r, c, v)); ref[i].add_coeff(r, c, v)
        elif k < 0.75:
            script.append("set %d %d %d %r" % (i, r, c, v)); ref[i].set_coeff(r, c, v)
        elif k < 0.9:
            script.append("rhs %d %d %r" % (i, r, v)); ref[i].add_rhs(r, v)
        else:
            script.append("scale %d %d %r" % (i, r, v)); ref[i].scale_row(r, v)
    script += ["+= 0 1", "-= 0 2", "*= 0 0.37", "dump 0", "dump 1"]
    ref[0].add_eq(ref[1]); ref[0].sub_eq(ref[2]); ref[0].scale(0.37)
    out = subprocess.run([exe], input="\n".join(script) + "\n", capture_output=True, text=True, check=True).stdout
    lines = out.strip().splitlines()
    for blk, rid in ((lines[:4], 0), (lines[4:8], 1)):
        got = {l.split()[0]: l.split()[1:] for l in blk}
        rp, ci, va, rhs = ref[rid].export()
        assert [int(x) for x in got["rowPtr"]] == list(rp)
        assert [int(x) for x in got["colInd"]] == list(ci)
        assert np.array_equal(np.array([float(x) for x in got["vals"]]), va)
        assert np.array_equal(np.array([float(x) for x in got["rhs"]]), rhs)
