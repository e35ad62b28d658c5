"""The oracle reproduces the committed golden fixtures (made by
tests/golden/make_golden.py in the container that holds the reference)."""
import glob
import os

import numpy as np
import pytest

import oracle as O
from tests.util import ORACLE_MESH_INT

GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")) if not os.path.basename(p).startswith("ref_"))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_oracle_matches_golden(path):
    g = np.load(path)
    mk = O.Mesh.rectilinear if str(g["kind"]) == "rect" else O.Mesh.triangulated
    om = mk(int(g["nx"]), int(g["ny"]), float(g["w"]), float(g["h"]))
    for k in ORACLE_MESH_INT:
        assert np.array_equal(om.array(k), g["mesh_" + k]), k
    fs = O.cavity(om, 1.0, 0.1)
    for k in ("ux", "uy", "gpx", "gpy", "p", "u0x", "u0y", "ufx", "ufy", "pf", "u0fx", "u0fy"):
        fs.view(k)[:] = g["state_" + k]
    dt = float(g["dt"])
    for tag, e in (("u", fs.assemble_u(dt)), ("p", fs.assemble_p(dt))):
        rp, ci, va, rhs = e.export()
        assert np.array_equal(rp, g[tag + "_rowPtr"]) and np.array_equal(ci, g[tag + "_colInd"])
        assert np.array_equal(va, g[tag + "_vals"]) and np.array_equal(rhs, g[tag + "_rhs"])
    part = g["part"]
    for r, loc in enumerate(om.partition(part, 3)):
        for k in ("globalId", "owner", "localRow", "globalRow", "bufPtr", "bufCell", "sendPtr", "sendCell"):
            assert np.array_equal(loc.array(k), g["part%d_%s" % (r, k)]), (r, k)


def test_golden_partition_maps_on_product_path():
    """I5 through the C ABI (host-only context) against the golden halo maps."""
    from phase_b200.api import Communicator, FiniteVolumeGrid2D as G
    for path in GOLD:
        g = np.load(path)
        hc = Communicator(Communicator.HOST_ONLY)
        mk = G.rectilinear if str(g["kind"]) == "rect" else G.triangulated
        gg = mk(hc, int(g["nx"]), int(g["ny"]), float(g["w"]), float(g["h"]))
        for r in range(3):
            gl = gg.local(g["part"], Communicator(Communicator.HOST_ONLY, r, 3))
            for k in ("globalId", "owner", "localRow", "globalRow", "bufPtr", "bufCell", "sendPtr", "sendCell", "faceL", "faceR"):
                assert np.array_equal(gl.i32(k), g["part%d_%s" % (r, k)]), (path, r, k)
