"""CUDA path against the committed golden fixtures (no oracle involved at run time)."""
import glob
import os

import numpy as np
import pytest

from tests.util import ORACLE_MESH_INT, rel_l2

pytestmark = pytest.mark.gpu
GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")) if not os.path.basename(p).startswith("ref_"))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_gpu_matches_golden(path):
    from phase_b200.api import Communicator, FiniteVolumeGrid2D as G, lid_driven_cavity
    g = np.load(path)
    comm = Communicator(0)
    mk = G.rectilinear if str(g["kind"]) == "rect" else G.triangulated
    grid = mk(comm, int(g["nx"]), int(g["ny"]), float(g["w"]), float(g["h"]))
    for k in ORACLE_MESH_INT:
        assert np.array_equal(grid.i32(k), g["mesh_" + k]), k
    fs = lid_driven_cavity(grid, 1.0, 0.1, solver=dict(tolerance=1e-12, maxIters=20000))
    s = lambda k: g["state_" + k]
    fs.u.set("cells", np.concatenate([s("ux"), s("uy")])); fs.u.set("faces", np.concatenate([s("ufx"), s("ufy")]))
    fs.gradP.set("cells", np.concatenate([s("gpx"), s("gpy")]))
    fs.p.set("cells", s("p")); fs.p.set("faces", s("pf"))
    fs.u.savePreviousTimeStep()
    fs.u.set("cells0", np.concatenate([s("u0x"), s("u0y")])); fs.u.set("faces0", np.concatenate([s("u0fx"), s("u0fy")]))
    dt = float(g["dt"])
    for tag, eq, layout in (("u", fs.assembleU(dt), 0), ("p", fs.assembleP(dt), 1)):
        rp, ci, va, rhs = eq.export(layout)
        assert np.array_equal(rp, g[tag + "_rowPtr"]) and np.array_equal(ci, g[tag + "_colInd"])      # bit-exact
        assert np.allclose(va, g[tag + "_vals"], rtol=2e-13, atol=2e-13 * np.abs(g[tag + "_vals"]).max())
        assert np.allclose(rhs, g[tag + "_rhs"], rtol=1e-11, atol=1e-12 * np.abs(g[tag + "_rhs"]).max())
    fs.close()
    fs = lid_driven_cavity(grid, 1.0, 0.1, solver=dict(tolerance=1e-12, maxIters=20000))
    for _ in range(int(g["K"])):
        fs.solve(dt)
    u, p = fs.u.get("cells"), fs.p.get("cells")
    assert rel_l2(u[0], g["final_ux"]) < 1e-6 and rel_l2(u[1], g["final_uy"]) < 1e-6
    assert rel_l2(p - p.mean(), g["final_p0"]) < 1e-6
    uf = fs.u.get("faces")
    assert rel_l2(uf[0], g["final_ufx"]) < 1e-6 and rel_l2(uf[1], g["final_ufy"]) < 1e-6
    fs.close(); grid.close(); comm.close()
