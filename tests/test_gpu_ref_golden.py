"""The CUDA path (through the C ABI) against golden vectors written by the REFERENCE ITSELF
(tests/golden/ref_*.npz; see tests/golden/make_ref_golden.py): connectivity, link tables and the CSR patterns at
the solver hand-off bit-exact; fields after K time steps within the north star's rel-L2 1e-6."""
import glob
import os

import numpy as np
import pytest

from tests.util import rel_l2

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "ref_cavity_*.npz")))
INT_KEYS = ["cptr", "cind", "faceN1", "faceN2", "faceL", "faceR", "ilPtr", "ilFace", "ilCell", "blPtr", "blFace", "dlPtr", "dlCell"]


@pytest.fixture(scope="module")
def comm():
    from phase_b200.api import Communicator
    c = Communicator(0)
    yield c
    c.close()


@pytest.mark.parametrize("pc", ["ilu0", "amg"])
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_cuda_path_reproduces_reference_golden(comm, path, pc):
    from phase_b200.api import FiniteVolumeGrid2D as Grid, lid_driven_cavity
    G = np.load(path)
    kind, nx, ny, w, h, K, dt = str(G["kind"]), int(G["nx"]), int(G["ny"]), float(G["w"]), float(G["h"]), int(G["K"]), float(G["dt"])
    g = (Grid.rectilinear if kind == "rect" else Grid.triangulated)(comm, nx, ny, w, h)
    for k in INT_KEYS:
        assert np.array_equal(g.i32(k), G["mesh_" + k]), k
    for k in ("vol", "cellCx", "cellCy", "faceCx", "faceCy"):
        assert np.allclose(g.f64(k), G["mesh_" + k], rtol=1e-13, atol=1e-15), k
    fs = lid_driven_cavity(g, float(G["rho"]), float(G["mu"]),
                           solver=dict(tolerance=1e-12, maxIters=20000, preconditioner=pc, amgCoarsest=30))
    for _ in range(K):
        st = fs.solve(dt)
    u, uf, p, pr = fs.u.get("cells"), fs.u.get("faces"), fs.p.get("cells"), G["field_p"]
    assert rel_l2(u[0], G["field_ux"]) < 1e-6 and rel_l2(u[1], G["field_uy"]) < 1e-6
    assert rel_l2(uf[0], G["field_ufx"]) < 1e-6 and rel_l2(uf[1], G["field_ufy"]) < 1e-6
    assert rel_l2(p - p.mean(), pr - pr.mean()) < 1e-6
    gp = fs.gradP.get("cells")
    assert rel_l2(gp[0], G["field_gpx"]) < 1e-6 and rel_l2(gp[1], G["field_gpy"]) < 1e-6
    # the systems of step K in the reference's own hand-off layouts
    for which, eq, layout in (("uEqn", fs.uEqn, 0), ("pEqn", fs.pEqn, 1)):
        rp, ci, va, rhs = eq.export(layout)
        assert np.array_equal(rp, G[which + "_rowPtr"]) and np.array_equal(ci, G[which + "_colInd"]), which
        assert np.abs(va - G[which + "_vals"]).max() <= 1e-9 * np.abs(va).max(), which
        assert np.abs(-rhs - G[which + "_b"]).max() <= 1e-7 * max(np.abs(rhs).max(), 1e-300), which
    assert abs(st["maxCourant"] - float(G["maxCourant"])) < 1e-8
    fs.close(); g.close()
