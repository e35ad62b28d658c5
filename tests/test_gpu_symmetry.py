"""SYMMETRY velocity patches: the tensor terms of the vector Laplacian (UD/Laplacian.cpp:33-41) and
add(cell, cell, Tensor2D) (UE/VectorFiniteVolumeEquation.cpp:48-66) on the device against a golden run of the
REFERENCE ITSELF (tests/golden/ref_symmetry_sheared_10x8.npz, written by oracle/_ref/libphase_ref_fv.so): a sheared
quad mesh whose x-/x+ boundaries are slanted, so the symmetry condition couples the velocity components (cross
entries in uEqn_)."""
import os

import numpy as np
import pytest

from tests.util import rel_l2

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def comm():
    from phase_b200.api import Communicator
    c = Communicator(0)
    yield c
    c.close()


@pytest.mark.parametrize("pc", ["ilu0", "jacobi", "amg"])
def test_symmetry_patches_against_reference_golden(comm, pc):
    from phase_b200.api import FIXED, NORMAL_GRADIENT, SYMMETRY, FiniteVolumeGrid2D as Grid, FractionalStep
    G = np.load(os.path.join(HERE, "golden", "ref_symmetry_sheared_10x8.npz"))
    g = Grid.from_cells(comm, G["xy"], G["cptr"], G["cind"])
    for pn in ("x-", "x+", "y-", "y+"):
        g.createPatchByNodes(pn, G["patch_" + pn])
    g.finalize()
    fs = FractionalStep(g, 1.0, 0.1)
    fs.u.setBoundary("x-", SYMMETRY, (0.0, 0.0)); fs.u.setBoundary("x+", SYMMETRY, (0.0, 0.0))
    fs.u.setBoundary("y-", FIXED, (0.0, 0.0)); fs.u.setBoundary("y+", FIXED, (1.0, 0.0))
    for pn in ("x-", "x+", "y-", "y+"):
        fs.p.setBoundary(pn, NORMAL_GRADIENT, 0.0)
    cfg = dict(solver="BICGSTAB", tolerance=1e-12, maxIters=20000, preconditioner=pc, amgCoarsest=30)
    fs.uEqn.solver.setup(cfg); fs.pEqn.solver.setup(cfg)
    fs.initialize()
    dt = float(G["dt"])
    for _ in range(int(G["K"])):
        st = fs.solve(dt)
        assert st["errorU"] <= 1e-10 and st["errorP"] <= 1e-10, st
    u, uf, p, pr = fs.u.get("cells"), fs.u.get("faces"), fs.p.get("cells"), G["field_p"]
    assert rel_l2(u[0], G["field_ux"]) < 1e-6 and rel_l2(u[1], G["field_uy"]) < 1e-6
    assert rel_l2(uf[0], G["field_ufx"]) < 1e-6 and rel_l2(uf[1], G["field_ufy"]) < 1e-6
    assert rel_l2(p - p.mean(), pr - pr.mean()) < 1e-6
    # uEqn_ of the last step in the reference's compact layout: cross-component entries at the end of the rows
    rp, ci, va, rhs = fs.uEqn.export(0)
    assert np.array_equal(rp, G["uEqn_rowPtr"]) and np.array_equal(ci, G["uEqn_colInd"])
    N = g.sizes()["nCells"]
    assert (ci[rp[0]:rp[1]] >= N).any()          # row (cell 0, x) reaches into the y block
    assert np.abs(va - G["uEqn_vals"]).max() <= 1e-9 * np.abs(va).max()
    assert np.abs(-rhs - G["uEqn_b"]).max() <= 1e-7 * np.abs(rhs).max()
    fs.close(); g.close()
