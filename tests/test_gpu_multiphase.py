"""FractionalStepMultiphase on the device (CICSAM advection, density / viscosity blend, gravity and CELESTE surface
tension, rho-weighted reconstructions, variable-coefficient pressure equation) against the REFERENCE'S OWN
implementation: golden vectors written by oracle/_ref/libphase_ref_fv.so (tests/golden/make_ref_golden.py) and,
where that library travelled, a live run of it on another mesh.  Fields after K time steps within rel-L2 1e-6."""
import os

import numpy as np
import pytest

from tests import mp_util as U
from tests.util import rel_l2

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def comm():
    from phase_b200.api import Communicator
    c = Communicator(0)
    yield c
    c.close()


def run_device(comm, nx, ny, w, h, radius, gamma_cells, K, dt):
    from phase_b200.api import FiniteVolumeGrid2D as G
    g = G.rectilinear(comm, nx, ny, w, h)
    mp = U.device_solver(g, radius)
    mp.gamma.set("cells", gamma_cells)
    mp.gamma.interpolateFaces()
    gamma_faces = mp.gamma.get("faces").copy()
    mp.initialize()
    init = {k: getattr(mp, k).get("cells").copy() for k in ("rho", "mu", "kappa", "gammaTilde")}
    init.update({k: getattr(mp, k).get("cells").copy() for k in ("n", "fst", "sg")})
    st = None
    for _ in range(K):
        st = mp.solve(dt)
    out = {k: getattr(mp, k).get("cells").copy() for k in U.SCALARS + U.VECTORS}
    out["uf"] = mp.u.get("faces").copy()
    mp.close(); g.close()
    return gamma_faces, init, out, st


def compare(out, ref, tol=1e-6):
    for k in U.SCALARS:
        scale = max(np.linalg.norm(ref[k]), 1e-300)
        assert np.linalg.norm(out[k] - ref[k]) / scale < tol, k
    for k in U.VECTORS:
        for c in (0, 1):
            scale = max(np.linalg.norm(ref[k][0]), np.linalg.norm(ref[k][1]), 1e-300)
            assert np.linalg.norm(out[k][c] - ref[k][c]) / scale < tol, (k, c)


def test_multiphase_step_against_reference_golden(comm):
    G = np.load(os.path.join(HERE, "golden", "ref_bubble_24x48.npz"))
    nx, ny, w, h, K, dt, radius = int(G["nx"]), int(G["ny"]), float(G["w"]), float(G["h"]), int(G["K"]), float(G["dt"]), float(G["radius"])
    gamma_faces, init, out, st = run_device(comm, nx, ny, w, h, radius, G["gamma0"], K, dt)
    assert np.allclose(gamma_faces, G["gamma0_faces"], rtol=0, atol=1e-14)
    for k in ("rho", "mu", "kappa", "gammaTilde"):
        assert rel_l2(init[k], G["init_" + k]) < 1e-8, k
    for k in ("n", "fst", "sg"):
        assert rel_l2(init[k], np.stack([G["init_" + k + "_x"], G["init_" + k + "_y"]])) < 1e-8, k
    ref = {k: G["field_" + k] for k in U.SCALARS}
    ref.update({k: np.stack([G["field_" + k + "_x"], G["field_" + k + "_y"]]) for k in U.VECTORS})
    compare(out, ref)
    assert rel_l2(out["uf"], np.stack([G["field_uf_x"], G["field_uf_y"]])) < 1e-6
    assert st["maxDivergence"] < 1e-9 and abs(st["maxCourant"] - float(G["maxCourant"])) < 1e-8
    assert np.abs(out["kappa"]).max() > 1.0 and np.abs(out["u"]).max() > 1e-3      # the sources are alive


def test_multiphase_step_against_live_reference(comm):
    from oracle import ref_fv as R
    if not R.available():
        pytest.skip("reference FV library not available")
    nx, ny, w, h, K, dt = 18, 30, 0.6, 1.0, 3, 5e-4
    radius = 2.1 * w / nx
    case = U.reference_case(R, nx, ny, w, h, radius)
    rg = R.Grid.rectilinear(case)
    gamma0 = U.initial_gamma(rg.array("cellCx"), rg.array("cellCy"), w, h)
    gamma_faces, init, out, st = run_device(comm, nx, ny, w, h, radius, gamma0, K, dt)
    R.use_direct_solver()
    fs = R.Multiphase(case, rg)
    fs.set_field("gamma", gamma0)
    fs.set_field("gamma", gamma_faces, faces=True)
    fs.initialize()
    for _ in range(K):
        fs.step(dt)
    ref = {k: fs.field(k) for k in U.SCALARS}
    ref.update({k: np.stack([fs.field(U.REF_NAME.get(k, k), 0), fs.field(U.REF_NAME.get(k, k), 1)]) for k in U.VECTORS})
    compare(out, ref)
    fs.close(); rg.close(); case.close()
