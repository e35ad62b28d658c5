"""phasePiso (README-era SIMPLE/PISO module): no reference implementation exists in
the snapshot, so the checks are self-consistency: the mass imbalance vanishes, the
run reaches a steady state, and that state agrees with the fractional-step
module's steady state of the same discrete operators."""
import numpy as np
import pytest

from tests.util import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def comm():
    from phase_b200.api import Communicator
    c = Communicator(0)
    yield c
    c.close()


def test_piso_cavity_steady_state(comm):
    from phase_b200.api import FiniteVolumeGrid2D as G, Piso, lid_driven_cavity, FIXED, NORMAL_GRADIENT
    n = 32
    g = G.rectilinear(comm, n, n, 1.0, 1.0)
    ps = Piso(g, 1.0, 0.1, numInnerIterations=1, numPressureCorrections=2, momentumRelaxation=0.8,
              pressureCorrectionRelaxation=0.3)
    for pt in ("x-", "x+", "y-"):
        ps.u.setBoundary(pt, FIXED, (0.0, 0.0))
    ps.u.setBoundary("y+", FIXED, (1.0, 0.0))
    for pt in ("x-", "x+", "y-", "y+"):
        ps.p.setBoundary(pt, NORMAL_GRADIENT, 0.0)
    cfg = dict(maxIters=5000, tolerance=1e-10, preconditioner="jacobi")
    ps.uSolver.setup(cfg); ps.pCorrSolver.setup(cfg)
    ps.initialize()
    hist = []
    uprev = None
    for k in range(400):
        st = ps.solve(0.05)
        hist.append(st["maxMassImbalance"])
    u = ps.u.get("cells")
    st = ps.solve(0.05)
    u2 = ps.u.get("cells")
    assert rel_l2(u2, u) < 1e-5                       # steady
    assert hist[-1] < 1e-6 and hist[-1] < 1e-3 * max(hist[:10])   # mass imbalance driven to zero
    # fractional step to its steady state (small dt: splitting error O(dt))
    fs = lid_driven_cavity(g, 1.0, 0.1, solver=dict(tolerance=1e-10, maxIters=5000))
    for k in range(1500):
        fs.solve(0.002)
    uf = fs.u.get("cells")
    assert rel_l2(u2, uf) < 3e-2
    # primary vortex: negative u_x near the bottom, positive under the lid
    ux = u2[0].reshape(n, n)
    assert ux[n - 2, n // 2] > 0.3 and ux[n // 4, n // 2] < 0.0
    ps.close(); fs.close(); g.close()


@pytest.mark.parametrize("kind,nx,ny,inner,corr", [("rect", 24, 20, 1, 2), ("tri", 10, 9, 2, 1)])
def test_piso_against_the_cpu_restatement(comm, kind, nx, ny, inner, corr):
    """K PISO steps of the device module against oracle/piso.py, an independent numpy / scipy transcription of the
    same equations (exact LU solves): u, p (minus mean), d, the mass imbalance."""
    import oracle as O
    from oracle.piso import Piso as OPiso
    from phase_b200.api import FiniteVolumeGrid2D as G, Piso, FIXED, NORMAL_GRADIENT
    om = (O.Mesh.rectilinear if kind == "rect" else O.Mesh.triangulated)(nx, ny, 1.0, 1.0)
    g = (G.rectilinear if kind == "rect" else G.triangulated)(comm, nx, ny, 1.0, 1.0)
    keys = dict(numInnerIterations=inner, numPressureCorrections=corr, momentumRelaxation=0.7, pressureCorrectionRelaxation=0.4)
    ps = Piso(g, 1.2, 0.05, **keys)
    ubc = {"x-": (FIXED, (0.0, 0.0)), "x+": (FIXED, (0.0, 0.0)), "y-": (FIXED, (0.0, 0.0)), "y+": (FIXED, (1.0, 0.0))}
    pbc = {pt: (NORMAL_GRADIENT, 0.0) for pt in ("x-", "x+", "y-", "y+")}
    for pt, (t, v) in ubc.items():
        ps.u.setBoundary(pt, t, v)
    for pt, (t, v) in pbc.items():
        ps.p.setBoundary(pt, t, v)
    cfg = dict(maxIters=20000, tolerance=1e-13, preconditioner="ilu0")
    ps.uSolver.setup(cfg); ps.pCorrSolver.setup(cfg)
    ps.initialize()
    op = OPiso(om, 1.2, 0.05, ubc, pbc, inner, corr, 0.7, 0.4)
    dt = 0.1
    for _ in range(6):
        st = ps.solve(dt)
        mi = op.step(dt)
    u, p = ps.u.get("cells"), ps.p.get("cells")
    assert rel_l2(u[0], op.u[0]) < 1e-6 and rel_l2(u[1], op.u[1]) < 1e-6
    assert rel_l2(p - p.mean(), op.p - op.p.mean()) < 1e-6
    assert rel_l2(ps.d.get("cells"), op.d) < 1e-9
    assert rel_l2(ps.u.get("faces"), op.uf) < 1e-6
    assert abs(st["maxMassImbalance"] - mi) < 1e-8
    ps.close(); g.close()
