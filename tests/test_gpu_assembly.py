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


@pytest.mark.parametrize("kind,nx,ny", [("rect", 33, 31), ("tri", 14, 12)])
def test_fused_assembly_equals_term_by_term(comm, kind, nx, ny):
    """phb_fs_assemble_u in one pass (k_momentum_fused) against the same equation assembled one operator at a time
    (fusedAssembly 0): identical pattern, coefficients and right-hand side to rounding."""
    om, ofs = oracle_cavity(kind, nx, ny, 1.0, 0.8)
    g, gfs = gpu_cavity(comm, kind, nx, ny, 1.0, 0.8)
    set_random_state(ofs, gfs, seed=11)
    dt = 0.007
    fused = gfs.assembleU(dt).export(0)
    gfs.setup(fusedAssembly=0)
    terms = gfs.assembleU(dt).export(0)
    gfs.setup(fusedAssembly=1)
    assert_eqn_equal(fused, terms, rtol=1e-14)
    assert np.abs(fused[3]).max() > 0
    assert_eqn_equal(fused, ofs.assemble_u(dt).export())
    fusedP = gfs.assembleP(dt).export(1)
    gfs.setup(fusedAssembly=0)
    termsP = gfs.assembleP(dt).export(1)
    gfs.setup(fusedAssembly=1)
    assert_eqn_equal(fusedP, termsP, rtol=1e-14)
    assert np.abs(fusedP[3]).max() > 0
    assert_eqn_equal(fusedP, ofs.assemble_p(dt).export())
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


def test_multiphase_momentum_equation(comm):
    """uEqn_ = (rho*fv::ddt(u,dt) + rho*fv::dive(u,u,0.5) == fv::laplacian(mu,u,0.5) + src::src(f)),
    US/FractionalStepMultiphase.cpp:111-112: ddt, dive (aliased history), field laplacian with theta,
    row scaling, vector source."""
    from phase_b200.api import FiniteVolumeField, FiniteVolumeEquation
    om, ofs = oracle_cavity("tri", 6, 7, 1.0, 1.2)
    g, gfs = gpu_cavity(comm, "tri", 6, 7, 1.0, 1.2)
    set_random_state(ofs, gfs, seed=9)
    rng = np.random.default_rng(10)
    N, F = om.sizes["nCells"], om.sizes["nFaces"]
    rho = rng.uniform(1.0, 900.0, N)
    mu, mu0 = rng.uniform(1e-3, 1.0, F), rng.uniform(1e-3, 1.0, F)
    fx, fy = rng.standard_normal(N), rng.standard_normal(N)
    dt = 0.004
    want = ofs.ueqn_multiphase(dt, rho, mu, mu0, fx, fy).export()
    rhoF, muF, fF = FiniteVolumeField(g, 1, "rho"), FiniteVolumeField(g, 1, "mu"), FiniteVolumeField(g, 2, "f")
    rhoF.set("cells", rho)
    muF.set("faces", mu0); muF.savePreviousTimeStep(); muF.set("faces", mu)
    fF.set("cells", np.concatenate([fx, fy]))
    eq = FiniteVolumeEquation(gfs.u).zero()
    eq.ddt(gfs.u, dt).dive(gfs.u, gfs.u, 0.5).scaleRows(rhoF)
    eq.laplacian(muF, gfs.u, 0.5, sign=-1.0).src(fF, sign=-1.0)
    assert_eqn_equal(eq.export(0), want, rtol=5e-13)
    for o in (eq, rhoF, muF, fF, gfs, g):
        o.close()


@pytest.mark.parametrize("theta", [1.0, 0.5])
def test_scalar_transport_with_density_field(comm, theta):
    """(fv::ddt(rho, phi, dt) + fv::div(u, phi, theta) == 0): implicit upwind matrix part,
    FIXED and NORMAL_GRADIENT boundary branches (UD/Divergence.h:27-45), rho/rho0 fields."""
    from phase_b200.api import FiniteVolumeGrid2D as G, FractionalStep, FiniteVolumeField, FiniteVolumeEquation, FIXED, NORMAL_GRADIENT
    om = O.Mesh.rectilinear(11, 6, 2.0, 1.0)
    ofs = O.FracStep(om, 1.0, 1.0)
    g = G.rectilinear(comm, 11, 6, 2.0, 1.0)
    gfs = FractionalStep(g, 1.0, 1.0)
    for pt, t, v in (("x-", FIXED, 1.0), ("x+", NORMAL_GRADIENT, 0.0), ("y-", NORMAL_GRADIENT, 0.0), ("y+", FIXED, 0.0)):
        ofs.set_bc("p", pt, t, v); gfs.p.setBoundary(pt, t, v)
    ofs.initialize(); gfs.initialize()
    set_random_state(ofs, gfs, seed=21)
    # the random state overwrote boundary faces of p on both sides identically: fine for parity
    rng = np.random.default_rng(22)
    N, F = om.sizes["nCells"], om.sizes["nFaces"]
    rho, rho0 = rng.uniform(1, 5, N), rng.uniform(1, 5, N)
    phi0, phi0f = rng.standard_normal(N), rng.standard_normal(F)
    dt = 0.01
    want = ofs.scalar_transport(dt, theta, rho, rho0, phi0, phi0f).export()
    rhoF = FiniteVolumeField(g, 1, "rho")
    rhoF.set("cells", rho0); rhoF.savePreviousTimeStep(); rhoF.set("cells", rho)
    pc, pf = gfs.p.get("cells").copy(), gfs.p.get("faces").copy()
    gfs.p.set("cells", phi0); gfs.p.set("faces", phi0f); gfs.p.savePreviousTimeStep()
    gfs.p.set("cells", pc); gfs.p.set("faces", pf)
    eq = FiniteVolumeEquation(gfs.p).zero().ddt(gfs.p, dt, rho=rhoF).div(gfs.u, gfs.p, theta)
    assert_eqn_equal(eq.export(0), want, rtol=5e-13)
    for o in (eq, rhoF, gfs, g):
        o.close()


def test_relax(comm):
    """relax(omega): a_PP /= omega; rhs_P -= (1-omega) a_PP phi_P (UE/ScalarFiniteVolumeEquation.cpp:45-55)."""
    om, ofs = oracle_cavity("rect", 7, 6)
    g, gfs = gpu_cavity(comm, "rect", 7, 6)
    set_random_state(ofs, gfs, seed=4)
    eq = gfs.assembleU(0.01)
    rp, ci, va, rhs = eq.export(0)
    u = gfs.u.get("cells").reshape(-1)
    eq.relax(0.8)
    rp2, ci2, va2, rhs2 = eq.export(0)
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2)
    diag = rp[:-1]
    va_want = va.copy(); va_want[diag] = va[diag] / 0.8
    assert np.allclose(va2, va_want, rtol=1e-15)
    assert np.allclose(rhs2, rhs - 0.2 * va_want[diag] * u, rtol=1e-14, atol=1e-14)
    gfs.close(); g.close()


@pytest.mark.parametrize("kind,nx,ny", [("rect", 14, 11), ("tri", 9, 8)])
def test_cicsam(comm, kind, nx, ny):
    """A9: cicsam::faceInterpolationWeights, cicsam::div and computeMomentumFlux (UD/Cicsam.cpp) vs the oracle:
    a diffuse circular interface advected by a rotating + random velocity field."""
    from phase_b200.api import (FiniteVolumeGrid2D as G, FractionalStep, FiniteVolumeField, FiniteVolumeEquation,
                                cicsam, FIXED, NORMAL_GRADIENT)
    om = (O.Mesh.rectilinear if kind == "rect" else O.Mesh.triangulated)(nx, ny, 1.0, 1.0)
    ofs = O.FracStep(om, 1.0, 1.0)
    g = (G.rectilinear if kind == "rect" else G.triangulated)(comm, nx, ny, 1.0, 1.0)
    gfs = FractionalStep(g, 1.0, 1.0)
    for pt, t, v in (("x-", FIXED, 0.0), ("x+", NORMAL_GRADIENT, 0.0), ("y-", NORMAL_GRADIENT, 0.0), ("y+", FIXED, 1.0)):
        ofs.set_bc("p", pt, t, v); gfs.p.setBoundary(pt, t, v)
    ofs.initialize(); gfs.initialize()
    rng = np.random.default_rng(31)
    N, F = om.sizes["nCells"], om.sizes["nFaces"]
    cx, cy, fx, fy = om.array("cellCx"), om.array("cellCy"), om.array("faceCx"), om.array("faceCy")
    gam = 0.5 * (1 + np.tanh((0.3 - np.hypot(cx - 0.5, cy - 0.45)) / 0.08)) + 0.05 * rng.standard_normal(N)
    gamf = 0.5 * (1 + np.tanh((0.3 - np.hypot(fx - 0.5, fy - 0.45)) / 0.08))
    ufx = -(fy - 0.5) + 0.1 * rng.standard_normal(F); ufy = (fx - 0.5) + 0.1 * rng.standard_normal(F)
    gx, gy = rng.standard_normal(N), rng.standard_normal(N)
    g0, g0f = gam + 0.02 * rng.standard_normal(N), gamf + 0.02 * rng.standard_normal(F)
    ofs.view("p")[:] = gam; ofs.view("pf")[:] = gamf; ofs.view("ufx")[:] = ufx; ofs.view("ufy")[:] = ufy
    dt = 0.02
    beta_o = ofs.cicsam_weights(dt, gx, gy)
    gfs.u.set("faces", np.concatenate([ufx, ufy]))
    gfs.p.set("cells", g0); gfs.p.set("faces", g0f); gfs.p.savePreviousTimeStep()
    gfs.p.set("cells", gam); gfs.p.set("faces", gamf)
    gradG, beta, rhoU = FiniteVolumeField(g, 2, "gradGamma"), FiniteVolumeField(g, 1, "beta"), FiniteVolumeField(g, 2, "rhoU")
    gradG.set("cells", np.concatenate([gx, gy]))
    cicsam.faceInterpolationWeights(gfs.u, gfs.p, gradG, dt, beta)
    b = beta.get("faces")
    assert b.min() >= 0.0 and b.max() <= 1.0 and (b > 0).sum() > 5
    assert np.allclose(b, beta_o, rtol=1e-9, atol=1e-12), np.abs(b - beta_o).max()
    want = ofs.cicsam_div(0.5, beta_o, g0, g0f).export()
    beta.set("faces", beta_o)            # identical weights on both sides for the operator comparison
    eq = FiniteVolumeEquation(gfs.p).zero().cicsamDiv(gfs.u, gfs.p, beta, 0.5)
    # cicsam::div alone leaves ELL-5 rows [donor/acceptor order]: compare as matrices (row-wise dict) + rhs
    rp, ci, va, rhs = eq.export(0)
    A = O.csr_to_scipy(rp, ci, va, N).toarray()
    Aw = O.csr_to_scipy(want[0], want[1], want[2], N).toarray()
    assert np.allclose(A, Aw, rtol=1e-13, atol=1e-15) and np.allclose(rhs, want[3], rtol=1e-12, atol=1e-14)
    ox, oy = ofs.cicsam_momentum_flux(998.0, 1.225, beta_o)
    cicsam.computeMomentumFlux(998.0, 1.225, gfs.u, gfs.p, beta, rhoU)
    rf = rhoU.get("faces")
    assert np.allclose(rf[0], ox, rtol=1e-13, atol=1e-13) and np.allclose(rf[1], oy, rtol=1e-13, atol=1e-13)
    for o in (eq, gradG, beta, rhoU, gfs, g):
        o.close()
