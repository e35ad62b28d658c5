"""Seam 1 through the C ABI: SpMV and BiCGStab against the oracle (scipy LU and
the C BiCGStab) on matrices in the reference's own hand-off layouts."""
import numpy as np
import pytest

import oracle as O
from tests.util import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def comm():
    from phase_b200.api import Communicator
    c = Communicator(0)
    yield c
    c.close()


def poisson_system(kind, nx, ny, fixed_top=True, seed=0):
    om = (O.Mesh.rectilinear if kind == "rect" else O.Mesh.triangulated)(nx, ny, 1.0, 1.0)
    fs = O.FracStep(om, 1.0, 1.0)
    if fixed_top:
        fs.set_bc("p", "y+", O.FIXED, 0.5)
    fs.initialize()
    rng = np.random.default_rng(seed)
    fs.view("ufx")[:] = rng.standard_normal(om.sizes["nFaces"])
    fs.view("ufy")[:] = rng.standard_normal(om.sizes["nFaces"])
    rp, ci, va, rhs = fs.assemble_p(0.05).export()       # padded [nb0,P,...,-1]
    return rp, ci, va, -rhs


@pytest.mark.parametrize("kind,nx,ny", [("rect", 20, 13), ("tri", 9, 11), ("rect", 1, 1), ("rect", 70, 1)])
def test_spmv_padded_csr(comm, kind, nx, ny):
    from phase_b200.api import SparseMatrixSolver
    rp, ci, va, b = poisson_system(kind, nx, ny)
    s = SparseMatrixSolver(comm)
    s.setRank(len(b))
    s.set(rp, ci, va)
    x = np.random.default_rng(1).standard_normal(len(b))
    y = s.spmv(x)
    ref = O.csr_to_scipy(rp, ci, va) @ x
    assert np.allclose(y, ref, rtol=1e-13, atol=1e-13)
    s.close()


@pytest.mark.parametrize("pc", ["none", "jacobi", "ilu0"])
@pytest.mark.parametrize("kind,nx,ny", [("rect", 40, 40), ("tri", 24, 20)])
def test_bicgstab_vs_direct(comm, kind, nx, ny, pc):
    from phase_b200.api import SparseMatrixSolver
    rp, ci, va, b = poisson_system(kind, nx, ny)
    xd = O.direct_solve(rp, ci, va, b)
    s = SparseMatrixSolver(comm).setup(dict(solver="BICGSTAB", maxIters=5000, tolerance=1e-11, preconditioner=pc))
    s.setRank(len(b))
    s.set(rp, ci, va)
    s.setRhs(b)
    err = s.solve()
    assert err <= 1e-11 and s.nIters() > 0
    x = s.x()
    A = O.csr_to_scipy(rp, ci, va)
    assert np.linalg.norm(b - A @ x) <= 1.0000001e-11 * np.linalg.norm(b) * 1.01
    assert rel_l2(x, xd) < 1e-8
    # same matrix again (pattern cache) with a new rhs and a warm start
    s.set(rp, ci, va * 1.0)
    s.setRhs(2.0 * b)
    s.setGuess(2.0 * x)
    s.solve()
    assert s.nIters() <= 2 and rel_l2(s.x(), 2 * xd) < 1e-8
    s.close()


def test_singular_neumann_system_converges(comm):
    from phase_b200.api import SparseMatrixSolver
    rp, ci, va, b = poisson_system("rect", 32, 32, fixed_top=False)
    b = b - b.mean()                      # compatible right-hand side
    s = SparseMatrixSolver(comm).setup(dict(maxIters=5000, tolerance=1e-10, preconditioner="jacobi"))
    s.set(rp, ci, va)
    s.setRhs(b)
    s.solve()
    x, xd = s.x(), O.direct_solve(rp, ci, va, b)
    assert rel_l2(x - x.mean(), xd - xd.mean()) < 1e-7
    s.close()


def test_coo_duplicates_summed_and_errors(comm):
    from phase_b200.api import SparseMatrixSolver, PhaseB200Error
    s = SparseMatrixSolver(comm).setup(dict(tolerance=1e-12, preconditioner="jacobi"))
    rows = [0, 0, 0, 1, 1, 2, 2, 2]
    cols = [0, 0, 1, 1, 0, 2, 1, 2]
    vals = [2.0, 2.0, -1.0, 4.0, -1.0, 2.0, -1.0, 2.0]   # duplicates (0,0) and (2,2) summed
    s.setEntries(3, rows, cols, vals)
    s.setRhs(np.array([1.0, 2.0, 3.0]))
    s.solve()
    A = np.array([[4.0, -1, 0], [-1, 4, 0], [0, -1, 4]])
    assert np.allclose(s.x(), np.linalg.solve(A, [1, 2, 3]), rtol=1e-10)
    with pytest.raises(PhaseB200Error):
        s.setup(dict(solver="GMRES"))
    with pytest.raises(PhaseB200Error):
        s.setRhs(np.zeros(5))
    s2 = SparseMatrixSolver(comm)
    with pytest.raises(PhaseB200Error):
        s2.solve()                        # nothing set
    s.close(); s2.close()


def test_matches_oracle_bicgstab_iteration_count(comm):
    """Same algorithm (right-preconditioned BiCGStab + Jacobi), same tolerance:
    iteration counts of the CUDA path and the C oracle agree to round-off effects."""
    from phase_b200.api import SparseMatrixSolver
    rp, ci, va, b = poisson_system("rect", 48, 48)
    xo, ito, rro = O.bicgstab(rp, ci, va, b, tol=1e-9, max_iters=5000, precond=1)
    s = SparseMatrixSolver(comm).setup(dict(maxIters=5000, tolerance=1e-9, preconditioner="jacobi"))
    s.set(rp, ci, va); s.setRhs(b); s.solve()
    assert abs(s.nIters() - ito) <= max(8, ito // 4)
    assert rel_l2(s.x(), xo) < 1e-6
    s.close()


def test_ilu0_levels_is_exact_on_a_chain(comm):
    """On a 1-D chain the pattern has no fill, so ILU(0) in the natural (wavefront-level) ordering is the
    exact LU: BiCGStab must converge in one iteration."""
    from phase_b200.api import SparseMatrixSolver
    rp, ci, va, b = poisson_system("rect", 70, 1)
    xd = O.direct_solve(rp, ci, va, b)
    s = SparseMatrixSolver(comm).setup(dict(maxIters=50, tolerance=1e-12, preconditioner="ilu0", ordering="levels"))
    s.set(rp, ci, va); s.setRhs(b); s.solve()
    assert s.nIters() <= 2 and rel_l2(s.x(), xd) < 1e-10
    s.close()


@pytest.mark.parametrize("kind,nx,ny", [("rect", 64, 64), ("tri", 40, 40)])
def test_ilu0_reduces_iterations(comm, kind, nx, ny):
    from phase_b200.api import SparseMatrixSolver
    rp, ci, va, b = poisson_system(kind, nx, ny)
    xd = O.direct_solve(rp, ci, va, b)
    its = {}
    for name, cfg in (("jacobi", dict(preconditioner="jacobi")), ("ilu0_mc", dict(preconditioner="ilu0")),
                      ("ilu0_lv", dict(preconditioner="ilu0", ordering="levels"))):
        s = SparseMatrixSolver(comm).setup(dict(maxIters=5000, tolerance=1e-10, **cfg))
        s.set(rp, ci, va); s.setRhs(b); s.solve()
        assert s.error() <= 1e-10 and rel_l2(s.x(), xd) < 1e-7, name
        its[name] = s.nIters()
        s.close()
    # same algorithm on the CPU oracle (natural-order ILU(0)) for the iteration count
    xo, ito, rro = O.bicgstab(rp, ci, va, b, tol=1e-10, max_iters=5000, precond=2)
    assert its["ilu0_mc"] < its["jacobi"]
    assert its["ilu0_lv"] <= its["ilu0_mc"]
    assert abs(its["ilu0_lv"] - ito) <= max(6, ito // 3), (its, ito)


def test_variable_coefficient_iteration_counts_match_oracle(comm):
    """Density-ratio-815 Poisson (config 4's pEqn_): same algorithm on the CPU oracle and the CUDA path
    (BiCGStab + multicolour ILU(0), and Jacobi) -> same solution, comparable iteration counts."""
    from phase_b200.api import SparseMatrixSolver
    om = O.Mesh.rectilinear(96, 96, 1.0, 1.0)
    fs = O.FracStep(om, 1.0, 1.0)
    fs.set_bc("p", "y+", O.FIXED, 0.0)
    fs.initialize()
    rng = np.random.default_rng(3)
    F = om.sizes["nFaces"]
    fs.view("ufx")[:] = rng.standard_normal(F); fs.view("ufy")[:] = rng.standard_normal(F)
    fx, fy = om.array("faceCx"), om.array("faceCy")
    alpha = 0.5 * (1 + np.tanh((0.25 - np.hypot(fx - 0.5, fy - 0.5)) / 0.04))
    rp, ci, va, rhs = fs.laplacian_field(1e-3 / (998.0 + alpha * (1.225 - 998.0))).export()
    b = -rhs
    xd = O.direct_solve(rp, ci, va, b)
    xo, ito, _ = O.bicgstab_ilu0_multicolor(rp, ci, va, b, tol=1e-10, max_iters=20000)
    s = SparseMatrixSolver(comm).setup(dict(maxIters=20000, tolerance=1e-10, preconditioner="ilu0"))
    s.set(rp, ci, va); s.setRhs(b); s.solve()
    assert s.error() <= 1e-10 and rel_l2(s.x(), xd) < 1e-6 and rel_l2(xo, xd) < 1e-6
    assert abs(s.nIters() - ito) <= max(10, ito // 2), (s.nIters(), ito)
    s.close()


# ------------------------------------------------------------------ smoothed-aggregation AMG (`preconditioner amg`)
@pytest.mark.parametrize("prec", ["single", "double"])
@pytest.mark.parametrize("kind,nx,ny,coarsest", [("rect", 48, 40, 40), ("tri", 30, 26, 40), ("rect", 12, 9, 400)])
def test_amg_vs_direct(comm, kind, nx, ny, coarsest, prec):
    from phase_b200.api import SparseMatrixSolver
    rp, ci, va, b = poisson_system(kind, nx, ny)
    xd = O.direct_solve(rp, ci, va, b)
    s = SparseMatrixSolver(comm).setup(dict(solver="BICGSTAB", maxIters=200, tolerance=1e-11, preconditioner="amg",
                                            amgCoarsest=coarsest, amgPrecision=prec))
    s.setRank(len(b))
    s.set(rp, ci, va)
    s.setRhs(b)
    err = s.solve()
    info = s.amgInfo()
    assert err <= 1e-11 and 0 < s.nIters() <= 30, (s.nIters(), info)
    assert info["levels"] >= (3 if coarsest == 40 else 1) and info["setups"] == 1
    assert rel_l2(s.x(), xd) < 1e-8
    # scaled matrix (a time-step change): hierarchy kept, still converges; new values: kept until iterations degrade
    s.set(rp, ci, 3.0 * va)
    s.setRhs(3.0 * b)
    s.solve()
    assert s.amgInfo()["setups"] == 1 and s.amgInfo()["stale"] == 0 and rel_l2(s.x(), xd) < 1e-8
    s.close()


def test_amg_fewer_iterations_than_ilu0(comm):
    from phase_b200.api import SparseMatrixSolver
    rp, ci, va, b = poisson_system("rect", 128, 128)
    its = {}
    for pc in ("ilu0", "amg"):
        s = SparseMatrixSolver(comm).setup(dict(maxIters=5000, tolerance=1e-9, preconditioner=pc))
        s.set(rp, ci, va)
        s.setRhs(b)
        s.solve()
        its[pc] = s.nIters()
        x = s.x()
        A = O.csr_to_scipy(rp, ci, va)
        assert np.linalg.norm(b - A @ x) <= 1.01e-9 * np.linalg.norm(b)
        s.close()
    assert its["amg"] <= 20 and its["amg"] * 4 < its["ilu0"], its


def test_amg_singular_neumann_system(comm):
    from phase_b200.api import SparseMatrixSolver
    rp, ci, va, b = poisson_system("rect", 64, 48, fixed_top=False)
    s = SparseMatrixSolver(comm).setup(dict(maxIters=200, tolerance=1e-10, preconditioner="amg", amgCoarsest=30,
                                            nullSpace="constant"))
    s.set(rp, ci, va)
    s.setRhs(b)                                  # 1^T b removed by the solver
    s.solve()
    assert s.nIters() <= 30
    bc = b - b.mean()
    x, xd = s.x(), O.direct_solve(rp, ci, va, bc)
    assert rel_l2(x - x.mean(), xd - xd.mean()) < 1e-7
    s.close()


def _upwind_like_system(n, shift, seed):
    """n x n system, every row [diagonal, i - shift, i + shift + 1] (wrapped): changing `shift` moves the
    columns while every row keeps exactly three entries -- the slot count of the sliced-ELL image is unchanged.
    Structurally NON-symmetric, as the upwind coefficients of a host-assembled CrsEquation are."""
    rng = np.random.default_rng(seed)
    rp = np.arange(0, 3 * n + 1, 3, dtype=np.int32)
    ci = np.empty(3 * n, np.int32)
    va = np.empty(3 * n)
    i = np.arange(n)
    ci[0::3], ci[1::3], ci[2::3] = i, (i - shift) % n, (i + shift + 1) % n
    va[1::3], va[2::3] = -rng.uniform(0.5, 1.0, n), -rng.uniform(0.1, 0.5, n)
    va[0::3] = 0.2 - va[1::3] - va[2::3]
    return rp, ci, va, rng.standard_normal(n)


@pytest.mark.parametrize("pc", ["ilu0", "amg", "jacobi"])
def test_pattern_change_at_constant_row_lengths(comm, pc):
    """A solver that lives as long as its equation (M/CrsEquation.cpp:16-26) sees the column pattern change
    between solves while row lengths -- and the padded slot count -- stay the same (IB stencils; `+=` dropping
    exact zeros).  Every cached analysis (ILU ordering and slot map, multigrid hierarchy, CUDA graph) has to go."""
    from phase_b200.api import SparseMatrixSolver
    n = 3000
    s = SparseMatrixSolver(comm).setup(dict(solver="BICGSTAB", maxIters=5000, tolerance=1e-11, preconditioner=pc,
                                            amgCoarsest=60))
    for shift, seed in ((1, 0), (7, 1), (1, 2), (40, 3)):
        rp, ci, va, b = _upwind_like_system(n, shift, seed)
        s.setRank(n); s.set(rp, ci, va); s.setRhs(b)
        err = s.solve()
        assert err <= 1e-11, (shift, err)
        xd = O.direct_solve(rp, ci, va, b)
        assert rel_l2(s.x(), xd) < 1e-8, (pc, shift)
    s.close()


def test_ilu0_refuses_rows_beyond_64_entries(comm):
    from phase_b200.api import PhaseB200Error, SparseMatrixSolver
    n, w = 200, 70
    rp = np.arange(0, w * n + 1, w, dtype=np.int32)
    ci = ((np.arange(n)[:, None] + np.arange(w)[None, :]) % n).astype(np.int32).reshape(-1)
    va = np.where(np.tile(np.arange(w), n) == 0, 100.0, -1.0)
    s = SparseMatrixSolver(comm).setup(dict(preconditioner="ilu0"))
    s.setRank(n); s.set(rp, ci, va); s.setRhs(np.ones(n))
    with pytest.raises(PhaseB200Error):
        s.solve()
    s.setup(dict(preconditioner="jacobi"))
    assert s.solve() <= 1e-8
    s.close()
