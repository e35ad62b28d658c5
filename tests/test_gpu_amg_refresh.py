"""Numeric re-setup of the multigrid hierarchy on the device (amg_refresh.cuh): with unchanged coefficients it has to
reproduce the host setup's level matrices, smoother weights and dense coarsest inverse; with coefficients changed on the
same pattern (variable-density pressure operator, US/FractionalStepMultiphase.cpp:129-148) the refreshed hierarchy has
to do what a fresh host setup does."""
import numpy as np
import pytest

from phase_b200.synthetic import beta_field, variable_laplacian
from tests.util import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def comm():
    from phase_b200.api import Communicator
    c = Communicator(0)
    yield c
    c.close()


def make_solver(comm, **keys):
    from phase_b200.api import SparseMatrixSolver
    d = dict(solver="BICGSTAB", maxIters=400, tolerance=1e-10, preconditioner="amg", amgCoarsest=300)
    d.update(keys)
    return SparseMatrixSolver(comm).setup(d)


def hierarchy_values(s):
    nl = int(s.amgInfo()["levels"])
    out = {}
    for l in range(nl):
        out[(l, 3)] = s.amgValues(l, 3)
        if l > 0:
            out[(l, 0)] = s.amgValues(l, 0)
        if l + 1 < nl:
            out[(l, 1)] = s.amgValues(l, 1); out[(l, 2)] = s.amgValues(l, 2)
    out["inv"] = s.amgValues(0, 4)
    return nl, out


@pytest.mark.parametrize("neumann", [True, False])
@pytest.mark.parametrize("precision,tol", [("double", 1e-12), ("single", 2e-6)])
def test_refresh_reproduces_host_setup(comm, neumann, precision, tol):
    nx, ny = 96, 70
    A = variable_laplacian(nx, ny, beta_field(nx, ny, 0.4, 100.0), neumann)
    b = np.random.default_rng(0).standard_normal(nx * ny)
    if neumann:
        b -= b.mean()
    s = make_solver(comm, amgPrecision=precision, nullSpace="constant" if neumann else "none")
    s.setRank(A.shape[0]); s.set(A.indptr, A.indices, A.data); s.setRhs(b)
    assert s.solve() <= 1e-10
    it0 = s.nIters()
    nl, ref = hierarchy_values(s)
    assert nl >= 3 and s.amgRefreshInfo()["resident"] == 1.0
    s.amgRefresh()
    _, got = hierarchy_values(s)
    for k, v in ref.items():
        scale = np.abs(v).max()
        t = 1e-7 if k == "inv" else tol          # the inverse: Gauss-Jordan here, LU on the host
        assert np.abs(got[k] - v).max() <= t * scale, (k, np.abs(got[k] - v).max() / scale)
    s.setRhs(b)
    assert s.solve() <= 1e-10 and s.nIters() <= it0 + 1
    info = s.amgInfo()
    assert info["setups"] == 1.0 and s.amgRefreshInfo()["refreshes"] == 1.0
    s.close()


@pytest.mark.parametrize("mode", ["always", "auto"])
def test_refresh_follows_moving_interface(comm, mode):
    """The bubble (density ratio 1000) crosses the domain in 12 steps: the refreshed hierarchy needs at most 1.2 x (+2)
    the iterations of a hierarchy set up from scratch for each matrix, without any further host setup."""
    nx, ny = 160, 120
    n = nx * ny
    rng = np.random.default_rng(1)
    s = make_solver(comm, amgRefresh=mode, nullSpace="constant")
    iters, fresh = [], []
    for step in range(12):
        A = variable_laplacian(nx, ny, beta_field(nx, ny, 0.25 + 0.045 * step, 1000.0))
        b = rng.standard_normal(n); b -= b.mean()
        s.setRank(n); s.set(A.indptr, A.indices, A.data); s.setRhs(b)
        assert s.solve() <= 1e-10
        iters.append(s.nIters())
        x = s.x().copy()
        f = make_solver(comm, amgRefresh="off", nullSpace="constant")
        f.setRank(n); f.set(A.indptr, A.indices, A.data); f.setRhs(b)
        assert f.solve() <= 1e-10
        fresh.append(f.nIters())
        xf = f.x().copy()
        f.close()
        assert rel_l2(x - x.mean(), xf - xf.mean()) < 1e-6
    info, rinfo = s.amgInfo(), s.amgRefreshInfo()
    print("refreshed", iters, "fresh", fresh, info, rinfo)
    assert info["setups"] == 1.0
    assert rinfo["refreshes"] >= (11 if mode == "always" else 1)
    if mode == "always":
        assert all(a <= 1.2 * b + 2 for a, b in zip(iters, fresh)), (iters, fresh)
    else:
        # worst case of `auto`: a stale hierarchy is given up inside a solve after max(2 f, f + 10) iterations, the solve
        # restarts from the current iterate with the refreshed one (<= 1.2 f + 2, the bound of `always`)
        f = max(fresh)
        assert max(iters) <= max(2 * f, f + 10) + 1.2 * f + 2, (iters, fresh)
    s.close()


def test_refresh_off_keeps_old_behaviour(comm):
    nx, ny = 64, 48
    A = variable_laplacian(nx, ny, beta_field(nx, ny, 0.4, 10.0), False)
    s = make_solver(comm, amgRefresh="off")
    s.setRank(A.shape[0]); s.set(A.indptr, A.indices, A.data); s.setRhs(np.ones(A.shape[0]))
    assert s.solve() <= 1e-10
    assert s.amgRefreshInfo()["resident"] == 0.0
    from phase_b200.api import PhaseB200Error
    with pytest.raises(PhaseB200Error):
        s.amgRefresh()
    s.close()
