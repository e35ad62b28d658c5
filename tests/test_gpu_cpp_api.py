"""The C++ mirror of the Phase API (include/phase): Seam 1 driven exactly like the
reference drives a backend, and the FractionalStep module written with the
reference's own statements, compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle as O
from tests.util import rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exes():
    from examples import build as eb
    return dict(zip(eb.TARGETS, eb.build()))


def test_seam1_crs_equation_solve(exes):
    r = subprocess.run([exes["seam1_solver"], "40"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "refused: SparseMatrixSolverFactory::create -> bad solver type \"eigen\"." in r.stdout
    assert "iterations =" in r.stdout


def test_fractional_step_module_matches_oracle(exes, tmp_path):
    out = tmp_path / "fields.bin"
    K = 10
    r = subprocess.run([exes["lid_driven_cavity"], os.path.join(ROOT, "examples", "LidDrivenCavity", "case"), str(K), str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(out, dtype=np.float64).reshape(3, -1)
    om = O.Mesh.rectilinear(40, 40, 1.0, 1.0)
    ofs = O.cavity(om, 1.0, 0.1)
    ofs.use_direct_solver()
    for _ in range(K):
        ofs.step(5e-3)
    assert rel_l2(got[0], ofs.view("ux")) < 1e-6
    assert rel_l2(got[1], ofs.view("uy")) < 1e-6
    p, po = got[2], ofs.view("p").copy()
    assert rel_l2(p - p.mean(), po - po.mean()) < 1e-6
    assert "FiniteVolumeEquation pEqn: Krylov iterations =" in r.stdout


def test_phase_piso_legacy_case(exes, tmp_path):
    """Config 1: the shipped cavity case in its legacy PISO format (timeStep 10, relaxation 0.8 / 0.2)."""
    out = tmp_path / "piso.bin"
    r = subprocess.run([exes["phase_piso"], os.path.join(ROOT, "examples", "LidDrivenCavityPiso", "case"), "100", str(out)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    a = np.fromfile(out, dtype=np.float64)
    n = 100 * 100
    ux = a[:n].reshape(100, 100)
    assert np.isfinite(a).all() and ux[98, 50] > 0.4 and ux[30, 50] < 0.0       # primary vortex under the lid
    assert "final mass imbalance" in r.stdout


def test_equation_ops_cell_groups_and_tensor_entries(exes, tmp_path):
    """fv::ddt(field, dt, cells), src::div(field, cells), src::laplacian and the per-entry Vector2D / Tensor2D
    coefficients through the C++ mirror: the operator expression bit for bit against the same expression through the
    Python operator interface (itself pinned on the reference in test_gpu_source_ops.py), the per-entry vector
    equation against a dense solve of the matrix the reference's add() semantics define."""
    from phase_b200.api import (Communicator, FiniteVolumeEquation, FiniteVolumeGrid2D as G, ScalarFiniteVolumeField,
                                VectorFiniteVolumeField)
    out = tmp_path / "ops.bin"
    r = subprocess.run([exes["equation_ops"], str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    a = np.fromfile(out, dtype=np.float64)
    n, nnz = int(a[0]), int(a[1])
    o = 2
    rp = a[o:o + n + 1].astype(np.int64); o += n + 1
    ci = a[o:o + nnz].astype(np.int64); o += nnz
    va = a[o:o + nnz]; o += nnz
    rhs = a[o:o + n]; o += n
    ux, uy = a[o:o + n], a[o + n:o + 2 * n]
    nx, ny, w, h = 12, 9, 1.2, 0.9
    assert n == nx * ny
    om = O.Mesh.rectilinear(nx, ny, w, h)
    cx, cy = om.array("cellCx"), om.array("cellCy")
    comm = Communicator(0)
    try:
        g = G.rectilinear(comm, nx, ny, w, h)
        u, p, phi = VectorFiniteVolumeField(g, "u"), ScalarFiniteVolumeField(g, "p"), ScalarFiniteVolumeField(g, "phi")
        u.set("cells", np.concatenate([np.sin(3. * cx) + cy, np.cos(2. * cy) * cx]))
        p.set("cells", cx * cx + 0.5 * cy)
        phi.set("cells", np.cos(cx + 2. * cy))
        u.interpolateFaces(); p.setBoundaryFaces(); phi.savePreviousTimeStep()
        cells = np.arange(0, n, 3, dtype=np.int32)
        e = FiniteVolumeEquation(phi)
        e.zero().ddtCells(phi, 0.01, cells).srcDivCells(u, cells, sign=-1.0).srcLaplacian(0.7, p, sign=1.0)
        rp2, ci2, va2, rhs2 = e.export(0)
        assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2)
        assert np.array_equal(va, va2) and np.array_equal(rhs, rhs2)
        assert np.abs(rhs).max() > 0 and np.count_nonzero(va) == len(cells)
    finally:
        comm.close()
    # the per-entry vector equation: rows (x block, y block), A x + rhs = 0
    A = np.zeros((2 * n, 2 * n)); b = np.zeros(2 * n)
    for c in range(n):
        A[c, c] += 6. + cx[c]; A[c, n + c] += 0.25 * cy[c]; A[n + c, c] += -0.5 if c % 2 else 0.; A[n + c, n + c] += 7. - cy[c]
        for nb in (c - 12, c + 12):
            if 0 <= nb < n:
                A[c, nb] += -0.75; A[n + c, n + nb] += -0.75
        if c % 12:
            A[c, c - 1] += -1.; A[n + c, n + c - 1] += -1.25
        b[c] += -np.sin(cx[c]); b[n + c] += 1. + cy[c]
    x = np.linalg.solve(A, -b)
    assert rel_l2(ux, x[:n]) < 1e-10 and rel_l2(uy, x[n:]) < 1e-10
    assert "get(17,17) = %.17g %.17g   get(17,16) = %.17g %.17g" % (A[17, 17], A[n + 17, n + 17], -1., -1.25) in r.stdout
