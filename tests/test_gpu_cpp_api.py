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
