"""src::laplacian and the cell-group overloads of fv::ddt / src::div (UD/Source.cpp:5-48, UD/TimeDerivative.h:50-62)
on the device against the reference's own functions (oracle/_ref/libphase_ref_fv.so, when it travelled) and a
numpy restatement of the same sums."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def comm():
    from phase_b200.api import Communicator
    c = Communicator(0)
    yield c
    c.close()


def _state(N, F, seed=4):
    rng = np.random.default_rng(seed)
    return dict(p=rng.standard_normal(N), pf=rng.standard_normal(F), ufx=rng.standard_normal(F), ufy=rng.standard_normal(F),
                gam=rng.uniform(0.5, 2.0, F))


@pytest.mark.parametrize("kind,nx,ny", [("rect", 9, 7), ("tri", 6, 5)])
def test_source_operators(comm, kind, nx, ny):
    from phase_b200.api import (FIXED, FiniteVolumeEquation, FiniteVolumeGrid2D as G, ScalarFiniteVolumeField,
                                VectorFiniteVolumeField)
    w, h = 1.3, 0.9
    om = (O.Mesh.rectilinear if kind == "rect" else O.Mesh.triangulated)(nx, ny, w, h)
    g = (G.rectilinear if kind == "rect" else G.triangulated)(comm, nx, ny, w, h)
    N, F = om.sizes["nCells"], om.sizes["nFaces"]
    st = _state(N, F)
    p, u, gam = ScalarFiniteVolumeField(g, "p"), VectorFiniteVolumeField(g, "u"), ScalarFiniteVolumeField(g, "gam")
    p.set("cells", st["p"]); p.set("faces", st["pf"])
    u.set("faces", np.concatenate([st["ufx"], st["ufy"]]))
    gam.set("faces", st["gam"])
    cells = np.arange(0, N, 3, dtype=np.int32)
    # ---- numpy restatement over the oracle's link tables
    ilPtr, ilFace, ilCell = om.array("ilPtr"), om.array("ilFace"), om.array("ilCell")
    blPtr, blFace = om.array("blPtr"), om.array("blFace")
    rc = np.stack([om.array("ilRcx"), om.array("ilRcy")]); sl = np.stack([om.array("ilSx"), om.array("ilSy")])
    rf = np.stack([om.array("blRfx"), om.array("blRfy")]); sb = np.stack([om.array("blSx"), om.array("blSy")])
    gi = (rc * sl).sum(0) / (rc ** 2).sum(0)
    gb = (rf * sb).sum(0) / (rf ** 2).sum(0)
    lap_s, lap_f, div = np.zeros(N), np.zeros(N), np.zeros(N)
    for c in range(N):
        for j in range(ilPtr[c], ilPtr[c + 1]):
            d = st["p"][ilCell[j]] - st["p"][c]
            lap_s[c] += d * 0.7 * gi[j]; lap_f[c] += d * st["gam"][ilFace[j]] * gi[j]
            div[c] += st["ufx"][ilFace[j]] * sl[0, j] + st["ufy"][ilFace[j]] * sl[1, j]
        for j in range(blPtr[c], blPtr[c + 1]):
            d = st["pf"][blFace[j]] - st["p"][c]
            lap_s[c] += d * 0.7 * gb[j]; lap_f[c] += d * st["gam"][blFace[j]] * gb[j]
            div[c] += st["ufx"][blFace[j]] * sb[0, j] + st["ufy"][blFace[j]] * sb[1, j]
    mask = np.zeros(N, bool); mask[cells] = True
    e = FiniteVolumeEquation(p)
    rhs = lambda: e.export(0)[3]
    close = lambda a, b: np.abs(a - b).max() <= 1e-12 * max(np.abs(b).max(), 1.0)
    e.zero().srcLaplacian(0.7, p)
    got_lap = rhs()
    assert close(got_lap, lap_s)
    e.zero().srcLaplacian(gam, p, sign=-1.0)
    assert close(rhs(), -lap_f)
    e.zero().srcDivCells(u, cells)
    got_div = rhs()
    assert close(got_div, np.where(mask, div, 0.0))
    dt = 0.05
    p.savePreviousTimeStep()
    e.zero().ddtCells(p, dt, cells)
    rp, ci, va, r = e.export(0)
    vol = om.array("vol")
    diag = np.zeros(N)
    for c in range(N):
        for k in range(rp[c], rp[c + 1]):
            if ci[k] == c:
                diag[c] += va[k]
    assert close(diag, np.where(mask, vol / dt, 0.0)) and close(r, np.where(mask, -vol * st["p"] / dt, 0.0))
    # ---- the reference's own functions
    from oracle import ref_fv as R
    if R.available() and kind == "rect":
        case = R.Case(nx, ny, w, h)
        rg = R.Grid.rectilinear(case)
        R.use_null_solver()
        fs = R.FracStep(case, rg)
        for k in ("p", "pf", "ufx", "ufy"):
            fs.set(k, st[k])
        assert close(got_lap, fs.src_laplacian(0.7))
        assert close(got_div, fs.src_div_cells(cells))
        d_ref, r_ref = fs.ddt_cells(dt, cells)
        assert close(diag, d_ref) and close(r, r_ref)
        fs.close(); rg.close(); case.close()
    for x in (e, p, u, gam):
        x.close()
    g.close()
