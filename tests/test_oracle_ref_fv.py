"""The oracle restatement against the REFERENCE'S OWN finite-volume code compiled in place
(oracle/_ref/libphase_ref_fv.so: FiniteVolumeGrid2D, fields, fv::/src:: operators, FiniteVolumeEquation::solve,
FractionalStep::solve from /root/reference/src over the stand-in headers of oracle/ref_stub).

Integer artefacts (connectivity, links, CSR patterns at the solver hand-off) must be bit-exact; coefficients,
right-hand sides and fields agree to round-off (the polygon area / centroid arithmetic behind cell volumes is
the stand-in's, every other operation is the reference's own statement order)."""
import numpy as np
import pytest

import oracle as O
from oracle import ref_fv as R

pytestmark = pytest.mark.skipif(not R.available(), reason="reference FV library not built (no /root/reference, no prebuilt .so)")

INT_KEYS = ["cptr", "cind", "faceN1", "faceN2", "faceL", "faceR", "ilPtr", "ilFace", "ilCell", "blPtr", "blFace", "dlPtr", "dlCell"]
F64_KEYS = ["nodeX", "nodeY", "vol", "cellCx", "cellCy", "faceCx", "faceCy", "faceNx", "faceNy", "ilRcx", "ilRcy", "ilSx",
            "ilSy", "blRfx", "blRfy", "blSx", "blSy"]
PATCHES = ("x-", "x+", "y-", "y+")


def ref_grid_like(om, case):
    """The reference grid with the oracle mesh's nodes / cells / patches (generic FiniteVolumeGrid2D path)."""
    xy = np.stack([om.array("nodeX"), om.array("nodeY")], axis=1)
    g = R.Grid.create(xy, om.array("cptr"), om.array("cind"))
    fp, n1, n2 = om.array("facePatch"), om.array("faceN1"), om.array("faceN2")
    for name in PATCHES:
        pid = om.patch_id(name)
        sel = np.flatnonzero(fp == pid)
        g.patch_by_nodes(name, np.stack([n1[sel], n2[sel]], axis=1).reshape(-1))
    return g


def check_mesh(g, om):
    for k in INT_KEYS:
        a, b = g.array(k), om.array(k)
        assert a.shape == b.shape and np.array_equal(a, b), k
    for k in F64_KEYS:
        a, b = g.array(k), om.array(k)
        assert a.shape == b.shape and np.allclose(a, b, rtol=1e-13, atol=1e-15), k
    fp = om.array("facePatch")
    for name in PATCHES:
        assert np.array_equal(np.sort(g.array("patch:" + name)), np.flatnonzero(fp == om.patch_id(name))), name


@pytest.mark.parametrize("nx,ny,w,h", [(3, 3, 1.0, 1.0), (6, 5, 1.0, 1.0), (16, 12, 2.0, 0.75), (1, 7, 1.0, 3.0)])
def test_rectilinear_grid_matches_reference(nx, ny, w, h):
    case = R.Case(nx, ny, w, h)
    g = R.Grid.rectilinear(case)
    om = O.Mesh.rectilinear(nx, ny, w, h)
    check_mesh(g, om)
    for k in (1, 2):       # IndexMap (UE/IndexMap.cpp:5-40), one rank
        loc, glo = g.index_map(k)
        lr, gr = om.array("localRow"), om.array("globalRow")
        n = len(lr)
        assert np.array_equal(loc, np.concatenate([lr + s * n for s in range(k)]))
        assert np.array_equal(glo, np.concatenate([gr + s * n for s in range(k)]))
    g.close(); case.close()


@pytest.mark.parametrize("nx,ny", [(4, 3), (9, 8)])
def test_triangulated_grid_matches_reference(nx, ny):
    case = R.Case(nx, ny)
    om = O.Mesh.triangulated(nx, ny, 1.0, 1.0)
    g = ref_grid_like(om, case)
    check_mesh(g, om)
    g.close(); case.close()


def _pair(kind, nx, ny, w=1.0, h=1.0, rho=1.0, mu=0.1, bcs=None, obcs=None):
    case = R.Case(nx, ny, w, h, rho, mu, bcs=bcs)
    if kind == "rect":
        g, om = R.Grid.rectilinear(case), O.Mesh.rectilinear(nx, ny, w, h)
    else:
        om = O.Mesh.triangulated(nx, ny, w, h)
        g = ref_grid_like(om, case)
    fs = R.FracStep(case, g)
    if obcs is None:
        ofs = O.cavity(om, rho, mu)
    else:
        ofs = O.FracStep(om, rho, mu)
        for (field, patch, typ, vx, vy) in obcs:
            ofs.set_bc(field, patch, typ, vx, vy)
        ofs.initialize()
    return case, g, fs, om, ofs


def _same_handoff(fs, which, oeq, tol=1e-13):
    rp, ci, va, b = fs.handoff(which)
    rp2, ci2, va2, rhs2 = oeq.export()
    assert np.array_equal(rp, rp2) and np.array_equal(ci, ci2), which + ": CSR pattern"
    scale = max(np.abs(va2).max(), 1e-300)
    assert np.abs(va - va2).max() <= tol * scale, which + ": coefficients"
    assert np.abs(b + rhs2).max() <= tol * max(np.abs(rhs2).max(), 1e-300), which + ": right-hand side (b = -rhs_)"


@pytest.mark.parametrize("kind,nx,ny,w,h", [("rect", 6, 5, 1.0, 1.0), ("rect", 12, 9, 2.0, 0.5), ("tri", 5, 4, 1.0, 1.0)])
def test_assembly_from_a_random_state_matches_reference(kind, nx, ny, w, h):
    """One FractionalStep::solve from the same pseudo-random state with a backend that returns x = 0: what
    FiniteVolumeEquation<T>::solve hands over for uEqn_ (compact layout) and pEqn_ (padded, neighbour first) --
    pattern bit-exact, values to round-off."""
    case, g, fs, om, ofs = _pair(kind, nx, ny, w, h)
    rng = np.random.default_rng(3)
    N, F = om.sizes["nCells"], om.sizes["nFaces"]
    for k in ("ux", "uy", "gpx", "gpy", "p"):
        v = rng.standard_normal(N)
        fs.set(k, v); ofs.view(k)[:] = v
    for k in ("ufx", "ufy", "pf"):
        v = rng.standard_normal(F)
        fs.set(k, v); ofs.view(k)[:] = v
    R.use_null_solver()
    null = O.SOLVE_CB(lambda n, rp, ci, va, b, x, user: 0)
    O.lib().or_fs_set_solver(ofs.h, null, None)
    dt = 0.37 / nx
    fs.step(dt)
    ofs.step(dt)
    _same_handoff(fs, "uEqn", O.Crs(handle=O.lib().or_fs_ueqn(ofs.h), own=False))
    _same_handoff(fs, "pEqn", O.Crs(handle=O.lib().or_fs_peqn(ofs.h), own=False))
    for k in ("ux", "uy", "ufx", "ufy", "p", "pf", "gpx", "gpy", "gpfx", "gpfy"):
        a, b = fs.view(k), ofs.view(k)
        assert np.abs(a - b).max() <= 1e-12 * max(np.abs(b).max(), 1.0), k
    fs.close(); g.close(); case.close()


@pytest.mark.parametrize("kind,nx,ny,K", [("rect", 8, 8, 6), ("rect", 20, 14, 5), ("tri", 8, 7, 5)])
def test_cavity_steps_match_reference(kind, nx, ny, K):
    """K converged steps (direct solves on both sides): the reference's fields vs the oracle's."""
    case, g, fs, om, ofs = _pair(kind, nx, ny)
    R.use_direct_solver()
    ofs.use_direct_solver()
    dt = 0.5 / nx
    for _ in range(K):
        fs.step(dt)
        ofs.step(dt)
    for k in ("ux", "uy", "ufx", "ufy", "gpx", "gpy"):
        a, b = fs.view(k), ofs.view(k)
        assert np.abs(a - b).max() <= 1e-10 * max(np.abs(b).max(), 1.0), k
    a, b = fs.view("p"), ofs.view("p")
    assert np.abs((a - a.mean()) - (b - b.mean())).max() <= 1e-9 * np.abs(b - b.mean()).max()
    assert abs(fs.max_divergence() - ofs.max_divergence()) < 1e-13
    assert abs(fs.max_courant(dt) - ofs.max_courant(dt)) < 1e-12
    fs.close(); g.close(); case.close()


def test_fixed_pressure_patch_and_outlet_style_bcs_match_reference():
    """p fixed on one patch (non-singular pEqn_), u normal_gradient on it: the boundary branches of div / laplacian."""
    bcs = {"u": {"*": ("fixed", "(0,0)"), "x-": ("fixed", "(1,0.25)"), "x+": ("normal_gradient", "(0,0)")},
           "p": {"*": ("normal_gradient", "0"), "x+": ("fixed", "0.5")}}
    obcs = [("u", "y-", O.FIXED, 0., 0.), ("u", "y+", O.FIXED, 0., 0.), ("u", "x-", O.FIXED, 1., 0.25),
            ("u", "x+", O.NORMAL_GRADIENT, 0., 0.),
            ("p", "x-", O.NORMAL_GRADIENT, 0., 0.), ("p", "y-", O.NORMAL_GRADIENT, 0., 0.), ("p", "y+", O.NORMAL_GRADIENT, 0., 0.),
            ("p", "x+", O.FIXED, 0.5, 0.)]
    case, g, fs, om, ofs = _pair("rect", 14, 6, 2.0, 1.0, 1.3, 0.05, bcs=bcs, obcs=obcs)
    R.use_direct_solver()
    ofs.use_direct_solver()
    dt = 0.02
    for _ in range(5):
        fs.step(dt)
        ofs.step(dt)
    _same_handoff(fs, "uEqn", O.Crs(handle=O.lib().or_fs_ueqn(ofs.h), own=False), tol=1e-11)
    _same_handoff(fs, "pEqn", O.Crs(handle=O.lib().or_fs_peqn(ofs.h), own=False), tol=1e-11)
    for k in ("ux", "uy", "ufx", "ufy", "p", "pf", "gpx", "gpy"):
        a, b = fs.view(k), ofs.view(k)
        assert np.abs(a - b).max() <= 1e-10 * max(np.abs(b).max(), 1.0), k
    fs.close(); g.close(); case.close()


def test_max_time_step_matches_reference():
    case, g, fs, om, ofs = _pair("rect", 10, 10)
    R.use_direct_solver()
    ofs.use_direct_solver()
    dt = 0.01
    for _ in range(3):
        fs.step(dt); ofs.step(dt)
    co = ofs.max_courant(dt)
    want = min(0.5 / co * dt, (1 + 0.1 * 0.5 / co) * dt, 1.2 * dt, 1.0)     # Solver.timeStep = 1 in the test case
    assert abs(fs.max_time_step(0.5, dt) - want) < 1e-15
    fs.close(); g.close(); case.close()
