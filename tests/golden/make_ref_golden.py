"""Writes tests/golden/ref_*.npz from the REFERENCE ITSELF: oracle/_ref/libphase_ref_fv.so is the reference's own
grid / field / operator / FractionalStep sources compiled in place (oracle/build.py, oracle/ref_fv_driver.cpp), so
every array below was produced by /root/reference/src code, not by the oracle restatement.  Run here (needs
/root/reference):   python tests/golden/make_ref_golden.py

Per case: connectivity and link tables (I1, I2), geometry (G1-G3), the cavity's fields after K converged time steps
(direct solves) and what FiniteVolumeEquation<T>::solve handed to the solver backend in step K: uEqn_ in the
compact layout, pEqn_ padded with the neighbour first (I4, A1-A8, S1).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from oracle import ref_fv as R  # noqa: E402
from tests.test_oracle_ref_fv import F64_KEYS, INT_KEYS, ref_grid_like  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [("ref_cavity_rect_16x12", "rect", 16, 12, 1.0, 1.0, 6), ("ref_cavity_rect_24x10", "rect", 24, 10, 2.0, 0.5, 4),
         ("ref_cavity_tri_8x7", "tri", 8, 7, 1.0, 1.0, 5)]


def make(name, kind, nx, ny, w, h, K):
    case = R.Case(nx, ny, w, h, 1.0, 0.1)
    if kind == "rect":
        g = R.Grid.rectilinear(case)
    else:   # node coordinates and cell -> node lists of the split lattice are inputs; everything derived is the reference's
        g = ref_grid_like(O.Mesh.triangulated(nx, ny, w, h), case)
    out = {"kind": kind, "nx": nx, "ny": ny, "w": w, "h": h, "K": K, "dt": 0.5 * w / nx, "rho": 1.0, "mu": 0.1}
    for k in INT_KEYS + F64_KEYS:
        out["mesh_" + k] = g.array(k)
    for p in ("x-", "x+", "y-", "y+"):
        out["patch_" + p] = np.sort(g.array("patch:" + p))
    R.use_direct_solver()
    fs = R.FracStep(case, g)
    for _ in range(K):
        fs.step(out["dt"])
    for k in ("ux", "uy", "ufx", "ufy", "p", "pf", "gpx", "gpy"):
        out["field_" + k] = fs.view(k)
    for which in ("uEqn", "pEqn"):
        rp, ci, va, b = fs.handoff(which)
        out[which + "_rowPtr"], out[which + "_colInd"], out[which + "_vals"], out[which + "_b"] = rp, ci, va, b
    out["maxDivergence"], out["maxCourant"] = fs.max_divergence(), fs.max_courant(out["dt"])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    fs.close(); g.close(); case.close()
    print(name, "written")


def make_bubble(name, nx, ny, w, h, K, dt):
    """FractionalStepMultiphase::solve (US/FractionalStepMultiphase.cpp) from a smooth bubble under a free surface:
    the state after initialize() and after K steps."""
    from tests import mp_util as U
    radius = 2.1 * w / nx
    case = U.reference_case(R, nx, ny, w, h, radius)
    g = R.Grid.rectilinear(case)
    cx, cy = g.array("cellCx"), g.array("cellCy")
    gamma0 = U.initial_gamma(cx, cy, w, h)
    # face values as FiniteVolumeField::interpolateFaces would set them (distance weights; boundary = cell value)
    fl, fr, fw = g.array("faceL"), g.array("faceR"), g.array("faceW")
    gf = np.where(fr >= 0, fw * gamma0[fl] + (1.0 - fw) * gamma0[np.maximum(fr, 0)], gamma0[fl])
    R.use_direct_solver()
    fs = R.Multiphase(case, g)
    fs.set_field("gamma", gamma0)
    fs.set_field("gamma", gf, faces=True)
    fs.initialize()
    out = {"nx": nx, "ny": ny, "w": w, "h": h, "K": K, "dt": dt, "radius": radius, "gamma0": gamma0, "gamma0_faces": gf}
    for k in ("rho", "mu", "kappa", "gammaTilde"):
        out["init_" + k] = fs.field(k)
    for k in ("n", "fst", "sg"):
        out["init_" + k + "_x"], out["init_" + k + "_y"] = fs.field(k, 0), fs.field(k, 1)
    for _ in range(K):
        fs.step(dt)
    for k in U.SCALARS:
        out["field_" + k] = fs.field(k)
    for k in U.VECTORS:
        rn = U.REF_NAME.get(k, k)
        out["field_" + k + "_x"], out["field_" + k + "_y"] = fs.field(rn, 0), fs.field(rn, 1)
    out["field_uf_x"], out["field_uf_y"] = fs.field("u", 0, faces=True), fs.field("u", 1, faces=True)
    for which in ("gammaEqn", "uEqn", "pEqn"):
        rp, ci, va, b = fs.handoff(which)
        out[which + "_rowPtr"], out[which + "_colInd"], out[which + "_vals"], out[which + "_b"] = rp, ci, va, b
    out["maxDivergence"], out["maxCourant"] = fs.max_divergence(), fs.max_courant(dt)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    fs.close(); g.close(); case.close()
    print(name, "written")


def sheared_mesh(nx, ny, shear):
    """nodes / cells / patch node pairs of an nx x ny quad lattice on the unit square sheared by x' = x + shear * y:
    the x-/x+ boundaries are not axis-aligned, so a SYMMETRY patch there couples the velocity components"""
    om = O.Mesh.rectilinear(nx, ny, 1.0, 1.0)
    xy = np.stack([om.array("nodeX") + shear * om.array("nodeY"), om.array("nodeY")], axis=1)
    fp, n1, n2 = om.array("facePatch"), om.array("faceN1"), om.array("faceN2")
    patches = {}
    for name in ("x-", "x+", "y-", "y+"):
        sel = np.flatnonzero(fp == om.patch_id(name))
        patches[name] = np.stack([n1[sel], n2[sel]], axis=1).reshape(-1)
    return xy, om.array("cptr"), om.array("cind"), patches


SYM_BCS = {"u": {"*": ("fixed", "(0,0)"), "y+": ("fixed", "(1,0)"), "x-": ("symmetry", "(0,0)"), "x+": ("symmetry", "(0,0)")},
           "p": {"*": ("normal_gradient", "0")}}


def make_symmetry(name, nx, ny, shear, K):
    """FractionalStep with SYMMETRY velocity patches on slanted boundaries: the tensor terms of the vector Laplacian
    (UD/Laplacian.cpp:33-41) and add(cell, cell, Tensor2D) (UE/VectorFiniteVolumeEquation.cpp:48-66)."""
    xy, cptr, cind, patches = sheared_mesh(nx, ny, shear)
    case = R.Case(nx, ny, 1.0, 1.0, 1.0, 0.1, bcs=SYM_BCS)
    g = R.Grid.create(xy, cptr, cind)
    for pn, nodes in patches.items():
        g.patch_by_nodes(pn, nodes)
    R.use_direct_solver()
    fs = R.FracStep(case, g)
    dt = 0.5 / nx
    out = {"nx": nx, "ny": ny, "shear": shear, "K": K, "dt": dt, "xy": xy, "cptr": cptr, "cind": cind}
    for pn, nodes in patches.items():
        out["patch_" + pn] = nodes
    for _ in range(K):
        fs.step(dt)
    for k in ("ux", "uy", "ufx", "ufy", "p", "gpx", "gpy"):
        out["field_" + k] = fs.view(k)
    for which in ("uEqn", "pEqn"):
        rp, ci, va, b = fs.handoff(which)
        out[which + "_rowPtr"], out[which + "_colInd"], out[which + "_vals"], out[which + "_b"] = rp, ci, va, b
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    fs.close(); g.close(); case.close()
    print(name, "written")


if __name__ == "__main__":
    if not R.available():
        sys.exit("the reference FV library is not available (needs /root/reference)")
    for c in CASES:
        make(*c)
    make_bubble("ref_bubble_24x48", 24, 48, 1.0, 2.0, 4, 1e-3)
    make_symmetry("ref_symmetry_sheared_10x8", 10, 8, 0.25, 5)
