"""The oracle restatement against golden vectors written by the REFERENCE ITSELF (tests/golden/ref_*.npz, produced
by tests/golden/make_ref_golden.py from oracle/_ref/libphase_ref_fv.so = the reference's own sources compiled in
place).  Runs anywhere: needs neither /root/reference nor the compiled reference."""
import glob
import os

import numpy as np
import pytest

import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "ref_cavity_*.npz")))
INT_KEYS = ["cptr", "cind", "faceN1", "faceN2", "faceL", "faceR", "ilPtr", "ilFace", "ilCell", "blPtr", "blFace", "dlPtr", "dlCell"]


def test_reference_goldens_are_committed():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_reproduces_reference_golden(path):
    G = np.load(path)
    kind, nx, ny, w, h, K, dt = str(G["kind"]), int(G["nx"]), int(G["ny"]), float(G["w"]), float(G["h"]), int(G["K"]), float(G["dt"])
    om = (O.Mesh.rectilinear if kind == "rect" else O.Mesh.triangulated)(nx, ny, w, h)
    for k in INT_KEYS:
        assert np.array_equal(om.array(k), G["mesh_" + k]), k
    for k in ("vol", "cellCx", "cellCy", "faceCx", "faceCy", "faceNx", "faceNy", "ilSx", "ilSy", "ilRcx", "ilRcy", "blRfx", "blSx"):
        assert np.allclose(om.array(k), G["mesh_" + k], rtol=1e-13, atol=1e-15), k
    fp = om.array("facePatch")
    for p in ("x-", "x+", "y-", "y+"):
        assert np.array_equal(np.flatnonzero(fp == om.patch_id(p)), G["patch_" + p]), p
    ofs = O.cavity(om, float(G["rho"]), float(G["mu"]))
    ofs.use_direct_solver()
    for _ in range(K):
        ofs.step(dt)
    for k in ("ux", "uy", "ufx", "ufy", "gpx", "gpy"):
        ref = G["field_" + k]
        assert np.abs(ofs.view(k) - ref).max() <= 1e-10 * max(np.abs(ref).max(), 1.0), k
    p, pr = ofs.view("p"), G["field_p"]
    assert np.abs((p - p.mean()) - (pr - pr.mean())).max() <= 1e-9 * np.abs(pr - pr.mean()).max()
    for which, eq in (("uEqn", O.lib().or_fs_ueqn(ofs.h)), ("pEqn", O.lib().or_fs_peqn(ofs.h))):
        rp, ci, va, rhs = O.Crs(handle=eq, own=False).export()
        assert np.array_equal(rp, G[which + "_rowPtr"]) and np.array_equal(ci, G[which + "_colInd"]), which
        assert np.abs(va - G[which + "_vals"]).max() <= 1e-11 * np.abs(va).max(), which
        assert np.abs(-rhs - G[which + "_b"]).max() <= 1e-9 * max(np.abs(rhs).max(), 1e-300), which
    assert abs(ofs.max_courant(dt) - float(G["maxCourant"])) < 1e-10
