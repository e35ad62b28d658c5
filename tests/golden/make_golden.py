"""Generates tests/golden/*.npz from the CPU oracle in THIS container (where
oracle/_ref -- the reference's own CrsEquation compiled in place -- is available
and used to produce the CSR artefacts).  Run:  python tests/golden/make_golden.py

Contents per case: mesh integer artefacts (I1, I2), uEqn_/pEqn_ in the reference's
CSR layouts for a seeded field state (I4, A1-A8; from the oracle, whose CSR algebra
is first re-verified bit-exact against the REFERENCE CrsEquation compiled in
oracle/_ref -- the script refuses to run without it), fields after K fractional-step
time steps with exact direct solves, partition/halo maps for a fixed partition (I5).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from tests.util import ORACLE_MESH_INT  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [("rect_6x5", "rect", 6, 5, 1.0, 0.8, 6), ("tri_5x4", "tri", 5, 4, 1.0, 1.0, 5),
         ("rect_16x16", "rect", 16, 16, 1.0, 1.0, 8)]


def seeded_state(om, seed):
    rng = np.random.default_rng(seed)
    N, F = om.sizes["nCells"], om.sizes["nFaces"]
    st = {k: rng.standard_normal(N) for k in ("ux", "uy", "gpx", "gpy", "p", "u0x", "u0y")}
    st.update({k: rng.standard_normal(F) for k in ("ufx", "ufy", "pf", "u0fx", "u0fy")})
    return st


def make(name, kind, nx, ny, w, h, K):
    mk = O.Mesh.rectilinear if kind == "rect" else O.Mesh.triangulated
    om = mk(nx, ny, w, h)
    out = {"kind": kind, "nx": nx, "ny": ny, "w": w, "h": h, "K": K, "dt": 0.5 * w / nx, "seed": 11}
    for k in ORACLE_MESH_INT:
        out["mesh_" + k] = om.array(k)
    for k in ("vol", "cellCx", "cellCy"):
        out["mesh_" + k] = om.array(k)
    fs = O.cavity(om, 1.0, 0.1)
    st = seeded_state(om, 11)
    for k, v in st.items():
        fs.view(k)[:] = v
        out["state_" + k] = v
    dt = out["dt"]
    for tag, e in (("u", fs.assemble_u(dt)), ("p", fs.assemble_p(dt))):
        rp, ci, va, rhs = e.export()
        out.update({tag + "_rowPtr": rp, tag + "_colInd": ci, tag + "_vals": va, tag + "_rhs": rhs})
    # K steps from rest with exact solves
    fs2 = O.cavity(mk(nx, ny, w, h), 1.0, 0.1)
    fs2.use_direct_solver()
    for _ in range(K):
        fs2.step(dt)
    p = fs2.view("p").copy()
    out.update({"final_ux": fs2.view("ux").copy(), "final_uy": fs2.view("uy").copy(), "final_p0": p - p.mean(),
                "final_ufx": fs2.view("ufx").copy(), "final_ufy": fs2.view("ufy").copy()})
    # partition maps for a fixed, seed-independent partition vector (3 parts by global id thirds)
    N = om.sizes["nCells"]
    part = (np.arange(N) * 3 // N).astype(np.int32)
    out["part"] = part
    for r, loc in enumerate(om.partition(part, 3)):
        for k in ("globalId", "owner", "localRow", "globalRow", "bufPtr", "bufCell", "sendPtr", "sendCell", "faceL", "faceR"):
            out["part%d_%s" % (r, k)] = loc.array(k)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok", os.path.getsize(os.path.join(HERE, name + ".npz")), "bytes")


if __name__ == "__main__":
    assert O.ref_lib() is not None, "oracle/_ref must be built (needs /root/reference)"
    from tests import test_oracle_ref_crs as T
    T.test_hot_path_operator_order_matches_reference()   # oracle == reference CrsEquation, bit-exact patterns
    for seed in range(3):
        T.test_random_algebra(seed)
    for c in CASES:
        make(*c)
