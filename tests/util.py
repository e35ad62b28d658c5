import numpy as np

import oracle as O

ORACLE_MESH_INT = ["cptr", "cind", "faceN1", "faceN2", "faceL", "faceR", "facePatch", "ilPtr", "ilFace",
                   "ilCell", "blPtr", "blFace", "dlPtr", "dlCell"]


def rel_l2(a, b):
    a, b = np.asarray(a, float).ravel(), np.asarray(b, float).ravel()
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (d if d > 0 else 1.0)


def oracle_cavity(kind, nx, ny, w=1.0, h=1.0, rho=1.0, mu=0.1):
    m = O.Mesh.rectilinear(nx, ny, w, h) if kind == "rect" else O.Mesh.triangulated(nx, ny, w, h)
    return m, O.cavity(m, rho, mu)


def set_random_state(ofs, gfs, seed=0):
    """Same pseudo-random fields in the oracle FracStep and the GPU FractionalStep."""
    rng = np.random.default_rng(seed)
    N, F = ofs.mesh.sizes["nCells"], ofs.mesh.sizes["nFaces"]
    vals = {k: rng.standard_normal(N) for k in ("ux", "uy", "gpx", "gpy", "p")}
    fvals = {k: rng.standard_normal(F) for k in ("ufx", "ufy", "pf")}
    for k, v in {**vals, **fvals}.items():
        ofs.view(k)[:] = v
    gfs.u.set("cells", np.concatenate([vals["ux"], vals["uy"]]))
    gfs.u.set("faces", np.concatenate([fvals["ufx"], fvals["ufy"]]))
    gfs.gradP.set("cells", np.concatenate([vals["gpx"], vals["gpy"]]))
    gfs.p.set("cells", vals["p"])
    gfs.p.set("faces", fvals["pf"])
    # old level = a different random state
    old = {k: rng.standard_normal(N) for k in ("u0x", "u0y")}
    fold = {k: rng.standard_normal(F) for k in ("u0fx", "u0fy")}
    for k, v in {**old, **fold}.items():
        ofs.view(k)[:] = v
    gfs.u.savePreviousTimeStep()
    gfs.u.set("cells0", np.concatenate([old["u0x"], old["u0y"]]))
    gfs.u.set("faces0", np.concatenate([fold["u0fx"], fold["u0fy"]]))
