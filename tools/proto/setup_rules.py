"""Iteration counts behind the two multigrid setup rules of round 2, CPU only: the library's host setup
(phb_amg_host_build_ex) + the scipy transcription of the V(1,1) cycle as preconditioner of scipy's BiCGStab, tolerance 1e-8.

    python tools/proto/setup_rules.py            # uniform Poisson at several mesh shapes, triangles, uEqn_, variable density
    python tools/proto/setup_rules.py --big      # adds 2000x2000 and 4000x2000 (minutes)

Columns: rules off (amgAggTheta 0, amgCoarseSmootherWeight 0 = round 1) | membership filter only | both (default)."""
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
import oracle as O
from tests.test_host_amg import HostAmg, bicgstab_iters, neumann_laplacian

rng = np.random.default_rng(0)


def variable_density(nx, ny, ratio):
    x = (np.arange(nx) + .5) / nx
    y = 2 * (np.arange(ny) + .5) / ny
    X, Y = np.meshgrid(x, y)
    rho = np.where(((X - .5) ** 2 + (Y - .5) ** 2 < .125 ** 2) | (Y > 1.5), 1.0, ratio).ravel()
    idx = np.arange(nx * ny).reshape(ny, nx)
    A = sp.csr_matrix((nx * ny, nx * ny))
    for a, b in ((idx[:, :-1].ravel(), idx[:, 1:].ravel()), (idx[:-1, :].ravel(), idx[1:, :].ravel())):
        w = 1. / (0.5 * (rho[a] + rho[b]))
        A = A + sp.coo_matrix((np.r_[-w, -w, w, w], (np.r_[a, b, a, b], np.r_[b, a, a, b])), shape=A.shape)
    return A.tocsr()


def row(name, A, b):
    out = []
    for kw in (dict(agg_theta=0.0, coarse_weight=0.0), dict(coarse_weight=0.0), {}):
        H = HostAmg(A, coarsest=1000, **kw)
        try:
            its = bicgstab_iters(A, H.cycle(), b)[0]
        except AssertionError:
            its = -1
        sizes = [H.mat(l, 0)[0].shape[0] for l in range(H.nLevels)]
        out.append(its)
        H.close()
    print("%-28s rules off %3d | filter %3d | filter + weights %3d   levels %s" % (name, out[0], out[1], out[2], sizes), flush=True)


def main():
    shapes = [(1000, 1000), (4000, 500), (3998, 500), (4002, 500)]
    if "--big" in sys.argv:
        shapes += [(2000, 2000), (4000, 2000)]
    for nx, ny in shapes:
        A = neumann_laplacian(nx, ny)
        b = rng.standard_normal(nx * ny)
        row("Poisson %dx%d" % (nx, ny), A, b - b.mean())
    nx = ny = 700
    fs = O.cavity(O.Mesh.triangulated(nx, ny, 1.0, 1.0), 1.0, 0.1)
    A = O.csr_to_scipy(*fs.assemble_p(0.25 / nx).export()[:3]).tocsr()
    A.eliminate_zeros()
    b = rng.standard_normal(A.shape[0])
    row("triangles pEqn_ 2x700x700", A, b - b.mean())
    nx = ny = 1000
    fs = O.cavity(O.Mesh.rectilinear(nx, ny, 1.0, 1.0), 1.0, 0.1)
    fs.use_ilu0_solver(tol=1e-6)
    fs.step(0.5 / nx)
    A = O.csr_to_scipy(*fs.assemble_u(0.5 / nx).export()[:3]).tocsr()[:nx * ny, :nx * ny].tocsr()
    A.eliminate_zeros()
    row("uEqn_ 1000x1000 (one block)", A, rng.standard_normal(nx * ny))
    A = variable_density(700, 1400, 1000.)
    b = rng.standard_normal(A.shape[0])
    row("density ratio 1000, 700x1400", A, b - b.mean())


if __name__ == "__main__":
    main()
