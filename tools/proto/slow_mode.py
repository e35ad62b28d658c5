"""Where does the slowest error mode of the V-cycle live?  (CPU only; the diagnosis behind `amgAggTheta`)

    python tools/proto/slow_mode.py [nx ny] [--old]

Power iteration on I - M^-1 A (M^-1 = the scipy transcription of the library's cycle), then the energy of the mode by
regions of the mesh and the sizes of the aggregates of every level.  With the round-1 rules (--old) on a 4000 x 500 mesh
the mode sits in the last rows of the natural order (top right corner), is smooth (Rayleigh quotient 4e-5) and coincides
with 16-24-member aggregates on the third level, where nine members are typical."""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from tests.test_host_amg import HostAmg, neumann_laplacian


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    nx, ny = (int(args[0]), int(args[1])) if len(args) >= 2 else (4000, 500)
    kw = dict(agg_theta=0.0, coarse_weight=0.0) if "--old" in sys.argv else {}
    A = neumann_laplacian(nx, ny)
    H = HostAmg(A, coarsest=1000, **kw)
    cyc = H.cycle()
    rng = np.random.default_rng(0)
    e = rng.standard_normal(nx * ny)
    e -= e.mean()
    f = 0.
    for _ in range(60):
        e2 = e - cyc(A @ e)
        e2 -= e2.mean()
        f = np.linalg.norm(e2) / np.linalg.norm(e)
        e = e2 / np.linalg.norm(e2)
    print("asymptotic convergence factor of the stationary cycle: %.3f" % f)
    E = (e.reshape(ny, nx)) ** 2
    bx, by = max(1, nx // 10), max(1, ny // 10)
    print("energy by tenths of x:", np.round([E[:, i * bx:(i + 1) * bx].sum() for i in range(10)], 3))
    print("energy by tenths of y:", np.round([E[i * by:(i + 1) * by, :].sum() for i in range(10)], 3))
    print("Rayleigh quotient e'Ae / e'De: %.2e" % ((e @ (A @ e)) / (e @ (A.diagonal() * e))))
    for l in range(H.nLevels - 1):
        P = H.mat(l, 1)[0]
        agg = np.asarray(abs(P).argmax(axis=1)).ravel()
        cnt = np.bincount(agg)
        print("level %d: %d rows, aggregates of %d..%d members (mean %.2f), histogram %s"
              % (l, P.shape[0], cnt.min(), cnt.max(), cnt.mean(), np.bincount(cnt)[:26].tolist()))
    H.close()


if __name__ == "__main__":
    main()
