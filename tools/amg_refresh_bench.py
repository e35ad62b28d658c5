"""Device-side numeric re-setup of the multigrid hierarchy at bench scale: the variable-density
pressure operator -div((1/rho) grad) on nx x nx cells with a bubble (density ratio 1000) that moves a little every step.
One solver keeps its hierarchy and refreshes the values on the device (`amgRefresh always`), a second one is set up from
scratch on the host for every matrix; prints per step the iterations of both, and the device time of the refresh.
    python tools/amg_refresh_bench.py [--nx 2000] [--steps 20]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phase_b200.synthetic import beta_field, variable_laplacian  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=2000)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--shift", type=float, default=0.01, help="bubble displacement per step (domain widths)")
    ap.add_argument("--mode", default="always", choices=["always", "auto"])
    ap.add_argument("--tol", type=float, default=1e-8)
    a = ap.parse_args()
    from phase_b200.api import Communicator, SparseMatrixSolver
    comm = Communicator(0)
    n = a.nx * a.nx
    keys = dict(solver="BICGSTAB", maxIters=2000, tolerance=a.tol, preconditioner="amg", nullSpace="constant")
    s = SparseMatrixSolver(comm).setup(dict(keys, amgRefresh=a.mode))
    rng = np.random.default_rng(0)
    rows = []
    for step in range(a.steps):
        A = variable_laplacian(a.nx, a.nx, beta_field(a.nx, a.nx, 0.3 + a.shift * step, 1000.0))
        b = rng.standard_normal(n); b -= b.mean()
        s.setRank(n); s.set(A.indptr, A.indices, A.data); s.setRhs(b)
        t0 = time.perf_counter()
        err = s.solve()
        t_solve = time.perf_counter() - t0
        it = s.nIters()
        f = SparseMatrixSolver(comm).setup(dict(keys, amgRefresh="off"))
        f.setRank(n); f.set(A.indptr, A.indices, A.data); f.setRhs(b)
        f.solve()
        rows.append({"step": step, "iters_refreshed": it, "iters_fresh_setup": f.nIters(), "relres": err,
                     "solve_incl_refresh_ms": 1e3 * t_solve, "refresh": s.amgRefreshInfo(), "fresh_setup_ms": f.amgInfo()["setupMs"]})
        f.close()
        print(json.dumps(rows[-1]), flush=True)
    ir = np.array([r["iters_refreshed"] for r in rows[1:]], float); iff = np.array([r["iters_fresh_setup"] for r in rows[1:]], float)
    print(json.dumps({"summary": "%dx%d variable-density pressure operator, bubble moving %.3g per step, %d steps" % (a.nx, a.nx, a.shift, a.steps),
                      "mode": a.mode, "host_setups": s.amgInfo()["setups"], "refreshes": s.amgRefreshInfo()["refreshes"],
                      "refresh_ms_mean": float(np.mean([r["refresh"]["refreshMs"] for r in rows[1:]])),
                      "symbolic_bytes": s.amgRefreshInfo()["bytes"],
                      "iters_refreshed_mean": float(ir.mean()), "iters_fresh_mean": float(iff.mean()),
                      "worst_ratio": float((ir / iff).max())}), flush=True)
    s.close(); comm.close()


if __name__ == "__main__":
    main()
