"""Config 4: rising bubble, FractionalStepMultiphase module (VOF + CICSAM + CELESTE surface tension), one B200.

    python tools/bubble_case.py [--nx 1414] [--steps 10] [--warmup 3] [--precond amg|ilu0]

Geometry, properties and boundary conditions of Examples/RisingBubble/case/*.info on an nx x 2nx grid of the 1 x 2
domain (1414 x 2828 = 4.0M cells): rho 998 / 1.225, mu 8.94e-4 / 1.84e-5, sigma 0.0762, g = (0, -9.8065); bubble of
radius 0.125 at (0.5, 0.5), free surface at y = 1.5; p fixed on y+.  The volume fraction is initialised from the exact
circle / half-plane signed distance (linear ramp over one cell).  smoothingKernelRadius and timeStep of the case file
(0.021, 2.5e-5 at h = 0.01) scale with h.  Prints one JSON line: time-steps/s, cell-updates/s, iterations per
equation, hierarchy rebuilds of the variable-density pressure equation.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=1414)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--precond", default="amg", choices=["amg", "ilu0"])
    ap.add_argument("--amg-refresh", default="auto", choices=["off", "auto", "always"],
                    help="numeric re-setup of the multigrid hierarchy on the device when the coefficients change")
    ap.add_argument("--dt-scale", type=float, default=1.0, help="multiplies the case's time step")
    a = ap.parse_args()
    import torch
    from phase_b200.api import FIXED, NORMAL_GRADIENT, Communicator, FiniteVolumeGrid2D as G, FractionalStepMultiphase
    comm = Communicator(0)
    nx, ny = a.nx, 2 * a.nx
    h = 1.0 / nx
    t0 = time.perf_counter()
    g = G.rectilinear(comm, nx, ny, 1.0, 2.0)
    t_mesh = time.perf_counter() - t0
    mp = FractionalStepMultiphase(g, 998.0, 1.225, 8.94e-4, 1.84e-5, 0.0762, (0.0, -9.8065), 2.1 * h)
    for pt in ("x-", "x+", "y-", "y+"):
        mp.u.setBoundary(pt, NORMAL_GRADIENT if pt == "y+" else FIXED, (0.0, 0.0))
        mp.p.setBoundary(pt, FIXED if pt == "y+" else NORMAL_GRADIENT, 0.0)
        mp.gamma.setBoundary(pt, NORMAL_GRADIENT, 0.0)
    cx, cy = g.f64("cellCx"), g.f64("cellCy")
    d = np.minimum(np.hypot(cx - 0.5, cy - 0.5) - 0.125, 1.5 - cy)       # < 0 inside the light phase
    mp.gamma.set("cells", np.clip(0.5 - d / h, 0.0, 1.0))
    mp.gamma.interpolateFaces()
    base = dict(solver="BICGSTAB", maxIters=5000, tolerance=a.tol)
    if a.precond == "amg":
        base["amgRefresh"] = a.amg_refresh
    mp.gammaEqn.solver.setup(dict(base, preconditioner="jacobi"))
    mp.uEqn.solver.setup(dict(base, preconditioner=a.precond))
    mp.pEqn.solver.setup(dict(base, preconditioner=a.precond))
    t0 = time.perf_counter()
    mp.initialize()
    t_init = time.perf_counter() - t0
    dt = 2.5e-5 * (h / 0.01) * a.dt_scale
    stream = torch.cuda.ExternalStream(comm.stream())
    stats = [mp.solve(dt) for _ in range(a.warmup)]
    torch.cuda.synchronize()
    l0 = comm.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
    timed = [mp.solve(dt) for _ in range(a.steps)]
    with torch.cuda.stream(stream):
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    mean = lambda k: float(np.mean([s[k] for s in timed]))
    gam = mp.gamma.get("cells")
    line = {"config": "RisingBubble (FractionalStepMultiphase), %dx%d = %d cells, dt = %.3e, tolerance %g, %s on uEqn_/pEqn_, jacobi on gammaEqn_"
                      % (nx, ny, nx * ny, dt, a.tol, a.precond),
            "time_steps_per_s": 1e3 / ms, "ms_per_step": ms, "cell_updates_per_s": nx * ny * 1e3 / ms, "steps": a.steps, "warmup": a.warmup,
            "iters": {"gammaEqn": mean("itersGamma"), "uEqn": mean("itersU"), "pEqn": mean("itersP")},
            "relres": {"gammaEqn": timed[-1]["errorGamma"], "uEqn": timed[-1]["errorU"], "pEqn": timed[-1]["errorP"]},
            "max_divergence": timed[-1]["maxDivergence"], "max_courant": timed[-1]["maxCourant"],
            "gamma_range": [float(gam.min()), float(gam.max())], "gamma_volume": float((gam * g.f64("vol")).sum()),
            "kappa_max": float(np.abs(mp.kappa.get("cells")).max()),
            "gpu_launches": int(comm.kernel_launches() - l0), "host_setup_s": {"mesh": t_mesh, "initialize_incl_celeste_stencils": t_init}}
    if a.precond == "amg":
        line["amg_pEqn"] = mp.pEqn.solver.amgInfo()
        line["amg_uEqn"] = mp.uEqn.solver.amgInfo()
        line["amg_refresh"] = {"mode": a.amg_refresh, "pEqn": mp.pEqn.solver.amgRefreshInfo(), "uEqn": mp.uEqn.solver.amgRefreshInfo()}
    print(json.dumps(line), flush=True)
    mp.close(); g.close(); comm.close()


if __name__ == "__main__":
    main()
