"""Config 5: pressure-Poisson solve sweep.

    python tools/poisson_sweep.py [--sizes 1,2,4,8,16,32,64] [--precond jacobi] [--variable]
    torchrun ... tools/poisson_sweep.py --sizes 8,16        (one rank per GPU, block partition)

pEqn_ of the cavity configuration with ONE fixed patch (y+ : p = 0, so the system is non-singular):
  constant coefficient   fv::laplacian(dt, p) == src::div(u)
  --variable             fv::laplacian(dt/rho_f, p) == src::div(u), rho_f a smooth two-fluid field
                         with density ratio 815 (RisingBubble properties, config 4)
u(x,y) = (sin 2 pi x cos 2 pi y, -cos 2 pi x sin 2 pi y) + 1e-3 U(-1,1) (seed 0) on the faces,
x0 = 0, tolerance 1e-8.  Prints one JSON line per size: iterations, time, achieved GB/s
(iters * bytes_per_iteration / time) and its fraction of the measured HBM peak.
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
    ap.add_argument("--sizes", default="1,2,4,8,16")
    ap.add_argument("--precond", default="jacobi")
    ap.add_argument("--variable", action="store_true")
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--max-iters", type=int, default=60000)
    ap.add_argument("--comm", default="peer")
    a = ap.parse_args()
    import torch
    from phase_b200.api import (Communicator, FiniteVolumeGrid2D as G, FiniteVolumeField, FiniteVolumeEquation,
                                SparseMatrixSolver, FIXED, NORMAL_GRADIENT)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    uid = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        box = [Communicator.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    comm = Communicator(lr, rank, world, uid)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    import bench
    px, py = bench.block_layout(world)
    for mcells in [float(s) for s in a.sizes.split(",")]:
        side = int(round(np.sqrt(mcells * 1e6)))
        nx, ny = side - side % px, side - side % py
        if world == 1:
            g = G.rectilinear(comm, nx, ny, 1.0, 1.0)
        else:
            g = G.rectilinear_block(comm, nx, ny, 1.0, 1.0, px, py)
            if a.comm == "peer":
                def ag(obj):
                    out = [None] * world
                    dist.all_gather_object(out, obj)
                    return out
                comm.enable_peer_memory(g, ag)
        p, u = FiniteVolumeField(g, 1, "p"), FiniteVolumeField(g, 2, "u")
        for pt in ("x-", "x+", "y-"):
            p.setBoundary(pt, NORMAL_GRADIENT, 0.0)
        p.setBoundary("y+", FIXED, 0.0)
        fx, fy = g.f64("faceCx"), g.f64("faceCy")
        gid = g.i32("globalId")
        rng = np.random.default_rng(0)
        # deterministic per-face noise independent of the partition: hash of the face centre
        noise = lambda s: 1e-3 * (2.0 * ((np.sin(12.9898 * fx * nx + 78.233 * fy * ny + s) * 43758.5453) % 1.0) - 1.0)
        uf = np.concatenate([np.sin(2 * np.pi * fx) * np.cos(2 * np.pi * fy) + noise(0.0),
                             -np.cos(2 * np.pi * fx) * np.sin(2 * np.pi * fy) + noise(1.0)])
        u.set("faces", uf)
        dt = 0.5 / nx
        eq = FiniteVolumeEquation(p).zero()
        if a.variable:
            gam = FiniteVolumeField(g, 1, "gamma")
            r2 = (fx - 0.5) ** 2 + (fy - 0.5) ** 2
            alpha = 0.5 * (1.0 + np.tanh((0.25 - np.sqrt(r2)) / (4.0 / nx)))       # smooth bubble of radius 0.25
            rho = 998.0 + alpha * (1.225 - 998.0)
            gam.set("faces", dt / rho)
            eq.laplacian(gam, p)
        else:
            eq.laplacian(dt, p)
        eq.srcDiv(u, sign=-1.0)
        s = SparseMatrixSolver(comm).setup(dict(maxIters=a.max_iters, tolerance=a.tol, preconditioner=a.precond))
        eq.solve(solver=s)                 # warm-up (graph capture, symbolic phases)
        p.fill(0.0)
        comm.sync()
        t0 = time.perf_counter()
        eq.solve(solver=s)
        comm.sync()
        dtm = time.perf_counter() - t0
        b_spmv, b_iter = s.bytes()
        if rank == 0:
            n = g.sizes()["nLocal"]
            print(json.dumps({"cells": nx * ny, "n_gpus": world, "rows_per_gpu": n, "variable_coefficient": a.variable,
                              "preconditioner": a.precond, "iterations": s.nIters(), "relres": s.error(),
                              "solve_s": dtm, "ms_per_iteration": 1e3 * dtm / max(1, s.nIters()),
                              "solve_GBps_per_gpu": s.nIters() * b_iter / dtm / 1e9,
                              "frac_of_measured_peak": s.nIters() * b_iter / dtm / 1e9 / peak,
                              "frac_of_8TBs": s.nIters() * b_iter / dtm / 1e9 / 8000.0}), flush=True)
        for o in (s, eq, p, u, g):
            o.close()
    comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
