"""Config 3: flow around a cylinder on an UNSTRUCTURED triangle mesh, fractional-step module.

    python tools/cylinder_case.py [--cells 1e6] [--steps 5] [--mesh path/to/CylinderMesh.cgns | --synthetic] [--refine k]
    torchrun ... tools/cylinder_case.py --cells 16e6          (one rank per GPU, METIS or RCB partition)

Physics and boundary conditions of Examples/UnstructuredFlowAroundCylinder/case/*.info:
rho 1.81, mu 1.81e-5, Inlet u = (15, 0) fixed, Cylinder u = 0, Outlet / TopBottom zero-gradient u and
fixed p = 0.  By default the shipped mesh is rebuilt from tests/golden/ref_cylinder_mesh.npz (written from the ADF-CGNS
file by tests/golden/make_cylinder_mesh.py; --mesh reads a CGNS file directly) and uniformly refined k times
(15 316 x 4^k triangles; k from --cells unless --refine is given: 16e6 -> k = 5, 15.7M cells); with --synthetic a channel [0,5]x[0,2] with a cylinder of radius 0.1 at (1,1) is
generated: a triangulated lattice with the cells inside the cylinder removed (irregular connectivity,
boundary faces classified by position into the same four patches).
Prints one JSON line: cells, time-steps/s, iterations, max divergence.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synthetic_cylinder_cells(ncells):
    L, H, R, cx, cy = 5.0, 2.0, 0.1, 1.0, 1.0
    ny = max(8, int(round(np.sqrt(ncells / 2 * H / L))))
    nx = int(round(ny * L / H))
    xs, ys = np.linspace(0, L, nx + 1), np.linspace(0, H, ny + 1)
    X, Y = np.meshgrid(xs, ys)
    xy = np.stack([X.ravel(), Y.ravel()], 1)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny))
    bl = (j * (nx + 1) + i).ravel()
    br, tl, tr = bl + 1, bl + nx + 1, bl + nx + 2
    even = ((i + j) % 2 == 0).ravel()
    t1 = np.where(even[:, None], np.stack([bl, br, tr], 1), np.stack([bl, br, tl], 1))
    t2 = np.where(even[:, None], np.stack([bl, tr, tl], 1), np.stack([br, tr, tl], 1))
    tris = np.empty((2 * len(bl), 3), np.int64)
    tris[0::2], tris[1::2] = t1, t2
    c = xy[tris].mean(1)
    keep = (c[:, 0] - cx) ** 2 + (c[:, 1] - cy) ** 2 > R * R
    tris = tris[keep]
    used = np.unique(tris)
    remap = -np.ones(len(xy), np.int64)
    remap[used] = np.arange(len(used))
    return xy[used], remap[tris], (L, H, R, cx, cy)


def classify_patches(grid, geom):
    L, H, R, cx, cy = geom
    fl, fr, n1, n2 = grid.i32("faceL"), grid.i32("faceR"), grid.i32("faceN1"), grid.i32("faceN2")
    x, y = grid.f64("nodeX"), grid.f64("nodeY")
    b = np.nonzero(fr < 0)[0]
    mx, my = 0.5 * (x[n1[b]] + x[n2[b]]), 0.5 * (y[n1[b]] + y[n2[b]])
    eps = 1e-9
    groups = {"Inlet": mx < eps, "Outlet": mx > L - eps, "TopBottom": (my < eps) | (my > H - eps)}
    rest = ~(groups["Inlet"] | groups["Outlet"] | groups["TopBottom"])
    groups["Cylinder"] = rest
    for name in ("Cylinder", "TopBottom", "Inlet", "Outlet"):
        sel = b[groups[name]]
        grid.createPatchByNodes(name, np.stack([n1[sel], n2[sel]], 1).ravel())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=float, default=1e6)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--mesh", default=None)
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--refine", type=int, default=None)
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--precond", default="amg")
    ap.add_argument("--partition", default="metis", choices=["metis", "rcb"])
    a = ap.parse_args()
    if a.refine is None:
        a.refine = max(0, int(round(np.log(max(a.cells, 1.0) / 15316.0) / np.log(4.0))))
    import torch
    from phase_b200.api import Communicator, FiniteVolumeGrid2D as G, FractionalStep, FIXED, NORMAL_GRADIENT
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
    build = comm if world == 1 else Communicator(Communicator.HOST_ONLY)
    t0 = time.perf_counter()
    fixture = os.path.join(ROOT, "tests", "golden", "ref_cylinder_mesh.npz")
    if a.mesh:
        g = G.from_cgns(build, a.mesh, refine=a.refine)
    elif not a.synthetic:
        d = np.load(fixture)
        g0 = G.from_cells(build, d["xy"], np.arange(0, 3 * len(d["tris"]) + 1, 3), d["tris"].ravel())
        for name in d["patch_order"]:
            g0.createPatchByNodes(str(name), d["patch_" + str(name)].ravel())
        if a.refine > 0:
            g = g0.refined(a.refine)
            g0.close()
        else:
            g = g0.finalize()
    else:
        xy, tris, geom = synthetic_cylinder_cells(a.cells)
        g = G.from_cells(build, xy, np.arange(0, 3 * len(tris) + 1, 3), tris.ravel())
        classify_patches(g, geom)
        g.finalize()
    n_global = g.sizes()["nCells"]
    t_part = 0.0
    if world > 1:
        tp = time.perf_counter()
        if rank == 0:                          # one rank partitions, everybody gets the vector
            part = g.partition_metis(world)[0] if a.partition == "metis" else g.partition_rcb(world)
            part_t = torch.from_numpy(np.ascontiguousarray(part, np.int32)).cuda()
        else:
            part_t = torch.empty(n_global, dtype=torch.int32, device="cuda")
        dist.broadcast(part_t, src=0)
        part = part_t.cpu().numpy()
        t_part = time.perf_counter() - tp
        gl = g.local(part, comm)
        g.close()
        g = gl

        def ag(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        comm.enable_peer_memory(g, ag)
    t_mesh = time.perf_counter() - t0
    fs = FractionalStep(g, 1.81, 1.81e-5)
    for pt, t, v in (("Inlet", FIXED, (15.0, 0.0)), ("Outlet", NORMAL_GRADIENT, (0.0, 0.0)),
                     ("Cylinder", FIXED, (0.0, 0.0)), ("TopBottom", NORMAL_GRADIENT, (0.0, 0.0))):
        fs.u.setBoundary(pt, t, v)
    for pt, t in (("Inlet", NORMAL_GRADIENT), ("Outlet", FIXED), ("Cylinder", NORMAL_GRADIENT), ("TopBottom", FIXED)):
        fs.p.setBoundary(pt, t, 0.0)
    cfg = dict(maxIters=5000, tolerance=a.tol, preconditioner=a.precond)
    fs.uEqn.solver.setup(cfg); fs.pEqn.solver.setup(cfg)
    fs.u.fill(15.0, 0.0)                       # initialConditions.info: uniform inlet velocity
    for pt, v in (("Inlet", (15.0, 0.0)), ("Cylinder", (0.0, 0.0))):
        fs.u.setBoundary(pt, FIXED, v)         # restore the fixed faces after the fill
    fs.initialize()
    vol = g.f64("vol")
    h = float(np.sqrt(2 * vol.min()))
    dt = 0.4 * h / 15.0                        # maxCo ~ 0.8 with the local speed-up around the cylinder
    stats = []
    for k in range(a.warmup):
        stats.append(fs.solve(dt))
    comm.sync()
    t0 = time.perf_counter()
    for k in range(a.steps):
        stats.append(fs.solve(dt))
    comm.sync()
    el = time.perf_counter() - t0
    if rank == 0:
        s = g.sizes()
        print(json.dumps({"config": "flow around a cylinder, unstructured triangles, fractional step",
                          "mesh": a.mesh or ("synthetic holed lattice" if a.synthetic else
                                             "shipped CylinderMesh (fixture), %d refinement rounds" % a.refine),
                          "cells": n_global, "cells_local": s["nLocal"], "n_gpus": world, "preconditioner": a.precond,
                          "partition": "none" if world == 1 else a.partition, "partition_s": t_part,
                          "amg_pEqn": fs.pEqn.solver.amgInfo() if a.precond == "amg" else None,
                          "dt": dt, "time_steps_per_s": a.steps / el, "ms_per_step": 1e3 * el / a.steps,
                          "iters_u": [x["itersU"] for x in stats], "iters_p": [x["itersP"] for x in stats],
                          "max_divergence": stats[-1]["maxDivergence"], "max_courant": stats[-1]["maxCourant"],
                          "mesh_build_s": t_mesh}), flush=True)
    fs.close(); g.close(); comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
