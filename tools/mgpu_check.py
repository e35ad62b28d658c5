"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/mgpu_check.py [--kind rect|tri] [--nx 48 --ny 40] [--steps 4] [--strip]

Every rank builds its local mesh (RCB partition of the global mesh, or the y-strip
fast path), runs K fractional-step time steps with NCCL halo exchange + all-reduced
dot products, and compares its OWNED cells with the oracle's single-domain result
(direct solves).  Exit code 0 = parity within 1e-6 rel-L2 on every rank.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kind", default="rect")
    ap.add_argument("--nx", type=int, default=48)
    ap.add_argument("--ny", type=int, default=40)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--strip", action="store_true")
    ap.add_argument("--precond", default="jacobi")
    ap.add_argument("--amg-scope", default="global", choices=["global", "local"])
    ap.add_argument("--amg-tail-rows", type=int, default=200)
    ap.add_argument("--peer", action="store_true", help="NVLink peer-memory exchanges inside the Krylov loop")
    ap.add_argument("--fused", action="store_true", help="peer pushes/waits inside the compute kernels")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import oracle as O
    from phase_b200.api import Communicator, FiniteVolumeGrid2D as G, lid_driven_cavity
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    lr = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    box = [Communicator.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    comm = Communicator(lr, rank, world, box[0])
    if a.strip:
        gl = G.rectilinear_strip(comm, a.nx, a.ny, 1.0, 1.0)
    else:
        host = Communicator(Communicator.HOST_ONLY)
        g = (G.rectilinear if a.kind == "rect" else G.triangulated)(host, a.nx, a.ny, 1.0, 1.0)
        gl = g.local(g.partition_rcb(world), comm)
    if a.peer:
        def all_gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        comm.enable_peer_memory(gl, all_gather)
    amg = a.precond == "amg"      # scalar systems only: pEqn_ gets the V-cycle, uEqn_ keeps ILU(0)
    fs = lid_driven_cavity(gl, 1.0, 0.1, solver=dict(tolerance=1e-11, maxIters=50000,
                                                     preconditioner="ilu0" if amg else a.precond,
                                                     peerFusion=1 if a.fused else 0),
                           pSolver=dict(preconditioner="amg", amgCoarsest=40, amgScope=a.amg_scope,
                                        amgTailRows=a.amg_tail_rows) if amg else None)
    om = (O.Mesh.rectilinear if a.kind == "rect" else O.Mesh.triangulated)(a.nx, a.ny, 1.0, 1.0)
    ofs = O.cavity(om, 1.0, 0.1)
    ofs.use_direct_solver()
    dt = 0.5 / a.nx
    if os.environ.get("PHB_CHECK_TRACE"):
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["PHB_CHECK_TRACE"]), exit=True)
    trace = lambda *m: print("[rank %d]" % rank, *m, flush=True) if os.environ.get("PHB_CHECK_TRACE") else None
    for k in range(a.steps):
        st = fs.solve(dt)
        trace("step", k, st)
        ofs.step(dt)
    owner, gid = gl.i32("owner"), gl.i32("globalId")
    mine = owner == rank
    u, p = fs.u.get("cells"), fs.p.get("cells")
    rel = lambda x, y: np.linalg.norm(x - y) / max(np.linalg.norm(y), 1e-300)
    eu = max(rel(u[0][mine], ofs.view("ux")[gid[mine]]), rel(u[1][mine], ofs.view("uy")[gid[mine]]))
    # p is defined up to a constant (all-Neumann): remove the GLOBAL mean
    s = torch.tensor([p[mine].sum(), float(mine.sum())], dtype=torch.float64, device="cuda")
    dist.all_reduce(s)
    pm = (s[0] / s[1]).item()
    po = ofs.view("p")
    ep = np.linalg.norm((p[mine] - pm) - (po[gid[mine]] - po.mean())) / np.linalg.norm(po - po.mean())
    # ghosts must hold their owners' values after the last sendMessages(u)
    eg = rel(u[0][~mine], ofs.view("ux")[gid[~mine]]) if (~mine).any() else 0.0
    # Seam 1 on a distributed system: the rank's rows with GLOBAL columns (reference IndexMap numbering),
    # host arrays, halo lists from the local mesh
    from phase_b200.api import SparseMatrixSolver
    rp, ci, va, rhs = fs.assembleP(dt).export(1)
    s1 = SparseMatrixSolver(comm).setup(dict(tolerance=1e-11, maxIters=50000, preconditioner=a.precond,
                                             nullSpace="constant"))
    s1.setHalo(gl)
    trace("seam1 set")
    s1.setRank(len(rhs)); s1.set(rp, ci, va); s1.setRhs(-rhs); s1.solve()
    trace("seam1 solved", s1.nIters())
    x1 = s1.x()
    fs.pEqn.solver.setup(dict(tolerance=1e-11))
    fs.p.fill(0.0)
    fs.pEqn.solve(warmStart=False)
    pd = fs.p.get("cells")[mine]
    lrow = gl.i32("localRow")[mine]
    t1 = torch.tensor([x1.sum(), float(len(x1)), pd.sum()], dtype=torch.float64, device="cuda")
    dist.all_reduce(t1)
    xm, pm2 = (t1[0] / t1[1]).item(), (t1[2] / t1[1]).item()
    es = np.linalg.norm((x1[lrow] - xm) - (pd - pm2)) / max(np.linalg.norm(pd - pm2), 1e-300)
    ok = eu < 1e-6 and ep < 1e-6 and eg < 1e-6 and es < 1e-6
    print("rank %d/%d: cells %d owned %d  relL2 u %.2e p %.2e ghosts %.2e seam1 %.2e  itersP %d  div %.1e  %s" %
          (rank, world, len(owner), int(mine.sum()), eu, ep, eg, es, st["itersP"], st["maxDivergence"],
           "OK" if ok else "FAIL"), flush=True)
    t = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(t)
    s1.close(); fs.close(); gl.close(); comm.close()
    dist.destroy_process_group()
    sys.exit(int(t.item() != 0))


if __name__ == "__main__":
    main()
