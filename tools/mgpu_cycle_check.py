"""ONE application of the device V-cycle on N GPUs against a scipy transcription of the same distributed hierarchy
(built by the library's own host setup with the ranks as threads), run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
        tools/mgpu_cycle_check.py [--nx 256 --ny 192] [--tail-rows 2000] [--coarsest 100] [--peer] [--precision double]

The converged fields of the parity tests cannot tell a correct preconditioner from a merely convergent one; this
can.  Exit code 0 = the cycle agrees (1e-10 in double, 1e-4 in single) and the iteration count matches the
transcription's.  With one rank it checks the serial cycle.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--ny", type=int, default=192)
    ap.add_argument("--tail-rows", type=int, default=2000)
    ap.add_argument("--coarsest", type=int, default=100)
    ap.add_argument("--peer", action="store_true")
    ap.add_argument("--precision", default="double")
    ap.add_argument("--fuse-rows", type=int, default=200000, help="amgFuseRows: levels up to this size run in the fused kernel")
    a = ap.parse_args()
    import scipy.sparse as sp
    import torch
    import torch.distributed as dist
    import oracle as O
    from phase_b200.api import Communicator, FiniteVolumeGrid2D as G, lid_driven_cavity
    from tests.test_host_amg import DistAmg, HostAmg, bicgstab_iters
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    lr = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(lr)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
        box = [Communicator.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    comm = Communicator(lr, rank, world, uid)
    px = 1
    while px * px * 4 <= world and world % (px * 2) == 0:
        px *= 2
    py = world // px
    if world == 1:
        g = G.rectilinear(comm, a.nx, a.ny, 1.0, 1.0)
    else:
        g = G.rectilinear_block(comm, a.nx, a.ny, 1.0, 1.0, px, py)

    def all_gather(obj):
        if world == 1:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out
    if a.peer and world > 1:
        comm.enable_peer_memory(g, all_gather)
    keys = dict(tolerance=1e-10, maxIters=500, preconditioner="amg", amgCoarsest=a.coarsest, amgTailRows=a.tail_rows,
                amgPrecision=a.precision, amgFuseRows=a.fuse_rows)
    fs = lid_driven_cavity(g, 1.0, 0.1, solver=keys)
    dt = 0.5 / a.nx
    fs.solve(dt)
    fs.p.fill(0.0)
    fs.pEqn.solve(warmStart=False)
    its_dev = fs.pEqn.solver.nIters()
    fs_info = fs.pEqn.solver.amgInfo()
    owner, gid, lrow = g.i32("owner"), g.i32("globalId"), g.i32("localRow")
    mine = np.flatnonzero(owner == rank)
    order = mine[np.argsort(lrow[mine])]           # owned cells in local row order
    ggid = gid[order]
    assert np.all(np.diff(ggid) > 0), "owned rows are not in ascending global order"
    N = a.nx * a.ny
    r = np.random.default_rng(5).standard_normal(N)
    r -= r.mean()
    z_loc = fs.pEqn.solver.applyPreconditioner(r[ggid])
    pieces = all_gather((ggid, z_loc))
    ok = True
    if rank == 0:
        z = np.zeros(N)
        part = np.zeros(N, np.int32)
        for q, (gg, zl) in enumerate(pieces):
            z[gg] = zl
            part[gg] = q
        ofs = O.cavity(O.Mesh.rectilinear(a.nx, a.ny, 1.0, 1.0), 1.0, 0.1)
        A = O.csr_to_scipy(*ofs.assemble_p(dt).export()[:3]).tocsr()
        if world == 1:
            H = HostAmg(A, coarsest=a.coarsest)
            M = H.cycle()
            info = "serial, %d levels" % H.nLevels
        else:
            H = DistAmg(A, part, coarsest=a.coarsest, tail_rows=a.tail_rows)
            levels, Al = [], A
            for l in range(H.nDist):
                nc = sum(H.rank_matrix(q, l, 2)[0].shape[0] - 1 for q in range(H.nRanks))
                P = H.global_matrix(l, 1, (Al.shape[0], nc))
                R = H.global_matrix(l, 2, (nc, Al.shape[0]))
                An = H.global_matrix(l + 1, 0, (nc, nc)) if l + 1 < H.nDist else H.tail(0).mat(0, 0)[0]
                levels.append((Al, P, R))
                Al = An
            tail = H.tail(0).cycle()

            def cyc(l, b):
                if l == H.nDist:
                    return tail(b)
                A_, P_, R_ = levels[l]
                rho = np.abs(sp.diags(1.0 / A_.diagonal()) @ A_).sum(axis=1).max()
                w = (1.8 / rho) / A_.diagonal()
                x = w * b
                x = x + P_ @ cyc(l + 1, R_ @ (b - A_ @ x))
                return x + w * (b - A_ @ x)
            M = lambda v: cyc(0, v)
            info = "%d distributed + %d replicated levels" % (H.nDist, H.nTail)
        z_ref = M(r)
        err = np.linalg.norm(z - z_ref) / np.linalg.norm(z_ref)
        its_ref = bicgstab_iters(A, M, r, tol=1e-10)[0]
        tol = 1e-10 if a.precision == "double" else 2e-4
        ok = err < tol
        info += ", %d launches per cycle" % int(fs_info["launchesPerCycle"])
        print("cycle check: %s, %d rank(s), rel. difference of one cycle %.3e (tolerance %.0e); BiCGStab iterations: device %d "
              "(pEqn_ rhs), transcription %d (random rhs)  %s" % (info, world, err, tol, its_dev, its_ref, "OK" if ok else "FAIL"),
              flush=True)
    flag = all_gather(ok)
    fs.close(); g.close(); comm.close()
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if all(flag) else 1)


if __name__ == "__main__":
    main()
