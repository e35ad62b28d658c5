"""I3 + I5 on the product path WITHOUT a GPU (host-only context): partition,
local numbering, IndexMap rows and halo maps of libphase_b200 are bit-exact with
the oracle's restatement of FiniteVolumeGrid2D::partition / initCommBuffers, and
a world_size-2 gloo run exchanges halos with them and reproduces the global SpMV."""
import os
import socket

import numpy as np
import pytest

import oracle as O

INT_ARRAYS = ["cptr", "cind", "faceN1", "faceN2", "faceL", "faceR", "facePatch", "ilPtr", "ilFace", "ilCell",
              "blPtr", "blFace", "dlPtr", "dlCell", "owner", "globalId", "localRow", "globalRow", "bufPtr",
              "bufCell", "sendPtr", "sendCell"]


def host_comm(rank=0, nprocs=1):
    from phase_b200.api import Communicator
    return Communicator(Communicator.HOST_ONLY, rank, nprocs)


@pytest.mark.parametrize("kind,nx,ny,P", [("rect", 9, 8, 3), ("tri", 6, 7, 4), ("rect", 16, 16, 8), ("rect", 5, 2, 2)])
def test_rcb_partition_local_meshes_match_oracle(kind, nx, ny, P):
    from phase_b200.api import FiniteVolumeGrid2D as G
    g = (G.rectilinear if kind == "rect" else G.triangulated)(host_comm(), nx, ny, 1.0, 1.0)
    om = (O.Mesh.rectilinear if kind == "rect" else O.Mesh.triangulated)(nx, ny, 1.0, 1.0)
    part = g.partition_rcb(P)
    counts = np.bincount(part, minlength=P)
    assert counts.min() > 0 and counts.max() - counts.min() <= 1      # balanced bisection
    assert np.array_equal(part, g.partition_rcb(P))                   # deterministic
    locs = om.partition(part, P)
    for r in range(P):
        gl = g.local(part, host_comm(r, P))
        for nm in INT_ARRAYS:
            assert np.array_equal(gl.i32(nm), locs[r].array(nm)), (r, nm)
        s = gl.sizes()
        assert s["nLocal"] == counts[r] and s["rowOffset"] == counts[:r].sum()
        # every ghost is a face- or node-neighbour of an owned cell; owned rows are contiguous
        owner, lrow = gl.i32("owner"), gl.i32("localRow")
        assert np.array_equal(np.sort(lrow[owner == r]), np.arange(counts[r]))
        gl.close()
    g.close()


@pytest.mark.parametrize("nx,ny,P", [(9, 8, 3), (4, 12, 4), (7, 5, 2), (3, 8, 8)])
def test_strip_mesh_equals_generic_path(nx, ny, P):
    from phase_b200.api import FiniteVolumeGrid2D as G
    g = G.rectilinear(host_comm(), nx, ny, 1.0, 2.0)
    part = ((np.arange(nx * ny) // nx) * P // ny).astype(np.int32)
    for r in range(P):
        gs = G.rectilinear_strip(host_comm(r, P), nx, ny, 1.0, 2.0)
        gl = g.local(part, host_comm(r, P))
        for nm in INT_ARRAYS + ["rowPtr", "colInd", "slotL", "slotR", "cell2dev"]:
            assert np.array_equal(gs.i32(nm), gl.i32(nm)), (r, nm)
        for nm in ("vol", "cellCx", "cellCy", "faceSx", "faceSy", "faceG", "faceW"):
            assert np.allclose(gs.f64(nm), gl.f64(nm), rtol=1e-13, atol=1e-15), nm
        gs.close(); gl.close()
    g.close()


def test_host_only_context_refuses_device_objects():
    from phase_b200.api import FiniteVolumeGrid2D as G, FiniteVolumeField, SparseMatrixSolver, PhaseB200Error
    c = host_comm()
    g = G.rectilinear(c, 3, 3)
    with pytest.raises(PhaseB200Error):
        FiniteVolumeField(g, 1, "p")
    with pytest.raises(PhaseB200Error):
        SparseMatrixSolver(c)
    g.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, nx, ny, q):
    import torch
    import torch.distributed as dist
    from phase_b200.api import Communicator, FiniteVolumeGrid2D as G
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        hc = Communicator(Communicator.HOST_ONLY)
        g = G.triangulated(hc, nx, ny, 1.0, 1.0)
        part = g.partition_rcb(world)
        gl = g.local(part, Communicator(Communicator.HOST_ONLY, rank, world))
        # global Laplacian-like operator on the canonical pattern: A = L (graph Laplacian + I)
        owner, gid, lrow = gl.i32("owner"), gl.i32("globalId"), gl.i32("localRow")
        ilPtr, ilCell = gl.i32("ilPtr"), gl.i32("ilCell")
        N = g.sizes()["nCells"]
        xg = np.sin(1.0 + np.arange(N) * 0.37)                 # global vector, known everywhere
        x = np.full(gl.sizes()["nCells"], np.nan)
        mine = owner == rank
        x[mine] = xg[gid[mine]]                                # ghosts unknown until the exchange
        bufPtr, bufCell = gl.i32("bufPtr"), gl.i32("bufCell")
        sendPtr, sendCell = gl.i32("sendPtr"), gl.i32("sendCell")
        reqs, recvs = [], {}
        for p in range(world):                                 # UG/FiniteVolumeGrid2D.tpp:9-48
            if p == rank:
                continue
            n = bufPtr[p + 1] - bufPtr[p]
            if n:
                recvs[p] = torch.empty(int(n), dtype=torch.float64)
                reqs.append(dist.irecv(recvs[p], src=p))
        for p in range(world):
            if p == rank:
                continue
            cells = sendCell[sendPtr[p]:sendPtr[p + 1]]
            if len(cells):
                reqs.append(dist.isend(torch.from_numpy(x[cells].copy()), dst=p))
        for r_ in reqs:
            r_.wait()
        for p, buf in recvs.items():
            x[bufCell[bufPtr[p]:bufPtr[p + 1]]] = buf.numpy()
        assert not np.isnan(x).any()
        assert np.array_equal(x, xg[gid])                      # halo delivered the right cells
        y = np.zeros(int(mine.sum()))
        for c in np.nonzero(mine)[0]:
            nb = ilCell[ilPtr[c]:ilPtr[c + 1]]
            y[lrow[c]] = (1.0 + len(nb)) * x[c] - x[nb].sum()
        # reference result from the global mesh
        gp, gc = g.i32("ilPtr"), g.i32("ilCell")
        yg = np.array([(1.0 + gp[c + 1] - gp[c]) * xg[c] - xg[gc[gp[c]:gp[c + 1]]].sum() for c in range(N)])
        ok = np.array_equal(y, yg[np.sort(gid[mine])])
        # dot product all-reduce (Krylov reductions)
        t = torch.tensor([float(y @ y)], dtype=torch.float64)
        dist.all_reduce(t)
        ok = ok and abs(t.item() - float(yg @ yg)) < 1e-9 * float(yg @ yg)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_halo_exchange_and_spmv():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 7, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


@pytest.mark.parametrize("nx,ny,px,py", [(8, 6, 2, 2), (9, 10, 2, 4), (7, 5, 3, 1), (6, 6, 1, 3)])
def test_block_mesh_equals_generic_path_and_oracle(nx, ny, px, py):
    from phase_b200.api import FiniteVolumeGrid2D as G
    P = px * py
    g = G.rectilinear(host_comm(), nx, ny, 2.0, 1.0)
    om = O.Mesh.rectilinear(nx, ny, 2.0, 1.0)
    cells = np.arange(nx * ny)
    part = (((cells // nx) * py // ny) * px + ((cells % nx) * px // nx)).astype(np.int32)
    locs = om.partition(part, P)
    for r in range(P):
        gb = G.rectilinear_block(host_comm(r, P), nx, ny, 2.0, 1.0, px, py)
        gl = g.local(part, host_comm(r, P))
        for nm in INT_ARRAYS + ["rowPtr", "colInd", "slotL", "slotR", "cell2dev"]:
            assert np.array_equal(gb.i32(nm), gl.i32(nm)), (r, nm)
        for nm in INT_ARRAYS:
            assert np.array_equal(gb.i32(nm), locs[r].array(nm)), (r, nm)
        for nm in ("vol", "cellCx", "cellCy", "faceSx", "faceSy", "faceG", "faceW"):
            assert np.allclose(gb.f64(nm), gl.f64(nm), rtol=1e-13, atol=1e-15), nm
        gb.close(); gl.close()
    g.close()
