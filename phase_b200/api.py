"""Host-side mirror of the reference's interface for the hot path, over the C ABI.

Names follow the reference (FiniteVolumeGrid2D, FiniteVolumeField,
FiniteVolumeEquation, SparseMatrixSolver, FractionalStep, fv.* / src.*); the
C++ mirror with the same names lives in include/phase/.  This Python layer is
the harness the tests and bench.py drive the library through -- it holds no
numerics of its own.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import PhaseB200Error, check  # noqa: F401

FIXED, NORMAL_GRADIENT, SYMMETRY = 0, 1, 2


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Communicator:
    """One GPU + streams (+ NCCL communicator): the S/Communicator analogue."""

    HOST_ONLY = -1   # mesh / partition / halo maps only (no kernels): host-logic tests

    def __init__(self, device=0, rank=0, nprocs=1, unique_id=None):
        self.L = _capi.lib()
        h = C.c_void_p()
        check(self.L.phb_ctx_create(device, C.byref(h)))
        self.h = h
        if nprocs > 1:
            buf = C.create_string_buffer(bytes(unique_id or b""), 128)
            check(self.L.phb_ctx_init_comm(self.h, rank, nprocs, buf))

    @staticmethod
    def unique_id():
        buf = C.create_string_buffer(128)
        check(_capi.lib().phb_comm_unique_id(buf))
        return bytes(buf.raw)

    def enable_peer_memory(self, grid, all_gather, max_solvers=4):
        """Switch the in-loop halo / all-reduce traffic of distributed solves to NVLink peer
        memory.  `all_gather(obj) -> list` is the launcher's object all-gather (e.g. a wrapper
        around torch.distributed.all_gather_object)."""
        n = self.nProcs()
        me = self.rank()
        info = all_gather((grid.sizes()["nCells"], [int(v) for v in grid.i32("recvOff")]))
        peer_ld = np.array([info[q][0] for q in range(n)], np.int32)
        peer_off = np.array([info[q][1][me] for q in range(n)], np.int32)
        check(self.L.phb_mesh_set_peer_layout(grid.h, _ip(peer_off), _ip(peer_ld)))
        buf = C.create_string_buffer(64)
        check(self.L.phb_ctx_peer_arena_create(self.h, int(peer_ld.max()), max_solvers, buf))
        handles = all_gather(bytes(buf.raw))
        allh = C.create_string_buffer(b"".join(handles), 64 * n)
        check(self.L.phb_ctx_peer_arena_open(self.h, allh))

    def rank(self):
        return self.L.phb_ctx_rank(self.h)

    def nProcs(self):
        return self.L.phb_ctx_nprocs(self.h)

    def sync(self):
        check(self.L.phb_ctx_sync(self.h))

    def kernel_launches(self):
        return self.L.phb_ctx_kernel_launches(self.h)

    def stream(self):
        return self.L.phb_ctx_stream(self.h)

    def close(self):
        if self.h:
            self.L.phb_ctx_destroy(self.h)
            self.h = None


class FiniteVolumeGrid2D:
    """UG/FiniteVolumeGrid2D: connectivity, geometry, canonical pattern, halo maps."""

    def __init__(self, comm, handle):
        self.comm, self.L, self.h = comm, comm.L, handle

    @classmethod
    def from_cells(cls, comm, xy, cptr, cind):
        xy = np.ascontiguousarray(xy, np.float64)
        cptr = np.ascontiguousarray(cptr, np.int32)
        cind = np.ascontiguousarray(cind, np.int32)
        h = C.c_void_p()
        check(comm.L.phb_mesh_create(comm.h, len(xy), _dp(xy), len(cptr) - 1, _ip(cptr), _ip(cind), C.byref(h)))
        return cls(comm, h)

    @classmethod
    def rectilinear(cls, comm, nx, ny, width=1.0, height=1.0, finalize=True):
        h = C.c_void_p()
        check(comm.L.phb_mesh_create_rectilinear(comm.h, nx, ny, width, height, C.byref(h)))
        g = cls(comm, h)
        return g.finalize() if finalize else g

    @classmethod
    def triangulated(cls, comm, nx, ny, width=1.0, height=1.0, finalize=True):
        h = C.c_void_p()
        check(comm.L.phb_mesh_create_triangulated(comm.h, nx, ny, width, height, C.byref(h)))
        g = cls(comm, h)
        return g.finalize() if finalize else g

    @classmethod
    def rectilinear_strip(cls, comm, nx, ny, width=1.0, height=1.0):
        """Local mesh of comm's rank for a y-strip partition (no global mesh is built)."""
        h = C.c_void_p()
        check(comm.L.phb_mesh_create_rect_strip(comm.h, nx, ny, width, height, C.byref(h)))
        return cls(comm, h)

    @classmethod
    def rectilinear_block(cls, comm, nx, ny, width, height, px, py):
        """Local mesh of comm's rank for a px x py block partition (no global mesh is built)."""
        h = C.c_void_p()
        check(comm.L.phb_mesh_create_rect_block(comm.h, nx, ny, width, height, px, py, C.byref(h)))
        return cls(comm, h)

    @classmethod
    def from_cgns(cls, comm, filename, refine=0, finalize=True):
        """CgnsUnstructuredGrid: ADF-format CGNS mesh (+ optional uniform refinement rounds)."""
        h = C.c_void_p()
        check(comm.L.phb_mesh_read_cgns(comm.h, str(filename).encode(), C.byref(h)))
        g = cls(comm, h)
        if refine > 0:
            h2 = C.c_void_p()
            check(comm.L.phb_mesh_refine(comm.h, g.h, refine, C.byref(h2)))
            g.close()
            g = cls(comm, h2)
        return g.finalize() if finalize else g

    def refined(self, levels=1, finalize=True):
        h2 = C.c_void_p()
        check(self.L.phb_mesh_refine(self.comm.h, self.h, levels, C.byref(h2)))
        g = FiniteVolumeGrid2D(self.comm, h2)
        return g.finalize() if finalize else g

    def patch_names(self):
        out = []
        for i in range(self.sizes()["nPatches"]):
            buf = C.create_string_buffer(64)
            check(self.L.phb_mesh_patch_name(self.h, i, buf, 64))
            out.append(buf.value.decode())
        return out

    def createPatchByNodes(self, name, pairs):
        pairs = np.ascontiguousarray(pairs, np.int32).reshape(-1)
        return check(self.L.phb_mesh_add_patch_by_nodes(self.h, name.encode(), len(pairs) // 2, _ip(pairs)))

    def finalize(self):
        check(self.L.phb_mesh_finalize(self.h))
        return self

    def sizes(self):
        out = (C.c_longlong * 11)()
        check(self.L.phb_mesh_sizes(self.h, out))
        k = ["nNodes", "nCells", "nFaces", "nPatches", "rank", "nProcs", "nLocal", "rowOffset",
             "nInteriorFaces", "nBoundaryFaces", "nnz"]
        return dict(zip(k, [int(v) for v in out]))

    def i32(self, name):
        n = check(self.L.phb_mesh_get_i32(self.h, name.encode(), None, 0))
        a = np.zeros(max(n, 1), np.int32)
        check(self.L.phb_mesh_get_i32(self.h, name.encode(), _ip(a), n))
        return a[:n]

    def f64(self, name):
        n = check(self.L.phb_mesh_get_f64(self.h, name.encode(), None, 0))
        a = np.zeros(max(n, 1), np.float64)
        check(self.L.phb_mesh_get_f64(self.h, name.encode(), _dp(a), n))
        return a[:n]

    def partition_rcb(self, nparts):
        part = np.zeros(self.sizes()["nCells"], np.int32)
        check(self.L.phb_partition_rcb(self.h, nparts, _ip(part)))
        return part

    def partition_metis(self, nparts, method="mesh_dual"):
        """METIS as the reference calls it: "mesh_dual" (FiniteVolumeGrid2D::partition) or "graph_recursive"
        (the PhasePartitionGrid utility).  Returns (cellPartition, edge cut)."""
        part = np.zeros(self.sizes()["nCells"], np.int32)
        obj = C.c_longlong()
        check(self.L.phb_partition_metis(self.h, nparts, 0 if method == "mesh_dual" else 1, _ip(part), C.byref(obj)))
        return part, obj.value

    def partition_file(self, part, proc, minBufferWidth=0.0):
        """The content of PhasePartitionGrid's solution/Proc<proc>/Grid.cgns (U/utilities/PhasePartitionGrid.cpp:56-153):
        dict with GlobalID, ProcNo, nodes (n x 2), eptr, eind (1-based), patches {name: 1-based node pairs}."""
        part = np.ascontiguousarray(part, np.int32)
        h = C.c_void_p()
        check(self.L.phb_partition_file_build(self.h, _ip(part), proc, minBufferWidth, C.byref(h)))
        sz = (C.c_longlong * 4)()
        check(self.L.phb_partition_file_sizes(h, sz))
        gid, pno = np.zeros(sz[0], np.int32), np.zeros(sz[0], np.int32)
        nodes, eptr, eind = np.zeros(2 * sz[1]), np.zeros(sz[0] + 1, np.int32), np.zeros(sz[2], np.int32)
        check(self.L.phb_partition_file_get(h, _ip(gid), _ip(pno), _dp(nodes), _ip(eptr), _ip(eind)))
        patches = {}
        for p in range(sz[3]):
            buf = C.create_string_buffer(64)
            n = check(self.L.phb_partition_file_patch(h, p, buf, 64, None))
            pairs = np.zeros(n, np.int32)
            check(self.L.phb_partition_file_patch(h, p, buf, 64, _ip(pairs)))
            patches[buf.value.decode()] = pairs
        self.L.phb_partition_file_destroy(h)
        return dict(GlobalID=gid, ProcNo=pno, nodes=nodes.reshape(-1, 2), eptr=eptr, eind=eind, patches=patches)

    def local(self, part, comm=None):
        comm = comm or self.comm
        part = np.ascontiguousarray(part, np.int32)
        h = C.c_void_p()
        check(self.L.phb_mesh_create_local(comm.h, self.h, _ip(part), C.byref(h)))
        return FiniteVolumeGrid2D(comm, h)

    def close(self):
        if self.h:
            self.L.phb_mesh_destroy(self.h)
            self.h = None


class SparseMatrixSolver:
    """Seam 1: the backend behind M/SparseMatrixSolver.h (type "b200")."""

    def __init__(self, comm, handle=None):
        self.comm, self.L = comm, comm.L
        self.own = handle is None
        if handle is None:
            handle = C.c_void_p()
            check(self.L.phb_solver_create(comm.h, C.byref(handle)))
        self.h = handle
        self._iters, self._err, self._n = 0, 0.0, 0

    def setup(self, parameters):
        for k, v in parameters.items():
            check(self.L.phb_solver_setup(self.h, str(k).encode(), str(v).encode()))
        return self

    def setRank(self, rows, cols=None):
        check(self.L.phb_solver_set_rank(self.h, rows, rows if cols is None else cols))

    def setHalo(self, grid):
        check(self.L.phb_solver_set_halo(self.h, grid.h, 1))

    def set(self, rowPtr, colInds, vals):
        rp = np.ascontiguousarray(rowPtr, np.int32)
        ci = np.ascontiguousarray(colInds, np.int32)
        va = np.ascontiguousarray(vals, np.float64)
        self._n = len(rp) - 1
        check(self.L.phb_solver_set_csr(self.h, self._n, _ip(rp), _ip(ci), _dp(va)))

    def setEntries(self, nrows, rows, cols, vals):
        r = np.ascontiguousarray(rows, np.int32)
        c = np.ascontiguousarray(cols, np.int32)
        v = np.ascontiguousarray(vals, np.float64)
        self._n = nrows
        check(self.L.phb_solver_set_coo(self.h, nrows, len(r), _ip(r), _ip(c), _dp(v)))

    def setRhs(self, b):
        b = np.ascontiguousarray(b, np.float64)
        check(self.L.phb_solver_set_rhs(self.h, _dp(b), len(b)))

    def setGuess(self, x0):
        x0 = np.ascontiguousarray(x0, np.float64)
        check(self.L.phb_solver_set_guess(self.h, _dp(x0), len(x0)))

    def solve(self):
        it, rr = C.c_int(), C.c_double()
        rc = self.L.phb_solver_solve(self.h, C.byref(it), C.byref(rr))
        self._iters, self._err = it.value, rr.value
        check(rc)
        return self._err

    def x(self, out=None):
        """the solution (M/SparseMatrixSolver.h: x(i)); `out` = caller's array to fill instead of a new one"""
        if out is None:
            out = np.zeros(self._n, np.float64)
        check(self.L.phb_solver_get_x(self.h, _dp(out), self._n))
        return out

    def nIters(self):
        return self._iters

    def error(self):
        return self._err

    def supportsMPI(self):
        return True

    def spmv(self, x):
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros_like(x)
        check(self.L.phb_solver_spmv(self.h, _dp(x), _dp(y), len(x)))
        return y

    def applyPreconditioner(self, r):
        """z = M^-1 r with the multigrid hierarchy of the last solve (owned rows)"""
        r = np.ascontiguousarray(r, np.float64)
        z = np.zeros_like(r)
        check(self.L.phb_solver_apply_preconditioner(self.h, _dp(r), _dp(z), len(r)))
        return z

    def time_spmv(self, reps):
        ms = C.c_double()
        check(self.L.phb_solver_time_spmv(self.h, reps, C.byref(ms)))
        return ms.value

    def bytes(self):
        out = (C.c_double * 2)()
        check(self.L.phb_solver_bytes(self.h, out))
        return float(out[0]), float(out[1])

    def amgInfo(self):
        out = (C.c_double * 8)()
        check(self.L.phb_solver_amg_info(self.h, out))
        keys = ("levels", "operatorComplexity", "setupMs", "setups", "coarsestRows", "launchesPerCycle",
                "itersAfterSetup", "stale")
        return dict(zip(keys, (float(v) for v in out)))

    def amgRefresh(self):
        """numeric re-setup of the hierarchy on the device from the resident matrix values"""
        check(self.L.phb_solver_amg_refresh(self.h))

    def amgRefreshInfo(self):
        out = (C.c_double * 8)()
        check(self.L.phb_solver_amg_refresh_info(self.h, out))
        keys = ("refreshes", "refreshMs", "itersAfterRefresh", "resident", "bytes")
        return dict(zip(keys, (float(v) for v in out)))

    def amgValues(self, level, which):
        """values of A (0), P (1), R (2), smoother weights (3) of a level or the dense coarsest inverse (4)"""
        n = check(self.L.phb_solver_amg_values(self.h, level, which, None, 0))
        out = np.zeros(max(int(n), 1), np.float64)
        check(self.L.phb_solver_amg_values(self.h, level, which, _dp(out), len(out)))
        return out[:int(n)]

    def timeAmg(self, reps=20):
        out = (C.c_double * 8)()
        check(self.L.phb_solver_time_amg(self.h, reps, out))
        keys = ("msResidual", "msRestriction", "msProlongation", "msJacobi", "msCycle", "bytesJacobi", "bytesCycle",
                "launchesPerCycle")
        return dict(zip(keys, (float(v) for v in out)))

    def close(self):
        if self.own and self.h:
            self.L.phb_solver_destroy(self.h)
        self.h = None


class FiniteVolumeField:
    """UF/FiniteVolumeField<T>: T = Scalar (nComp 1) or Vector2D (nComp 2)."""

    def __init__(self, grid, nComp=1, name="", handle=None):
        self.grid, self.L, self.nComp, self.name = grid, grid.L, nComp, name
        self.own = handle is None
        if handle is None:
            handle = C.c_void_p()
            check(self.L.phb_field_create(grid.h, nComp, name.encode(), C.byref(handle)))
        self.h = handle

    def setBoundary(self, patch, type_, value=(0.0, 0.0)):
        v = (value, 0.0) if np.isscalar(value) else value
        check(self.L.phb_field_set_bc(self.h, patch.encode(), type_, float(v[0]), float(v[1])))

    def _len(self, part):
        s = self.grid.sizes()
        return self.nComp * (s["nCells"] if part.startswith("cells") else s["nFaces"])

    def set(self, part, v):
        v = np.ascontiguousarray(v, np.float64).reshape(-1)
        check(self.L.phb_field_set(self.h, part.encode(), _dp(v), len(v)))

    def get(self, part, out=None):
        """`out`: caller's buffer (e.g. pinned host memory) to receive the values instead of a new array"""
        if out is None:
            out = np.zeros(self._len(part), np.float64)
        else:
            assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == self._len(part)
        check(self.L.phb_field_get(self.h, part.encode(), _dp(out), out.size))
        return out.reshape(self.nComp, -1) if self.nComp > 1 else out

    def fill(self, vx, vy=0.0):
        check(self.L.phb_field_fill(self.h, vx, vy))

    def savePreviousTimeStep(self):
        check(self.L.phb_field_save_previous(self.h))

    def interpolateFaces(self):
        check(self.L.phb_field_interpolate_faces(self.h))

    def setBoundaryFaces(self):
        check(self.L.phb_field_set_boundary_faces(self.h))

    def sendMessages(self):
        check(self.L.phb_field_send_messages(self.h))

    def close(self):
        if self.own and self.h:
            self.L.phb_field_destroy(self.h)
        self.h = None


def ScalarFiniteVolumeField(grid, name=""):
    return FiniteVolumeField(grid, 1, name)


def VectorFiniteVolumeField(grid, name=""):
    return FiniteVolumeField(grid, 2, name)


class ScalarGradient(FiniteVolumeField):
    """UF/ScalarGradient: face gradient + FACE_TO_CELL reconstruction."""

    def __init__(self, phi):
        super().__init__(phi.grid, 2, "grad" + phi.name)
        self.phi = phi

    def compute(self):
        check(self.L.phb_field_gradient(self.phi.h, self.h))


class cicsam:
    """namespace cicsam (UD/Cicsam.cpp)."""

    @staticmethod
    def faceInterpolationWeights(u, gamma, gradGamma, timeStep, beta):
        check(u.L.phb_cicsam_weights(u.h, gamma.h, gradGamma.h, timeStep, beta.h))
        return beta

    @staticmethod
    def computeMomentumFlux(rho1, rho2, u, gamma, beta, rhoU):
        check(u.L.phb_cicsam_momentum_flux(rho1, rho2, u.h, gamma.h, beta.h, rhoU.h))
        return rhoU


class FiniteVolumeEquation:
    """UE/FiniteVolumeEquation<T> on the canonical pattern, assembled on the device."""

    REFERENCE_COMPACT, REFERENCE_PADDED_NB_FIRST, REFERENCE_PADDED_DIAG_FIRST = 0, 1, 2

    def __init__(self, field, name="", handle=None):
        self.field, self.grid, self.L, self.name = field, field.grid, field.L, name
        self.own = handle is None
        if handle is None:
            handle = C.c_void_p()
            check(self.L.phb_eqn_create(self.grid.h, field.nComp, C.byref(handle)))
        self.h = handle
        self.solver = None

    def zero(self):
        check(self.L.phb_eqn_zero(self.h))
        return self

    def configureSparseSolver(self, parameters):
        self.solver = SparseMatrixSolver(self.grid.comm).setup(parameters)
        return self

    def ddt(self, phi, dt, rho=1.0, sign=1.0):
        rf = rho.h if isinstance(rho, FiniteVolumeField) else None
        rc = 1.0 if rf is not None else float(rho)
        check(self.L.phb_assemble_ddt(self.h, phi.h, rc, rf, dt, sign))
        return self

    def div(self, u, phi, theta=1.0, sign=1.0):
        check(self.L.phb_assemble_div(self.h, u.h, phi.h, theta, sign))
        return self

    def dive(self, u, phi, theta, sign=1.0):
        check(self.L.phb_assemble_dive(self.h, u.h, phi.h, theta, sign))
        return self

    def laplacian(self, gamma, phi, theta=None, sign=1.0):
        gf = gamma.h if isinstance(gamma, FiniteVolumeField) else None
        gc = 0.0 if gf is not None else float(gamma)
        check(self.L.phb_assemble_laplacian(self.h, gc, gf, phi.h, -1.0 if theta is None else theta, sign))
        return self

    def src(self, f, sign=1.0):
        check(self.L.phb_assemble_src(self.h, f.h, sign))
        return self

    def srcDiv(self, u, sign=1.0):
        check(self.L.phb_assemble_src_div(self.h, u.h, sign))
        return self

    def ddtCells(self, phi, dt, cells, sign=1.0):
        """fv::ddt(field, timeStep, cells) (UD/TimeDerivative.h:50-62)"""
        c = np.ascontiguousarray(cells, np.int32)
        check(self.L.phb_assemble_ddt_cells(self.h, phi.h, dt, sign, len(c), _ip(c)))
        return self

    def srcDivCells(self, u, cells, sign=1.0):
        """src::div(field, cells) (UD/Source.cpp:5-21)"""
        c = np.ascontiguousarray(cells, np.int32)
        check(self.L.phb_assemble_src_div_cells(self.h, u.h, sign, len(c), _ip(c)))
        return self

    def srcLaplacian(self, gamma, phi, sign=1.0):
        """src::laplacian(gamma, phi) (UD/Source.cpp:27-75); gamma scalar or ScalarFiniteVolumeField"""
        gf = gamma.h if isinstance(gamma, FiniteVolumeField) else None
        check(self.L.phb_assemble_src_laplacian(self.h, 0.0 if gf is not None else float(gamma), gf, phi.h, sign))
        return self

    def cicsamDiv(self, u, gamma, beta, theta, sign=1.0):
        check(self.L.phb_assemble_cicsam_div(self.h, u.h, gamma.h, beta.h, theta, sign))
        return self

    def scaleRows(self, rho):
        check(self.L.phb_eqn_scale_rows(self.h, rho.h))
        return self

    def relax(self, omega):
        check(self.L.phb_eqn_relax(self.h, self.field.h, omega))
        return self

    def export(self, layout=0):
        """(rowPtr, colInd, vals, rhs) in the reference's own CSR layout (I4)."""
        nnz = check(self.L.phb_eqn_export_csr(self.h, layout, None, None, None, None))
        n = self.field.nComp * self.grid.sizes()["nLocal"]
        rp, ci = np.zeros(n + 1, np.int32), np.zeros(max(nnz, 1), np.int32)
        va, rhs = np.zeros(max(nnz, 1), np.float64), np.zeros(n, np.float64)
        check(self.L.phb_eqn_export_csr(self.h, layout, _ip(rp), _ip(ci), _dp(va), _dp(rhs)))
        return rp, ci[:nnz], va[:nnz], rhs

    def solve(self, warmStart=False, solver=None):
        s = solver or self.solver
        it, rr = C.c_int(), C.c_double()
        rc = self.L.phb_eqn_solve(self.h, s.h, self.field.h, int(warmStart), C.byref(it), C.byref(rr))
        s._iters, s._err = it.value, rr.value
        check(rc)
        return rr.value

    def close(self):
        if self.own and self.h:
            self.L.phb_eqn_destroy(self.h)
        self.h = None


class FractionalStep:
    """US/FractionalStep: device-resident time step (the hot path's caller)."""

    def __init__(self, grid, rho=1.0, mu=1.0):
        self.grid, self.L = grid, grid.L
        h = C.c_void_p()
        check(self.L.phb_fs_create(grid.h, rho, mu, C.byref(h)))
        self.h = h
        f = lambda n, nc: FiniteVolumeField(grid, nc, n, handle=C.c_void_p(self.L.phb_fs_field(h, n.encode())))
        self.u, self.p, self.gradP = f("u", 2), f("p", 1), f("gradP", 2)
        self.uEqn = FiniteVolumeEquation(self.u, "uEqn", handle=C.c_void_p(self.L.phb_fs_eqn(h, b"uEqn")))
        self.pEqn = FiniteVolumeEquation(self.p, "pEqn", handle=C.c_void_p(self.L.phb_fs_eqn(h, b"pEqn")))
        self.uEqn.solver = SparseMatrixSolver(grid.comm, handle=C.c_void_p(self.L.phb_fs_solver(h, b"uEqn")))
        self.pEqn.solver = SparseMatrixSolver(grid.comm, handle=C.c_void_p(self.L.phb_fs_solver(h, b"pEqn")))

    def setup(self, **keys):
        for k, v in keys.items():
            check(self.L.phb_fs_setup(self.h, k.encode(), float(v)))
        return self

    def initialize(self):
        check(self.L.phb_fs_initialize(self.h))

    def assembleU(self, dt):
        check(self.L.phb_fs_assemble_u(self.h, dt))
        return self.uEqn

    def assembleP(self, dt):
        check(self.L.phb_fs_assemble_p(self.h, dt))
        return self.pEqn

    def solve(self, dt):
        st = (C.c_double * 6)()
        check(self.L.phb_fs_step(self.h, dt, st))
        return dict(itersU=int(st[0]), itersP=int(st[1]), errorU=st[2], errorP=st[3],
                    maxDivergence=st[4], maxCourant=st[5])

    def computeGradP(self):
        """gradP_.compute (UF/ScalarGradient.cpp:34-74) from the current p: what a step leaves behind, rebuilt after p
        has been replaced from the host."""
        check(self.L.phb_field_gradient(self.p.h, self.gradP.h))

    def rebuildFaces(self, dtPrev):
        """State from cell values alone (the reference's restart, US/Solver.cpp:544-581): ghosts, boundary faces, gradP and
        the face velocities of the step that ended with time step `dtPrev`, from the cells of u and p on the device."""
        check(self.L.phb_fs_rebuild_faces(self.h, float(dtPrev)))

    def computeMaxTimeStep(self, maxCo, prevDt, maxDt):
        out = C.c_double()
        check(self.L.phb_fs_max_time_step(self.h, maxCo, prevDt, maxDt, C.byref(out)))
        return out.value

    def close(self):
        if self.h:
            self.L.phb_fs_destroy(self.h)
            self.h = None


class FractionalStepMultiphase:
    """US/FractionalStepMultiphase: device-resident VOF time step (CICSAM + CELESTE surface tension), config 4.
    Set `gamma` (cells and faces) and the boundary conditions, configure the three solvers, then initialize()."""

    FIELDS = {"u": 2, "p": 1, "gradP": 2, "gamma": 1, "gradGamma": 2, "rho": 1, "mu": 1, "beta": 1, "sg": 2, "fst": 2,
              "kappa": 1, "gammaTilde": 1, "gradGammaTilde": 2, "n": 2, "gradRho": 2}

    def __init__(self, grid, rho1, rho2, mu1, mu2, sigma, g=(0.0, 0.0), smoothingKernelRadius=1.0, **keys):
        self.grid, self.L = grid, grid.L
        h = C.c_void_p()
        check(self.L.phb_mp_create(grid.h, rho1, rho2, mu1, mu2, sigma, float(g[0]), float(g[1]), smoothingKernelRadius,
                                   C.byref(h)))
        self.h = h
        for name, nc in self.FIELDS.items():
            setattr(self, name, FiniteVolumeField(grid, nc, name, handle=C.c_void_p(self.L.phb_mp_field(h, name.encode()))))
        for e, fld in (("gammaEqn", self.gamma), ("uEqn", self.u), ("pEqn", self.p)):
            eq = FiniteVolumeEquation(fld, e, handle=C.c_void_p(self.L.phb_mp_eqn(h, e.encode())))
            eq.solver = SparseMatrixSolver(grid.comm, handle=C.c_void_p(self.L.phb_mp_solver(h, e.encode())))
            setattr(self, e, eq)
        for k, v in keys.items():
            self.setup(k, v)

    def setup(self, key, value):
        check(self.L.phb_mp_setup(self.h, key.encode(), float(value)))

    def initialize(self):
        check(self.L.phb_mp_initialize(self.h))

    def solve(self, dt):
        st = (C.c_double * 8)()
        check(self.L.phb_mp_step(self.h, dt, st))
        return dict(itersGamma=int(st[0]), itersU=int(st[1]), itersP=int(st[2]), errorGamma=st[3], errorU=st[4],
                    errorP=st[5], maxDivergence=st[6], maxCourant=st[7])

    def close(self):
        if self.h:
            self.L.phb_mp_destroy(self.h)
            self.h = None


class Piso:
    """phasePiso: device-resident PISO/SIMPLE-type time step (README.md:26-37; self-consistent parity)."""

    def __init__(self, grid, rho=1.0, mu=1.0, **keys):
        self.grid, self.L = grid, grid.L
        h = C.c_void_p()
        check(self.L.phb_piso_create(grid.h, rho, mu, C.byref(h)))
        self.h = h
        f = lambda n, nc: FiniteVolumeField(grid, nc, n, handle=C.c_void_p(self.L.phb_piso_field(h, n.encode())))
        self.u, self.p, self.pCorr, self.gradP, self.d = f("u", 2), f("p", 1), f("pCorr", 1), f("gradP", 2), f("d", 1)
        self.uSolver = SparseMatrixSolver(grid.comm, handle=C.c_void_p(self.L.phb_piso_solver(h, b"uEqn")))
        self.pCorrSolver = SparseMatrixSolver(grid.comm, handle=C.c_void_p(self.L.phb_piso_solver(h, b"pCorrEqn")))
        for k, v in keys.items():
            check(self.L.phb_piso_setup(h, k.encode(), float(v)))

    def initialize(self):
        check(self.L.phb_piso_initialize(self.h))

    def solve(self, dt):
        st = (C.c_double * 6)()
        check(self.L.phb_piso_step(self.h, dt, st))
        return dict(itersU=int(st[0]), itersPCorr=int(st[1]), errorU=st[2], errorPCorr=st[3],
                    maxMassImbalance=st[4], maxCourant=st[5])

    def close(self):
        if self.h:
            self.L.phb_piso_destroy(self.h)
            self.h = None


def lid_driven_cavity(grid, rho=1.0, mu=0.1, lid=1.0, solver=None, pSolver=None):
    """Examples/LidDrivenCavity/case/boundaries.info on any grid with x-/x+/y-/y+ patches.
    `solver` = LinearAlgebra keys of both equations, `pSolver` = overrides for pEqn (e.g. preconditioner amg)."""
    fs = FractionalStep(grid, rho, mu)
    for pt in ("x-", "x+", "y-"):
        fs.u.setBoundary(pt, FIXED, (0.0, 0.0))
    fs.u.setBoundary("y+", FIXED, (lid, 0.0))
    for pt in ("x-", "x+", "y-", "y+"):
        fs.p.setBoundary(pt, NORMAL_GRADIENT, 0.0)
    cfg = dict(solver="BICGSTAB", maxIters=20000, tolerance=1e-10, preconditioner="ilu0")
    cfg.update(solver or {})
    fs.uEqn.solver.setup(cfg)
    pcfg = dict(cfg)
    pcfg.update(pSolver or {})
    fs.pEqn.solver.setup(pcfg)
    fs.initialize()
    return fs
