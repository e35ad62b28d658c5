"""CPU oracle for phase_b200 -- TEST INFRASTRUCTURE ONLY.

ctypes front-end over oracle/phase_oracle.c (our flat-array restatement of the
reference's hot path) and, when available, oracle/_ref (the reference's own
``src/Math`` translation units compiled in place).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this package; the product (``phase_b200``) never does.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

FIXED, NORMAL_GRADIENT, SYMMETRY = 0, 1, 2

_lib = None
_ref = None

SOLVE_CB = C.CFUNCTYPE(C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                       C.POINTER(C.c_double), C.POINTER(C.c_double),
                       C.POINTER(C.c_double), C.c_void_p)


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def lib():
    global _lib
    if _lib is None:
        so = _build.ORACLE_SO
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(
                os.path.join(_build.HERE, "phase_oracle.c")):
            so = _build.build_oracle()
        L = C.CDLL(so)
        vp = C.c_void_p
        L.or_mesh_create.restype = vp
        L.or_mesh_create.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int,
                                     C.POINTER(C.c_int), C.POINTER(C.c_int)]
        for f in (L.or_mesh_rectilinear, L.or_mesh_triangulated):
            f.restype = vp
            f.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
        L.or_mesh_destroy.argtypes = [vp]
        L.or_mesh_add_patch_by_nodes.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        L.or_mesh_patch_id.argtypes = [vp, C.c_char_p]
        L.or_mesh_array.restype = C.c_long
        L.or_mesh_array.argtypes = [vp, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
        L.or_mesh_partition_local.restype = vp
        L.or_mesh_partition_local.argtypes = [vp, C.POINTER(C.c_int), C.c_int, C.c_int]
        L.or_mesh_init_comm.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.or_crs_create.restype = vp
        L.or_crs_create.argtypes = [C.c_int, C.c_int]
        L.or_crs_clone.restype = vp
        L.or_crs_clone.argtypes = [vp]
        L.or_crs_destroy.argtypes = [vp]
        L.or_crs_add_coeff.argtypes = [vp, C.c_int, C.c_int, C.c_double]
        L.or_crs_set_coeff.argtypes = [vp, C.c_int, C.c_int, C.c_double]
        L.or_crs_add_rhs.argtypes = [vp, C.c_int, C.c_double]
        L.or_crs_scale_row.argtypes = [vp, C.c_int, C.c_double]
        L.or_crs_add_eq.argtypes = [vp, vp]
        L.or_crs_sub_eq.argtypes = [vp, vp]
        L.or_crs_sub_vec.argtypes = [vp, C.POINTER(C.c_double)]
        L.or_crs_add_vec.argtypes = [vp, C.POINTER(C.c_double)]
        L.or_crs_scale.argtypes = [vp, C.c_double]
        L.or_crs_rank.argtypes = [vp]
        L.or_crs_nnz.argtypes = [vp]
        L.or_crs_export.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                    C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.or_fs_create.restype = vp
        L.or_fs_create.argtypes = [vp, C.c_double, C.c_double]
        L.or_fs_destroy.argtypes = [vp]
        L.or_fs_set_bc.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int, C.c_double, C.c_double]
        L.or_fs_initialize.argtypes = [vp]
        L.or_fs_set_solver.argtypes = [vp, SOLVE_CB, vp]
        L.or_fs_set_solver_params.argtypes = [vp, C.c_double, C.c_int, C.c_int]
        L.or_fs_step.argtypes = [vp, C.c_double]
        L.or_fs_assemble_u.argtypes = [vp, C.c_double]
        L.or_fs_assemble_p.argtypes = [vp, C.c_double]
        L.or_fs_ueqn.restype = vp
        L.or_fs_ueqn.argtypes = [vp]
        L.or_fs_peqn.restype = vp
        L.or_fs_peqn.argtypes = [vp]
        L.or_fs_array.restype = C.c_long
        L.or_fs_array.argtypes = [vp, C.c_char_p, C.POINTER(C.POINTER(C.c_double))]
        L.or_fs_max_divergence.restype = C.c_double
        L.or_fs_max_divergence.argtypes = [vp]
        L.or_fs_max_courant.restype = C.c_double
        L.or_fs_max_courant.argtypes = [vp, C.c_double]
        L.or_fs_last_iters.argtypes = [vp, C.c_int]
        L.or_op_laplacian_field.restype = vp
        L.or_op_laplacian_field.argtypes = [vp, C.POINTER(C.c_double)]
        L.or_op_ueqn_multiphase.restype = vp
        L.or_op_ueqn_multiphase.argtypes = [vp, C.c_double] + [C.POINTER(C.c_double)] * 5
        L.or_op_scalar_transport.restype = vp
        L.or_op_scalar_transport.argtypes = [vp, C.c_double, C.c_double] + [C.POINTER(C.c_double)] * 4
        L.or_bicgstab.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_double), C.POINTER(C.c_double),
                                  C.POINTER(C.c_double), C.c_double, C.c_int, C.c_int,
                                  C.POINTER(C.c_double)]
        L.or_cicsam_weights.argtypes = [vp, C.c_double] + [C.POINTER(C.c_double)] * 3
        L.or_op_cicsam_div.restype = vp
        L.or_op_cicsam_div.argtypes = [vp, C.c_double] + [C.POINTER(C.c_double)] * 3
        L.or_cicsam_momentum_flux.argtypes = [vp, C.c_double, C.c_double] + [C.POINTER(C.c_double)] * 3
        L.or_multicolor_order.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.POINTER(C.c_int), C.c_int]
        L.or_bicgstab_blocks.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                         C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_int, C.c_int,
                                         C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.or_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def set_num_threads(n=None):
    """OpenMP threads of the oracle's solver (default: every host core, whatever OMP_NUM_THREADS the launcher
    exported).  Returns the count in effect."""
    return lib().or_set_num_threads(int(n or os.cpu_count() or 1))


def ref_lib():
    """The reference's own CrsEquation algebra (oracle/_ref) or None."""
    global _ref
    if _ref is None:
        so = _build.build_ref()
        if so is None:
            return None
        R = C.CDLL(so)
        vp = C.c_void_p
        R.ref_crs_create.restype = vp
        R.ref_crs_create.argtypes = [C.c_int, C.c_int]
        R.ref_crs_clone.restype = vp
        R.ref_crs_clone.argtypes = [vp]
        R.ref_crs_destroy.argtypes = [vp]
        R.ref_crs_add_coeff.argtypes = [vp, C.c_int, C.c_int, C.c_double]
        R.ref_crs_set_coeff.argtypes = [vp, C.c_int, C.c_int, C.c_double]
        R.ref_crs_add_rhs.argtypes = [vp, C.c_int, C.c_double]
        R.ref_crs_scale_row.argtypes = [vp, C.c_int, C.c_double]
        R.ref_crs_add_eq.argtypes = [vp, vp]
        R.ref_crs_sub_eq.argtypes = [vp, vp]
        R.ref_crs_sub_vec.argtypes = [vp, C.POINTER(C.c_double), C.c_int]
        R.ref_crs_scale.argtypes = [vp, C.c_double]
        R.ref_crs_rank.argtypes = [vp]
        R.ref_crs_nnz.argtypes = [vp]
        R.ref_crs_export.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.POINTER(C.c_double), C.POINTER(C.c_double)]
        R.ref_crs_solve_handoff.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                            C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _ref = R
    return _ref


class Mesh:
    """Flat mesh in the reference's numbering (I1, I2, G1-G4)."""

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def create(cls, xy, cptr, cind):
        xy = np.ascontiguousarray(xy, dtype=np.float64)
        cptr = np.ascontiguousarray(cptr, dtype=np.int32)
        cind = np.ascontiguousarray(cind, dtype=np.int32)
        return cls(lib().or_mesh_create(len(xy), _dp(xy), len(cptr) - 1, _ip(cptr), _ip(cind)))

    @classmethod
    def rectilinear(cls, nx, ny, width=1.0, height=1.0):
        return cls(lib().or_mesh_rectilinear(nx, ny, width, height))

    @classmethod
    def triangulated(cls, nx, ny, width=1.0, height=1.0):
        return cls(lib().or_mesh_triangulated(nx, ny, width, height))

    def add_patch_by_nodes(self, name, pairs):
        pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1)
        return lib().or_mesh_add_patch_by_nodes(self.h, name.encode(), len(pairs) // 2, _ip(pairs))

    def patch_id(self, name):
        return lib().or_mesh_patch_id(self.h, name.encode())

    def array(self, name):
        ptr = C.c_void_p()
        isd = C.c_int()
        n = lib().or_mesh_array(self.h, name.encode(), C.byref(ptr), C.byref(isd))
        if n < 0:
            raise KeyError(name)
        if n == 0:
            return np.zeros(0, dtype=np.float64 if isd.value else np.int32)
        ct = C.c_double if isd.value else C.c_int
        buf = C.cast(ptr, C.POINTER(ct * n)).contents
        return np.frombuffer(buf, dtype=np.float64 if isd.value else np.int32).copy()

    @property
    def sizes(self):
        s = self.array("sizes")
        return dict(nNodes=int(s[0]), nCells=int(s[1]), nFaces=int(s[2]), nPatches=int(s[3]),
                    rank=int(s[4]), nProcs=int(s[5]), nLocal=int(s[6]), rowOffset=int(s[7]))

    def partition(self, part, nprocs):
        """All local meshes + halo maps for a given cell-partition vector (I5)."""
        part = np.ascontiguousarray(part, dtype=np.int32)
        locs = [Mesh(lib().or_mesh_partition_local(self.h, _ip(part), r, nprocs))
                for r in range(nprocs)]
        arr = (C.c_void_p * nprocs)(*[m.h for m in locs])
        lib().or_mesh_init_comm(arr, nprocs)
        return locs

    def __del__(self):
        try:
            if self.h:
                lib().or_mesh_destroy(self.h)
                self.h = None
        except Exception:
            pass


class _CrsBase:
    def export(self):
        n, nnz = self.rank(), self.nnz()
        rp = np.zeros(n + 1, np.int32)
        ci = np.zeros(max(nnz, 1), np.int32)
        va = np.zeros(max(nnz, 1), np.float64)
        rhs = np.zeros(max(n, 1), np.float64)
        self._export(rp, ci, va, rhs)
        return rp, ci[:nnz], va[:nnz], rhs[:n]


class Crs(_CrsBase):
    """Our restatement of CrsEquation (M/CrsEquation.cpp)."""

    def __init__(self, n=None, nnz=5, handle=None, own=True):
        self.h = handle if handle is not None else lib().or_crs_create(n, nnz)
        self.own = own

    def clone(self):
        return Crs(handle=lib().or_crs_clone(self.h))

    def add_coeff(self, r, c, v): lib().or_crs_add_coeff(self.h, r, c, v)
    def set_coeff(self, r, c, v): lib().or_crs_set_coeff(self.h, r, c, v)
    def add_rhs(self, r, v): lib().or_crs_add_rhs(self.h, r, v)
    def scale_row(self, r, v): lib().or_crs_scale_row(self.h, r, v)
    def add_eq(self, o): lib().or_crs_add_eq(self.h, o.h)
    def sub_eq(self, o): lib().or_crs_sub_eq(self.h, o.h)

    def sub_vec(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        lib().or_crs_sub_vec(self.h, _dp(v))

    def scale(self, s): lib().or_crs_scale(self.h, s)
    def rank(self): return lib().or_crs_rank(self.h)
    def nnz(self): return lib().or_crs_nnz(self.h)
    def _export(self, rp, ci, va, rhs): lib().or_crs_export(self.h, _ip(rp), _ip(ci), _dp(va), _dp(rhs))

    def __del__(self):
        try:
            if self.own and self.h:
                lib().or_crs_destroy(self.h)
        except Exception:
            pass


class RefCrs(_CrsBase):
    """The reference's own CrsEquation (compiled in place, oracle/_ref)."""

    def __init__(self, n=None, nnz=5, handle=None):
        self.h = handle if handle is not None else ref_lib().ref_crs_create(n, nnz)

    def clone(self):
        return RefCrs(handle=ref_lib().ref_crs_clone(self.h))

    def add_coeff(self, r, c, v): ref_lib().ref_crs_add_coeff(self.h, r, c, v)
    def set_coeff(self, r, c, v): ref_lib().ref_crs_set_coeff(self.h, r, c, v)
    def add_rhs(self, r, v): ref_lib().ref_crs_add_rhs(self.h, r, v)
    def scale_row(self, r, v): ref_lib().ref_crs_scale_row(self.h, r, v)
    def add_eq(self, o): ref_lib().ref_crs_add_eq(self.h, o.h)
    def sub_eq(self, o): ref_lib().ref_crs_sub_eq(self.h, o.h)

    def sub_vec(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        ref_lib().ref_crs_sub_vec(self.h, _dp(v), len(v))

    def scale(self, s): ref_lib().ref_crs_scale(self.h, s)
    def rank(self): return ref_lib().ref_crs_rank(self.h)
    def nnz(self): return ref_lib().ref_crs_nnz(self.h)
    def _export(self, rp, ci, va, rhs): ref_lib().ref_crs_export(self.h, _ip(rp), _ip(ci), _dp(va), _dp(rhs))

    def solve_handoff(self):
        n, nnz = self.rank(), self.nnz()
        rp = np.zeros(n + 1, np.int32)
        ci = np.zeros(max(nnz, 1), np.int32)
        va = np.zeros(max(nnz, 1), np.float64)
        b = np.zeros(max(n, 1), np.float64)
        ref_lib().ref_crs_solve_handoff(self.h, _ip(rp), _ip(ci), _dp(va), _dp(b))
        return rp, ci[:nnz], va[:nnz], b[:n]

    def __del__(self):
        try:
            if self.h:
                ref_lib().ref_crs_destroy(self.h)
        except Exception:
            pass


def csr_to_scipy(rp, ci, va, n=None):
    """CSR with -1 padding (what SparseMatrixSolver::set receives) -> scipy."""
    import scipy.sparse as sp
    n = len(rp) - 1 if n is None else n
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    keep = ci >= 0
    return sp.csr_matrix((va[keep], (rows[keep], ci[keep])), shape=(len(rp) - 1, n))


_lu_cache = {}


def _factor(A, key):
    """splu with the minimum-degree ordering of A^T + A (the matrices here are structurally symmetric; the default
    COLAMD ordering needs 30x longer on a 1M-row 5-point system); the last few factorisations are kept, keyed on
    the matrix values, because uEqn_ and pEqn_ of the fractional step do not change between time steps."""
    import scipy.sparse.linalg as spl
    lu = _lu_cache.get(key)
    if lu is None:
        if len(_lu_cache) >= 4:
            _lu_cache.pop(next(iter(_lu_cache)))
        lu = _lu_cache[key] = spl.splu(A.tocsc(), permc_spec="MMD_AT_PLUS_A")
    return lu


def direct_solve(rp, ci, va, b):
    """Exact sparse LU (SuperLU): stand-in for the snapshot's Eigen SparseLU
    (M/EigenSparseMatrixSolver.cpp:60-64).  Singular all-Neumann systems are
    regularised by pinning the constant mode (hard part 3): bordered system [A 1; 1^T 0]
    (zero-mean solution) for small systems, first unknown pinned to zero for large ones (a dense border row
    defeats the fill-reducing ordering) -- callers compare such fields minus their mean.
    A vector system [x-block | y-block] whose two diagonal blocks are identical and uncoupled (every equation
    here without SYMMETRY / PARTIAL_SLIP patches) is factorised once."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    A = csr_to_scipy(rp, ci, va).tocsr()
    n = A.shape[0]
    key = (n, A.nnz, hash(A.data.tobytes()), hash(A.indices.tobytes()))
    if n % 2 == 0 and n >= 20000:
        h = n // 2
        Axx, Ayy = A[:h, :h], A[h:, h:]
        if Axx.nnz + Ayy.nnz == A.nnz and Axx.nnz == Ayy.nnz and np.array_equal(Axx.indices, Ayy.indices) \
                and np.array_equal(Axx.data, Ayy.data):
            lu = _factor(Axx, key)
            return np.concatenate([lu.solve(b[:h]), lu.solve(b[h:])])
    rs = np.abs(A @ np.ones(n)).max()
    if rs < 1e-12 * np.abs(A.diagonal()).max():
        if n >= 50000:
            lu = _factor(A[1:, 1:], key)
            return np.concatenate([[0.0], lu.solve(b[1:] - b.mean())])
        # constant null space: bordered system  [A 1; 1^T 0]
        one = sp.csc_matrix(np.ones((n, 1)))
        K = sp.bmat([[A, one], [one.T, None]], format="csc")
        x = spl.splu(K).solve(np.concatenate([b, [0.0]]))
        return x[:n]
    if n >= 20000:
        return _factor(A, key).solve(b)
    return spl.splu(A.tocsc()).solve(b)


class FracStep:
    """FractionalStep::solve restated (US/FractionalStep.cpp)."""

    def __init__(self, mesh, rho=1.0, mu=1.0):
        self.mesh = mesh
        self.h = lib().or_fs_create(mesh.h, rho, mu)
        self._cb = None

    def set_bc(self, field, patch, type_, vx=0.0, vy=0.0):
        rc = lib().or_fs_set_bc(self.h, field.encode(), patch.encode(), type_, vx, vy)
        if rc:
            raise ValueError("bad bc %s %s" % (field, patch))

    def initialize(self):
        lib().or_fs_initialize(self.h)

    def use_direct_solver(self):
        def cb(n, rp, ci, va, b, x, user):
            rp_ = np.ctypeslib.as_array(rp, (n + 1,))
            nnz = int(rp_[n])
            ci_ = np.ctypeslib.as_array(ci, (nnz,))
            va_ = np.ctypeslib.as_array(va, (nnz,))
            b_ = np.ctypeslib.as_array(b, (n,))
            x_ = np.ctypeslib.as_array(x, (n,))
            x_[:] = direct_solve(rp_, ci_, va_, b_)
            return 1
        self._cb = SOLVE_CB(cb)
        lib().or_fs_set_solver(self.h, self._cb, None)

    def use_ilu0_solver(self, tol=1e-8, max_iters=20000, null_space=True, guesses=None):
        """Every solve of step() = right-preconditioned BiCGStab + ILU(0) (the Belos/Ifpack2 RILUK(0) role,
        M/TrilinosBelosSparseMatrixSolver.cpp:52-83) in the multicolour ordering, OpenMP over the rows of a colour,
        converged to `tol` on ||r||/||b||, warm-started from the equation's previous solution as the Tpetra
        solution vector of the reference backend is (M/TrilinosSparseMatrixSolver.cpp).  The permutation is
        cached per pattern.  `guesses` = {rows: x0} seeds the first solve of an equation (a state taken over from
        elsewhere).  self.solve_log collects (rows, iterations, relres, seconds) per solve."""
        import time
        cache, self.solve_log = {}, []
        guesses = dict(guesses or {})

        def cb(n, rp, ci, va, b, x, user):
            t0 = time.perf_counter()
            rp_ = np.ctypeslib.as_array(rp, (n + 1,))
            nnz = int(rp_[n])
            ci_ = np.ctypeslib.as_array(ci, (nnz,))
            va_ = np.ctypeslib.as_array(va, (nnz,))
            b_ = np.ctypeslib.as_array(b, (n,))
            x_ = np.ctypeslib.as_array(x, (n,))
            c = cache.get(n)
            if c is None or c["nnz"] != nnz or not np.array_equal(c["ci"], ci_):
                rp2, ci2, _, new2old, bp = multicolor_permute(rp_, ci_, va_)
                rows = np.repeat(np.arange(n, dtype=np.int32), np.diff(rp_))
                keep = np.flatnonzero(ci_ >= 0)
                old2new = np.empty(n, np.int32)
                old2new[new2old] = np.arange(n, dtype=np.int32)
                order = np.argsort(old2new[rows[keep]], kind="stable")
                c = cache[n] = dict(nnz=nnz, ci=ci_.copy(), rp2=rp2, ci2=ci2, new2old=new2old, bp=bp,
                                    src=np.ascontiguousarray(keep[order]), x=np.zeros(n))
                if n in guesses:
                    c["x"] = np.ascontiguousarray(np.asarray(guesses.pop(n), np.float64)[new2old])
            va2 = np.ascontiguousarray(va_[c["src"]])
            b2 = np.ascontiguousarray(b_[c["new2old"]])
            if null_space:
                rs = np.abs(np.add.reduceat(np.where(ci_ >= 0, va_, 0.0), rp_[:-1])).max()
                if rs < 1e-12 * np.abs(va_).max():
                    b2 -= b2.mean()                # all-Neumann pressure: compatible right-hand side
            x2 = c["x"]
            rr = C.c_double()
            it = lib().or_bicgstab_blocks(n, _ip(c["rp2"]), _ip(c["ci2"]), _dp(va2), _dp(b2), _dp(x2), tol, max_iters,
                                          len(c["bp"]) - 1, _ip(c["bp"]), C.byref(rr))
            x_[c["new2old"]] = x2
            self.solve_log.append((n, it, rr.value, time.perf_counter() - t0))
            return it
        self._cb = SOLVE_CB(cb)
        lib().or_fs_set_solver(self.h, self._cb, None)

    def set_solver_params(self, tol=1e-10, max_iters=20000, precond=1):
        lib().or_fs_set_solver_params(self.h, tol, max_iters, precond)

    def step(self, dt):
        return lib().or_fs_step(self.h, dt)

    def assemble_u(self, dt):
        lib().or_fs_assemble_u(self.h, dt)
        return Crs(handle=lib().or_fs_ueqn(self.h), own=False)

    def assemble_p(self, dt):
        lib().or_fs_assemble_p(self.h, dt)
        return Crs(handle=lib().or_fs_peqn(self.h), own=False)

    def laplacian_field(self, gamma_face):
        g = np.ascontiguousarray(gamma_face, dtype=np.float64)
        return Crs(handle=lib().or_op_laplacian_field(self.h, _dp(g)))

    def ueqn_multiphase(self, dt, rho_cell, mu_face, mu0_face, fx, fy):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (rho_cell, mu_face, mu0_face, fx, fy)]
        return Crs(handle=lib().or_op_ueqn_multiphase(self.h, dt, *[_dp(v) for v in a]))

    def scalar_transport(self, dt, theta, rho, rho0, phi0, phi0f):
        a = [np.ascontiguousarray(v, dtype=np.float64) for v in (rho, rho0, phi0, phi0f)]
        return Crs(handle=lib().or_op_scalar_transport(self.h, dt, theta, *[_dp(v) for v in a]))

    def cicsam_weights(self, dt, gx, gy):
        gx, gy = np.ascontiguousarray(gx, np.float64), np.ascontiguousarray(gy, np.float64)
        beta = np.zeros(self.mesh.sizes["nFaces"])
        lib().or_cicsam_weights(self.h, dt, _dp(gx), _dp(gy), _dp(beta))
        return beta

    def cicsam_div(self, theta, beta, gamma0, gamma0f):
        a = [np.ascontiguousarray(v, np.float64) for v in (beta, gamma0, gamma0f)]
        return Crs(handle=lib().or_op_cicsam_div(self.h, theta, *[_dp(v) for v in a]))

    def cicsam_momentum_flux(self, rho1, rho2, beta):
        beta = np.ascontiguousarray(beta, np.float64)
        F = self.mesh.sizes["nFaces"]
        ox, oy = np.zeros(F), np.zeros(F)
        lib().or_cicsam_momentum_flux(self.h, rho1, rho2, _dp(beta), _dp(ox), _dp(oy))
        return ox, oy

    def view(self, name):
        """Writable numpy view of a field array (ux, uy, ufx, ufy, p, pf, gpx, ...)."""
        ptr = C.POINTER(C.c_double)()
        n = lib().or_fs_array(self.h, name.encode(), C.byref(ptr))
        if n < 0:
            raise KeyError(name)
        return np.ctypeslib.as_array(ptr, (n,))

    def max_divergence(self):
        return lib().or_fs_max_divergence(self.h)

    def max_courant(self, dt):
        return lib().or_fs_max_courant(self.h, dt)

    def last_iters(self, which):
        return lib().or_fs_last_iters(self.h, which)

    def __del__(self):
        try:
            if self.h:
                lib().or_fs_destroy(self.h)
                self.h = None
        except Exception:
            pass


def bicgstab(rp, ci, va, b, x0=None, tol=1e-8, max_iters=10000, precond=1):
    rp = np.ascontiguousarray(rp, np.int32)
    ci = np.ascontiguousarray(ci, np.int32)
    va = np.ascontiguousarray(va, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    x = np.zeros_like(b) if x0 is None else np.array(x0, dtype=np.float64)
    rr = C.c_double()
    it = lib().or_bicgstab(len(b), _ip(rp), _ip(ci), _dp(va), _dp(b), _dp(x), tol, max_iters,
                           precond, C.byref(rr))
    return x, it, rr.value


def cavity(mesh, rho=1.0, mu=0.1, lid=1.0):
    """Lid-driven cavity BCs of Examples/LidDrivenCavity/case/boundaries.info."""
    fs = FracStep(mesh, rho, mu)
    for pt in ("x-", "x+", "y-"):
        fs.set_bc("u", pt, FIXED, 0.0, 0.0)
    fs.set_bc("u", "y+", FIXED, lid, 0.0)
    for pt in ("x-", "x+", "y-", "y+"):
        fs.set_bc("p", pt, NORMAL_GRADIENT, 0.0)
    fs.initialize()
    return fs


def multicolor_permute(rp, ci, va):
    """Compact the CSR (drop -1 padding) and permute it symmetrically into the multicolour ordering the
    CUDA path uses for ILU(0).  Returns (rp, ci, va, new2old, blockPtr)."""
    rp = np.ascontiguousarray(rp, np.int32)
    ci = np.ascontiguousarray(ci, np.int32)
    va = np.ascontiguousarray(va, np.float64)
    n = len(rp) - 1
    new2old = np.zeros(n, np.int32)
    bp = np.zeros(66, np.int32)
    nc = lib().or_multicolor_order(n, _ip(rp), _ip(ci), _ip(new2old), _ip(bp), 64)
    if nc < 0:
        raise RuntimeError("more than 64 colours")
    old2new = np.empty(n, np.int32)
    old2new[new2old] = np.arange(n, dtype=np.int32)
    rows = np.repeat(np.arange(n, dtype=np.int32), np.diff(rp))
    keep = ci >= 0
    r2, c2, v2 = old2new[rows[keep]], old2new[ci[keep]], va[keep]
    order = np.argsort(r2, kind="stable")
    r2, c2, v2 = r2[order], c2[order], v2[order]
    rp2 = np.zeros(n + 1, np.int32)
    np.add.at(rp2, r2 + 1, 1)
    rp2 = np.cumsum(rp2).astype(np.int32)
    return rp2, np.ascontiguousarray(c2), np.ascontiguousarray(v2), new2old, np.ascontiguousarray(bp[:nc + 1])


def bicgstab_ilu0_multicolor(rp, ci, va, b, tol=1e-8, max_iters=10000):
    """BiCGStab + ILU(0) in the multicolour ordering (the CUDA path's default algorithm) on the CPU."""
    rp2, ci2, va2, new2old, bp = multicolor_permute(rp, ci, va)
    b2 = np.ascontiguousarray(np.asarray(b, np.float64)[new2old])
    x2 = np.zeros_like(b2)
    rr = C.c_double()
    it = lib().or_bicgstab_blocks(len(b2), _ip(rp2), _ip(ci2), _dp(va2), _dp(b2), _dp(x2), tol, max_iters,
                                  len(bp) - 1, _ip(bp), C.byref(rr))
    x = np.empty_like(x2)
    x[new2old] = x2
    return x, it, rr.value
