"""The REFERENCE'S OWN finite-volume path, compiled in place (oracle/_ref/libphase_ref_fv.so) -- TEST INFRASTRUCTURE.

ctypes front end over oracle/ref_fv_driver.cpp: StructuredRectilinearGrid / FiniteVolumeGrid2D, the fields, the
fv:: / src:: operators and FractionalStep::solve exactly as /root/reference/src compiles them (over the stand-in
headers of oracle/ref_stub; one MPI rank; area / centroid of a polygon are the stand-in's shoelace formulas).
Used to pin oracle/phase_oracle.c (tests/test_oracle_ref_fv.py) and to write tests/golden/ref_*.npz
(tests/golden/make_ref_golden.py).  Only available where /root/reference is mounted or a prebuilt .so travelled.
"""
import ctypes as C
import os
import tempfile

import numpy as np

from . import build as _build
from . import SOLVE_CB, direct_solve

_lib = None


def lib():
    """The library, or None when it cannot be had (no reference sources and no prebuilt .so)."""
    global _lib
    if _lib is None:
        so = _build.build_ref_fv()
        if so is None or not os.path.exists(so):
            return None
        L = C.CDLL(so)
        vp = C.c_void_p
        L.rfv_last_error.restype = C.c_char_p
        L.rfv_set_solver.argtypes = [SOLVE_CB, vp]
        L.rfv_case_open.restype = vp
        L.rfv_case_open.argtypes = [C.c_char_p]
        L.rfv_case_close.argtypes = [vp]
        L.rfv_grid_rectilinear.restype = vp
        L.rfv_grid_rectilinear.argtypes = [vp]
        L.rfv_grid_create.restype = vp
        L.rfv_grid_create.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.rfv_grid_patch_by_nodes.restype = C.c_long
        L.rfv_grid_patch_by_nodes.argtypes = [vp, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
        L.rfv_grid_close.argtypes = [vp]
        L.rfv_grid_get.restype = C.c_long
        L.rfv_grid_get.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.rfv_index_map.restype = C.c_long
        L.rfv_index_map.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.rfv_fs_create.restype = vp
        L.rfv_fs_create.argtypes = [vp, vp]
        L.rfv_fsm_create.restype = vp
        L.rfv_fsm_create.argtypes = [vp, vp]
        L.rfv_fs_initialize.restype = C.c_long
        L.rfv_fs_initialize.argtypes = [vp]
        L.rfv_fs_any_field.restype = C.c_long
        L.rfv_fs_any_field.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int]
        L.rfv_src_laplacian_scalar.restype = C.c_long
        L.rfv_src_laplacian_scalar.argtypes = [vp, C.c_double, C.POINTER(C.c_double)]
        L.rfv_src_div_cells.restype = C.c_long
        L.rfv_src_div_cells.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        L.rfv_ddt_cells.restype = C.c_long
        L.rfv_ddt_cells.argtypes = [vp, C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.rfv_partition_run.restype = C.c_long
        L.rfv_partition_run.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int]
        L.rfv_partition_get.restype = C.c_long
        L.rfv_partition_get.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int)]
        L.rfv_fs_close.argtypes = [vp]
        L.rfv_fs_step.restype = C.c_long
        L.rfv_fs_step.argtypes = [vp, C.c_double]
        for f in (L.rfv_fs_max_divergence,):
            f.restype = C.c_double
            f.argtypes = [vp]
        L.rfv_fs_max_courant.restype = C.c_double
        L.rfv_fs_max_courant.argtypes = [vp, C.c_double]
        L.rfv_fs_max_time_step.restype = C.c_double
        L.rfv_fs_max_time_step.argtypes = [vp, C.c_double, C.c_double]
        L.rfv_fs_field.restype = C.c_long
        L.rfv_fs_field.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double), C.c_int]
        L.rfv_fs_handoff.restype = C.c_long
        L.rfv_fs_handoff.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double),
                                     C.POINTER(C.c_double), C.POINTER(C.c_long)]
        _lib = L
    return _lib


def available():
    return lib() is not None


def _check(rc):
    if rc is None or (isinstance(rc, int) and rc < 0):
        raise RuntimeError("reference: " + lib().rfv_last_error().decode())
    return rc


INT_ARRAYS = ("sizes", "cptr", "cind", "faceN1", "faceN2", "faceL", "faceR", "ilPtr", "ilFace", "ilCell", "blPtr",
              "blFace", "dlPtr", "dlCell")


class Case:
    """A case directory in the reference's INFO format (case.info, boundaries.info, ...), written to a temp dir."""

    def __init__(self, nx, ny, width=1.0, height=1.0, rho=1.0, mu=0.1, time_step=1.0, bcs=None, properties=None,
                 solver=None):
        """bcs: {field: {patch or "*": (type, value string)}}; default = Examples/LidDrivenCavity/case/boundaries.info.
        properties / solver: extra `Properties` / `Solver` keys (rho1, sigma, g, smoothingKernelRadius, ...)"""
        self.dir = tempfile.mkdtemp(prefix="phase_ref_case_")
        if bcs is None:
            bcs = {"u": {"*": ("fixed", "(0,0)"), "y+": ("fixed", "(1,0)")}, "p": {"*": ("normal_gradient", "0")}}
        with open(os.path.join(self.dir, "case.info"), "w") as f:
            sol = "".join("  %s %s\n" % (k, v) for k, v in (solver or {}).items())
            prop = "".join("  %s %s\n" % (k, v if isinstance(v, str) else "%.17g" % v) for k, v in (properties or {}).items())
            f.write("Solver\n{\n  timeStep %.17g\n%s}\nProperties\n{\n  rho %.17g\n  mu %.17g\n%s}\n" % (time_step, sol, rho, mu, prop))
            f.write("Grid\n{\n  type rectilinear\n  nCellsX %d\n  nCellsY %d\n  width %.17g\n  height %.17g\n}\n" % (nx, ny, width, height))
            f.write("LinearAlgebra\n{\n" + "".join("  %s\n  {\n    lib recording\n  }\n" % e for e in ("uEqn", "pEqn", "gammaEqn")) + "}\n")
        with open(os.path.join(self.dir, "boundaries.info"), "w") as f:
            f.write("Boundaries\n{\n")
            for field, patches in bcs.items():
                f.write("  %s\n  {\n" % field)
                for patch, (typ, val) in patches.items():
                    f.write("    %s\n    {\n      type %s\n      value %s\n    }\n" % (patch, typ, val))
                f.write("  }\n")
            f.write("}\n")
        for name in ("initialConditions.info", "postProcessing.info"):
            with open(os.path.join(self.dir, name), "w") as f:
                f.write("; empty\n")
        self.h = _check(lib().rfv_case_open(self.dir.encode()))

    def close(self):
        if self.h:
            lib().rfv_case_close(self.h)
            self.h = None


class Grid:
    def __init__(self, handle):
        self.h = _check(handle)

    @classmethod
    def rectilinear(cls, case):
        """StructuredRectilinearGrid(input) (UG/StructuredRectilinearGrid.cpp:3-95)"""
        return cls(lib().rfv_grid_rectilinear(case.h))

    @classmethod
    def create(cls, xy, cptr, cind):
        """FiniteVolumeGrid2D(nodes, cptr, cind, origin) (UG/FiniteVolumeGrid2D.cpp:20-35)"""
        xy = np.ascontiguousarray(xy, np.float64)
        cptr = np.ascontiguousarray(cptr, np.int32)
        cind = np.ascontiguousarray(cind, np.int32)
        return cls(lib().rfv_grid_create(len(xy), xy.ctypes.data_as(C.POINTER(C.c_double)), len(cptr) - 1,
                                         cptr.ctypes.data_as(C.POINTER(C.c_int)), cind.ctypes.data_as(C.POINTER(C.c_int))))

    def patch_by_nodes(self, name, nodes):
        nodes = np.ascontiguousarray(nodes, np.int32).reshape(-1)
        _check(lib().rfv_grid_patch_by_nodes(self.h, name.encode(), len(nodes), nodes.ctypes.data_as(C.POINTER(C.c_int))))

    def array(self, name):
        n = _check(lib().rfv_grid_get(self.h, name.encode(), None, None))
        if name in INT_ARRAYS or name.startswith("patch:"):
            a = np.zeros(max(n, 1), np.int32)
            _check(lib().rfv_grid_get(self.h, name.encode(), a.ctypes.data_as(C.POINTER(C.c_int)), None))
        else:
            a = np.zeros(max(n, 1), np.float64)
            _check(lib().rfv_grid_get(self.h, name.encode(), None, a.ctypes.data_as(C.POINTER(C.c_double))))
        return a[:n]

    def index_map(self, n_indices):
        n = self.array("sizes")[1] * n_indices
        loc, glo = np.zeros(n, np.int32), np.zeros(n, np.int32)
        _check(lib().rfv_index_map(self.h, n_indices, loc.ctypes.data_as(C.POINTER(C.c_int)),
                                   glo.ctypes.data_as(C.POINTER(C.c_int))))
        return loc, glo

    def close(self):
        if self.h:
            lib().rfv_grid_close(self.h)
            self.h = None


_cb_keep = []


def use_direct_solver():
    """Every FiniteVolumeEquation<T>::solve of the reference hands its system to scipy's sparse LU."""
    def cb(n, rp, ci, va, b, x, user):
        rp_ = np.ctypeslib.as_array(rp, (n + 1,))
        nnz = int(rp_[n])
        x_ = np.ctypeslib.as_array(x, (n,))
        x_[:] = direct_solve(rp_, np.ctypeslib.as_array(ci, (nnz,)), np.ctypeslib.as_array(va, (nnz,)),
                             np.ctypeslib.as_array(b, (n,)))
        return 1
    fn = SOLVE_CB(cb)
    _cb_keep.append(fn)
    lib().rfv_set_solver(fn, None)


def use_null_solver():
    """Assembly only: the recording backend returns x = 0."""
    fn = SOLVE_CB(lambda n, rp, ci, va, b, x, user: 0)
    _cb_keep.append(fn)
    lib().rfv_set_solver(fn, None)


class FracStep:
    """FractionalStep(input, grid) + initialize() (US/FractionalStep.cpp:7-28)"""

    def __init__(self, case, grid):
        self.case, self.grid = case, grid
        self.h = _check(lib().rfv_fs_create(case.h, grid.h))
        s = grid.array("sizes")
        self.N, self.F = int(s[1]), int(s[2])

    def step(self, dt):
        _check(lib().rfv_fs_step(self.h, dt))

    def view(self, name):
        n = self.F if name in ("ufx", "ufy", "pf", "gpfx", "gpfy") else self.N
        out = np.zeros(n)
        _check(lib().rfv_fs_field(self.h, name.encode(), out.ctypes.data_as(C.POINTER(C.c_double)), 0))
        return out

    def set(self, name, v):
        v = np.ascontiguousarray(v, np.float64)
        _check(lib().rfv_fs_field(self.h, name.encode(), v.ctypes.data_as(C.POINTER(C.c_double)), 1))

    def handoff(self, which):
        """(rowPtr, colInd, vals, b) as FiniteVolumeEquation<T>::solve handed them to the backend in its last solve"""
        nnz = C.c_long()
        n = _check(lib().rfv_fs_handoff(self.h, which.encode(), None, None, None, None, C.byref(nnz)))
        rp, ci = np.zeros(n + 1, np.int32), np.zeros(max(nnz.value, 1), np.int32)
        va, b = np.zeros(max(nnz.value, 1)), np.zeros(n)
        _check(lib().rfv_fs_handoff(self.h, which.encode(), rp.ctypes.data_as(C.POINTER(C.c_int)),
                                    ci.ctypes.data_as(C.POINTER(C.c_int)), va.ctypes.data_as(C.POINTER(C.c_double)),
                                    b.ctypes.data_as(C.POINTER(C.c_double)), None))
        return rp, ci[:nnz.value], va[:nnz.value], b

    def field(self, name, comp=-1, faces=False):
        """any registered field by its reference name; comp -1 = scalar, 0 / 1 = vector component"""
        out = np.zeros(self.F if faces else self.N)
        _check(lib().rfv_fs_any_field(self.h, name.encode(), comp, int(faces), out.ctypes.data_as(C.POINTER(C.c_double)), 0))
        return out

    def set_field(self, name, v, comp=-1, faces=False):
        v = np.ascontiguousarray(v, np.float64)
        assert len(v) == (self.F if faces else self.N)
        _check(lib().rfv_fs_any_field(self.h, name.encode(), comp, int(faces), v.ctypes.data_as(C.POINTER(C.c_double)), 1))

    def initialize(self):
        _check(lib().rfv_fs_initialize(self.h))

    # ---- operator probes on u_, p_, co_ (the reference's own src:: / fv:: functions)
    def src_laplacian(self, gamma):
        """src::laplacian(Scalar gamma, p_) (UD/Source.cpp:27-48); the field overload (:50-75) cannot be run: it
        indexes the scalar index map out of bounds"""
        out = np.zeros(self.N)
        _check(lib().rfv_src_laplacian_scalar(self.h, float(gamma), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def src_div_cells(self, cells):
        c = np.ascontiguousarray(cells, np.int32)
        out = np.zeros(self.N)
        _check(lib().rfv_src_div_cells(self.h, len(c), c.ctypes.data_as(C.POINTER(C.c_int)), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def ddt_cells(self, dt, cells):
        """(diagonal, rhs_) of fv::ddt(p_, dt, cells) after p_.savePreviousTimeStep"""
        c = np.ascontiguousarray(cells, np.int32)
        d, r = np.zeros(self.N), np.zeros(self.N)
        _check(lib().rfv_ddt_cells(self.h, dt, len(c), c.ctypes.data_as(C.POINTER(C.c_int)),
                                   d.ctypes.data_as(C.POINTER(C.c_double)), r.ctypes.data_as(C.POINTER(C.c_double))))
        return d, r

    def max_divergence(self):
        return lib().rfv_fs_max_divergence(self.h)

    def max_courant(self, dt):
        return lib().rfv_fs_max_courant(self.h, dt)

    def max_time_step(self, max_co, prev_dt):
        return lib().rfv_fs_max_time_step(self.h, max_co, prev_dt)

    def close(self):
        if self.h:
            lib().rfv_fs_close(self.h)
            self.h = None


class Multiphase(FracStep):
    """FractionalStepMultiphase(input, grid) (US/FractionalStepMultiphase.cpp); set gamma, then initialize()."""

    def __init__(self, case, grid):
        self.case, self.grid = case, grid
        self.h = _check(lib().rfv_fsm_create(case.h, grid.h))
        s = grid.array("sizes")
        self.N, self.F = int(s[1]), int(s[2])


def partition(case, part, n_ranks, n_indices=2):
    """FiniteVolumeGrid2D::partition (UG/FiniteVolumeGrid2D.cpp:276-392) + initCommBuffers (:459-511) + IndexMap
    (UE/IndexMap.cpp:5-40) of the reference, run on `n_ranks` MPI ranks (threads of this process) for the rectilinear
    grid of `case` with the given cell-partition vector.  Returns one dict of int arrays per rank."""
    part = np.ascontiguousarray(part, np.int32)
    _check(lib().rfv_partition_run(case.h, n_ranks, part.ctypes.data_as(C.POINTER(C.c_int)), len(part), n_indices))
    out = []
    for r in range(n_ranks):
        d = {}
        for k in ("globalId", "owner", "bufPtr", "bufCell", "sendPtr", "sendCell", "local", "global", "faceL", "faceR"):
            n = _check(lib().rfv_partition_get(r, k.encode(), None))
            a = np.zeros(max(n, 1), np.int32)
            _check(lib().rfv_partition_get(r, k.encode(), a.ctypes.data_as(C.POINTER(C.c_int))))
            d[k] = a[:n]
        out.append(d)
    return out
