"""Build recipes for the CPU oracle (TEST INFRASTRUCTURE ONLY).

* ``build_oracle()``  -> oracle/_build/libphase_oracle.so from oracle/phase_oracle.c
* ``build_ref()``     -> oracle/_ref/libphase_ref_crs.so: the reference's own
  src/Math/{CrsEquation,Vector,SparseMatrixSolver}.cpp + src/System/Exception.cpp
  compiled WHERE THEY LIE under /root/reference (nothing is copied), plus
  oracle/ref_driver.cpp (ours).  Only possible where /root/reference is mounted
  (this container); the GPU box uses the prebuilt .so that travels with gpurun.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
ORACLE_SO = os.path.join(HERE, "_build", "libphase_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libphase_ref_crs.so")
REF_FV_SO = os.path.join(HERE, "_ref", "libphase_ref_fv.so")


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def build_oracle(force=False):
    srcs = [os.path.join(HERE, "phase_oracle.c"), os.path.join(HERE, "phase_oracle.h")]
    if not force and _newer(ORACLE_SO, srcs):
        return ORACLE_SO
    os.makedirs(os.path.dirname(ORACLE_SO), exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-Wall", "-o", ORACLE_SO,
           srcs[0], "-lm"]
    subprocess.check_call(cmd)
    return ORACLE_SO


def build_ref(force=False):
    """Returns the path of the .so, or None when neither sources nor a prebuilt exist."""
    ref_files = [os.path.join(REF_SRC, "Math", f) for f in
                 ("CrsEquation.cpp", "Vector.cpp", "SparseMatrixSolver.cpp")]
    ref_files.append(os.path.join(REF_SRC, "System", "Exception.cpp"))
    driver = os.path.join(HERE, "ref_driver.cpp")
    if not all(os.path.exists(f) for f in ref_files):
        return REF_SO if os.path.exists(REF_SO) else None
    if not force and _newer(REF_SO, ref_files + [driver]):
        return REF_SO
    os.makedirs(os.path.dirname(REF_SO), exist_ok=True)
    cmd = ["g++", "-std=c++11", "-O2", "-fPIC", "-shared", "-w",
           "-I" + os.path.join(HERE, "ref_stub"), "-I" + REF_SRC,
           "-o", REF_SO, driver] + ref_files
    subprocess.check_call(cmd)
    return REF_SO


def ref_fv_sources():
    """The reference translation units behind oracle/_ref/libphase_ref_fv.so, compiled where they lie: system +
    CSR algebra, 2-D geometry, the unstructured grid, fields, equations, the fv:: / src:: / cicsam:: operators and
    the fractional-step solver module.  Left out: everything that needs CGNS, Trilinos / Eigen (the backends; a
    recording backend takes their place in ref_fv_driver.cpp), the command line and the immersed-boundary /
    post-processing subsystems."""
    import glob
    R = REF_SRC
    U = os.path.join(R, "2D", "Unstructured")
    files = [os.path.join(R, "System", f) for f in ("Exception.cpp", "Communicator.cpp", "Input.cpp")]
    files += [os.path.join(R, "Math", f) for f in ("Vector.cpp", "CrsEquation.cpp", "SparseMatrixSolver.cpp", "Matrix.cpp")]
    files += sorted(glob.glob(os.path.join(R, "2D", "Geometry", "*.cpp")))
    for root, _, names in sorted(os.walk(os.path.join(U, "FiniteVolumeGrid2D"))):
        files += [os.path.join(root, n) for n in sorted(names)
                  if n.endswith(".cpp") and "Cgns" not in n and "Factory" not in n]
    files += sorted(glob.glob(os.path.join(U, "FiniteVolume", "Field", "*.cpp")))
    files += sorted(glob.glob(os.path.join(U, "FiniteVolume", "Equation", "*.cpp")))
    files += [os.path.join(U, "FiniteVolume", "Discretization", f) for f in
              ("Laplacian.cpp", "Source.cpp", "Divergence.cpp", "Cicsam.cpp", "Axisymmetric.cpp")]
    files += [os.path.join(U, "FiniteVolume", "Multiphase", f) for f in
              ("Celeste.cpp", "CelesteStencil.cpp", "SurfaceTensionForce.cpp", "SurfaceTensionForceSmoothingKernel.cpp")]
    files += [os.path.join(U, "Solvers", f) for f in ("Solver.cpp", "FractionalStep.cpp", "FractionalStepMultiphase.cpp")]
    return files


def _openblas():
    """The LP64 OpenBLAS inside scipy (symbols prefixed scipy_, see oracle/ref_stub/{cblas,lapacke}.h)."""
    import glob
    import scipy
    libs = glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas-*.so"))
    return libs[0] if libs else None


def build_ref_fv(force=False):
    """oracle/_ref/libphase_ref_fv.so, or None when neither the reference sources nor a prebuilt exist."""
    import concurrent.futures
    driver = os.path.join(HERE, "ref_fv_driver.cpp")
    if not os.path.exists(os.path.join(REF_SRC, "2D", "Unstructured", "Solvers", "FractionalStep.cpp")):
        return REF_FV_SO if os.path.exists(REF_FV_SO) else None
    srcs = ref_fv_sources() + [driver, os.path.join(HERE, "ref_mpi_threads.cpp")]
    stubs = [os.path.join(r, n) for r, _, ns in os.walk(os.path.join(HERE, "ref_stub")) for n in ns]
    if not force and _newer(REF_FV_SO, srcs + stubs):
        return REF_FV_SO
    blas = _openblas()
    if blas is None:
        return REF_FV_SO if os.path.exists(REF_FV_SO) else None
    objdir = os.path.join(HERE, "_ref", "obj")
    os.makedirs(objdir, exist_ok=True)
    inc = ["-I" + os.path.join(HERE, "ref_stub"), "-I" + REF_SRC, "-I" + os.path.join(REF_SRC, "2D"),
           "-I" + os.path.join(REF_SRC, "2D", "Unstructured")]

    def cc(src):
        obj = os.path.join(objdir, os.path.relpath(src, "/").replace("/", "_") + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(f) for f in [src] + stubs):
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-w", "-c"] + inc + [src, "-o", obj])
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(cc, srcs))
    subprocess.check_call(["g++", "-shared", "-o", REF_FV_SO] + objs +
                          [blas, "-Wl,-rpath," + os.path.dirname(blas), "-lstdc++fs", "-lpthread"])
    return REF_FV_SO


if __name__ == "__main__":
    print(build_oracle(force=True))
    print(build_ref(force=True))
    print(build_ref_fv(force=True))
