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


if __name__ == "__main__":
    print(build_oracle(force=True))
    print(build_ref(force=True))
