"""Builds the C++ examples against include/phase and libphase_b200.so (g++ only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_build")
TARGETS = ["lid_driven_cavity", "seam1_solver", "phase_piso", "equation_ops"]


def build(force=False):
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(ROOT, "phase_b200", "libphase_b200.so")
    hdrs = [os.path.join(ROOT, "include", "phase", f) for f in os.listdir(os.path.join(ROOT, "include", "phase"))]
    hdrs.append(os.path.join(ROOT, "include", "phase_b200.h"))
    newest = max(os.path.getmtime(h) for h in hdrs)
    for t in TARGETS:
        src, exe = os.path.join(HERE, t + ".cpp"), os.path.join(OUT, t)
        if (not force and os.path.exists(exe) and os.path.getmtime(exe) >= max(newest, os.path.getmtime(src))
                and os.path.getmtime(exe) >= os.path.getmtime(lib)):
            continue
        subprocess.check_call(["g++", "-std=c++14", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), src, "-o", exe,
                               "-L" + os.path.join(ROOT, "phase_b200"), "-lphase_b200",
                               "-Wl,-rpath," + os.path.join("$ORIGIN", "..", "..", "phase_b200")])
    return [os.path.join(OUT, t) for t in TARGETS]


if __name__ == "__main__":
    print(build(force=True))
