"""In-tree build of libphase_b200.so (hand-written CUDA for sm_100a + the C ABI).

    python -m phase_b200.build [--force]

nvcc cross-compiles without a GPU; the .so stays in the tree (git-ignored) so it
travels to the GPU box with the gpurun snapshot.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libphase_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-diag-suppress", "177"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "phase_b200.h"))
    return hs


# METIS ships with the CUDA toolkit as a static library (cuSOLVER's; 64-bit idx_t): the optional "same algorithm
# family as the reference" partitioner (partition.cu)
METIS = os.path.join(os.path.dirname(os.path.dirname(NVCC)), "targets", "x86_64-linux", "lib", "libmetis_static.a")


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + (["-DPHB_HAVE_METIS"] if os.path.exists(METIS) else []) + \
        ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    hdr_t = max(os.path.getmtime(h) for h in headers())
    todo, objs = [], []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_t):
            todo.append(s)
    logs = []
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for obj, log in ex.map(lambda s: _compile(s, verbose), todo):
                logs.append(log)
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ([METIS] if os.path.exists(METIS) else []) + \
            ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-ldl"]
        subprocess.check_call(cmd)
    return LIB, "\n".join(logs)


if __name__ == "__main__":
    lib, log = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    if log.strip():
        print(log)
    print(lib)
