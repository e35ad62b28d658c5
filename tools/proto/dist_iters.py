"""BiCGStab iteration counts of the library's serial and distributed multigrid hierarchies (ranks as threads, CPU
only): how much does the rank-local aggregation / Petrov-Galerkin restriction cost?   python tools/proto/dist_iters.py N"""
import sys
import numpy as np, scipy.sparse as sp
sys.path.insert(0, '/root/repo')
from tests.test_host_amg import HostAmg, DistAmg, neumann_laplacian, block_partition, bicgstab_iters

def dist_cycle(A, part, coarsest, tail_rows, omega=1.8):
    H = DistAmg(A, part, coarsest=coarsest, tail_rows=tail_rows)
    levels = []; Al = A
    for l in range(H.nDist):
        nc = sum(H.rank_matrix(r, l, 2)[0].shape[0] - 1 for r in range(H.nRanks))
        P = H.global_matrix(l, 1, (Al.shape[0], nc)); R = H.global_matrix(l, 2, (nc, Al.shape[0]))
        An = H.global_matrix(l + 1, 0, (nc, nc)) if l + 1 < H.nDist else H.tail(0).mat(0, 0)[0]
        levels.append((Al, P, R)); Al = An
    tail = H.tail(0).cycle()
    def cyc(l, b):
        if l == H.nDist: return tail(b)
        A_, P_, R_ = levels[l]
        rho = np.abs(sp.diags(1.0 / A_.diagonal()) @ A_).sum(axis=1).max()
        w = (omega / rho) / A_.diagonal()
        x = w * b
        x = x + P_ @ cyc(l + 1, R_ @ (b - A_ @ x))
        return x + w * (b - A_ @ x)
    return (lambda v: cyc(0, v)), H

n = int(sys.argv[1])
A = neumann_laplacian(n, n)
b = np.random.default_rng(2).standard_normal(n * n); b -= b.mean()
S = HostAmg(A, coarsest=1000)
print("serial levels", S.nLevels, "iters", bicgstab_iters(A, S.cycle(), b)[0], flush=True)
for px, py in [tuple(int(v) for v in a.split("x")) for a in (sys.argv[2:] or ["1x2", "2x2", "2x4"])]:
    for tail in (50000,):
        M, H = dist_cycle(A, block_partition(n, n, px, py), 1000, tail)
        print("ranks %dx%d tailRows %d: dist levels %d tail levels %d iters %d" % (px, py, tail, H.nDist, H.nTail, bicgstab_iters(A, M, b)[0]), flush=True)
