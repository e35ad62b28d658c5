"""Prototype: how should the prolongator treat rank boundaries?  (global scipy matrices, partition only affects
aggregation and which entries of P / R are cut)"""
import sys, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, '/root/repo/tools/proto'); sys.path.insert(0, '/root/repo')
from sa_proto import neumann_lap, build, make_cycle
from tests.test_host_amg import block_partition, bicgstab_iters


def aggregate_ranked(A, part):
    """greedy aggregation that never crosses a rank boundary"""
    n = A.shape[0]; rp, ci = A.indptr, A.indices
    agg = -np.ones(n, dtype=np.int64); nc = 0
    nb = lambda i: [j for j in ci[rp[i]:rp[i + 1]] if j != i and part[j] == part[i]]
    for i in range(n):
        if agg[i] >= 0: continue
        N = nb(i)
        if any(agg[j] >= 0 for j in N): continue
        agg[i] = nc
        for j in N: agg[j] = nc
        nc += 1
    a2 = agg.copy()
    for i in range(n):
        if agg[i] >= 0: continue
        for j in nb(i):
            if agg[j] >= 0: a2[i] = agg[j]; break
    agg = a2
    for i in range(n):
        if agg[i] >= 0: continue
        agg[i] = nc
        for j in nb(i):
            if agg[j] < 0: agg[j] = nc
        nc += 1
    return agg, nc


def level(A, part, variant):
    n = A.shape[0]
    agg, nc = aggregate_ranked(A, part)
    cpart = np.zeros(nc, dtype=np.int64); cpart[agg] = part
    T = sp.csr_matrix((np.ones(n), (np.arange(n), agg)), shape=(n, nc))
    D = A.diagonal()
    same = sp.csr_matrix((np.ones(A.nnz), A.indices, A.indptr), shape=A.shape).tocoo()
    loc = part[same.row] == part[same.col]
    Aloc = sp.csr_matrix((A.tocoo().data[loc], (same.row[loc], same.col[loc])), shape=A.shape)
    if variant == 'block':       # current: ghost couplings lumped, P block diagonal
        lump = np.asarray((A - Aloc).sum(axis=1)).ravel()
        Af = Aloc + sp.diags(lump); Df = Af.diagonal()
        rho = np.abs(sp.diags(1 / Df) @ Af).sum(axis=1).max()
        P = (T - (4 / 3 / rho) * (sp.diags(1 / Df) @ Af @ T)).tocsr(); R = P.T.tocsr()
    else:
        rho = np.abs(sp.diags(1 / D) @ A).sum(axis=1).max()
        P = (T - (4 / 3 / rho) * (sp.diags(1 / D) @ A @ T)).tocsr()
        if variant == 'full':
            R = P.T.tocsr()
        else:                    # 'cut': restriction keeps only the entries inside the owner rank of the coarse row
            Pc = P.tocoo(); keep = part[Pc.row] == cpart[Pc.col]
            R = sp.csr_matrix((Pc.data[keep], (Pc.col[keep], Pc.row[keep])), shape=(nc, n))
    return P, R, (R @ A @ P).tocsr(), cpart


def run(n, px, py, ndist, variant):
    A = neumann_lap(n, n)
    part = block_partition(n, n, px, py)
    lv = []; Al = A; pl = part
    for _ in range(ndist):
        P, R, Ac, cp = level(Al, pl, variant)
        lv.append((Al, P, R)); Al = Ac; pl = cp
    tail = make_cycle(build(Al), 1)
    def cyc(l, b):
        if l == ndist: return tail(b)
        A_, P_, R_ = lv[l]
        rho = np.abs(sp.diags(1 / A_.diagonal()) @ A_).sum(axis=1).max()
        w = (4 / 3 / rho) / A_.diagonal()
        x = w * b; x = x + P_ @ cyc(l + 1, R_ @ (b - A_ @ x)); return x + w * (b - A_ @ x)
    b = np.random.default_rng(2).standard_normal(n * n); b -= b.mean()
    return bicgstab_iters(A, lambda v: cyc(0, v), b)[0]


if __name__ == '__main__':
    n = int(sys.argv[1]); nd = int(sys.argv[2])
    for v in ('block', 'full', 'cut'):
        print(n, nd, v, run(n, 2, 4, nd, v), flush=True)
