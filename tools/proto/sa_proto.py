"""Prototype (CPU, scipy) of the smoothed-aggregation hierarchy used to pick parameters for amg.cu."""
import sys, time
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla


def neumann_lap(nx, ny, fixed_side=False):
    def d1(n):
        e = np.ones(n)
        T = sp.diags([-e[:-1], 2 * e, -e[:-1]], [-1, 0, 1]).tolil()
        T[0, 0] = 1; T[n - 1, n - 1] = 1
        return T.tocsr()
    A = sp.kron(sp.eye(ny), d1(nx)) + sp.kron(d1(ny), sp.eye(nx))
    A = A.tocsr()
    if fixed_side:
        d = np.zeros(nx * ny); d[::nx] += 2.0
        A = A + sp.diags(d)
    return A.tocsr()


def aggregate(A, theta):
    n = A.shape[0]
    rp, ci, v = A.indptr, A.indices, A.data
    d = np.abs(A.diagonal())
    agg = -np.ones(n, dtype=np.int64)
    strong = [None] * n
    rows = np.repeat(np.arange(n), np.diff(rp))
    mask = (ci != rows) & (v * v >= theta * theta * d[rows] * d[ci]) & (v != 0)
    S = sp.csr_matrix((np.ones(mask.sum()), (rows[mask], ci[mask])), shape=(n, n))
    srp, sci = S.indptr, S.indices
    nc = 0
    for i in range(n):
        if agg[i] >= 0: continue
        nb = sci[srp[i]:srp[i + 1]]
        if len(nb) and (agg[nb] >= 0).any(): continue
        agg[i] = nc; agg[nb] = nc; nc += 1
    agg2 = agg.copy()
    for i in range(n):
        if agg[i] >= 0: continue
        nb = sci[srp[i]:srp[i + 1]]
        a = agg[nb]; a = a[a >= 0]
        if len(a): agg2[i] = a[0]
    agg = agg2
    for i in range(n):
        if agg[i] >= 0: continue
        nb = sci[srp[i]:srp[i + 1]]
        agg[i] = nc
        for j in nb:
            if agg[j] < 0: agg[j] = nc
        nc += 1
    return agg, nc


def build(A, theta=0.0, coarsest=400, maxlev=12, omega_p=4. / 3):
    levels = []
    while True:
        n = A.shape[0]
        D = A.diagonal()
        if n <= coarsest or len(levels) >= maxlev - 1:
            levels.append(dict(A=A, D=D)); break
        agg, nc = aggregate(A, theta)
        T = sp.csr_matrix((np.ones(n), (np.arange(n), agg)), shape=(n, nc))
        DinvA = sp.diags(1. / D) @ A
        rho = np.abs(DinvA).sum(axis=1).max()   # Gershgorin
        P = (T - (omega_p / rho) * (DinvA @ T)).tocsr()
        R = P.T.tocsr()
        Ac = (R @ A @ P).tocsr(); Ac.eliminate_zeros()
        levels.append(dict(A=A, D=D, P=P, R=R, rho=rho))
        A = Ac
    return levels


def make_cycle(levels, nu=1, omega_s=4. / 3, singular=True):
    Ac = levels[-1]['A'].toarray()
    nC = Ac.shape[0]
    if singular:
        Ac = Ac + np.ones((nC, nC)) * (np.abs(np.diag(Ac)).mean() / nC)
    Ainv = np.linalg.inv(Ac)

    def cyc(l, b):
        L = levels[l]
        if l == len(levels) - 1: return Ainv @ b
        w = omega_s / L['rho']
        x = w * b / L['D']
        for _ in range(nu - 1): x = x + w * (b - L['A'] @ x) / L['D']
        r = b - L['A'] @ x
        x = x + L['P'] @ cyc(l + 1, L['R'] @ r)
        for _ in range(nu): x = x + w * (b - L['A'] @ x) / L['D']
        return x
    return lambda b: cyc(0, b)


if __name__ == '__main__':
    nx = int(sys.argv[1]); theta = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    nu = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    A = neumann_lap(nx, nx)
    t = time.time(); lv = build(A, theta); print('setup', time.time() - t)
    print('levels', [(l['A'].shape[0], l['A'].nnz) for l in lv], 'opcx', sum(l['A'].nnz for l in lv) / A.nnz)
    M = make_cycle(lv, nu)
    rng = np.random.default_rng(0); b = rng.standard_normal(A.shape[0]); b -= b.mean()
    its = [0]
    def cb(x): its[0] += 1
    x, info = spla.bicgstab(A, b, rtol=1e-8, atol=0, M=spla.LinearOperator(A.shape, matvec=M), callback=cb, maxiter=500)
    print('bicgstab its', its[0], info, np.linalg.norm(b - A @ x) / np.linalg.norm(b))
