"""Prototype: strength threshold on the variable-density pEqn_ (config 4).  Run under `timeout`."""
import sys, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
sys.path.insert(0, 'tools/proto')
from sa_proto import build, make_cycle
def varlap(nx, ny):
    x = (np.arange(nx) + .5) / nx; y = 2 * (np.arange(ny) + .5) / ny
    X, Y = np.meshgrid(x, y)
    gam = ((X - .5) ** 2 + (Y - .5) ** 2 < .125 ** 2) | (Y > 1.5)
    rho = np.where(gam, 1.225, 998.).ravel()
    idx = np.arange(nx * ny).reshape(ny, nx)
    I = []; J = []; V = []
    def add(a, b):
        w = 1. / (0.5 * (rho[a] + rho[b]))
        I.extend([a, b, a, b]); J.extend([b, a, a, b]); V.extend([-w, -w, w, w])
    a = idx[:, :-1].ravel(); b = idx[:, 1:].ravel()
    w = 1. / (0.5 * (rho[a] + rho[b]))
    A = sp.coo_matrix((np.r_[-w, -w, w, w], (np.r_[a, b, a, b], np.r_[b, a, a, b])), shape=(nx * ny,) * 2)
    a = idx[:-1, :].ravel(); b = idx[1:, :].ravel()
    w = 1. / (0.5 * (rho[a] + rho[b]))
    A = A + sp.coo_matrix((np.r_[-w, -w, w, w], (np.r_[a, b, a, b], np.r_[b, a, a, b])), shape=(nx * ny,) * 2)
    d = np.zeros(nx * ny); t = idx[-1, :]; d[t] = 2. / rho[t]   # p fixed on y+
    return (A + sp.diags(d)).tocsr()
nx = int(sys.argv[1])
A = varlap(nx, 2 * nx)
for theta in (0.0, 0.08):   # 0.25 stalls the coarsening of this prototype (no stall guard here; amg.cu has one)
    lv = build(A, theta)
    M = make_cycle(lv, 1, singular=False)
    rng = np.random.default_rng(0); b = rng.standard_normal(A.shape[0])
    its = [0]
    def cb(x): its[0] += 1
    x, info = spla.bicgstab(A, b, rtol=1e-8, atol=0, M=spla.LinearOperator(A.shape, matvec=M), callback=cb, maxiter=500)
    print(theta, [l['A'].shape[0] for l in lv], 'its', its[0], info, np.linalg.norm(b - A @ x) / np.linalg.norm(b))
