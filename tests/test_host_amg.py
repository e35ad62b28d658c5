"""CPU checks of the smoothed-aggregation setup behind `preconditioner amg` (phase_b200/csrc/amg.cu).

The hierarchy is built on the host, so everything but the device cycle can be verified here: aggregates cover
every row, R = P^T, the coarse operators are the Galerkin products R A P, the constant is reproduced by P
(singular all-Neumann systems) and a scipy transcription of the V(1,1) cycle preconditions BiCGStab into tens of
iterations, independent of the mesh size."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from phase_b200 import _capi


def neumann_laplacian(nx, ny, fixed_left=False, sign=-1.0):
    def d1(n):
        e = np.ones(n)
        T = sp.diags([-e[:-1], 2 * e, -e[:-1]], [-1, 0, 1]).tolil()
        T[0, 0] = 1
        T[n - 1, n - 1] = 1
        return T.tocsr()
    A = (sp.kron(sp.eye(ny), d1(nx)) + sp.kron(d1(ny), sp.eye(nx))).tocsr()
    if fixed_left:
        d = np.zeros(nx * ny)
        d[::nx] = 2.0
        A = (A + sp.diags(d)).tocsr()
    return (sign * A).tocsr()


class HostAmg:
    def __init__(self, A, theta=0.0, coarsest=400, agg_theta=-1.0, coarse_weight=-1.0):
        """agg_theta / coarse_weight: `amgAggTheta` / `amgCoarseSmootherWeight` (negative: library default, 0: rule off)"""
        self.L = _capi.lib()
        A = A.tocsr()
        A.sort_indices()
        rp = A.indptr.astype(np.int32)
        ci = A.indices.astype(np.int32)
        v = A.data.astype(np.float64)
        h = C.c_void_p()
        rc = self.L.phb_amg_host_build_ex(A.shape[0], rp.ctypes.data_as(_capi.pi), ci.ctypes.data_as(_capi.pi),
                                          v.ctypes.data_as(_capi.pd), theta, agg_theta, coarse_weight, coarsest, C.byref(h))
        assert rc == 0, self.L.phb_last_error()
        self.h = h
        n, s, d = C.c_int(), C.c_int(), C.c_int()
        assert self.L.phb_amg_host_levels(h, C.byref(n), C.byref(s), C.byref(d)) == 0
        self.nLevels, self.singular, self.dense = n.value, bool(s.value), bool(d.value)

    def mat(self, level, which):
        nr, nc, nnz, rho = C.c_int(), C.c_int(), C.c_longlong(), C.c_double()
        rc = self.L.phb_amg_host_level_size(self.h, level, which, C.byref(nr), C.byref(nc), C.byref(nnz), C.byref(rho))
        assert rc == 0, self.L.phb_last_error()
        rp = np.zeros(nr.value + 1, np.int32)
        ci = np.zeros(nnz.value, np.int32)
        v = np.zeros(nnz.value)
        rc = self.L.phb_amg_host_level_csr(self.h, level, which, rp.ctypes.data_as(_capi.pi),
                                           ci.ctypes.data_as(_capi.pi), v.ctypes.data_as(_capi.pd))
        assert rc == 0
        return sp.csr_matrix((v, ci, rp), shape=(nr.value, nc.value)), rho.value

    def coarse_inverse(self):
        n = self.mat(self.nLevels - 1, 0)[0].shape[0]
        inv = np.zeros((n, n))
        assert self.L.phb_amg_host_coarse_inverse(self.h, inv.ctypes.data_as(_capi.pd)) == 0
        return inv

    def weight_scale(self, level):
        f = C.c_double()
        assert self.L.phb_amg_host_level_weight(self.h, level, C.byref(f)) == 0
        return f.value

    def cycle(self, nu=1):
        """scipy transcription of amg_apply() (amg.cu)"""
        lv = []
        for l in range(self.nLevels):
            A, rho = self.mat(l, 0)
            e = dict(A=A, w=self.weight_scale(l) / A.diagonal())
            if l + 1 < self.nLevels:
                e["P"] = self.mat(l, 1)[0]
                e["R"] = self.mat(l, 2)[0]
            lv.append(e)
        inv = self.coarse_inverse()

        def cyc(l, b):
            L = lv[l]
            if l == self.nLevels - 1:
                return inv @ b
            x = L["w"] * b
            for _ in range(nu - 1):
                x = x + L["w"] * (b - L["A"] @ x)
            x = x + L["P"] @ cyc(l + 1, L["R"] @ (b - L["A"] @ x))
            for _ in range(nu):
                x = x + L["w"] * (b - L["A"] @ x)
            return x
        return lambda b: cyc(0, b)

    def close(self):
        self.L.phb_amg_host_destroy(self.h)


def bicgstab_iters(A, M, b, tol=1e-8):
    its = [0]

    def cb(_):
        its[0] += 1
    x, info = spla.bicgstab(A, b, rtol=tol, atol=0, M=spla.LinearOperator(A.shape, matvec=M), callback=cb, maxiter=200)
    assert info == 0
    return its[0], np.linalg.norm(b - A @ x) / np.linalg.norm(b)


@pytest.mark.parametrize("fixed", [False, True])
def test_hierarchy_is_galerkin(fixed):
    A = neumann_laplacian(60, 50, fixed_left=fixed)
    H = HostAmg(A, coarsest=50)
    assert H.nLevels >= 3 and H.dense
    assert H.singular == (not fixed)
    A0, _ = H.mat(0, 0)
    assert abs(A0 - A).max() == 0.0
    prev = A0
    for l in range(H.nLevels - 1):
        P, _ = H.mat(l, 1)
        R, _ = H.mat(l, 2)
        Ac, _ = H.mat(l + 1, 0)
        assert P.shape == (prev.shape[0], Ac.shape[0])
        assert abs(R - P.T).max() == 0.0
        G = (R @ prev @ P).tocsr()
        assert abs(G - Ac).max() <= 1e-12 * abs(Ac).max()
        assert Ac.shape[0] < 0.5 * prev.shape[0]                      # coarsening does coarsen
        assert np.all(np.diff(P.indptr) >= 1)                         # every row belongs to an aggregate
        if not fixed:                                                 # constant reproduced level by level
            assert np.abs(P @ np.ones(P.shape[1]) - 1.0).max() < 1e-12
            assert np.abs(Ac @ np.ones(Ac.shape[0])).max() < 1e-10 * abs(Ac).max()
        prev = Ac
    H.close()


def test_cycle_preconditions_bicgstab_mesh_independently():
    rng = np.random.default_rng(0)
    counts = []
    for n in (64, 128, 256):
        A = neumann_laplacian(n, n)
        H = HostAmg(A)
        b = rng.standard_normal(n * n)
        b -= b.mean()
        its, rel = bicgstab_iters(A, H.cycle(), b)
        assert rel < 2e-8
        counts.append(its)
        H.close()
    assert max(counts) <= 20, counts


def test_iterations_do_not_depend_on_the_row_length_of_the_mesh():
    """Round 1's wide-mesh pathology (a 4000-cell-wide mesh needed twice the iterations of a 3998-cell-wide one): with every
    small Galerkin entry a member-maker, oversized aggregates formed at the end of the natural order on the coarse
    levels.  The row-relative membership filter (`amgAggTheta`) and the power-iteration smoother weights on the Galerkin
    levels (`amgCoarseSmootherWeight`) bring both meshes to the same count; switched off, the old behaviour is back."""
    rng = np.random.default_rng(0)
    counts = {}
    for nx in (4000, 3998):
        A = neumann_laplacian(nx, 250)
        b = rng.standard_normal(A.shape[0])
        b -= b.mean()
        for tag, kw in (("new", {}), ("old", dict(agg_theta=0.0, coarse_weight=0.0))):
            H = HostAmg(A, coarsest=1000, **kw)
            counts[nx, tag] = bicgstab_iters(A, H.cycle(), b)[0]
            if tag == "new":
                ws = [H.weight_scale(l) for l in range(H.nLevels)]
                assert abs(ws[0] - 0.9) < 1e-12 and all(0.95 < w < 1.35 for w in ws[1:]), ws
            else:
                assert all(abs(H.weight_scale(l) - 0.9) < 1e-12 for l in range(H.nLevels))
            H.close()
    assert counts[4000, "new"] <= 9 and counts[3998, "new"] <= 9, counts
    assert counts[4000, "old"] >= counts[4000, "new"] + 2, counts


def test_smoother_weights_keep_jacobi_a_contraction():
    """`amgCoarseSmootherWeight` / lambda_est on the Galerkin levels: the power-iteration estimate approaches lambda_max
    from below, so the weight must stay safely under 2 / lambda_max(D^-1 A) (the Jacobi sweep would amplify the top modes
    otherwise) and -- the point of the rule -- above the Gershgorin weight it replaces on these levels."""
    nx, ny = 96, 192
    x = (np.arange(nx) + .5) / nx
    y = 2 * (np.arange(ny) + .5) / ny
    X, Y = np.meshgrid(x, y)
    rho = np.where((X - .5) ** 2 + (Y - .5) ** 2 < .125 ** 2, 1.0, 1000.).ravel()
    idx = np.arange(nx * ny).reshape(ny, nx)
    V = sp.csr_matrix((nx * ny, nx * ny))
    for a, b in ((idx[:, :-1].ravel(), idx[:, 1:].ravel()), (idx[:-1, :].ravel(), idx[1:, :].ravel())):
        w = 1. / (0.5 * (rho[a] + rho[b]))
        V = V + sp.coo_matrix((np.r_[-w, -w, w, w], (np.r_[a, b, a, b], np.r_[b, a, a, b])), shape=V.shape)
    for A in (neumann_laplacian(150, 120), neumann_laplacian(90, 70, fixed_left=True), V.tocsr()):
        H = HostAmg(A, coarsest=60)
        assert H.nLevels >= 3
        for l in range(1, H.nLevels - 1):
            M, rho_g = H.mat(l, 0)
            d = M.diagonal()
            Dh = sp.diags(1. / np.sqrt(np.abs(d)))
            S = (Dh @ M @ Dh).tocsc() * np.sign(d[0])          # symmetric, same spectrum as D^-1 A
            lam = float(spla.eigsh(S, k=1, which="LA", return_eigenvectors=False, tol=1e-8)[0])
            f = H.weight_scale(l)
            assert f * lam < 1.9, (l, f, lam)                   # contraction on every mode, with a margin
            assert f > 1.8 / rho_g, (l, f, rho_g)               # on these Galerkin levels stronger than the rule it replaces
            assert abs(f - 1.6 / min(rho_g, max(0.5 * rho_g, 1.05 * lam))) < 0.1 * f, (l, f, lam)   # the estimate is close
        H.close()


def test_variable_coefficient_and_threshold():
    """density ratio 815 (config 4's pEqn): both theta = 0 and the classical 0.08 give a usable hierarchy"""
    nx, ny = 96, 192
    x = (np.arange(nx) + .5) / nx
    y = 2 * (np.arange(ny) + .5) / ny
    X, Y = np.meshgrid(x, y)
    rho = np.where(((X - .5) ** 2 + (Y - .5) ** 2 < .125 ** 2) | (Y > 1.5), 1.225, 998.).ravel()
    idx = np.arange(nx * ny).reshape(ny, nx)
    A = sp.csr_matrix((nx * ny, nx * ny))
    for a, b in ((idx[:, :-1].ravel(), idx[:, 1:].ravel()), (idx[:-1, :].ravel(), idx[1:, :].ravel())):
        w = 1. / (0.5 * (rho[a] + rho[b]))
        A = A + sp.coo_matrix((np.r_[-w, -w, w, w], (np.r_[a, b, a, b], np.r_[b, a, a, b])), shape=A.shape)
    d = np.zeros(nx * ny)
    d[idx[-1, :]] = 2. / rho[idx[-1, :]]
    A = (A + sp.diags(d)).tocsr()
    b = np.random.default_rng(1).standard_normal(nx * ny)
    for theta in (0.0, 0.08):
        H = HostAmg(A, theta=theta)
        assert not H.singular
        its, rel = bicgstab_iters(A, H.cycle(), b)
        assert rel < 2e-8 and its <= 30, (theta, its)
        H.close()


def test_tiny_matrix_is_one_dense_level():
    A = neumann_laplacian(6, 5, fixed_left=True)
    H = HostAmg(A)
    assert H.nLevels == 1 and H.dense
    assert np.abs(H.coarse_inverse() @ A.toarray() - np.eye(30)).max() < 1e-10
    H.close()


# ------------------------------------------------------------------ distributed setup (ranks as threads)
class DistAmg:
    def __init__(self, A, part, theta=0.0, coarsest=40, tail_rows=150):
        self.L = _capi.lib()
        A = A.tocsr()
        A.sort_indices()
        self.n, self.nRanks = A.shape[0], int(part.max()) + 1
        rp, ci, v = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
        pt = np.ascontiguousarray(part, np.int32)
        h = C.c_void_p()
        rc = self.L.phb_amg_dist_build(self.nRanks, self.n, rp.ctypes.data_as(_capi.pi), ci.ctypes.data_as(_capi.pi),
                                       v.ctypes.data_as(_capi.pd), pt.ctypes.data_as(_capi.pi), theta, coarsest,
                                       tail_rows, C.byref(h))
        assert rc == 0, self.L.phb_last_error()
        self.h = h
        a, b, s = C.c_int(), C.c_int(), C.c_int()
        assert self.L.phb_amg_dist_info(h, C.byref(a), C.byref(b), C.byref(s)) == 0
        self.nDist, self.nTail, self.singular = a.value, b.value, bool(s.value)

    def rank_matrix(self, rank, level, which):
        nr, nnz = C.c_int(), C.c_longlong()
        assert self.L.phb_amg_dist_matrix_size(self.h, rank, level, which, C.byref(nr), C.byref(nnz)) == 0
        rp, ci, v = np.zeros(nr.value + 1, np.int32), np.zeros(nnz.value, np.int32), np.zeros(nnz.value)
        gid = np.zeros(nr.value, np.int32)
        assert self.L.phb_amg_dist_matrix(self.h, rank, level, which, rp.ctypes.data_as(_capi.pi),
                                          ci.ctypes.data_as(_capi.pi), v.ctypes.data_as(_capi.pd),
                                          gid.ctypes.data_as(_capi.pi)) == 0
        return rp, ci, v, gid

    def global_matrix(self, level, which, shape):
        """rows of all ranks stacked by their global ids"""
        I, J, V = [], [], []
        for r in range(self.nRanks):
            rp, ci, v, gid = self.rank_matrix(r, level, which)
            I.append(np.repeat(gid, np.diff(rp)))
            J.append(ci)
            V.append(v)
        return sp.csr_matrix((np.concatenate(V), (np.concatenate(I), np.concatenate(J))), shape=shape)

    def halo(self, rank, level, n_send_max):
        sp_, si, rp_ = np.zeros(self.nRanks + 1, np.int32), np.zeros(n_send_max, np.int32), np.zeros(self.nRanks + 1, np.int32)
        assert self.L.phb_amg_dist_halo(self.h, rank, level, sp_.ctypes.data_as(_capi.pi), si.ctypes.data_as(_capi.pi),
                                        rp_.ctypes.data_as(_capi.pi)) == 0
        return sp_, si[:sp_[-1]], rp_

    def ghost_gids(self, rank, level):
        n = int(self.halo(rank, level, 100000)[2][-1])
        out = np.zeros(max(n, 1), np.int32)
        assert self.L.phb_amg_dist_ghost_gids(self.h, rank, level, out.ctypes.data_as(_capi.pi)) == 0
        return out[:n]

    def tail(self, rank):
        t = HostAmg.__new__(HostAmg)
        t.L = self.L
        t.h = C.c_void_p(self.L.phb_amg_dist_tail(self.h, rank))
        n, s, d = C.c_int(), C.c_int(), C.c_int()
        assert self.L.phb_amg_host_levels(t.h, C.byref(n), C.byref(s), C.byref(d)) == 0
        t.nLevels, t.singular, t.dense = n.value, bool(s.value), bool(d.value)
        return t

    def close(self):
        self.L.phb_amg_dist_destroy(self.h)


def block_partition(nx, ny, px, py):
    j, i = np.divmod(np.arange(nx * ny), nx)
    return (np.minimum(j * py // ny, py - 1) * px + np.minimum(i * px // nx, px - 1)).astype(np.int32)


@pytest.mark.parametrize("px,py,fixed", [(2, 1, False), (2, 2, True), (1, 3, False), (2, 4, False)])
def test_distributed_hierarchy_is_the_global_galerkin_hierarchy(px, py, fixed):
    nx, ny = 48, 36
    A = neumann_laplacian(nx, ny, fixed_left=fixed)
    part = block_partition(nx, ny, px, py)
    H = DistAmg(A, part)
    assert H.nDist >= 2 and H.nTail >= 1 and H.singular == (not fixed)
    Al = H.global_matrix(0, 0, A.shape)
    assert abs(Al - A).max() == 0.0
    levels = []
    for l in range(H.nDist):
        nc = sum(H.rank_matrix(r, l, 2)[0].shape[0] - 1 for r in range(H.nRanks))    # coarse rows of all ranks
        P = H.global_matrix(l, 1, (Al.shape[0], nc))
        R = H.global_matrix(l, 2, (nc, Al.shape[0]))
        assert np.all(np.diff(P.indptr) >= 1)
        if not fixed:
            assert np.abs(P @ np.ones(nc) - 1.0).max() < 1e-12          # constant reproduced across rank boundaries
        # R = the part of P^T inside the rank that owns the coarse row; P itself reaches across rank boundaries
        D = (P.T - R).tocoo()
        D.eliminate_zeros()
        assert R.nnz > 0 and D.nnz > 0 and abs(R - R.multiply(P.T != 0)).max() == 0.0
        G = (R @ Al @ P).tocsr()
        if l + 1 < H.nDist:
            An = H.global_matrix(l + 1, 0, (nc, nc))
        else:
            An = H.tail(0).mat(0, 0)[0]                                   # gathered level, replicated on every rank
            for r in range(1, H.nRanks):
                assert abs(H.tail(r).mat(0, 0)[0] - An).max() == 0.0
        assert An.shape == G.shape and abs(G - An).max() <= 1e-12 * abs(An).max()
        levels.append((Al, P, R))
        Al = An
    # halo lists: what r sends to q is what q expects from r, level by level
    for l in range(H.nDist):
        for r in range(H.nRanks):
            sp_r, si_r, rp_r = H.halo(r, l, 100000)
            gid_r = H.rank_matrix(r, l, 0)[3]
            for q in range(H.nRanks):
                if q == r:
                    continue
                sent = gid_r[si_r[sp_r[q]:sp_r[q + 1]]]
                rp_q = H.halo(q, l, 100000)[2]
                assert rp_q[r + 1] - rp_q[r] == len(sent)
                # ... cell by cell, in order: the k-th value r packs for q lands in q's k-th ghost slot of r
                assert np.array_equal(sent, H.ghost_gids(q, l)[rp_q[r]:rp_q[r + 1]])
    # a V(1,1) cycle over the global levels + the replicated tail preconditions BiCGStab like the serial one
    tail = H.tail(0).cycle()

    def cyc(l, b):
        if l == H.nDist:
            return tail(b)
        A_, P_, R_ = levels[l]
        rho = np.abs(sp.diags(1.0 / A_.diagonal()) @ A_).sum(axis=1).max()
        w = (4.0 / 3.0 / rho) / A_.diagonal()
        x = w * b
        x = x + P_ @ cyc(l + 1, R_ @ (b - A_ @ x))
        return x + w * (b - A_ @ x)
    b = np.random.default_rng(2).standard_normal(A.shape[0])
    if not fixed:
        b -= b.mean()
    its, rel = bicgstab_iters(A, lambda v: cyc(0, v), b)
    assert rel < 2e-8 and its <= 14, its
    H.close()


def test_nonsymmetric_momentum_like_matrix():
    """uEqn_-type operator: V/dt I + upwind convection - nu Laplacian (non-symmetric, strictly diagonally dominant);
    restriction = P^T is still a good transfer and the cycle preconditions BiCGStab into a handful of iterations"""
    nx = ny = 96
    h = 1.0 / nx
    L = neumann_laplacian(nx, ny, sign=1.0)                    # positive semi-definite 5-point operator (unit weights)
    idx = np.arange(nx * ny).reshape(ny, nx)
    # upwind convection with u = (1, 0.5): flux h*u through each face, donor = upstream cell
    I, J, V = [], [], []
    for (a, b, f) in ((idx[:, :-1].ravel(), idx[:, 1:].ravel(), 1.0 * h), (idx[:-1, :].ravel(), idx[1:, :].ravel(), 0.5 * h)):
        I += [a, b]; J += [a, a]; V += [np.full(len(a), f), np.full(len(a), -f)]   # outflow of a = inflow of b
    Cv = sp.csr_matrix((np.concatenate(V), (np.concatenate(I), np.concatenate(J))), shape=L.shape)
    A = (sp.eye(nx * ny) * (h * h / (0.5 * h)) + 0.5 * Cv + 0.05 * L).tocsr()
    assert abs(A - A.T).max() > 0
    H = HostAmg(A, coarsest=60)
    assert not H.singular and H.nLevels >= 3
    b = np.random.default_rng(3).standard_normal(nx * ny)
    its, rel = bicgstab_iters(A, H.cycle(), b)
    assert rel < 2e-8 and its <= 12, its
    H.close()


def test_setup_does_not_depend_on_the_thread_count(monkeypatch):
    A = neumann_laplacian(150, 120)
    mats = []
    for t in ("1", "5"):
        monkeypatch.setenv("PHB_HOST_THREADS", t)
        H = HostAmg(A, coarsest=50)
        mats.append([H.mat(l, w)[0] for l in range(H.nLevels - 1) for w in (0, 1, 2)])
        H.close()
    assert len(mats[0]) == len(mats[1])
    for a, b in zip(*mats):
        assert a.shape == b.shape and abs(a - b).max() == 0.0


@pytest.mark.parametrize("px,py", [(2, 1), (2, 2), (2, 4)])
def test_rank_local_cycle_with_halo_exchanges_equals_the_global_cycle(px, py):
    """numpy transcription of the DEVICE cycle (Cycle::run in amg.cu) on exactly the per-rank data it uses -- local
    matrices with ghost columns, send/receive lists, rank-local restriction, gathered tail -- against the same cycle
    on the assembled global matrices."""
    nx, ny = 48, 40
    A = neumann_laplacian(nx, ny)
    H = DistAmg(A, block_partition(nx, ny, px, py))
    NR, ND = H.nRanks, H.nDist
    toff = np.zeros(NR + 1, np.int32)
    assert H.L.phb_amg_dist_tail_offsets(H.h, toff.ctypes.data_as(_capi.pi)) == 0

    def local(rank, level, which, ncols):
        rp, ci, v, _ = H.rank_matrix(rank, level, which + 10)
        return sp.csr_matrix((v, ci, rp), shape=(len(rp) - 1, ncols))

    lev = []
    for l in range(ND):
        ranks = []
        for r in range(NR):
            sp_, si, rp_ = H.halo(r, l, 100000)
            n = H.rank_matrix(r, l, 10)[0].shape[0] - 1
            g = int(rp_[-1])
            ncr = H.rank_matrix(r, l, 12)[0].shape[0] - 1
            gcn = int(toff[-1]) if l + 1 == ND else None
            ranks.append(dict(n=n, g=g, sp=sp_, si=si, rp=rp_, nc=ncr, A=local(r, l, 0, n + g), Pc=gcn))
        lev.append(ranks)
    for l in range(ND):
        for r in range(NR):
            e = lev[l][r]
            mcols = e["Pc"] if e["Pc"] is not None else lev[l + 1][r]["n"] + lev[l + 1][r]["g"]
            e["P"] = local(r, l, 1, mcols)
            e["R"] = local(r, l, 2, e["n"])
            d = e["A"].diagonal()
            rho = (abs(e["A"]).sum(axis=1).A1 / abs(d)).max()
            e["w"] = (4.0 / 3.0 / rho) / d
    tail = H.tail(0).cycle()

    def halo(l, X):
        for r in range(NR):
            for q in range(NR):
                if q == r:
                    continue
                src, dst = lev[l][q], lev[l][r]
                vals = X[q][src["si"][src["sp"][r]:src["sp"][r + 1]]]
                X[r][dst["n"] + dst["rp"][q]:dst["n"] + dst["rp"][q + 1]] = vals

    def dist_cycle(b_ranks):
        xs, bs = [], [b_ranks]
        for l in range(ND):                                    # down
            X = []
            for r in range(NR):
                e = lev[l][r]
                x = np.zeros(e["n"] + e["g"])
                x[:e["n"]] = e["w"] * bs[l][r]
                X.append(x)
            halo(l, X)
            bc = [lev[l][r]["R"] @ (bs[l][r] - lev[l][r]["A"] @ X[r]) for r in range(NR)]
            xs.append(X)
            bs.append(bc)
        xg = tail(np.concatenate(bs[ND]))                      # gathered, replicated tail
        for l in range(ND - 1, -1, -1):                        # up
            X = xs[l]
            if l + 1 < ND:
                halo(l + 1, xs[l + 1])
            for r in range(NR):
                e = lev[l][r]
                xc = xg if l + 1 == ND else xs[l + 1][r]
                X[r][:e["n"]] += e["P"] @ xc
            halo(l, X)
            for r in range(NR):
                e = lev[l][r]
                X[r][:e["n"]] = X[r][:e["n"]] + e["w"] * (bs[l][r] - e["A"] @ X[r])
        return [X[r][:lev[0][r]["n"]] for r, X_ in enumerate(xs[0])]

    # the same cycle on the assembled global matrices, with the per-rank smoother weights scattered by global id
    glob, size = [], A.shape[0]
    Al = A
    for l in range(ND):
        nc = sum(lev[l][r]["nc"] for r in range(NR))
        P = H.global_matrix(l, 1, (size, nc))
        R = H.global_matrix(l, 2, (nc, size))
        w = np.zeros(size)
        for r in range(NR):
            w[H.rank_matrix(r, l, 0)[3]] = lev[l][r]["w"]
        glob.append((Al, P, R, w))
        Al, size = (R @ Al @ P).tocsr(), nc

    def glob_cycle(l, b):
        if l == ND:
            return tail(b)
        A_, P_, R_, w_ = glob[l]
        x = w_ * b
        x = x + P_ @ glob_cycle(l + 1, R_ @ (b - A_ @ x))
        return x + w_ * (b - A_ @ x)

    b = np.random.default_rng(5).standard_normal(A.shape[0])
    b -= b.mean()
    gids = [H.rank_matrix(r, 0, 0)[3] for r in range(NR)]
    xd = dist_cycle([b[g] for g in gids])
    xg = glob_cycle(0, b)
    for r in range(NR):
        assert np.abs(xd[r] - xg[gids[r]]).max() <= 1e-11 * np.abs(xg).max()
    H.close()


def test_dense_coarsest_level_of_a_thousand_rows():
    """the coarsest level may hold up to 1024 rows: LU + column solves (threaded) give the inverse to round-off,
    for a regular and for a singular (regularised) operator"""
    A = neumann_laplacian(31, 30, fixed_left=True)
    H = HostAmg(A, coarsest=1000)
    assert H.nLevels == 1 and H.dense
    assert np.abs(H.coarse_inverse() @ A.toarray() - np.eye(930)).max() < 1e-9
    H.close()
    A = neumann_laplacian(31, 30)
    H = HostAmg(A, coarsest=1000)
    inv = H.coarse_inverse()
    b = np.random.default_rng(0).standard_normal(930)
    b -= b.mean()
    x = inv @ b
    assert np.abs(A @ x - b).max() < 1e-9 * np.abs(b).max() * 930          # solves compatible systems of the singular operator
    H.close()
