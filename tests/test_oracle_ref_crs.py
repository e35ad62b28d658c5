"""Our CrsEquation restatement (oracle/phase_oracle.c) against the reference's
own M/CrsEquation.cpp compiled in place (oracle/_ref): bit-exact patterns and
values on random operation sequences and on the hot path's operator orders."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.skipif(O.ref_lib() is None, reason="oracle/_ref not built and /root/reference absent")


def same(a, b):
    ra, rb = a.export(), b.export()
    for x, y in zip(ra, rb):
        assert x.shape == y.shape
        assert np.array_equal(x, y)


def random_eq(rng, n, nnz, fill):
    a, b = O.Crs(n, nnz), O.RefCrs(n, nnz)
    for _ in range(fill):
        r, c = int(rng.integers(n)), int(rng.integers(n))
        v = float(rng.choice([0.0, 1.0, -2.5, rng.standard_normal()]))
        if rng.random() < 0.8:
            a.add_coeff(r, c, v); b.add_coeff(r, c, v)
        else:
            a.set_coeff(r, c, v); b.set_coeff(r, c, v)
        if rng.random() < 0.3:
            w = float(rng.standard_normal())
            a.add_rhs(r, w); b.add_rhs(r, w)
    return a, b


@pytest.mark.parametrize("seed", range(6))
def test_random_algebra(seed):
    rng = np.random.default_rng(seed)
    n = 17
    a1, b1 = random_eq(rng, n, 3, 80)   # overflows rows -> insert fallback
    a2, b2 = random_eq(rng, n, 5, 60)
    a3, b3 = random_eq(rng, n, 2, 40)
    same(a1, b1); same(a2, b2)
    a1.add_eq(a2); b1.add_eq(b2); same(a1, b1)
    a1.sub_eq(a3); b1.sub_eq(b3); same(a1, b1)
    v = rng.standard_normal(n)
    a1.sub_vec(v); b1.sub_vec(v)
    a1.scale(0.37); b1.scale(0.37)
    a1.scale_row(3, -2.0); b1.scale_row(3, -2.0)
    same(a1, b1)
    c1, d1 = a1.clone(), b1.clone()
    c1.add_coeff(0, 16, 9.0); d1.add_coeff(0, 16, 9.0)
    same(c1, d1); same(a1, b1)


def test_hot_path_operator_order_matches_reference():
    """Replay the add/addSource call sequence of uEqn_ / pEqn_ on the REFERENCE
    CrsEquation and compare with the oracle's assembled equations."""
    m = O.Mesh.rectilinear(5, 4, 1.0, 0.8)
    fs = O.cavity(m, 1.0, 0.1)
    rng = np.random.default_rng(1)
    for nm in ("ux", "uy", "ufx", "ufy", "gpx", "gpy"):
        fs.view(nm)[:] = rng.standard_normal(fs.view(nm).shape)
    for nm in ("u0x", "u0y", "u0fx", "u0fy"):
        fs.view(nm)[:] = rng.standard_normal(fs.view(nm).shape)
    dt = 0.05
    N = m.sizes["nCells"]
    ilPtr, ilFace, ilCell = m.array("ilPtr"), m.array("ilFace"), m.array("ilCell")
    blPtr, blFace = m.array("blPtr"), m.array("blFace")
    ilS = np.stack([m.array("ilSx"), m.array("ilSy")], 1)
    ilRc = np.stack([m.array("ilRcx"), m.array("ilRcy")], 1)
    blS = np.stack([m.array("blSx"), m.array("blSy")], 1)
    blRf = np.stack([m.array("blRfx"), m.array("blRfy")], 1)
    vol = m.array("vol")
    ux, uy, ufx, ufy = (fs.view(k).copy() for k in ("ux", "uy", "ufx", "ufy"))
    u0x, u0y, u0fx, u0fy = (fs.view(k).copy() for k in ("u0x", "u0y", "u0fx", "u0fy"))
    gpx, gpy = fs.view("gpx").copy(), fs.view("gpy").copy()

    def vadd(e, c, nb, v):
        e.add_coeff(c, nb, v); e.add_coeff(N + c, N + nb, v)

    def vsrc(e, c, sx, sy):
        e.add_rhs(c, sx); e.add_rhs(N + c, sy)

    # --- pEqn on the reference class
    pe = O.RefCrs(N, 5)
    for c in range(N):
        for j in range(ilPtr[c], ilPtr[c + 1]):
            coeff = dt * (ilRc[j] @ ilS[j]) / (ilRc[j] @ ilRc[j])
            pe.add_coeff(c, int(ilCell[j]), coeff)
            pe.add_coeff(c, c, -coeff)
    div = np.zeros(N)
    for c in range(N):
        d = 0.0
        for j in range(ilPtr[c], ilPtr[c + 1]):
            d += ufx[ilFace[j]] * ilS[j, 0] + ufy[ilFace[j]] * ilS[j, 1]
        for j in range(blPtr[c], blPtr[c + 1]):
            d += ufx[blFace[j]] * blS[j, 0] + ufy[blFace[j]] * blS[j, 1]
        div[c] = d
    pe.sub_vec(div)
    same(fs.assemble_p(dt), pe)
    rp, ci, va, b = pe.solve_handoff()       # set() + setRhs(-rhs_)
    assert np.array_equal(b, -pe.export()[3])

    # --- uEqn on the reference class (all u patches FIXED in the cavity)
    e1, e2, e3 = O.RefCrs(2 * N, 5), O.RefCrs(2 * N, 5), O.RefCrs(2 * N, 5)
    th = 0.0
    for c in range(N):
        vadd(e1, c, c, vol[c] / dt)
        vsrc(e1, c, -vol[c] * u0x[c] / dt, -vol[c] * u0y[c] / dt)
    for c in range(N):
        for j in range(ilPtr[c], ilPtr[c + 1]):
            f, nb = ilFace[j], int(ilCell[j])
            flux = ufx[f] * ilS[j, 0] + ufy[f] * ilS[j, 1]
            flux0 = u0fx[f] * ilS[j, 0] + u0fy[f] * ilS[j, 1]
            vadd(e2, c, c, th * max(flux, 0.0))
            vadd(e2, c, nb, th * min(flux, 0.0))
            a = (1.0 - th) * max(flux0, 0.0)
            vsrc(e2, c, a * u0x[c], a * u0y[c])
            b_ = (1.0 - th) * min(flux0, 0.0)
            vsrc(e2, c, b_ * u0x[nb], b_ * u0y[nb])
        for j in range(blPtr[c], blPtr[c + 1]):
            f = blFace[j]
            flux = ufx[f] * blS[j, 0] + ufy[f] * blS[j, 1]
            flux0 = u0fx[f] * blS[j, 0] + u0fy[f] * blS[j, 1]
            vsrc(e2, c, th * flux * ufx[f], th * flux * ufy[f])
            vsrc(e2, c, (1.0 - th) * flux0 * u0fx[f], (1.0 - th) * flux0 * u0fy[f])
    e1.add_eq(e2)
    th, gamma = 0.5, 0.1 / 1.0
    for c in range(N):
        for j in range(ilPtr[c], ilPtr[c + 1]):
            nb = int(ilCell[j])
            coeff = gamma * (ilRc[j] @ ilS[j]) / (ilRc[j] @ ilRc[j])
            vadd(e3, c, nb, th * coeff)
            vadd(e3, c, c, th * -coeff)
            a = (1.0 - th) * coeff
            vsrc(e3, c, a * (u0x[nb] - u0x[c]), a * (u0y[nb] - u0y[c]))
        for j in range(blPtr[c], blPtr[c + 1]):
            f = blFace[j]
            coeff = gamma * (blRf[j] @ blS[j]) / (blRf[j] @ blRf[j])
            vadd(e3, c, c, th * -coeff)
            vsrc(e3, c, th * coeff * ufx[f], th * coeff * ufy[f])
            a = (1.0 - th) * coeff
            vsrc(e3, c, a * (u0fx[f] - u0x[c]), a * (u0fy[f] - u0y[c]))
    src = np.concatenate([gpx * vol, gpy * vol])
    e3.sub_vec(src)
    e1.sub_eq(e3)
    ours = fs.assemble_u(dt)
    ra, rb = ours.export(), e1.export()
    assert np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])   # pattern: bit-exact
    assert np.allclose(ra[2], rb[2], rtol=1e-14, atol=0) and np.allclose(ra[3], rb[3], rtol=1e-12, atol=1e-14)
