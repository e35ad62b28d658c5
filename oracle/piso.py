"""CPU restatement of the phasePiso time step -- TEST INFRASTRUCTURE ONLY.

The mounted snapshot ships no PISO/SIMPLE module any more: what is left of it is README.md:26-37
(uEqn_ = ddt(rho,u,dt) + div(rho*u,u) == laplacian(mu,u) - grad(p), relax, solve; pCorrEqn_ = laplacian(rho*d, pCorr)
== m; corrections), the commented body of relax() (UE/ScalarFiniteVolumeEquation.cpp:45-55) and the legacy case keys
(Examples/LidDrivenCavity/case/case.info:12-15).  PARITY UNPINNED: there is no reference implementation to run; this
file is an INDEPENDENT numpy / scipy transcription of those equations (dense link loops over the oracle's mesh tables,
exact sparse LU solves) against which the CUDA time step (phase_b200/csrc/piso.cu) is compared field by field.  The
operators it is built from (ddt, div, laplacian, gradient, face interpolation) are the ones pinned on the reference by
tests/test_oracle_ref_fv.py.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl

FIXED, NORMAL_GRADIENT = 0, 1


class Piso:
    def __init__(self, mesh, rho, mu, u_bc, p_bc, num_inner=1, num_corr=1, omega_u=0.8, omega_p=0.2):
        """mesh: oracle.Mesh; u_bc / p_bc: {patch name: (type, value)} (value = (ux, uy) or scalar)"""
        m = self.m = mesh
        self.rho, self.mu, self.nI, self.nC, self.wu, self.wp = rho, mu, num_inner, num_corr, omega_u, omega_p
        a = m.array
        self.N, self.F = m.sizes["nCells"], m.sizes["nFaces"]
        N, F = self.N, self.F
        self.vol = a("vol")
        self.ilPtr, self.ilFace, self.ilCell = a("ilPtr"), a("ilFace"), a("ilCell")
        self.blPtr, self.blFace = a("blPtr"), a("blFace")
        self.ilS = np.stack([a("ilSx"), a("ilSy")])
        rc = np.stack([a("ilRcx"), a("ilRcy")])
        self.ilG = (rc * self.ilS).sum(0) / (rc ** 2).sum(0)
        self.ilQ = rc / (rc ** 2).sum(0)
        self.blS = np.stack([a("blSx"), a("blSy")])
        rf = np.stack([a("blRfx"), a("blRfy")])
        self.blG = (rf * self.blS).sum(0) / (rf ** 2).sum(0)
        self.blQ = rf / (rf ** 2).sum(0)
        self.ilRow = np.repeat(np.arange(N), np.diff(self.ilPtr))
        self.blRow = np.repeat(np.arange(N), np.diff(self.blPtr))
        fl, fr = a("faceL"), a("faceR")
        self.fl, self.fr = fl, fr
        cc = np.stack([a("cellCx"), a("cellCy")])
        fc = np.stack([a("faceCx"), a("faceCy")])
        self.interior = fr >= 0
        frs = np.maximum(fr, 0)
        l1 = np.linalg.norm(fc - cc[:, frs], axis=0)
        l2 = np.linalg.norm(fc - cc[:, fl], axis=0)
        self.fw = np.where(self.interior, l1 / (l1 + l2), 1.0)          # weight of lCell (UG/Face/Face.cpp:60-64)
        # face geometry seen from lCell: S_f, r/|r|^2 (interior: c_r - c_l; boundary: c_f - c_l)
        r = np.where(self.interior, cc[:, frs] - cc[:, fl], fc - cc[:, fl])
        self.fQ = r / (r ** 2).sum(0)
        fp = a("facePatch")
        self.uType = np.full(F, NORMAL_GRADIENT); self.pType = np.full(F, NORMAL_GRADIENT)
        self.uRef = np.zeros((2, F)); self.pRef = np.zeros(F)
        for name, (t, v) in u_bc.items():
            sel = fp == m.patch_id(name)
            self.uType[sel] = t; self.uRef[0, sel], self.uRef[1, sel] = v
        for name, (t, v) in p_bc.items():
            sel = fp == m.patch_id(name)
            self.pType[sel] = t; self.pRef[sel] = v
        self.u, self.uf = np.zeros((2, N)), np.zeros((2, F))
        self.p, self.pf, self.pCorr = np.zeros(N), np.zeros(F), np.zeros(N)
        self.gP, self.gPf = np.zeros((2, N)), np.zeros((2, F))
        self.d = np.zeros(N)
        bnd = ~self.interior
        fixed_u = bnd & (self.uType == FIXED)
        self.uf[:, fixed_u] = self.uRef[:, fixed_u]
        fixed_p = bnd & (self.pType == FIXED)
        self.pf[fixed_p] = self.pRef[fixed_p]
        self.singular = not fixed_p.any()
        self.initialize()

    # ---- field glue (pinned operators)
    def interp(self, c):
        frs = np.maximum(self.fr, 0)
        return self.fw * c[..., self.fl] + (1.0 - self.fw) * c[..., frs]

    def u_faces(self):
        f = self.interp(self.u)
        ng = ~self.interior & (self.uType == NORMAL_GRADIENT)
        keep = ~self.interior & (self.uType == FIXED)
        f[:, keep] = self.uf[:, keep]
        f[:, ng] = self.u[:, self.fl[ng]]
        self.uf = f

    def scalar_boundary_faces(self, c, f, types):
        ng = ~self.interior & (types != FIXED)
        f[ng] = c[self.fl[ng]]

    def gradient(self, c, f):
        frs = np.maximum(self.fr, 0)
        dphi = np.where(self.interior, c[frs], f) - c[self.fl]
        gf = dphi * self.fQ
        num, den = np.zeros((2, self.N)), np.zeros((2, self.N))
        ai, ab = np.abs(self.ilS), np.abs(self.blS)
        for k in (0, 1):
            np.add.at(num[k], self.ilRow, gf[k, self.ilFace] * ai[k]); np.add.at(den[k], self.ilRow, ai[k])
            np.add.at(num[k], self.blRow, gf[k, self.blFace] * ab[k]); np.add.at(den[k], self.blRow, ab[k])
        return num / den, gf

    def initialize(self):
        self.u_faces()
        self.scalar_boundary_faces(self.p, self.pf, self.pType)
        self.gP, self.gPf = self.gradient(self.p, self.pf)

    def solve_system(self, A, b, singular=False):
        A = A.tocsc()
        if singular:
            n = A.shape[0]
            x = np.zeros(n)
            x[1:] = spl.splu(A[1:, 1:]).solve((b - b.mean())[1:])
            return x
        return spl.splu(A).solve(b)

    def step(self, dt):
        N, rho, mu = self.N, self.rho, self.mu
        u0 = self.u.copy()
        for _ in range(self.nI):
            Fi = rho * (self.uf[:, self.ilFace] * self.ilS).sum(0)
            Fb = rho * (self.uf[:, self.blFace] * self.blS).sum(0)
            diag = rho * self.vol / dt
            rhs = -rho * self.vol * u0 / dt                       # equation A u + rhs = 0
            off = np.minimum(Fi, 0.0) - mu * self.ilG            # div (theta = 1, upwind) - laplacian
            diag = diag + np.bincount(self.ilRow, np.maximum(Fi, 0.0) + mu * self.ilG, N)
            bt = self.uType[self.blFace]
            fx = bt == FIXED
            diag = diag + np.bincount(self.blRow, np.where(fx, mu * self.blG, Fb), N)
            for k in (0, 1):
                rhs[k] += np.bincount(self.blRow, np.where(fx, (Fb - mu * self.blG) * self.uf[k, self.blFace], 0.0), N)
            rhs += self.gP * self.vol
            a = diag / self.wu                                    # relax(omega)
            rhs -= (1.0 - self.wu) * a * self.u
            diag = a
            self.d = self.vol / diag
            A = sp.csr_matrix((np.concatenate([diag, off]), (np.concatenate([np.arange(N), self.ilRow]),
                                                            np.concatenate([np.arange(N), self.ilCell]))), shape=(N, N))
            lu = spl.splu(A.tocsc())
            self.u = np.stack([lu.solve(-rhs[0]), lu.solve(-rhs[1])])
            df = self.interp(self.d)
            bnd = ~self.interior
            df[bnd] = self.d[self.fl[bnd]]
            for _ in range(self.nC):
                self.u_faces()
                it = self.interior
                frs = np.maximum(self.fr, 0)
                gbar = self.fw * self.gP[:, self.fl] + (1.0 - self.fw) * self.gP[:, frs]
                self.uf[:, it] -= (df * (self.gPf - gbar))[:, it]              # Rhie-Chow
                ci = rho * df[self.ilFace] * self.ilG
                pdiag = -np.bincount(self.ilRow, ci, N)
                pfx = self.pType[self.blFace] == FIXED
                pdiag = pdiag - np.bincount(self.blRow, np.where(pfx, rho * df[self.blFace] * self.blG, 0.0), N)
                mass = np.bincount(self.ilRow, rho * (self.uf[:, self.ilFace] * self.ilS).sum(0), N) + \
                    np.bincount(self.blRow, rho * (self.uf[:, self.blFace] * self.blS).sum(0), N)
                Ap = sp.csr_matrix((np.concatenate([pdiag, ci]), (np.concatenate([np.arange(N), self.ilRow]),
                                                                  np.concatenate([np.arange(N), self.ilCell]))), shape=(N, N))
                self.pCorr = self.solve_system(Ap, mass, self.singular)
                pcf = np.zeros(self.F)
                self.scalar_boundary_faces(self.pCorr, pcf, self.pType)
                gC, gCf = self.gradient(self.pCorr, pcf)
                self.p = self.p + self.wp * self.pCorr
                self.u = self.u - self.d * gC
                self.uf = self.uf - df * gCf
                self.scalar_boundary_faces(self.p, self.pf, self.pType)
                self.gP, self.gPf = self.gradient(self.p, self.pf)
        mass = np.bincount(self.ilRow, (self.uf[:, self.ilFace] * self.ilS).sum(0), N) + \
            np.bincount(self.blRow, (self.uf[:, self.blFace] * self.blS).sum(0), N)
        return float(np.abs(mass).max())
