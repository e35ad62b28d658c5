"""Shared set-up of the multiphase (rising-bubble style) parity cases: the same boundary conditions, properties and
initial volume fraction for the reference library (oracle/ref_fv.py) and the CUDA path."""
import numpy as np

PROPS = dict(rho1=998.0, rho2=1.225, mu1=8.94e-4, mu2=1.84e-5, sigma=0.0762, g=(0.0, -9.8065))
BCS = {"u": {"*": ("fixed", "(0,0)"), "y+": ("normal_gradient", "(0,0)")},
       "p": {"*": ("normal_gradient", "0"), "y+": ("fixed", "0")},
       "gamma": {"*": ("normal_gradient", "0")}}


def initial_gamma(cx, cy, width, height):
    """a smooth bubble of the light phase under a free surface (Examples/RisingBubble geometry, tanh profiles)"""
    r = np.hypot(cx - 0.5 * width, cy - 0.25 * height)
    bubble = 0.5 * (1.0 - np.tanh((r - 0.125 * height) / (0.02 * height)))
    surface = 0.5 * (1.0 + np.tanh((cy - 0.75 * height) / (0.02 * height)))
    return np.maximum(bubble, surface)


def reference_case(R, nx, ny, width, height, radius):
    return R.Case(nx, ny, width, height, 1.0, 1.0, time_step=1.0, bcs=BCS,
                  properties=dict(rho1=PROPS["rho1"], rho2=PROPS["rho2"], mu1=PROPS["mu1"], mu2=PROPS["mu2"],
                                  sigma=PROPS["sigma"], g="(%.17g,%.17g)" % PROPS["g"]),
                  solver=dict(smoothingKernelRadius="%.17g" % radius))


def device_solver(grid, radius, tol=1e-12):
    from phase_b200.api import FIXED, NORMAL_GRADIENT, FractionalStepMultiphase
    mp = FractionalStepMultiphase(grid, PROPS["rho1"], PROPS["rho2"], PROPS["mu1"], PROPS["mu2"], PROPS["sigma"], PROPS["g"], radius)
    for pt in ("x-", "x+", "y-", "y+"):
        mp.u.setBoundary(pt, NORMAL_GRADIENT if pt == "y+" else FIXED, (0.0, 0.0))
        mp.p.setBoundary(pt, FIXED if pt == "y+" else NORMAL_GRADIENT, 0.0)
        mp.gamma.setBoundary(pt, NORMAL_GRADIENT, 0.0)
    cfg = dict(solver="BICGSTAB", maxIters=20000, tolerance=tol, preconditioner="ilu0")
    for e in (mp.gammaEqn, mp.uEqn, mp.pEqn):
        e.solver.setup(cfg)
    return mp


SCALARS = ("gamma", "rho", "mu", "kappa", "gammaTilde", "p")
VECTORS = ("u", "sg", "fst", "n", "gradGamma", "gradP")
# reference field names (Solver::scalarField / vectorField): ScalarGradient fields are called "grad" + name
REF_NAME = {"gradGamma": "gradgamma", "gradP": "gradp"}
