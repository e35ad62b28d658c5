import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from phase_b200.api import Communicator, FiniteVolumeGrid2D, lid_driven_cavity
comm = Communicator(0)
n = 2000
g = FiniteVolumeGrid2D.rectilinear(comm, n, n, 1.0, 1.0)
fs = lid_driven_cavity(g, 1.0, 0.1, solver=dict(tolerance=1e-8, preconditioner="amg"))
dt = 0.5 / n
for _ in range(3): fs.solve(dt)
s = g.sizes(); N, F = s["nCells"], s["nFaces"]
host = {k: torch.empty(m, dtype=torch.float64).pin_memory() for k, m in (("uc", 2*N), ("uf", 2*F), ("pc", N))}
pag = {k: np.empty(v.numel()) for k, v in host.items()}
parts = {"uc": (fs.u, "cells"), "uf": (fs.u, "faces"), "pc": (fs.p, "cells")}
for k, (fld, part) in parts.items(): host[k].numpy()[:] = fld.get(part).reshape(-1)
for rep in range(3):
    for k, (fld, part) in parts.items():
        a = host[k].numpy()
        t0 = time.perf_counter(); fld.set(part, a); t1 = time.perf_counter(); fld.get(part, out=a); t2 = time.perf_counter()
        print(rep, k, "MB %.0f  set %.2f ms (%.1f GB/s)  get %.2f ms (%.1f GB/s)" % (a.nbytes/1e6, (t1-t0)*1e3, a.nbytes/(t1-t0)/1e9, (t2-t1)*1e3, a.nbytes/(t2-t1)/1e9))
t0 = time.perf_counter(); fs.p.sendMessages(); fs.p.setBoundaryFaces(); fs.computeGradP(); comm.sync(); print("glue ms", (time.perf_counter()-t0)*1e3)
for _ in range(3):
    t0 = time.perf_counter(); st = fs.solve(dt); print("solve ms", (time.perf_counter()-t0)*1e3, st["itersU"], st["itersP"])
