import numpy as np, sys
sys.path.insert(0,'/root/repo')
from phase_b200.api import *
comm=Communicator(0)
def run(n,dt,corr,wu,wp,steps):
    g=FiniteVolumeGrid2D.rectilinear(comm,n,n)
    ps=Piso(g,1.0,0.1,numInnerIterations=1,numPressureCorrections=corr,momentumRelaxation=wu,pressureCorrectionRelaxation=wp)
    for pt in ("x-","x+","y-"): ps.u.setBoundary(pt,FIXED,(0,0))
    ps.u.setBoundary("y+",FIXED,(1.0,0))
    for pt in ("x-","x+","y-","y+"): ps.p.setBoundary(pt,NORMAL_GRADIENT,0.0)
    cfg=dict(maxIters=5000,tolerance=1e-10,preconditioner="jacobi")
    ps.uSolver.setup(cfg); ps.pCorrSolver.setup(cfg)
    ps.initialize()
    for k in range(steps):
        try:
            st=ps.solve(dt)
        except Exception as e:
            print("ERR at",k,e); break
        if k%20==0 or k<5:
            u=ps.u.get("cells"); p=ps.p.get("cells")
            print(n,dt,corr,wu,wp,k,"umax %.3e pmax %.3e m %.2e itU %d itP %d"%(np.abs(u).max(),np.abs(p).max(),st["maxMassImbalance"],st["itersU"],st["itersPCorr"]))
    ps.close(); g.close()
run(32,0.05,2,0.8,0.3,120)
run(32,0.05,1,0.8,0.2,120)
run(32,10.0,1,0.8,0.2,100)
