#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark on B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port, all host cores)

Workload (BASELINE.json configs[1]): lid-driven cavity, rho=1, mu=0.1, on a synthetic 2000x2000 quad mesh
(4M cells); one "step" = one whole time step of the snapshot's solver module (FractionalStep::solve: assemble
uEqn_, BiCGStab, interpolate, assemble pEqn_, BiCGStab, gradient, correct).  The north star calls the time step
"PISO"; the mounted snapshot ships only the fractional-step successor, which is what is timed
and parity-checked.  With N > 1 the same 4M-cell problem is split over the GPUs (strong scaling, `value`) and the
line also carries the 8M-cells-per-GPU series that ends in the north star's 64M-cell / 8-GPU point
(`weak_8M_per_gpu`) and a multi-GPU parity check against the oracle (`parity`).

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "PISO time-steps/s at 4M cells; BiCGStab SpMV HBM GB/s vs B200 peak"
UNIT = "time-steps/s"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.proc = index, [], set(), None
        self.max_mhz = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            p = [x.strip() for x in line.split(",")]
            try:
                active = [n for n, v in zip(names, p[2:6]) if v.lower().startswith("active")]
                self.samples.append((time.perf_counter(), float(p[0]), active))
                self.max_mhz = float(p[1])
            except Exception:
                pass

    def mark(self):
        """start of the timed region (the sampler itself is started earlier: nvidia-smi needs ~0.1 s to come up)"""
        self.t0 = time.perf_counter()

    def stop(self):
        t1 = time.perf_counter()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        t0 = getattr(self, "t0", 0.0)
        inside = [x for x in self.samples if t0 <= x[0] <= t1]
        # a timed region shorter than the sampling period: fall back to the samples closest to it (under the same load:
        # the last warm-up step) and say so
        used = inside if inside else self.samples[-2:]
        mhz = sorted(x[1] for x in used)
        reasons = sorted({r for x in used for r in x[2]})
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(inside), "samples_used": len(used), "period_ms": 50}


# ------------------------------------------------------------------------------------------------ workload
def block_layout(nprocs):
    """px x py blocks, as square as possible, the extra factor of two going to y (1 -> 1x1, 2 -> 1x2, 4 -> 2x2,
    8 -> 2x4): interfaces parallel to the row-major cell numbering cost the rank-local aggregation of the multigrid
    setup fewer iterations than interfaces across it (4M cells, pEqn_: 11 vs 16 iterations on 2 ranks;
    tools/proto/dist_iters.py)."""
    px = 1
    while px * px * 4 <= nprocs and nprocs % (px * 2) == 0:
        px *= 2
    return px, nprocs // px


def workload_config(args):
    """What is computed -- identical for both arms and (strong scaling) for every N.  How it is computed
    (preconditioner, precision, partition, exchange mechanism) is each arm's `algorithm` record."""
    import math
    if args.mesh == "tri":
        tn = int(round(args.n / math.sqrt(2.0)))
        mesh = "unstructured triangles: %dx%d lattice split along alternating diagonals = %d cells" % (tn, tn, 2 * tn * tn)
        cells = 2 * tn * tn
    else:
        mesh = "%dx%d quads = %d cells" % (args.n, args.n, args.n * args.n)
        cells = args.n * args.n
    per = " per GPU (weak scaling: the global grid grows with N)" if args.scaling == "weak" else ""
    return {"workload": "lid-driven cavity (rho=1, mu=0.1, lid u=1), %s%s, FractionalStep time step (the snapshot's PISO "
                        "successor), dt = 0.5 h (maxCo 0.5), both equations solved by right-preconditioned BiCGStab to "
                        "||r||/||b|| <= %g, warm start from the previous step" % (mesh, per, args.tol),
            "cells": cells, "mesh": args.mesh, "tolerance": args.tol, "max_iters": args.max_iters,
            "l2": "inputs larger than L2 (matrix + vectors ~0.6 GB per solve vs 126 MB L2); no flush needed"}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_converged_steps(nx, ny, dt, tol, max_iters, n_warm, n_timed, budget_s, state=None):
    """The oracle (CPU port of the reference path: cell-loop assembly into CrsEquation-style CSR, BiCGStab + ILU(0),
    field glue) advancing the SAME cavity with every solve converged to `tol`: real wall time per step, nothing
    scaled.  `state` (u, p, gradP of another run) replaces the rest state; timed steps stop at the wall budget."""
    import numpy as np
    import oracle as O
    cores = O.set_num_threads()            # all host cores, whatever OMP_NUM_THREADS the launcher exported
    t0 = time.perf_counter()
    om = O.Mesh.rectilinear(nx, ny, 1.0, 1.0)
    ofs = O.cavity(om, 1.0, 0.1)
    t_mesh = time.perf_counter() - t0
    guesses = None
    if state is not None:
        for k in ("ux", "uy", "ufx", "ufy", "p", "pf", "gpx", "gpy"):
            ofs.view(k)[:] = state[k]
        guesses = {2 * nx * ny: np.concatenate([state["ux"], state["uy"]]), nx * ny: state["p"]}
    ofs.use_ilu0_solver(tol=tol, max_iters=max_iters, guesses=guesses)
    for _ in range(n_warm):
        ofs.step(dt)
    n0 = len(ofs.solve_log)
    times = []
    t_start = time.perf_counter()
    for _ in range(n_timed):
        t0 = time.perf_counter()
        ofs.step(dt)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    log = ofs.solve_log[n0:]
    detail = {"threads": cores, "t_mesh_s": t_mesh, "steps_timed": len(times), "warmup_steps": n_warm,
              "s_per_step": times, "preconditioner": "ilu0 (multicolour ordering, OpenMP over the rows of a colour)",
              "assembly": "serial cell loops (the reference has no threads)",
              "iters_u": [it for n, it, rr, t in log if n == 2 * nx * ny],
              "iters_p": [it for n, it, rr, t in log if n == nx * ny],
              "relres_max": max([rr for n, it, rr, t in log] or [0.0]),
              "solve_s_per_step": sum(t for n, it, rr, t in log) / max(1, len(times))}
    return len(times) / sum(times), cores, detail


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nx = ny = args.n
    dt = 0.5 / nx
    n_warm = 1 if args.warmup > 0 else 0
    v, cores, detail = cpu_converged_steps(nx, ny, dt, args.tol, args.max_iters, n_warm, max(1, args.steps),
                                           budget_s=args.ref_budget)
    k = detail["steps_timed"]
    sample = ("%d fully converged time steps (steps %d..%d from rest) of the %dx%d cavity, tolerance %g, after %d converged "
              "warm-up step; wall clock per step, nothing extrapolated" % (k, n_warm + 1, n_warm + k, nx, ny, args.tol, n_warm))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": k, "steps_requested": args.steps, "warmup": n_warm, "warmup_requested": args.warmup,
            "ms_per_step": 1e3 / v, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "algorithm": {"preconditioner": "ilu0", "what": "the reference's path: host CrsEquation-style assembly + right-"
                          "preconditioned BiCGStab with ILU(0) (the Belos/Ifpack2 RILUK(0) default, "
                          "M/TrilinosBelosSparseMatrixSolver.cpp:52-83) on all host cores", "threads": cores},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "detail": detail},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
class Job:
    """torch.distributed + one phase_b200 context per sub-problem."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def communicator(self):
        from phase_b200.api import Communicator
        uid = None
        if self.world > 1:
            box = [Communicator.unique_id() if self.rank == 0 else None]
            self.dist.broadcast_object_list(box, src=0)
            uid = box[0]
        return Communicator(self.local_rank, self.rank, self.world, uid)

    def all_gather(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return float(v)
        t = self.torch.tensor([v], device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, vals):
        if self.world == 1:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t)
        return [float(v) for v in t.tolist()]


def solver_keys(args, precond, extra=None):
    cfg = dict(solver="BICGSTAB", maxIters=args.max_iters, tolerance=args.tol, preconditioner=precond,
               peerFusion=1 if args.comm == "peer-fused" else 0)
    for kv in args.solver_key:
        k, v = kv.split("=", 1)
        cfg[k] = v
    cfg.update(extra or {})
    return cfg


def make_cavity(job, comm, args, nx, ny, width, height, px, py, u_pc, p_pc, extra=None, mesh="quad"):
    """Grid (one block per rank) + FractionalStep with the lid-driven-cavity boundary conditions."""
    from phase_b200.api import Communicator, FiniteVolumeGrid2D, lid_driven_cavity
    world = job.world
    if mesh == "tri":
        if world == 1:
            grid = FiniteVolumeGrid2D.triangulated(comm, nx, ny, width, height)
        else:                                                # generic path: global mesh on the host, RCB, local mesh
            hostc = Communicator(Communicator.HOST_ONLY)
            gg = FiniteVolumeGrid2D.triangulated(hostc, nx, ny, width, height)
            grid = gg.local(gg.partition_rcb(world), comm)
            gg.close()
    elif world == 1:
        grid = FiniteVolumeGrid2D.rectilinear(comm, nx, ny, width, height)
    else:
        grid = FiniteVolumeGrid2D.rectilinear_block(comm, nx, ny, width, height, px, py)
    if world > 1 and args.comm.startswith("peer"):
        comm.enable_peer_memory(grid, job.all_gather)
    fs = lid_driven_cavity(grid, 1.0, 0.1, solver=solver_keys(args, u_pc, extra),
                           pSolver=dict(preconditioner=p_pc) if p_pc != u_pc else None)
    fs.setup(guessOrder=args.guess_order)
    return grid, fs


def timed_steps(job, comm, fs, dt, warmup, steps, sampler=None):
    """W untimed + K timed steps, CUDA events on the library's stream, barrier + synchronize on both sides,
    max over ranks."""
    torch = job.torch
    stream = torch.cuda.ExternalStream(comm.stream())
    for _ in range(warmup):
        fs.solve(dt)
    job.barrier()
    if sampler:
        sampler.mark()
    launches0 = comm.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
    stats = [fs.solve(dt) for _ in range(steps)]
    with torch.cuda.stream(stream):
        e1.record()
    job.barrier()
    ms = job.max_over_ranks(e0.elapsed_time(e1))
    return ms / steps, stats, comm.kernel_launches() - launches0


def pressure_solve_record(job, comm, fs, peak):
    """pEqn_ as assembled by the last step, solved from a zero guess: whole-solve algorithmic throughput =
    iterations x bytes per iteration (2 SpMV + 2 preconditioner applies + vector passes)."""
    torch = job.torch
    stream = torch.cuda.ExternalStream(comm.stream())
    fs.pEqn.solve(warmStart=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    job.barrier()
    with torch.cuda.stream(stream):
        e0.record()
    fs.pEqn.solve(warmStart=False)
    with torch.cuda.stream(stream):
        e1.record()
    job.barrier()
    ms = job.max_over_ranks(e0.elapsed_time(e1))
    its = fs.pEqn.solver.nIters()
    _, b_iter = fs.pEqn.solver.bytes()
    gbs = its * b_iter / (ms * 1e-3) / 1e9
    return {"what": "pEqn_ of the last step solved from a zero guess, CUDA events around phb_eqn_solve (max over ranks)",
            "iterations": its, "ms": ms, "ms_per_iteration": ms / max(1, its), "relres": fs.pEqn.solver.error(),
            "algorithmic_bytes_per_iteration_per_gpu": b_iter, "achieved_GBps_per_gpu": gbs,
            "frac_of_measured_peak": gbs / peak, "frac_of_8TBs_nominal": gbs / 8000.0}


def mean(xs):
    xs = list(xs)
    return float(sum(xs)) / max(1, len(xs))


def sub_record(job, args, u_pc, p_pc, extra, warmup, steps, what):
    """The same 4M-cell workload under another algorithm choice (single GPU), device-resident, same timing rules."""
    comm = job.communicator()
    grid, fs = make_cavity(job, comm, args, args.n, args.n, 1.0, 1.0, 1, 1, u_pc, p_pc, extra)
    dt = 0.5 / args.n
    ms, stats, launches = timed_steps(job, comm, fs, dt, warmup, steps)
    peak, _ = measured_peak()
    rec = {"what": what, "value": 1e3 / ms, "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup,
           "iters_per_solve": {"uEqn": mean(s["itersU"] for s in stats), "pEqn": mean(s["itersP"] for s in stats)},
           "relres_p": stats[-1]["errorP"], "gpu_launches": int(launches),
           "pressure_solve": pressure_solve_record(job, comm, fs, peak)}
    fs.close(); grid.close(); comm.close()
    return rec


def parity_record(job, args):
    """A small cavity on the same communicator layout and default solver keys against the oracle's single-domain
    direct solve (4 steps, rel-L2 of the owned cells; p minus its global mean).  N > 1: RCB partition, the
    distributed multigrid hierarchy forced onto three levels, peer-memory exchanges."""
    import numpy as np
    import oracle as O
    from phase_b200.api import Communicator, FiniteVolumeGrid2D as G, lid_driven_cavity
    nx, ny, steps = 64, 48, 4
    comm = job.communicator()
    if job.world == 1:
        gl = G.rectilinear(comm, nx, ny, 1.0, 1.0)
    else:
        host = Communicator(Communicator.HOST_ONLY)
        g = G.rectilinear(host, nx, ny, 1.0, 1.0)
        gl = g.local(g.partition_rcb(job.world), comm)
        g.close()
        if args.comm.startswith("peer"):
            comm.enable_peer_memory(gl, job.all_gather)
    amg = args.precond == "amg"
    keys = dict(tolerance=1e-11, maxIters=50000, preconditioner=args.u_precond if amg else args.precond)
    if amg:
        keys.update(amgCoarsest=40, amgTailRows=200)
    fs = lid_driven_cavity(gl, 1.0, 0.1, solver=keys, pSolver=dict(preconditioner="amg") if amg else None)
    ofs = O.cavity(O.Mesh.rectilinear(nx, ny, 1.0, 1.0), 1.0, 0.1)
    ofs.use_direct_solver()
    dt = 0.5 / nx
    for _ in range(steps):
        st = fs.solve(dt)
        ofs.step(dt)
    owner, gid = gl.i32("owner"), gl.i32("globalId")
    mine = owner == job.rank
    u, p, po = fs.u.get("cells"), fs.p.get("cells"), ofs.view("p")
    ou = (ofs.view("ux"), ofs.view("uy"))
    num = sum(float(np.sum((u[k][mine] - ou[k][gid[mine]]) ** 2)) for k in (0, 1))
    psum, cnt = job.sum_over_ranks([float(p[mine].sum()), float(mine.sum())])
    pm = psum / cnt
    nump = float(np.sum(((p[mine] - pm) - (po[gid[mine]] - po.mean())) ** 2))
    num, nump = job.sum_over_ranks([num, nump])
    eu = (num / float(np.sum(ou[0] ** 2) + np.sum(ou[1] ** 2))) ** 0.5
    ep = (nump / float(np.sum((po - po.mean()) ** 2))) ** 0.5
    rec = {"what": "%dx%d cavity, %d steps, %d rank(s)%s, fields against the oracle's single-domain direct solve"
                   % (nx, ny, steps, job.world, ", RCB partition" if job.world > 1 else ""),
           "relL2_u": eu, "relL2_p": ep, "tolerance": 1e-6, "ok": bool(eu < 1e-6 and ep < 1e-6),
           "preconditioner": "amg (3+ levels, distributed)" if amg else args.precond,
           "itersP_last": st["itersP"], "max_divergence": st["maxDivergence"]}
    fs.close(); gl.close(); comm.close()
    return rec


def refresh_record(args):
    import numpy as np
    """Variable-density pressure operator (config 4's pEqn_) at the headline size with a bubble that moves every step:
    the multigrid values are recomputed on the device (amg_refresh.cuh), against a host setup from scratch per matrix."""
    from phase_b200.api import Communicator, SparseMatrixSolver
    from phase_b200.synthetic import beta_field, variable_laplacian
    nx = args.n
    n = nx * nx
    comm = Communicator(0)
    keys = dict(solver="BICGSTAB", maxIters=2000, tolerance=args.tol, preconditioner="amg", nullSpace="constant")
    s = SparseMatrixSolver(comm).setup(dict(keys, amgRefresh="always"))
    rng = np.random.default_rng(0)
    it_r, it_f, ms_r, ms_setup = [], [], [], []
    for step in range(4):
        A = variable_laplacian(nx, nx, beta_field(nx, nx, 0.3 + 0.01 * step, 1000.0))
        b = rng.standard_normal(n); b -= b.mean()
        s.setRank(n); s.set(A.indptr, A.indices, A.data); s.setRhs(b)
        s.solve()
        f = SparseMatrixSolver(comm).setup(dict(keys, amgRefresh="off"))
        f.setRank(n); f.set(A.indptr, A.indices, A.data); f.setRhs(b)
        f.solve()
        if step:
            it_r.append(s.nIters()); it_f.append(f.nIters()); ms_r.append(s.amgRefreshInfo()["refreshMs"])
        ms_setup.append(f.amgInfo()["setupMs"])
        f.close()
    rec = {"what": "pEqn_ = -div((1/rho) grad p) on %dx%d cells, density ratio 1000, bubble displaced 0.01 widths per step: values of "
                   "every multigrid level recomputed on the device (aggregates and patterns kept) vs a host setup from scratch" % (nx, nx),
           "refresh_ms": float(np.mean(ms_r)), "host_setup_ms": float(np.mean(ms_setup)), "host_setups_refreshed_solver": s.amgInfo()["setups"],
           "iters_refreshed": it_r, "iters_fresh_setup": it_f, "symbolic_bytes": s.amgRefreshInfo()["bytes"]}
    s.close(); comm.close()
    return rec


def weak_record(job, args, peak):
    """The series that ends in the north star's 64M-cell / 8-GPU point: 4000x2000 = 8M cells PER GPU (h = 1/4000
    everywhere), cavity time steps and the zero-guess pressure solve.  Weak-scaling efficiency follows from the
    N = 1 line's record of the same name."""
    px, py = block_layout(job.world)
    bx, by = args.weak_block
    nx, ny = bx * px, by * py
    comm = job.communicator()
    t0 = time.perf_counter()
    grid, fs = make_cavity(job, comm, args, nx, ny, float(px), float(py) * by / bx, px, py,
                           args.u_precond if args.precond == "amg" else args.precond, args.precond)
    dt = 0.5 / bx
    t_setup = time.perf_counter() - t0
    ms, stats, launches = timed_steps(job, comm, fs, dt, 2, args.weak_steps)
    rec = {"what": "lid-driven cavity on %dx%d = %d cells, %dx%d blocks of %dx%d cells (one per GPU), dt = 0.5 h, same "
                   "solver keys as the headline run" % (nx, ny, nx * ny, px, py, bx, by),
           "global_cells": nx * ny, "cells_per_gpu": bx * by, "n_gpus": job.world,
           "time_steps_per_s": 1e3 / ms, "ms_per_step": ms, "steps": args.weak_steps, "warmup": 2,
           "cell_updates_per_s": nx * ny * 1e3 / ms,
           "iters_per_solve": {"uEqn": mean(s["itersU"] for s in stats), "pEqn": mean(s["itersP"] for s in stats)},
           "pressure_solve": pressure_solve_record(job, comm, fs, peak),
           "host_setup_s": t_setup}
    if args.precond == "amg":
        rec["amg"] = fs.pEqn.solver.amgInfo()
    fs.close(); grid.close(); comm.close()
    return rec


def e2e_record(job, comm, fs, dt, steps, sizes, weak):
    """The same step through the public API with HOST state in pinned memory.  The state is what the reference itself
    persists between runs: the CELL values of u and p (Solver::readLatestCgnsFlowSolution, US/Solver.cpp:544-581, reads
    cell fields and re-derives the faces).  Per step: H2D of those cells, FractionalStep.rebuildFaces (ghosts, boundary
    faces, grad p and the face velocities exactly as the previous step left them -- tests/test_gpu_fracstep.py::
    test_state_from_cells_alone), the step, D2H of the same arrays -- a serial chain, since a host-owned state feeds
    every step with the previous step's output."""
    torch = job.torch
    N = sizes["nCells"]
    host = {k: torch.empty(n, dtype=torch.float64).pin_memory() for k, n in (("uc", 2 * N), ("pc", N))}
    parts = {"uc": (fs.u, "cells"), "pc": (fs.p, "cells")}
    for k, (fld, part) in parts.items():
        host[k].numpy()[:] = fld.get(part).reshape(-1)
    nbytes = sum(v.numel() for v in host.values()) * 8

    def one_step():
        for k, (fld, part) in parts.items():
            fld.set(part, host[k].numpy())
        fs.rebuildFaces(dt)
        fs.solve(dt)
        for k, (fld, part) in parts.items():
            fld.get(part, out=host[k].numpy())

    for _ in range(2):                  # the first transfers of a pinned buffer run at a third of the link rate
        one_step()
    job.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    job.barrier()
    s = job.max_over_ranks((time.perf_counter() - t0) / steps)
    return {"value": (job.world if weak else 1) / s, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
            "steps": steps, "warmup": 2, "ms_per_step": 1e3 * s,
            "what": "pinned host state (cells of u and p: what the reference's restart persists) copied in, faces and grad p "
                    "rebuilt on the device (phb_fs_rebuild_faces), FractionalStep.solve, the same arrays copied out, every step; "
                    "wall clock, max over ranks"}


def seam1_record(comm, fs, dt, args):
    """Seam 1 alone (single GPU): the reference-facing backend call with HOST CSR arrays, exactly what
    FiniteVolumeEquation<T>::solve hands to a SparseMatrixSolver: set(rowPtr,colInd,vals) + setRhs(-rhs_) + solve + x."""
    from phase_b200.api import SparseMatrixSolver
    rp, ci, va, rhs = fs.assembleP(dt).export(1)      # reference layout: ELL-5 padded, nb before diagonal
    b = -rhs
    s1 = SparseMatrixSolver(comm).setup(solver_keys(args, args.precond))
    s1.setup(dict(nullSpace="constant"))
    s1.setRank(len(b)); s1.set(rp, ci, va); s1.setRhs(b); s1.solve()   # warm-up: pattern analysis, graph capture
    ts = []
    xs = s1.x()
    for _ in range(3):
        t0 = time.perf_counter()
        s1.setRank(len(b)); s1.set(rp, ci, va); s1.setRhs(b); s1.solve(); s1.x(out=xs)
        ts.append(time.perf_counter() - t0)
    rec = {"what": "pEqn_ (%d rows, %d padded entries) through set(rowPtr,colInd,vals)+setRhs+solve+x with host arrays "
                   "(pageable, as a std::vector is); best of 3" % (len(b), len(ci)),
           "seconds_per_solve": min(ts), "iterations": s1.nIters(), "relres": s1.error(),
           "h2d_bytes": int(rp.nbytes + ci.nbytes + va.nbytes + b.nbytes), "d2h_bytes": int(xs.nbytes)}
    s1.close()
    return rec


def committed_traffic(name, rows):
    """DRAM bytes per launch from the committed `ncu --set full` capture -- only when it was taken on this size."""
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            d = json.load(f)
        return d.get("dram_bytes_per_launch") if int(d.get("rows", -1)) == int(rows) else None
    except Exception:
        return None


def roofline_records(fs, peak, peak_src, ms_per_step, iters_u, iters_p, world, amg, nrows):
    """Live timing of the dominant kernels on the resident pEqn_ system (CUDA events on the library's stream)."""
    spmv_ms = fs.pEqn.solver.time_spmv(50)
    b_spmv, _ = fs.pEqn.solver.bytes()
    ach = b_spmv / (spmv_ms * 1e-3) / 1e9
    spmv = {"bound": "hbm", "kernel": "k_spmv<1,0> (fp64 sliced-ELL SpMV of the BiCGStab loop, pEqn_, %d rows per GPU)" % nrows,
            "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "frac_of_8TBs_nominal": ach / 8000.0,
            "traffic": committed_traffic("spmv_traffic.json", nrows) if world == 1 else None,
            "algorithmic_bytes_per_launch": b_spmv, "ms_per_launch": spmv_ms, "peak_source": peak_src,
            "bytes_formula": "12 nnz + 4 (n+1) + 16 n  (SURVEY 8d)",
            "share_of_step": 2.0 * iters_p * spmv_ms / ms_per_step}
    if not amg or world > 1:
        return spmv, None
    try:
        ta = fs.pEqn.solver.timeAmg(20)
    except Exception as exc:           # keep the line (SpMV roofline) rather than lose the run
        print("bench: timeAmg failed: %s" % exc, file=sys.stderr)
        return spmv, None
    aj = ta["bytesJacobi"] / (ta["msJacobi"] * 1e-3) / 1e9
    cyc = ta["bytesCycle"] / (ta["msCycle"] * 1e-3) / 1e9
    roof = {"bound": "hbm",
            "kernel": "k_amg_spmv<MODE 2> (damped-Jacobi sweep on multigrid level 0 of pEqn_, %d rows, single-precision "
                      "matrix, fp64 Krylov vectors in and out): the largest launch of the V-cycle" % nrows,
            "achieved": aj, "peak": peak, "unit": "GB/s", "frac": aj / peak, "frac_of_8TBs_nominal": aj / 8000.0,
            "traffic": committed_traffic("amg_traffic.json", nrows),
            "algorithmic_bytes_per_launch": ta["bytesJacobi"], "ms_per_launch": ta["msJacobi"], "peak_source": peak_src,
            "bytes_formula": "8 nnz + 4 (n/32+1) [slice offsets] + 4 n [w] + 4 n [x, once] + 8 n [b] + 8 n [y]",
            "share_of_step": 2.0 * iters_p * ta["msJacobi"] / ms_per_step,
            "level0_ms": {"residual": ta["msResidual"], "restriction": ta["msRestriction"],
                          "prolongation": ta["msProlongation"], "jacobi": ta["msJacobi"]},
            "cycle": {"ms": ta["msCycle"], "launches": ta["launchesPerCycle"], "algorithmic_bytes": ta["bytesCycle"],
                      "achieved_GBps": cyc, "frac_of_measured_peak": cyc / peak,
                      "share_of_step_pEqn": 2.0 * iters_p * ta["msCycle"] / ms_per_step}}
    return spmv, roof


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--side", dest="n", type=int, default=2000, help="cells per side (2000 -> 4M cells)")
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--max-iters", type=int, default=20000)
    ap.add_argument("--precond", default="amg", choices=["ilu0", "jacobi", "none", "amg"],
                    help="amg = smoothed-aggregation V-cycle on pEqn_; uEqn_ per --u-precond")
    ap.add_argument("--u-precond", default="amg", choices=["ilu0", "amg"], help="uEqn_ preconditioner when --precond amg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-sub", action="store_true", help="skip the ilu0 / amg_double sub-records (N = 1)")
    ap.add_argument("--no-weak", action="store_true", help="skip the 8M-cells-per-GPU record")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-line parity check")
    ap.add_argument("--weak-block", type=int, nargs=2, default=[4000, 2000], metavar=("BX", "BY"))
    ap.add_argument("--weak-steps", type=int, default=5)
    ap.add_argument("--ref-budget", type=float, default=150.0, help="reference arm: wall budget of the timed steps, s")
    ap.add_argument("--solver-key", action="append", default=[], metavar="KEY=VALUE",
                    help="extra LinearAlgebra key for both equations (e.g. amgPrecision=double, amgSweeps=2)")
    ap.add_argument("--mesh", default="quad", choices=["quad", "tri"],
                    help="quad: side x side quads; tri: the same cell count as triangles (each quad of a "
                         "side/sqrt(2) lattice split along alternating diagonals, unstructured connectivity)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = the same side x side problem split over the GPUs; weak = side x side cells per GPU")
    ap.add_argument("--guess-order", type=int, default=1, help="pEqn_ initial guess: 0 previous p, 1 extrapolated")
    ap.add_argument("--comm", default="peer", choices=["peer", "peer-fused", "nccl"],
                    help="in-loop halo/all-reduce: NVLink peer-memory kernels (default) or NCCL calls")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import math
    import numpy as np
    job = Job()
    rank, world = job.rank, job.world
    comm = job.communicator()
    px, py = block_layout(world)
    weak = args.scaling == "weak"
    amg = args.precond == "amg"
    u_pc = args.u_precond if amg else args.precond
    if args.mesh == "tri":
        nx = ny = int(round(args.n / math.sqrt(2.0)))        # 2 tn^2 triangles ~ side^2 cells
        width = height = 1.0
        dt = 0.25 / nx                                       # triangles: half the lattice spacing
    elif weak:
        nx, ny, width, height = args.n * px, args.n * py, float(px), float(py)
        dt = 0.5 / args.n                                    # same cell size h = 1/side in every block
    else:
        nx, ny, width, height = args.n, args.n, 1.0, 1.0
        dt = 0.5 / nx                                        # maxCo 0.5 with the unit lid speed, h = 1/nx
    grid, fs = make_cavity(job, comm, args, nx, ny, width, height, px, py, u_pc, args.precond, mesh=args.mesh)
    sizes = grid.sizes()
    sampler = ClockSampler(job.local_rank)
    if rank == 0:
        sampler.start()                    # before the warm-up, so that it is sampling when the timed region starts
    ms_per_step, timed, launches = timed_steps(job, comm, fs, dt, args.warmup, args.steps, sampler if rank == 0 else None)
    clocks = sampler.stop() if rank == 0 else None
    steps_per_s = 1e3 / ms_per_step
    # strong scaling: the job advances ONE problem, value = its time steps per second.  weak scaling: every rank
    # advances a side x side block, the job processes `world` such block-steps per step
    value = steps_per_s * (world if weak else 1)
    iters_u, iters_p = mean(s["itersU"] for s in timed), mean(s["itersP"] for s in timed)
    peak, peak_src = measured_peak()
    spmv_roof, amg_roof = roofline_records(fs, peak, peak_src, ms_per_step, iters_u, iters_p, world, amg, sizes["nLocal"])
    pressure = pressure_solve_record(job, comm, fs, peak)
    amg_info = None
    if amg:
        amg_info = fs.pEqn.solver.amgInfo()
        amg_info["uEqn"] = fs.uEqn.solver.amgInfo() if u_pc == "amg" else u_pc
        amg_info["note"] = ("hierarchy built in the first (warm-up) solve and reused: pEqn_ = laplacian(dt, p) is constant up to "
                            "the scalar dt; setupMs is that one-off cost, outside the timed steps")
    try:        # the device-resident `value` above must survive a failure of the legs that follow
        e2e = e2e_record(job, comm, fs, dt, max(1, min(args.steps, 5)), sizes, weak)
    except Exception as exc:
        print("bench: e2e leg failed: %s" % exc, file=sys.stderr)
        e2e = {"error": str(exc)}
    try:
        seam1 = seam1_record(comm, fs, dt, args) if world == 1 else None
    except Exception as exc:
        print("bench: Seam 1 leg failed: %s" % exc, file=sys.stderr)
        seam1 = {"error": str(exc)}
    state = None
    if world == 1 and rank == 0 and not args.no_cpu and args.mesh == "quad":
        u, uf, g = fs.u.get("cells"), fs.u.get("faces"), fs.gradP.get("cells")
        state = {"ux": u[0].copy(), "uy": u[1].copy(), "ufx": uf[0].copy(), "ufy": uf[1].copy(), "p": fs.p.get("cells").copy(),
                 "pf": fs.p.get("faces").copy(), "gpx": g[0].copy(), "gpy": g[1].copy()}
    nlocal = sizes["nLocal"]
    fs.close(); grid.close(); comm.close()

    extra = {}
    if world == 1 and not args.no_sub and amg:
        extra["ilu0"] = sub_record(job, args, "ilu0", "ilu0", None, 2, 3,
                                   "like for like with the CPU arm: multicolour ILU(0) (the preconditioner family north_star names) "
                                   "on both equations, everything else as the headline run")
        extra["amg_double"] = sub_record(job, args, "amg", "amg", dict(amgPrecision="double"), 3, 10,
                                         "the headline algorithm with the V-cycle in fp64 (no single-precision arithmetic anywhere)")
    if world == 1 and not args.no_sub and amg and args.mesh == "quad":
        try:
            extra["amg_refresh"] = refresh_record(args)
        except Exception as exc:
            extra["amg_refresh"] = {"error": str(exc)}
    if not args.no_parity:
        extra["parity"] = parity_record(job, args)
    if not args.no_weak and args.mesh == "quad":
        try:
            extra["weak_8M_per_gpu"] = weak_record(job, args, peak)
        except Exception as exc:
            extra["weak_8M_per_gpu"] = {"error": str(exc)}

    if rank != 0:
        if job.dist:
            job.dist.destroy_process_group()
        return
    single = not any(kv.startswith("amgPrecision=double") for kv in args.solver_key)
    algorithm = {"preconditioner": ("smoothed-aggregation AMG V(1,1), damped Jacobi, on pEqn_; %s on uEqn_" % u_pc) if amg else args.precond,
                 "precision": ("fp64 assembly, Krylov recurrences, residual test and true-residual recheck; %s multigrid cycle"
                               % ("fp32" if single else "fp64")) if amg else "fp64 throughout",
                 "n_gpus": world, "cells_per_gpu": nlocal,
                 "partition": "none" if world == 1 else ("RCB, %d parts" % world if args.mesh == "tri" else
                                                         "%dx%d blocks of %dx%d cells" % (px, py, nx // px, ny // py)),
                 "comm": "single GPU" if world == 1 else (
                     "NVLink peer-memory kernels inside the Krylov loop (halo + all-reduce one-CTA kernels in the CUDA graph); "
                     "NCCL outside the loop" if args.comm == "peer" else
                     "NVLink peer memory, halo push/wait and the sigma all-reduce inside the compute kernels"
                     if args.comm == "peer-fused" else "NCCL send/recv + all-reduce"),
                 "departure_from_north_star": "north_star names level-scheduled ILU(0)/Jacobi; AMG is this backend's MueLu-role "
                                              "preconditioner (M/TrilinosMueluSparseMatrixSolver.cpp:27-32); the ILU(0) run of the "
                                              "same workload is the `ilu0` record" if amg else None}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None,
            "dtype": ("f64 (assembly, BiCGStab, residuals) + f32 (multigrid preconditioner cycle)" if amg and single else "f64"),
            "data": "synthetic", "config": workload_config(args), "algorithm": algorithm,
            "value_definition": "block time-steps per second summed over GPUs = n_gpus x (time steps of the global problem per second)"
            if weak else "time steps per second of the %dx%d problem (split over the GPUs)" % (nx, ny),
            "global_time_steps_per_s": steps_per_s,
            "cell_updates_per_s": steps_per_s * nlocal * world,
            "ms_per_bicgstab_iteration": ms_per_step / max(1.0, iters_p + iters_u),
            "iters_per_solve": {"uEqn": iters_u, "pEqn": iters_p,
                                "relres_p": timed[-1]["errorP"], "relres_u": timed[-1]["errorU"]},
            "max_divergence": timed[-1]["maxDivergence"], "max_courant": timed[-1]["maxCourant"],
            "roofline": amg_roof or spmv_roof, "pressure_solve": pressure,
            "e2e": e2e, "e2e_seam1": seam1, "gpu_launches": int(launches), "clocks": clocks}
    if amg_roof:
        line["roofline_spmv"] = spmv_roof
    if amg_info:
        line["amg"] = amg_info
    line.update(extra)
    if state is not None:
        # one fully converged CPU step of the reference algorithm from the state the B200 arm ended in
        v, cores, detail = cpu_converged_steps(args.n, args.n, dt, args.tol, args.max_iters, 0, 1, 1e9, state=state)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "1 fully converged time step (BiCGStab + ILU(0), tolerance %g) continuing from the state the "
                                          "B200 arm ended in (warm-up + timed + end-to-end steps); wall clock, nothing extrapolated"
                                          % args.tol, "detail": detail}
    print(json.dumps(line), flush=True)
    if job.dist:
        job.dist.destroy_process_group()


if __name__ == "__main__":
    main()
