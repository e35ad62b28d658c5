#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark on B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA arm
    python bench.py --impl reference --gpus N --steps K ...  # CPU reference arm (oracle port)

Workload (BASELINE.json configs[1]): lid-driven cavity, rho=1, mu=0.1, on a
synthetic 2000x2000 quad mesh (4M cells) per GPU; one "step" = one whole time step
of the snapshot's solver module (FractionalStep::solve: assemble uEqn_, BiCGStab,
interpolate, assemble pEqn_, BiCGStab, gradient, correct).  The north star calls
the time step "PISO"; the mounted snapshot ships only the fractional-step
successor, which is what is timed and parity-checked.

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


def cpu_reference_sample(nx, ny, dt, iters_u, iters_p, cap=12, precond="ilu0"):
    """Time the oracle (CPU port of the reference path) on a bounded sample of the
    SAME 4M-cell step: full assembly + field glue once, and `cap` BiCGStab+Jacobi
    iterations of each solve on all host cores; the step time is then scaled to
    the iteration counts the tolerance needs."""
    import ctypes as C
    import numpy as np
    import oracle as O
    t0 = time.perf_counter()
    om = O.Mesh.rectilinear(nx, ny, 1.0, 1.0)
    ofs = O.cavity(om, 1.0, 0.1)
    t_mesh = time.perf_counter() - t0
    # a non-trivial state: one cheap capped step so u, p, gradP are not all zero
    ofs.set_solver_params(tol=1e-30, max_iters=2, precond=1)
    ofs.step(dt)
    t0 = time.perf_counter()
    ue = ofs.assemble_u(dt)
    t_asm_u = time.perf_counter() - t0
    t0 = time.perf_counter()
    pe = ofs.assemble_p(dt)
    t_asm_p = time.perf_counter() - t0
    per_iter, t_setup = [], []
    for e in (ue, pe):
        rp, ci, va, rhs = e.export()
        if precond == "ilu0":
            # the CUDA path's algorithm: ILU(0) in the multicolour ordering, sweeps parallel inside a colour
            rp2, ci2, va2, new2old, bp = O.multicolor_permute(rp, ci, va)
            b2 = np.ascontiguousarray((-rhs)[new2old])
            tt = []
            for k in (cap, 2 * cap):
                x2 = np.zeros_like(b2)
                rr = C.c_double()
                t0 = time.perf_counter()
                O.lib().or_bicgstab_blocks(len(b2), O._ip(rp2), O._ip(ci2), O._dp(va2), O._dp(b2), O._dp(x2), 1e-30, k,
                                           len(bp) - 1, O._ip(bp), C.byref(rr))
                tt.append(time.perf_counter() - t0)
            per_iter.append(max((tt[1] - tt[0]) / cap, 0.25 * tt[1] / (2 * cap)))
            t_setup.append(max(0.0, tt[0] - cap * per_iter[-1]))      # factorisation + initial residual
        else:
            t0 = time.perf_counter()
            O.bicgstab(rp, ci, va, -rhs, tol=1e-30, max_iters=cap, precond=1 if precond == "jacobi" else 0)
            per_iter.append((time.perf_counter() - t0) / cap)
            t_setup.append(0.0)
    # field glue (interpolate, gradient, correct) is part of step(); measure via a capped step
    ofs.set_solver_params(tol=1e-30, max_iters=1, precond=1)
    t0 = time.perf_counter()
    ofs.step(dt)
    t_step1 = time.perf_counter() - t0
    t_glue = max(0.0, t_step1 - t_asm_u - t_asm_p - per_iter[0] - per_iter[1])
    t_full = t_asm_u + t_asm_p + t_glue + sum(t_setup) + iters_u * per_iter[0] + iters_p * per_iter[1]
    detail = dict(t_mesh_s=t_mesh, t_precond_setup_s=sum(t_setup), preconditioner=precond, t_assemble_u_s=t_asm_u, t_assemble_p_s=t_asm_p, t_glue_s=t_glue,
                  s_per_iter_u=per_iter[0], s_per_iter_p=per_iter[1], iters_u=iters_u, iters_p=iters_p)
    return 1.0 / t_full, detail


def typical_iters(precond="ilu0"):
    """Iteration counts per solve at tolerance 1e-8 on the 4M-cell cavity, measured
    on the GPU arm (same algorithm: right-preconditioned BiCGStab + Jacobi) and
    committed under profiles/ so the CPU arm can scale its bounded sample."""
    try:
        with open(os.path.join(ROOT, "profiles", "iters_4M.json")) as f:
            d = json.load(f)
        d = d.get(precond, d)
        return float(d["iters_u"]), float(d["iters_p"]), "profiles/iters_4M.json"
    except Exception:
        return 60.0, 3000.0, "default estimate (profiles/iters_4M.json missing)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    nx = ny = args.n
    dt = 0.5 / nx
    config = workload_config(args, 1)     # the same workload description as the B200 arm prints
    if args.precond == "amg":
        args.precond = "ilu0"        # the reference arm always runs the reference's algorithm
    iu, ip, src = typical_iters(args.precond)
    vals = []
    detail = None
    for _ in range(max(1, min(args.steps, 2))):
        v, detail = cpu_reference_sample(nx, ny, dt, iu, ip, precond=args.precond)
        vals.append(v)
    v = max(vals)
    cores = O.lib().or_num_threads()
    sample = ("1 assembled 4M-cell step + 12/24 BiCGStab(%s) iterations per solve on %d OpenMP threads, scaled to "
              "%.0f (uEqn) / %.0f (pEqn) iterations per solve (%s)" % (args.precond, cores, iu, ip, src))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "reference_algorithm": "the reference's path: host CrsEquation-style assembly + right-preconditioned BiCGStab with "
                                   "ILU(0) (Belos/Ifpack2 RILUK(0) role) on all host cores, same mesh, time step and tolerance",
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "detail": detail},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def block_layout(nprocs):
    """px x py blocks, as square as possible (1 -> 1x1, 2 -> 1x2, 4 -> 2x2, 8 -> 2x4)."""
    px = 1
    while px * px * 2 <= nprocs and nprocs % (px * 2) == 0:
        px *= 2
    return px, nprocs // px


def workload_config(args, nprocs):
    px, py = block_layout(nprocs)
    strong = getattr(args, "scaling", "strong") == "strong"
    gx, gy = (args.n, args.n) if strong else (args.n * px, args.n * py)
    if getattr(args, "mesh", "quad") == "tri":
        import math
        tn = int(round(args.n / math.sqrt(2.0)))
        return {"workload": "lid-driven cavity (rho=1, mu=0.1, lid u=1), unstructured triangles: %dx%d lattice split along "
                            "alternating diagonals = %d cells, FractionalStep time step, dt = 0.5 h, BiCGStab + %s, tolerance %g, "
                            "warm start" % (tn, tn, 2 * tn * tn, args.precond, args.tol),
                "cells_per_gpu": 2 * tn * tn // nprocs, "global_cells": 2 * tn * tn,
                "partition": "none" if nprocs == 1 else "RCB, %d parts" % nprocs,
                "l2": "inputs larger than L2; no flush needed", "tolerance": args.tol, "max_iters": args.max_iters,
                "preconditioner": args.precond, "comm": "single GPU" if nprocs == 1 else args.comm}
    return {"workload": "lid-driven cavity (rho=1, mu=0.1, lid u=1), %dx%d quads = %d cells%s, "
                        "FractionalStep time step (the snapshot's PISO successor), dt = 0.5 h (maxCo 0.5), "
                        "BiCGStab + %s, tolerance %g on ||r||/||b||, warm start from the previous step"
                        % (gx, gy, gx * gy, "" if strong or nprocs == 1 else " (%dx%d per GPU)" % (args.n, args.n),
                           "smoothed-aggregation AMG V(1,1) on pEqn_ / %s on uEqn_" % getattr(args, "u_precond", "ilu0") if args.precond == "amg"
                           else args.precond, args.tol),
            "cells_per_gpu": gx * gy // nprocs, "global_cells": gx * gy,
            "partition": "none" if nprocs == 1 else "%dx%d blocks of %dx%d cells, one per GPU (global grid %dx%d)" % (
                px, py, gx // px, gy // py, gx, gy),
            "l2": "inputs larger than L2 (matrix + vectors ~0.6 GB per solve vs 126 MB L2); no flush needed",
            "tolerance": args.tol, "max_iters": args.max_iters, "preconditioner": args.precond,
            "comm": "single GPU" if nprocs == 1 else (
                "NVLink peer-memory kernels inside the Krylov loop (2 halo + 2 all-reduce one-CTA kernels per iteration, "
                "in the CUDA graph); NCCL outside the loop" if args.comm == "peer" else
                "NVLink peer memory, halo push/wait and the sigma all-reduce inside the compute kernels" if args.comm == "peer-fused" else
                "NCCL send/recv + all-reduce")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--side", dest="n", type=int, default=2000, help="cells per side per GPU (2000 -> 4M cells)")
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--max-iters", type=int, default=20000)
    ap.add_argument("--precond", default="amg", choices=["ilu0", "jacobi", "none", "amg"],
                    help="amg = smoothed-aggregation V-cycle on pEqn_; uEqn_ per --u-precond")
    ap.add_argument("--u-precond", default="amg", choices=["ilu0", "amg"], help="uEqn_ preconditioner when --precond amg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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

    import numpy as np
    import torch
    from phase_b200.api import Communicator, FiniteVolumeGrid2D, lid_driven_cavity

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    uid = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        box = [Communicator.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    comm = Communicator(local_rank, rank, world, uid)
    px, py = block_layout(world)
    if args.scaling == "weak":
        nx, ny, width, height = args.n * px, args.n * py, float(px), float(py)
    else:
        nx, ny, width, height = args.n, args.n, 1.0, 1.0
    if args.mesh == "tri":
        import math
        tn = int(round(args.n / math.sqrt(2.0)))            # 2 tn^2 triangles ~ side^2 cells
        if world == 1:
            grid = FiniteVolumeGrid2D.triangulated(comm, tn, tn, 1.0, 1.0)
        else:                                                # generic path: global mesh on the host, RCB, local mesh
            hostc = Communicator(Communicator.HOST_ONLY)
            gg = FiniteVolumeGrid2D.triangulated(hostc, tn, tn, 1.0, 1.0)
            grid = gg.local(gg.partition_rcb(world), comm)
            gg.close()
        nx = ny = tn
    elif world == 1:
        grid = FiniteVolumeGrid2D.rectilinear(comm, nx, ny, 1.0, 1.0)
    else:
        grid = FiniteVolumeGrid2D.rectilinear_block(comm, nx, ny, width, height, px, py)
    if world > 1 and args.comm.startswith("peer"):
        def all_gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        comm.enable_peer_memory(grid, all_gather)
    amg = args.precond == "amg"
    cfg = dict(solver="BICGSTAB", maxIters=args.max_iters, tolerance=args.tol,
               preconditioner=args.u_precond if amg else args.precond, peerFusion=1 if args.comm == "peer-fused" else 0)
    for kv in args.solver_key:
        k, v = kv.split("=", 1)
        cfg[k] = v
    fs = lid_driven_cavity(grid, 1.0, 0.1, solver=cfg, pSolver=dict(preconditioner="amg") if amg else None)
    fs.setup(guessOrder=args.guess_order)
    dt = 0.5 / nx                                            # maxCo 0.5 with the unit lid speed, h = 1/nx
    if args.scaling == "weak" and args.mesh == "quad":
        dt = 0.5 / args.n                                     # same cell size h = 1/side in every block
    if args.mesh == "tri":
        dt = 0.25 / nx                                        # triangles: half the lattice spacing
    stream = torch.cuda.ExternalStream(comm.stream())
    sizes = grid.sizes()
    N, F = sizes["nCells"], sizes["nFaces"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                    # before the warm-up, so that it is sampling when the timed region starts
    stats = []
    for _ in range(args.warmup):
        stats.append(fs.solve(dt))
    barrier()
    sampler.mark()
    launches0 = comm.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
    timed = []
    for _ in range(args.steps):
        timed.append(fs.solve(dt))
    with torch.cuda.stream(stream):
        e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = comm.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    steps_per_s = 1e3 / ms_per_step
    # strong scaling: the job advances ONE side x side problem, value = its time steps per second.
    # weak scaling: every rank advances a side x side block, so the job processes `world` such
    # block-steps per step (aggregate; equals time-steps/s at N = 1)
    value = steps_per_s * (world if args.scaling == "weak" else 1)

    # ---- dominant kernel: the SpMV inside BiCGStab, timed alone on the resident pEqn matrix
    spmv_ms = fs.pEqn.solver.time_spmv(50)
    b_spmv, b_iter = fs.pEqn.solver.bytes()
    peak, peak_src = measured_peak()
    achieved = b_spmv / (spmv_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "spmv_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass

    # ---- with the multigrid preconditioner the V-cycle, not the Krylov SpMV, is where the step's time goes: time its
    # level-0 kernels and one whole cycle live on the resident pEqn_ hierarchy
    amg_roof, ta = None, None
    if amg and world == 1:
        try:
            ta = fs.pEqn.solver.timeAmg(20)
        except Exception as exc:           # keep the line (SpMV roofline) rather than lose the run
            print("bench: timeAmg failed: %s" % exc, file=sys.stderr)
    if ta:
        atraffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "amg_traffic.json")) as f:
                atraffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            pass
        aj = ta["bytesJacobi"] / (ta["msJacobi"] * 1e-3) / 1e9
        amg_roof = {"bound": "hbm",
                    "kernel": "k_amg_spmv<MODE 2> (damped-Jacobi sweep on multigrid level 0 of pEqn_, 4M rows, single-precision "
                              "matrix, fp64 Krylov vectors): the largest launch of the V-cycle that dominates the step",
                    "achieved": aj, "peak": peak, "unit": "GB/s", "frac": aj / peak, "frac_of_8TBs_nominal": aj / 8000.0,
                    "traffic": atraffic, "algorithmic_bytes_per_launch": ta["bytesJacobi"], "ms_per_launch": ta["msJacobi"],
                    "peak_source": peak_src,
                    "level0_ms": {"residual": ta["msResidual"], "restriction": ta["msRestriction"],
                                  "prolongation": ta["msProlongation"], "jacobi": ta["msJacobi"]},
                    "cycle": {"ms": ta["msCycle"], "launches": ta["launchesPerCycle"], "algorithmic_bytes": ta["bytesCycle"],
                              "achieved_GBps": ta["bytesCycle"] / (ta["msCycle"] * 1e-3) / 1e9,
                              "frac_of_measured_peak": ta["bytesCycle"] / (ta["msCycle"] * 1e-3) / 1e9 / peak}}

    # ---- the pressure solve alone (pEqn_ as assembled by the last step, zero initial guess): whole-solve
    # algorithmic throughput = iterations x bytes per iteration (2 SpMV + 2 preconditioner applies + vector passes)
    fs.pEqn.solve(warmStart=False)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        pe0.record()
    fs.pEqn.solve(warmStart=False)
    with torch.cuda.stream(stream):
        pe1.record()
    barrier()
    p_ms, p_its = pe0.elapsed_time(pe1), fs.pEqn.solver.nIters()
    pressure_solve = {"what": "pEqn_ of the last step solved from a zero guess, CUDA events around phb_eqn_solve",
                      "iterations": p_its, "ms": p_ms, "relres": fs.pEqn.solver.error(),
                      "algorithmic_bytes_per_iteration": b_iter,
                      "achieved_GBps_per_gpu": p_its * b_iter / (p_ms * 1e-3) / 1e9,
                      "frac_of_measured_peak": p_its * b_iter / (p_ms * 1e-3) / 1e9 / peak}

    # ---- e2e: the same step through the public API with HOST state in pinned memory:
    # H2D of the step's input state, the step, D2H of the resulting u and p
    host = {k: torch.empty(n, dtype=torch.float64).pin_memory() for k, n in
            (("uc", 2 * N), ("uf", 2 * F), ("pc", N), ("pf", F), ("gc", 2 * N))}
    for k, (fld, part) in {"uc": (fs.u, "cells"), "uf": (fs.u, "faces"), "pc": (fs.p, "cells"),
                           "pf": (fs.p, "faces"), "gc": (fs.gradP, "cells")}.items():
        host[k].numpy()[:] = fld.get(part).reshape(-1)
    h2d = sum(v.numel() for v in host.values()) * 8
    d2h = (2 * N + N) * 8
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 2))
    for _ in range(e2e_steps):
        fs.u.set("cells", host["uc"].numpy()); fs.u.set("faces", host["uf"].numpy())
        fs.p.set("cells", host["pc"].numpy()); fs.p.set("faces", host["pf"].numpy())
        fs.gradP.set("cells", host["gc"].numpy())
        st = fs.solve(dt)
        fs.u.get("cells", out=host["uc"].numpy()); fs.p.get("cells", out=host["pc"].numpy())
        fs.u.get("faces", out=host["uf"].numpy()); fs.p.get("faces", out=host["pf"].numpy())
        fs.gradP.get("cells", out=host["gc"].numpy())
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    d2h_all = (2 * N + N + 2 * F + F + 2 * N) * 8

    # ---- Seam 1 alone (single GPU): the reference-facing backend call with HOST CSR arrays, exactly what
    # FiniteVolumeEquation<T>::solve hands to a SparseMatrixSolver: set(rowPtr,colInd,vals) + setRhs(-rhs_) + solve + x
    seam1 = None
    if world == 1:
        from phase_b200.api import SparseMatrixSolver
        rp, ci, va, rhs = fs.assembleP(dt).export(1)      # reference layout: ELL-5 padded, nb before diagonal
        b = -rhs
        s1 = SparseMatrixSolver(comm).setup(cfg)
        s1.setup(dict(nullSpace="constant", preconditioner=args.precond))
        s1.setRank(len(b)); s1.set(rp, ci, va); s1.setRhs(b); s1.solve()   # warm-up: pattern analysis, graph capture
        t0 = time.perf_counter()
        s1.setRank(len(b)); s1.set(rp, ci, va); s1.setRhs(b); s1.solve(); xs = s1.x()
        t_s1 = time.perf_counter() - t0
        seam1 = {"what": "pEqn_ (4M rows, 20M padded entries) through set(rowPtr,colInd,vals)+setRhs+solve+x with host arrays",
                 "seconds_per_solve": t_s1, "iterations": s1.nIters(), "relres": s1.error(),
                 "h2d_bytes": int(rp.nbytes + ci.nbytes + va.nbytes + b.nbytes), "d2h_bytes": int(xs.nbytes)}
        s1.close()

    if rank != 0:
        fs.close(); grid.close(); comm.close()
        dist.destroy_process_group()
        return
    iters_u = float(np.mean([s["itersU"] for s in timed]))
    iters_p = float(np.mean([s["itersP"] for s in timed]))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, world),
            "value_definition": ("time steps per second of the %dx%d-cell problem (split over the GPUs)" % (nx, ny)) if args.scaling == "strong"
            else "4M-cell block time-steps per second summed over GPUs = n_gpus x (time steps of the global problem per second)",
            "global_time_steps_per_s": steps_per_s,
            "cell_updates_per_s": steps_per_s * sizes["nLocal"] * world,
            "ms_per_bicgstab_iteration": ms_per_step / max(1.0, float(np.mean([s["itersP"] + s["itersU"] for s in timed]))),
            "iters_per_solve": {"uEqn": iters_u, "pEqn": iters_p,
                                "relres_p": timed[-1]["errorP"], "relres_u": timed[-1]["errorU"]},
            "max_divergence": timed[-1]["maxDivergence"], "max_courant": timed[-1]["maxCourant"],
            "roofline_spmv" if amg_roof else "roofline":
                        {"bound": "hbm", "kernel": "k_spmv (fp64 sliced-ELL SpMV inside BiCGStab, pEqn_ 4M rows)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_8TBs_nominal": achieved / 8000.0, "traffic": traffic,
                         "algorithmic_bytes_per_launch": b_spmv, "ms_per_launch": spmv_ms, "peak_source": peak_src},
            "pressure_solve": pressure_solve,
            "bicgstab": {"bytes_per_iteration": b_iter,
                         "note": "whole-solve GB/s = iters * bytes_per_iteration / solve time; see profiles/"},
            "e2e": {"value": (world if args.scaling == "weak" else 1) / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_all,
                    "what": "host state (u, p, gradP cells+faces) copied in, FractionalStep.solve, state copied out, per step"},
            "e2e_seam1": seam1, "gpu_launches": int(launches), "clocks": clocks}
    if amg_roof:
        amg_roof["share_of_step"] = {"pEqn_cycles": 2.0 * iters_p * amg_roof["cycle"]["ms"] / ms_per_step,
                                     "note": "2 cycles per BiCGStab iteration; uEqn_ runs the two-component variant of the same kernels"}
        line["roofline"] = amg_roof
    if amg:
        line["amg"] = fs.pEqn.solver.amgInfo()
        line["amg"]["uEqn"] = fs.uEqn.solver.amgInfo() if args.u_precond == "amg" else "ilu0"
        line["amg"]["note"] = ("hierarchy built on the host in the first (warm-up) solve and reused: pEqn_ = laplacian(dt, p) "
                               "is constant up to the scalar dt; setupMs is that one-off cost, outside the timed steps")
    if not args.no_cpu and world == 1:
        # the CPU arm runs the reference's algorithm (BiCGStab + ILU(0)); with AMG on the GPU arm its bounded
        # sample is scaled by the ILU(0) iteration counts measured for this workload (profiles/iters_4M.json)
        cpc = "ilu0" if amg else args.precond
        ciu, cip, csrc = (typical_iters("ilu0") if amg else (iters_u, iters_p, "this run"))
        v, detail = cpu_reference_sample(args.n, args.n, 0.5 / args.n, ciu, cip, precond=cpc)
        import oracle as O
        cores = O.lib().or_num_threads()
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "1 assembled 4M-cell step + 12/24 BiCGStab(%s) iterations per solve on %d "
                                          "OpenMP threads, scaled to %.0f/%.0f iterations per solve (%s)" %
                                          (cpc, cores, ciu, cip, csrc), "detail": detail}
    print(json.dumps(line), flush=True)
    fs.close(); grid.close(); comm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
