#!/usr/bin/env python
"""bench.py -- BE step throughput of the diphasic cut-cell heat problem (BASELINE.json configs[1], Heat_2ph_2D).

    python bench.py --gpus N --steps K --warmup W            # this repo (libpenguin_b200.so through its C ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port), rank 0 only

Workload (benchmark/Heat_2ph_2D.jl:64-111): 2048 x 2048 cells per GPU on [0,8] x [0,8N], one circular
interface (r = 2) per GPU slab, phase 1 inside / phase 2 outside, D1 = D2 = 1, ScalarJump(1,1,0), FluxJump(1,1,0),
empty BorderConditions, u0 = [1,1,0,0], backward Euler with dt = 0.5 h^2, Krylov to ||r|| <= 1e-10 ||b||.
A "step" is one BE step: RHS assembly + the linear solve + state update.  metric = DOF*steps/s, DOF = active bulk
unknowns of both phases (the rows the reference keeps after remove_zero_rows_cols!).
One JSON line on stdout (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "BE step throughput, 2D diphasic cut-cell heat (Heat_2ph_2D)"
UNIT = "DOF*steps/s"
L2_MB = 126


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md, clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# the CPU arm: the reference's algorithm (sparse assembly + direct solve per step) restated in oracle/
# ---------------------------------------------------------------------------------------------------------------------
def cpu_sample(nx, steps, warmup=0):
    """DOF*steps/s of the oracle port on an nx^2 sample of the same workload (1 core: SciPy SuperLU is serial)."""
    from oracle import geom, penguin_oracle as po
    mesh = po.Mesh((nx, nx), (8.0, 8.0))
    ls = geom.LevelSet.ball((4.0, 4.0), 2.0)
    c1, c2 = geom.capacity(mesh, ls), geom.capacity(mesh, ls.flipped())
    f = lambda x, y, z, t: 0.0 * x
    p1, p2 = po.Phase(c1, po.DiffusionOps(c1), f, 1.0), po.Phase(c2, po.DiffusionOps(c2), f, 1.0)
    n = mesh.n
    u0 = np.concatenate([np.ones(2 * n), np.zeros(2 * n)])
    dt = 0.5 * (8.0 / nx) ** 2
    ic = po.InterfaceConditions(po.ScalarJump(1.0, 1.0, 0.0), po.FluxJump(1.0, 1.0, 0.0))
    bc = po.BorderConditions()
    s = po.DiffusionUnsteadyDiph(p1, p2, bc, ic, dt, u0, "BE")
    s.x = po.solve_system(s.A, s.b)                       # the constructor's step (solve_DiffusionUnsteadyDiph!, diffusion.jl:429)
    Ti = s.x
    s.A = po.A_diph_unstead_diff(p1.operator, p2.operator, c1, c2, p1.D, p2.D, ic, dt, "BE")
    dof = int(np.count_nonzero((c1.V != 0) | np.any(np.stack(c1.B) != 0, axis=0)) +
              np.count_nonzero((c2.V != 0) | np.any(np.stack(c2.B) != 0, axis=0)))
    t, t0 = 0.0, None
    for k in range(warmup + steps):
        if k == warmup:
            t0 = time.perf_counter()
        t += dt
        s.b = po.b_diph_unstead_diff(p1.operator, p2.operator, p1.source, p2.source, c1, c2, p1.D, p2.D, ic, Ti, dt, t, "BE")
        s.A, s.b = po.BC_border_diph(s.A, s.b, bc, c1, c2)
        s.x = po.solve_system(s.A, s.b)                   # remove_zero_rows_cols! + sparse LU, every step (solver.jl:158-188)
        Ti = s.x
    el = time.perf_counter() - t0
    return dof * steps / el, dof, el


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nx = 256 if args.steps <= 30 else 128
    v, dof, el = cpu_sample(nx, args.steps, args.warmup)
    sample = f"{nx}x{nx} sample of the workload, {args.steps} BE steps, sparse LU per step (SciPy SuperLU), {dof} DOF"
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": {"workload": "Heat_2ph_2D diphasic BE step (BASELINE.json configs[1])", "sample": sample},
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(out)


# ---------------------------------------------------------------------------------------------------------------------
# the GPU arm
# ---------------------------------------------------------------------------------------------------------------------
KCLASS = ["apply", "update", "pupdate", "band_apply", "band_prec", "exchange", "prologue", "epilogue"]
KNAMES = {"apply": "operator apply fused with the search-direction / solution update (kf3_apply MODE 5: TMA-staged tile + halo)",
          "update": "residual update + fused dots (kf2_update)",
          "pupdate": "pointwise p / x update of ghost-class tiles and interface unknowns (kf2_pupd)",
          "band_apply": "interface-band part of the operator (kf_apply_band)", "band_prec": "interface-band preconditioner (kf_band_poly)",
          "exchange": "halo exchange + scalar reductions between ranks", "prologue": "step prologue", "epilogue": "back-transform + state write-back"}


class Harness:
    """process group, context, timing helpers shared by the workloads"""

    def __init__(self, args):
        import torch
        import penguin_b200 as pb
        from penguin_b200 import _lib as L
        self.torch, self.pb, self.L = torch, pb, L
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries exactly one JSON line
        if self.world != args.gpus and self.world > 1:
            raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={self.world}")
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

            def bcast(ident):
                obj = [ident]
                dist.broadcast_object_list(obj, src=0)
                return obj[0]
            self.ctx = pb.init_distributed(self.rank, self.world, self.local, bcast)
        else:
            self.ctx = pb.init(self.local)
        self.lib = L.lib()
        self.ext = torch.cuda.ExternalStream(self.ctx.stream)     # the library's launching stream, for torch.cuda.Event timing

    def barrier(self):
        self.torch.cuda.synchronize()
        self.ctx.sync()
        if self.dist is not None:
            self.dist.barrier()

    def _red(self, v, op):
        if self.dist is None:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def allmax(self, v):
        return self._red(v, self.dist.ReduceOp.MAX) if self.dist is not None else v

    def allsum(self, v):
        return self._red(v, self.dist.ReduceOp.SUM) if self.dist is not None else v

    def timed_steps(self, step, n, st):
        """n steps bracketed by barriers + device events on the library's stream; max over ranks"""
        torch = self.torch
        iters, setup_ms, solve_ms = [], 0.0, 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.ext)
        for _ in range(n):
            step()
            iters.append(st.iters)
            setup_ms += st.setup_ms
            solve_ms += st.solve_ms
        e1.record(self.ext)
        self.barrier()
        return self.allmax(e0.elapsed_time(e1)), iters, setup_ms, solve_ms

    def profile_steps(self, step, n, st):
        """the same steps with every launch bracketed by CUDA events (no graph replay, one stream): device ms and launches per kernel class"""
        ms, cnt = [0.0] * 8, [0] * 8
        self.lib.pb200_set_profiling(self.ctx.h, 1)
        for _ in range(n):
            step()
            for q in range(8):
                ms[q] += st.kernel_ms[q]
                cnt[q] += st.kernel_launches[q]
        self.lib.pb200_set_profiling(self.ctx.h, 0)
        self.barrier()
        return ms, cnt


def kernel_table(H, ms, cnt, cells, cells_general, band_rows, ndim, peak, fused=True, poly=False):
    """per kernel class of the Krylov loop: launches, mean device time, algorithmic bytes and fraction of the measured HBM peak.
    Bytes per launch on THIS rank; cells = cells of the active tiles, cells_general = those with streamed coefficient arrays."""
    nblk = 1 + 2 * ndim
    # poly: every iteration launches the fused apply (6 passes) AND one polynomial step z = c_r r + c_A M^ r (read r, write z: 2 passes + the
    # coefficient arrays again): the class average is what the per-launch time is compared with
    abytes = {"apply": 8 * ((4 if poly else 6 if fused else 2) * cells + ndim * cells_general),      # fused: read z, p, x; write p, x, v (+ N coefficient arrays on general tiles)
              "update": 8 * 3 * cells,                                                 # read v, r; write r
              "band_apply": 8 * band_rows * (9 * nblk + 3 * nblk + 4),                 # 3 x 3 blocks + gathered unknowns + RMW of v
              "band_prec": 8 * band_rows * (9 * nblk + 3 * nblk + 6) // 3}             # band cells only (~1/3 of the rows), columns restricted to the band
    # band heads (one rank, small bands): the two band kernels run INSIDE the apply / the update launch -- their bytes belong to those classes
    # (the band_prec class then has no launches of its own; band_apply keeps the one band-polynomial launch that opens every solve)
    heads = cnt[KCLASS.index("band_prec")] == 0 and 0 < cnt[KCLASS.index("band_apply")] * 2 < cnt[KCLASS.index("apply")]
    per_iter = abytes["apply"] + abytes["update"] + abytes["band_apply"] + abytes["band_prec"]
    if heads:
        abytes["apply"] += abytes["band_apply"]
        abytes["update"] += abytes["band_prec"]
        abytes["band_apply"] = abytes["band_prec"]      # (what is left in that class is a band-polynomial launch)
        abytes["band_prec"] = 0
    abytes["per_iteration"] = per_iter
    tot = sum(ms)
    rows = []
    for q, name in enumerate(KCLASS):
        if cnt[q] == 0:
            continue
        us = 1e3 * ms[q] / cnt[q]
        kname = KNAMES[name]
        if heads and name == "apply":
            kname += " + apply head (band couplings, interface unknowns)"
        if heads and name == "update":
            kname = "residual update + fused dots + update head (band polynomial) (kf2_update_b)"
        if heads and name == "band_apply":
            kname = "band polynomial of the first residual (kf_band_poly, once per solve)"
        row = {"class": name, "kernel": kname, "launches_timed": int(cnt[q]), "avg_launch_us": us, "share_of_timed_kernels": ms[q] / tot if tot > 0 else 0.0}
        if name in abytes and us > 0:
            row["algorithmic_bytes_per_launch"] = int(abytes[name])
            row["achieved_gbs"] = abytes[name] / (us * 1e-6) / 1e9
            row["frac"] = row["achieved_gbs"] / peak
        rows.append(row)
    return rows, abytes


def run_heat3d(H, args, kind):
    """3-D workloads of BASELINE.json: kind 'diph' = configs[3] weak-scaled (1024 x 1024 x 128 N cells: N = 8 is the 1024^3 north-star problem; the
    sub-box is centred on the sphere, so N = 1 holds the equatorial slab -- the one with the most interface cells), kind 'mono' = configs[2]
    (Heat3D 512^3, exterior phase, BE then CN) STRONG-scaled over the ranks."""
    pb, L, lib, ctx = H.pb, H.L, H.lib, H.ctx
    N = H.world
    peak, peak_src = peaks()
    if kind == "diph":
        nz = args.nz3d * N
        hh = 4.0 / args.nx3d
        Lz = nz * hh
        mesh = pb.Mesh((args.nx3d, args.nx3d, nz), (4.0, 4.0, Lz), (0.0, 0.0, 2.0 - 0.5 * Lz))
        body = pb.Sphere((2.0, 2.0, 2.0), 1.0)
        t0 = time.perf_counter()
        c1, c2 = pb.Capacity(body, mesh, compute_centroids=False), pb.Capacity(-body, mesh, compute_centroids=False)
        ctx.sync()
        cap_s = time.perf_counter() - t0
        p1, p2 = pb.Phase(c1, pb.DiffusionOps(c1), 0.0, 1.0), pb.Phase(c2, pb.DiffusionOps(c2), 0.0, 1.0)
        n = c1.nloc
        dt = 0.5 * hh * hh
        ic = pb.InterfaceConditions(pb.ScalarJump(1.0, 2.0, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
        u0 = np.concatenate([np.ones(2 * n), np.zeros(2 * n)])
        s = pb.DiffusionUnsteadyDiph(p1, p2, pb.BorderConditions(), ic, dt, u0, "BE")
        del u0
        scheme, workload, scaling = 0, (f"3-D diphasic heat, sphere interface, ScalarJump(1,2,0), FluxJump(1,1,0), BE, dt = 0.5 h^2 (BASELINE.json configs[3]; "
                                        "Robin borders are no-ops in the reference, SURVEY 8d-4)"), "weak"
        grid = [args.nx3d, args.nx3d, nz]
    else:
        nx = args.nxmono
        hh = 4.0 / nx
        mesh = pb.Mesh((nx, nx, nx), (4.0, 4.0, 4.0))
        body = -pb.Sphere((2.01, 2.01, 2.01), 1.0)
        t0 = time.perf_counter()
        c1 = pb.Capacity(body, mesh, compute_centroids=False)
        ctx.sync()
        cap_s = time.perf_counter() - t0
        p1 = pb.Phase(c1, pb.DiffusionOps(c1), 0.0, 1.0)
        n = c1.nloc
        dt = 0.75 * hh * hh
        keys = ("left", "right", "top", "bottom", "forward", "backward")
        s = pb.DiffusionUnsteadyMono(p1, pb.BorderConditions({k: pb.Dirichlet(1.0) for k in keys}), pb.Dirichlet(1.0), dt, np.zeros(2 * n), "BE")
        scheme, workload, scaling = 1, "Heat3D monophasic, sphere embedded boundary (exterior phase), Dirichlet borders, BE then CN, dt = 0.75 h^2 (BASELINE.json configs[2])", "strong"
        grid = [nx, nx, nx]
    opts = L.KrylovOpts()
    opts.method, opts.rtol, opts.atol, opts.maxit, opts.warm_start, opts.check_every = 0, 1e-10, 0.0, 5000, args.warm, 8
    si = L.StepIn()
    si.dt = dt
    si.g_const[0] = si.g_const[1] = 0.0 if kind == "diph" else 1.0
    st = L.StepStats()

    def step(sch=scheme):
        si.scheme = sch
        L.check(lib.pb200_solver_step(s._h, C.byref(si), C.byref(opts), C.byref(st)), ctx.h)
    step(0)                                      # the constructor's BE step (builds the folded system, captures the graphs)
    for _ in range(max(3, args.warmup3d)):
        step()
    ms, iters, setup_ms, solve_ms = H.timed_steps(step, args.steps3d, st)
    dof = int(st.dof_bulk)
    kms, kn = H.profile_steps(step, min(args.steps3d, 3), st) if not args.no_profile else ([0.0] * 8, [0] * 8)
    cu, cg, rows_b = int(st.apply_cells_uniform), int(st.apply_cells_general), int(st.band_rows)
    table, abytes = kernel_table(H, kms, kn, cu + cg, cg, rows_b, 3, peak, poly=(kind == "mono"))
    # whole-step algorithmic bytes over ALL ranks: iterations x (fused apply + update + band kernels) + prologue (8 passes) + epilogue (6 passes)
    it_mean = float(np.mean(iters))
    per_iter = abytes["per_iteration"] + (abytes["apply"] if kind == "mono" else 0)   # (mono: fused apply + polynomial step)
    step_bytes = H.allsum(it_mean * per_iter + 8 * 14 * (cu + cg))
    agg = step_bytes / (ms / args.steps3d * 1e-3) / 1e9
    out = {"workload": workload, "grid": grid, "cells_per_gpu": [grid[0], grid[1], grid[2] // N if kind == "diph" else grid[2] / N], "n_gpus": N, "scaling": scaling,
           "dof": dof, "steps": args.steps3d, "ms_per_step": ms / args.steps3d, "value": dof * args.steps3d / (ms * 1e-3), "unit": UNIT,
           "iters_per_step": it_mean, "rtol": 1e-10, "krylov": "CG on the folded system" + (", polynomial (Chebyshev degree 1) preconditioner" if kind == "mono" else ", interface-band preconditioner"),
           "solve_ms_per_step": solve_ms / args.steps3d, "prologue_ms_per_step": setup_ms / args.steps3d, "capacity_build_s": cap_s,
           "final_rel_residual": st.rnorm / st.bnorm if st.bnorm > 0 else 0.0,
           "algorithmic_bytes_per_step_all_ranks": step_bytes, "aggregate_gbs": agg, "frac_of_measured_hbm": agg / (N * peak), "frac_of_nominal_8TBs": agg / (N * 8000.0),
           "peak_source": peak_src, "kernels_rank0": table,
           "band_rows_rank0": rows_b, "cells_constant_coef_tiles_rank0": cu, "cells_streamed_coef_tiles_rank0": cg}
    del s
    return out


POISSON_SEED = 20261018


def random_spheres(K=64, L=4.0, seed=POISSON_SEED, rmin=0.1, rmax=0.3, margin=0.5, gap=0.11):
    """BASELINE.json configs[4] geometry: K spheres, centres U[margin, L - margin]^3, radii U[rmin, rmax], rejection-sampled so that any two
    surfaces are at least `gap` apart (two cell diagonals at 128^3: no cell, and no staggered volume between two barycentres, meets two spheres)."""
    rng = np.random.default_rng(seed)
    cen, rad = [], []
    while len(cen) < K:
        c = rng.uniform(margin, L - margin, 3)
        r = rng.uniform(rmin, rmax)
        if all(np.linalg.norm(c - c2) >= r + r2 + gap for c2, r2 in zip(cen, rad)):
            cen.append(c)
            rad.append(r)
    return np.array(cen), np.array(rad)


def run_poisson3d(H, nx, precond="default", spheres=64, rtol=1e-10, maxit=40000, check_every=16, repeat=2, path="auto"):
    """BASELINE.json configs[4]: steady cut-cell Poisson (src/solver/diffusion.jl:14-72) around `spheres` random disjoint spheres (fluid outside), f = 1, D = 1,
    Dirichlet 0 on the interface and on the six recognised border keys, CG to ||r|| <= rtol ||b|| through the public API (solve_DiffusionSteadyMono_, the D2H read
    of solver.x included in time_to_tolerance_ms).  A "step" of this workload is one Krylov iteration: the system has no V / dt shift, kappa = O(n^2).
    precond = "mg": geometric multigrid V-cycle (csrc/mg.cuh, one GPU).  `repeat` solves: the first builds the folded system (and the hierarchy), the last is timed."""
    pb, torch = H.pb, H.torch
    N = H.world
    peak, peak_src = peaks()
    mesh = pb.Mesh((nx, nx, nx), (4.0, 4.0, 4.0))
    cen, rad = random_spheres(spheres)
    body = pb.Balls(cen, rad, fluid_inside=False)
    t0 = time.perf_counter()
    cap = pb.Capacity(body, mesh, compute_centroids=False)
    H.ctx.sync()
    cap_s = time.perf_counter() - t0
    phase = pb.Phase(cap, pb.DiffusionOps(cap), 1.0, 1.0)
    keys = ("left", "right", "top", "bottom", "forward", "backward")
    bc_b = pb.BorderConditions({k: pb.Dirichlet(0.0) for k in keys})
    s = pb.DiffusionSteadyMono(phase, bc_b, pb.Dirichlet(0.0))
    n = cap.nloc
    kw = dict(reltol=rtol, maxiter=maxit, check_every=check_every, path=path, precond=precond)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms, first_ms = 0.0, 0.0
    for k in range(max(1, repeat)):
        H.barrier()
        e0.record(H.ext)
        pb.solve_DiffusionSteadyMono_(s, method="cg", **kw)
        e1.record(H.ext)
        H.barrier()
        ms = H.allmax(e0.elapsed_time(e1))
        if k == 0:
            first_ms = ms
    ch = s.ch[-1]
    it, dof = int(ch["iters"]), int(ch["dof_bulk"])
    solve_ms = H.allmax(float(ch["solve_ms"]))
    cu, cg = int(ch["apply_cells_uniform"]), int(ch["apply_cells_general"])
    cells = cu + cg
    # per iteration of the fused CG: apply 6 passes + N coefficient arrays on the general tiles, update 3 passes (the byte model of the PLAIN iteration)
    it_bytes = H.allsum(8.0 * (9 * cells + 3 * cg))
    agg = it_bytes * it / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 and precond != "mg" else None
    T = s.x[:n]
    out = {"workload": "steady Poisson 3-D, union of random disjoint spheres (fluid outside), f = 1, Dirichlet 0 on interface and borders (BASELINE.json configs[4])",
           "grid": [nx, nx, nx], "spheres": int(spheres), "seed": POISSON_SEED, "n_gpus": N, "dof": dof, "rtol": rtol,
           "krylov": "CG on the folded (block-Jacobi-scaled) system, " + ("multigrid V-cycle preconditioner (levels rediscretised by the capacity kernels, cell-aggregation transfers, "
                                                                         "Chebyshev smoothers)" if precond == "mg" else "no multigrid"),
           "iterations": it, "converged": bool(ch["converged"]), "final_rel_residual": ch["rnorm"] / ch["bnorm"] if ch["bnorm"] else 0.0,
           "time_to_tolerance_ms": ms, "first_solve_ms_incl_setup": first_ms, "krylov_loop_ms": solve_ms, "prologue_ms": H.allmax(float(ch["setup_ms"])), "capacity_build_s": cap_s,
           "ms_per_iteration": solve_ms / max(it, 1), "dof_iterations_per_s": dof * it / (solve_ms * 1e-3) if solve_ms > 0 else 0.0,
           "plain_iteration_algorithmic_bytes_all_ranks": it_bytes, "aggregate_gbs": agg, "frac_of_measured_hbm": agg / (N * peak) if agg else None, "peak_source": peak_src,
           "cells_constant_coef_tiles_rank0": cu, "cells_streamed_coef_tiles_rank0": cg, "launches": int(ch["launches"]),
           "max_T_rank0": float(T.max()), "min_T_rank0": float(T.min())}
    del s, phase, cap
    return out


def run_gpu(args):
    H = Harness(args)
    torch, pb, L, lib, ctx = H.torch, H.pb, H.L, H.lib, H.ctx
    rank, world, local = H.rank, H.world, H.local
    barrier, allmax, allsum = H.barrier, H.allmax, H.allsum

    nx = args.nx
    N = world
    mesh = pb.Mesh((nx, nx * N), (8.0, 8.0 * N))
    body = pb.Balls([[4.0, 4.0 + 8.0 * k] for k in range(N)], [2.0] * N)
    t0 = time.perf_counter()
    c1, c2 = pb.Capacity(body, mesh, compute_centroids=False), pb.Capacity(-body, mesh, compute_centroids=False)
    ctx.sync()
    cap_s = time.perf_counter() - t0
    f = 0.0
    p1, p2 = pb.Phase(c1, pb.DiffusionOps(c1), f, 1.0), pb.Phase(c2, pb.DiffusionOps(c2), f, 1.0)
    nloc = c1.nloc
    h = 8.0 / nx
    dt = 0.5 * h * h
    ic = pb.InterfaceConditions(pb.ScalarJump(1.0, 1.0, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
    u0 = np.concatenate([np.ones(2 * nloc), np.zeros(2 * nloc)])
    s = pb.DiffusionUnsteadyDiph(p1, p2, pb.BorderConditions(), ic, dt, u0, "BE")

    opts = L.KrylovOpts()
    opts.method, opts.rtol, opts.atol, opts.maxit, opts.warm_start, opts.check_every = args.method, 1e-10, 0.0, 5000, args.warm, args.check_every
    si = L.StepIn()
    si.scheme, si.dt = 0, dt
    st = L.StepStats()

    def step():
        rc = lib.pb200_solver_step(s._h, C.byref(si), C.byref(opts), C.byref(st))
        L.check(rc, ctx.h)

    lib.pb200_set_profiling(ctx.h, 0)
    # Spin-up: a fresh box idles at low clocks and the first launches build the folded system and capture the CUDA graphs.  Run the
    # workload untimed for ~0.75 s, then RESET the state to the initial condition so that the W warm-up steps and the K timed steps are
    # steps 1..W and W+1..W+K of the run, exactly as without spin-up (later steps of the transient need fewer Krylov iterations).
    # (seen once on a fresh box: the first process ran its first ~100 ms of steps 4x slower than every later one -- host side still paging in --
    # so the spin-up also waits until blocks of 25 steps have stopped getting faster, for at most 4 s)
    t_spin = time.perf_counter()
    best, settled = None, 0
    while True:
        tb = time.perf_counter()
        for _ in range(25):
            step()
        ctx.sync()
        blk = time.perf_counter() - tb
        settled = settled + 1 if (best is not None and blk <= 1.1 * best) else 0
        best = blk if best is None else min(best, blk)
        el = time.perf_counter() - t_spin
        if allmax(0.0 if (el >= args.spinup and settled >= 3) or el >= 4.0 else 1.0) == 0.0:     # a COLLECTIVE decision: every rank runs the same number of steps
            break
    ctx.sync()
    L.check(lib.pb200_solver_set_state(s._h, u0.ctypes.data_as(L.dp)), ctx.h)
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    ms, iters, setup_ms, solve_ms = H.timed_steps(step, args.steps, st)
    launches = int(allsum(ctx.launches - launches0))
    dof = int(st.dof_bulk)                          # all ranks (allreduced inside the library)
    value = dof * args.steps / (ms * 1e-3)
    rnorm_rel = st.rnorm / st.bnorm if st.bnorm > 0 else 0.0

    # ---- roofline pass: the same K steps again with every launch of the Krylov loop bracketed by CUDA events on the launching stream
    # (pb200_set_profiling).  Kept out of the headline region because per-launch events force plain launches instead of graph replay.
    kms, kn = ([0.0] * 8, [0] * 8) if args.no_profile else H.profile_steps(step, args.steps, st)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the public API with HOST buffers: per step the jump data g, h go host -> device from pinned
    # memory and the new state comes back device -> host (the reference pushes every state to solver.states) ---------------
    pin = lambda n: torch.empty(n, dtype=torch.float64).pin_memory().numpy()
    g_host, h_host, x_host = pin(nloc), pin(nloc), pin(4 * nloc)
    g_host[:] = 0.0
    h_host[:] = 0.0
    dp = L.dp
    si.g_arr[0] = g_host.ctypes.data_as(dp)
    si.g_arr[1] = h_host.ctypes.data_as(dp)
    e2e_steps = max(1, min(args.steps, 50))
    x_bufs = [x_host, pin(4 * nloc)]
    for k in range(2):
        step()
        L.check(lib.pb200_solver_get_state_async(s._h, x_bufs[k & 1].ctypes.data_as(dp)), ctx.h)
    L.check(lib.pb200_solver_wait_state(s._h), ctx.h)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        step()                                                       # uploads g, h (pinned host arrays) and solves
        L.check(lib.pb200_solver_get_state_async(s._h, x_bufs[k & 1].ctypes.data_as(dp)), ctx.h)   # state k streams out under step k + 1
    L.check(lib.pb200_solver_wait_state(s._h), ctx.h)
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    state_absmax = allmax(float(np.max(np.abs(x_bufs[(e2e_steps - 1) & 1]))))   # the reference prints max|x| every step (diffusion.jl:448)
    e2e = {"value": dof * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(allsum(2 * 8 * nloc)),
           "d2h_bytes_per_step": int(allsum(4 * 8 * nloc)), "steps": e2e_steps, "max_abs_state": state_absmax, "timing": "host wall clock between device syncs, max over ranks; per step: H2D of the jump data g, h from pinned memory, "
                     "the solve, D2H of the full state [T_w1; T_g1; T_w2; T_g2] into pinned memory (double-buffered: it overlaps the next step)"}
    si.g_arr[0] = None
    si.g_arr[1] = None

    # ---- roofline: every kernel class of the Krylov loop, the whole iteration and the whole step ------------------------------------
    peak, peak_src = peaks()
    roof = None
    cu, cg, rows_b = int(st.apply_cells_uniform), int(st.apply_cells_general), int(st.band_rows)
    cells = cu + cg
    if sum(kn) and sum(kms) > 0:
        table, abytes = kernel_table(H, kms, kn, cells, cg, rows_b, mesh.N, peak)
        # the line's roofline object describes the kernel that is FURTHEST below the roofline among those that matter (>= 15 % of the timed kernel time)
        cand = [r for r in table if "frac" in r and r["share_of_timed_kernels"] >= 0.15]
        dom = min(cand, key=lambda r: r["frac"]) if cand else max((r for r in table if "frac" in r), key=lambda r: r["share_of_timed_kernels"])
        it_mean = float(np.mean(iters))
        per_iter = abytes["per_iteration"]
        iter_us = 1e3 * (solve_ms / args.steps) / max(it_mean, 1e-9)
        step_bytes = it_mean * per_iter + 8 * 14 * cells          # + prologue (V, T, sc, b, b^, x^0, 3 older states, mask) and epilogue (x^, x, T, ufix, ...) passes
        traffic, dram_frac = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("nx") == nx and tj.get("n_gpus", 1) == world and tj.get("round") == 2:
                traffic = tj.get("dram_bytes_per_launch", {}).get(dom["class"])
                if traffic:
                    dram_frac = traffic / (dom["avg_launch_us"] * 1e-6) / 1e9 / peak
        roof = {"bound": "hbm", "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"], "traffic": traffic,
                "dram_side_frac": dram_frac, "frac_of_nominal_8TBs": dom["achieved_gbs"] / 8000.0,
                "kernel": dom["kernel"], "kernel_class": dom["class"],
                "dominant_rule": "lowest fraction among the kernel classes with >= 15 % of the timed kernel time",
                "algorithmic_bytes_per_dof": dom["algorithmic_bytes_per_launch"] / (dof / world), "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_launch"],
                "cells_constant_coef_tiles": cu, "cells_streamed_coef_tiles": cg, "band_rows": rows_b, "launches_timed": dom["launches_timed"],
                "avg_launch_us": dom["avg_launch_us"], "peak_source": peak_src,
                "timed": "CUDA events around every launch of the Krylov loop, separate pass of the same K steps (plain launches on one stream; the timed region replays CUDA graphs)",
                "whole_iteration": {"algorithmic_bytes": int(per_iter), "us": iter_us, "achieved_gbs": per_iter / (iter_us * 1e-6) / 1e9,
                                    "frac": per_iter / (iter_us * 1e-6) / 1e9 / peak, "note": "solve time of the timed region / iterations (graph replay, both streams)"},
                "whole_step": {"algorithmic_bytes": int(step_bytes), "ms": ms / args.steps, "achieved_gbs": step_bytes / (ms / args.steps * 1e-3) / 1e9,
                               "frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak, "frac_of_nominal_8TBs": step_bytes / (ms / args.steps * 1e-3) / 1e9 / 8000.0},
                "kernels": table}

    vec_len = int(allsum(4 * nloc))     # collective: every rank calls it
    del s, p1, p2, c1, c2
    import gc
    gc.collect()
    extra = {}
    if not args.no_3d:
        for kind, key in (("diph", "heat3d_diph_weak"), ("mono", "heat3d_mono_512_strong")):
            try:
                extra[key] = run_heat3d(H, args, kind)
            except Exception as e:          # the headline line must survive a failure of the additional workloads
                extra[key] = {"error": repr(e)[:300]}
                barrier()
            gc.collect()
    if not args.no_3d and not args.no_poisson:
        # configs[4] at a single-GPU size: the plain folded CG and the multigrid-preconditioned one on the same problem (one rank: csrc/mg.cuh)
        if world == 1:
            for pre, key in (("default", "poisson3d_steady_cg"), ("mg", "poisson3d_steady_mgcg")):
                try:
                    extra[key] = run_poisson3d(H, args.nxpoisson, pre)
                except Exception as e:
                    extra[key] = {"error": repr(e)[:300]}
                gc.collect()
        else:
            extra["poisson3d_steady_mgcg"] = {"skipped": "the multigrid preconditioner runs on one rank (the slab decomposition does not coarsen with the grid); tools/run_poisson3d.py --gpus N (not yet measured) runs the plain CG on N ranks"}
    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:
            v, cdof, el = cpu_sample(256, 12)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"256x256 sample of the workload (cross-size, DOF-normalised), 12 BE steps in {el:.1f} s, sparse LU per step (SciPy SuperLU, serial), {cdof} DOF"}
        fields_mb = 8 * nloc / 1e6
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": {"workload": "Heat_2ph_2D diphasic BE step (BASELINE.json configs[1])", "grid": [nx, nx * N], "cells_per_gpu": [nx, nx],
                          "dof": dof, "vector_length_4n": vec_len, "scheme": "BE", "dt": dt, "krylov": "BiCGSTAB" if args.method == 2 else "CG (symmetrised, block-Jacobi-scaled system), fused iteration",
                          "rtol": 1e-10, "initial_guess": "zero" if args.warm == 0 else f"polynomial extrapolation through the last {args.warm} states", "iters_per_step": float(np.mean(iters)), "rhs_assembly_ms_per_step": setup_ms / args.steps, "solve_ms_per_step": solve_ms / args.steps, "final_rel_residual": rnorm_rel,
                          "parallelism": f"y-slab x{world}" if world > 1 else "single GPU",
                          "l2": f"inputs larger than L2: the Krylov loop streams x, r, p (two buffers), v ({5 * fields_mb:.0f} MB on the active tiles) plus coefficient and band arrays "
                                f"every iteration vs {L2_MB} MB of L2; no flush between steps",
                          "capacity_build_s": cap_s},
               "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
        out.update(extra)
        emit(out)
    if H.dist is not None:
        H.dist.barrier()
        H.dist.destroy_process_group()
    pb.finalize()


_REAL_STDOUT = None


def emit(obj):
    """the ONE JSON line, on the process's original stdout"""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    # Libraries (NCCL's version banner, torch warnings) write to fd 1: point it at stderr for the whole run and keep the original
    # stdout for the result line only.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nx", type=int, default=2048)
    ap.add_argument("--method", type=int, default=0, help="0 auto (CG on the folded system), 1 CG, 2 BiCGSTAB")
    ap.add_argument("--check-every", type=int, default=8)
    ap.add_argument("--warm", type=int, default=4, help="initial guess: 0 zero, 1 previous state, 2 linear, 3 quadratic extrapolation of the previous states")
    ap.add_argument("--spinup", type=float, default=0.75, help="seconds of untimed steps before the state is reset and the W warm-up steps start")
    ap.add_argument("--no-profile", action="store_true", help="do not bracket the apply launches with CUDA events")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-3d", action="store_true", help="skip the additional 3-D workloads (configs[3] weak-scaled, configs[2] strong-scaled)")
    ap.add_argument("--nx3d", type=int, default=1024, help="configs[3]: cells in x and y")
    ap.add_argument("--nz3d", type=int, default=128, help="configs[3]: planes per GPU (128 x 8 GPUs = the 1024^3 north-star problem)")
    ap.add_argument("--nxmono", type=int, default=512, help="configs[2]: cells per direction (strong-scaled)")
    ap.add_argument("--no-poisson", action="store_true", help="skip the configs[4] workload (steady Poisson around 64 spheres, plain and multigrid-preconditioned CG; N = 1 only)")
    ap.add_argument("--nxpoisson", type=int, default=384, help="configs[4]: cells per direction")
    ap.add_argument("--steps3d", type=int, default=10)
    ap.add_argument("--warmup3d", type=int, default=3)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
