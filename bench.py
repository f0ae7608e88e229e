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
def run_gpu(args):
    import torch
    import penguin_b200 as pb
    from penguin_b200 import _lib as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries exactly one JSON line
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

        def bcast(ident):
            obj = [ident]
            dist.broadcast_object_list(obj, src=0)
            return obj[0]
        ctx = pb.init_distributed(rank, world, local, bcast)
    else:
        ctx = pb.init(local)
    lib = L.lib()

    def barrier():
        torch.cuda.synchronize()
        ctx.sync()
        if dist is not None:
            dist.barrier()

    def allmax(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

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

    ext = torch.cuda.ExternalStream(ctx.stream)     # the library's launching stream, for torch.cuda.Event timing
    lib.pb200_set_profiling(ctx.h, 0)
    # Spin-up: a fresh box idles at low clocks and the first launches build the folded system and capture the CUDA graphs.  Run the
    # workload untimed for ~0.75 s, then RESET the state to the initial condition so that the W warm-up steps and the K timed steps are
    # steps 1..W and W+1..W+K of the run, exactly as without spin-up (later steps of the transient need fewer Krylov iterations).
    t_spin = time.perf_counter()
    while allmax(time.perf_counter() - t_spin) < args.spinup:     # a COLLECTIVE decision: every rank runs the same number of steps
        step()
    ctx.sync()
    L.check(lib.pb200_solver_set_state(s._h, u0.ctypes.data_as(L.dp)), ctx.h)
    for _ in range(args.warmup):
        step()
    iters = []
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    setup_ms, solve_ms = 0.0, 0.0
    for _ in range(args.steps):
        step()
        iters.append(st.iters)
        setup_ms += st.setup_ms
        solve_ms += st.solve_ms
    e1.record(ext)
    barrier()
    ms = allmax(e0.elapsed_time(e1))
    launches = int(allsum(ctx.launches - launches0))
    dof = int(st.dof_bulk)                          # all ranks (allreduced inside the library)
    value = dof * args.steps / (ms * 1e-3)
    rnorm_rel = st.rnorm / st.bnorm if st.bnorm > 0 else 0.0

    # ---- roofline pass: the same K steps again with every operator-apply launch bracketed by CUDA events on the launching stream
    # (pb200_set_profiling).  Kept out of the headline region because per-launch events force plain launches instead of graph replay.
    kms, kn = [0.0, 0.0, 0.0], [0, 0, 0]
    if not args.no_profile:
        lib.pb200_set_profiling(ctx.h, 1)
        for _ in range(args.steps):
            step()
            for q in range(3):
                kms[q] += st.kernel_ms[q]
                kn[q] += st.kernel_launches[q]
        lib.pb200_set_profiling(ctx.h, 0)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the public API with HOST buffers: per step the jump data g, h go host -> device from pinned
    # memory and the new state comes back device -> host (the reference pushes every state to solver.states) ---------------
    lib.pb200_set_profiling(ctx.h, 0)
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

    # ---- roofline of the dominant kernel (the operator apply inside the Krylov loop) ---------------------------------------------
    peak, peak_src = peaks()
    roof = None
    if kn[0] and kms[0] > 0:
        # Algorithmic bytes per launch on this rank, from the tile census of the folded system:
        #   apply   : x and y for every cell of an active tile + the N coefficient arrays for the cells of tiles whose coefficients are
        #             not constants (interface band, domain border ring)
        #   update  : r -= a v with fused dots: read v, r, write r -> 3 passes
        #   p-update: x += a p, p = z + b p: read r, p, x, write p, x -> 5 passes
        cu, cg = int(st.apply_cells_uniform), int(st.apply_cells_general)
        cells = cu + cg
        names = ["operator apply, dense part (kf_apply_dense)", "residual update + fused dots (kf_cg_update)", "solution + search direction update (kf_cg_p)"]
        abytes = [8 * (2 * cells + mesh.N * cg), 8 * 3 * cells, 8 * 5 * cells]
        table = []
        for q in range(3):
            if kn[q] and kms[q] > 0:
                us = 1e3 * kms[q] / kn[q]
                table.append({"kernel": names[q], "launches_timed": int(kn[q]), "avg_launch_us": us, "algorithmic_bytes_per_launch": abytes[q],
                              "achieved_gbs": abytes[q] / (us * 1e-6) / 1e9, "frac": abytes[q] / (us * 1e-6) / 1e9 / peak,
                              "share_of_timed_kernels": kms[q] / sum(kms)})
        dom = max(range(3), key=lambda q: kms[q])
        us = 1e3 * kms[dom] / kn[dom]
        achieved = abytes[dom] / (us * 1e-6) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("nx") == nx and tj.get("n_gpus", 1) == world:
                traffic = tj.get("dram_bytes_per_launch", {}).get(["apply", "update", "pupdate"][dom])
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "frac_of_nominal_8TBs": achieved / 8000.0,     # BASELINE.json quotes the metric against the 8 TB/s datasheet figure
                "kernel": names[dom], "algorithmic_bytes_per_dof": abytes[dom] / (dof / world), "algorithmic_bytes_per_launch": abytes[dom],
                "cells_constant_coef_tiles": cu, "cells_streamed_coef_tiles": cg, "launches_timed": int(kn[dom]),
                "avg_launch_us": us, "peak_source": peak_src, "timed": "CUDA events around every launch, separate pass of the same K steps",
                "kernels": table}

    vec_len = int(allsum(4 * nloc))     # collective: every rank calls it
    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:
            v, cdof, el = cpu_sample(256, 12)
            cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": f"256x256 sample of the workload, 12 BE steps in {el:.1f} s, sparse LU per step (SciPy SuperLU, serial), {cdof} DOF"}
        fields_mb = 8 * nloc / 1e6
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": {"workload": "Heat_2ph_2D diphasic BE step (BASELINE.json configs[1])", "grid": [nx, nx * N], "cells_per_gpu": [nx, nx],
                          "dof": dof, "vector_length_4n": vec_len, "scheme": "BE", "dt": dt, "krylov": "BiCGSTAB" if args.method == 2 else "CG (symmetrised, block-Jacobi-scaled system)",
                          "rtol": 1e-10, "initial_guess": "zero" if args.warm == 0 else f"polynomial extrapolation through the last {args.warm} states", "iters_per_step": float(np.mean(iters)), "rhs_assembly_ms_per_step": setup_ms / args.steps, "solve_ms_per_step": solve_ms / args.steps, "final_rel_residual": rnorm_rel,
                          "parallelism": f"y-slab x{world}" if world > 1 else "single GPU",
                          "l2": f"inputs larger than L2: the Krylov loop streams x, r, p, v ({4 * fields_mb:.0f} MB on the active tiles) plus coefficient and band arrays "
                                f"every iteration vs {L2_MB} MB of L2; no flush between steps",
                          "capacity_build_s": cap_s},
               "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
        emit(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
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
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
