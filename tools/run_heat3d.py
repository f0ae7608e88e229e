"""BASELINE.json configs[2] / configs[3] at single-GPU sizes: 3-D monophasic Heat3D (benchmark/Heat3D.jl:53-74): sphere embedded boundary (exterior phase),
interface Dirichlet 1, borders Dirichlet 1 on the six recognised keys, u0 = 0, first step BE then CN, dt = 0.75 h^2.
    python tools/run_heat3d.py --nx 512 --steps 50
Prints one JSON line (not the bench line: bench.py measures configs[1])."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import penguin_b200 as pb                       # noqa: E402
from penguin_b200 import _lib as L              # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--interior", action="store_true", help="fluid inside the sphere (the script's phase, 6.5 % active) instead of outside")
    ap.add_argument("--diph", action="store_true", help="configs[3]: diphasic BE, sphere interface r = 1 centre (2,2,2), ScalarJump(1,2,0), FluxJump(1,1,0), "
                                                        "u0 = [1,1,0,0], dt = 0.5 h^2, empty borders (examples/3D/Diffusion/Heat_2ph.jl)")
    ap.add_argument("--warm", type=int, default=4)
    ap.add_argument("--nz", type=int, default=0, help="--diph: planes in z (default nx); the box is [0,4]^2 x nz h centred on the sphere, as bench.py's configs[3] slab")
    args = ap.parse_args()
    import torch
    ctx = pb.init(0)
    lib = L.lib()
    nx = args.nx
    mesh = pb.Mesh((nx, nx, nx), (4.0, 4.0, 4.0))
    if args.diph and args.nz:
        Lz = args.nz * 4.0 / nx
        mesh = pb.Mesh((nx, nx, args.nz), (4.0, 4.0, Lz), (0.0, 0.0, 2.0 - 0.5 * Lz))
    if args.diph:
        body = pb.Sphere((2.0, 2.0, 2.0), 1.0)
        t0 = time.perf_counter()
        cap = pb.Capacity(body, mesh, compute_centroids=False)
        cap2 = pb.Capacity(-body, mesh, compute_centroids=False)
        ctx.sync()
        cap_s = time.perf_counter() - t0
        ph, ph2 = pb.Phase(cap, pb.DiffusionOps(cap), 0.0, 1.0), pb.Phase(cap2, pb.DiffusionOps(cap2), 0.0, 1.0)
        n = cap.nloc
        h = 4.0 / nx
        dt = 0.5 * h * h
        ic = pb.InterfaceConditions(pb.ScalarJump(1.0, 2.0, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
        u0 = np.concatenate([np.ones(2 * n), np.zeros(2 * n)])
        s = pb.DiffusionUnsteadyDiph(ph, ph2, pb.BorderConditions(), ic, dt, u0, "BE")
        del u0
    else:
        body = pb.Sphere((2.01, 2.01, 2.01), 1.0)
        if not args.interior:
            body = -body
        t0 = time.perf_counter()
        cap = pb.Capacity(body, mesh, compute_centroids=False)
        ctx.sync()
        cap_s = time.perf_counter() - t0
        ph = pb.Phase(cap, pb.DiffusionOps(cap), 0.0, 1.0)
        n = cap.nloc
        h = 4.0 / nx
        dt = 0.75 * h * h
        keys = ("left", "right", "top", "bottom", "forward", "backward")
        bc = pb.BorderConditions({k: pb.Dirichlet(1.0) for k in keys})
        s = pb.DiffusionUnsteadyMono(ph, bc, pb.Dirichlet(1.0), dt, np.zeros(2 * n), "BE")
    opts = L.KrylovOpts()
    opts.method, opts.rtol, opts.atol, opts.maxit, opts.warm_start, opts.check_every = 0, 1e-10, 0.0, 5000, args.warm, 8
    si = L.StepIn()
    si.dt = dt
    si.g_const[0] = si.g_const[1] = 0.0 if args.diph else 1.0
    cn = 0 if args.diph else 1
    st = L.StepStats()

    def step(scheme):
        si.scheme = scheme
        L.check(lib.pb200_solver_step(s._h, C.byref(si), C.byref(opts), C.byref(st)), ctx.h)
    step(0)                                      # the constructor's BE step
    for _ in range(2):
        step(cn)
    ext = torch.cuda.ExternalStream(ctx.stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.sync()
    e0.record(ext)
    iters = []
    for _ in range(args.steps):
        step(cn)
        iters.append(st.iters)
    e1.record(ext)
    ctx.sync()
    ms = e0.elapsed_time(e1)
    dof = int(st.dof_bulk)
    lib.pb200_set_profiling(ctx.h, 1)
    kms, kn = [0.0] * 3, [0] * 3
    for _ in range(min(args.steps, 5)):
        step(cn)
        for q in range(3):
            kms[q] += st.kernel_ms[q]; kn[q] += st.kernel_launches[q]
    lib.pb200_set_profiling(ctx.h, 0)
    cu, cg = int(st.apply_cells_uniform), int(st.apply_cells_general)
    cells = cu + cg
    abytes = [8 * (2 * cells + 3 * cg), 8 * 3 * cells, 8 * 5 * cells]
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    kern = []
    for q, nm in enumerate(["apply", "update", "p-update"]):
        if kn[q]:
            us = 1e3 * kms[q] / kn[q]
            kern.append({"kernel": nm, "avg_us": us, "achieved_gbs": abytes[q] / (us * 1e-6) / 1e9, "frac_of_measured_peak": abytes[q] / (us * 1e-6) / 1e9 / peak})
    x = np.empty((4 if args.diph else 2) * n)
    L.check(lib.pb200_solver_get_state(s._h, x.ctypes.data_as(L.dp)), ctx.h)
    print(json.dumps({"workload": (f"3-D diphasic heat {nx}^3 BE, sphere interface" if args.diph else f"Heat3D {nx}^3 monophasic CN ({'interior' if args.interior else 'exterior'} phase)"), "dof": dof, "padded_cells": int(n),
                      "steps": args.steps, "ms_per_step": ms / args.steps, "dof_steps_per_s": dof * args.steps / (ms * 1e-3),
                      "iters_per_step": float(np.mean(iters)), "capacity_build_s": cap_s, "cells_constant_coef_tiles": cu,
                      "cells_streamed_coef_tiles": cg, "kernels": kern, "max_T": float(x[:n].max()), "min_T": float(x[:n].min()),
                      "final_rel_residual": st.rnorm / st.bnorm if st.bnorm else 0.0}))
    pb.finalize()


if __name__ == "__main__":
    main()
