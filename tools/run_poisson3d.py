"""BASELINE.json configs[4]: steady cut-cell Poisson in 3-D, union of K random disjoint spheres (fluid outside), f = 1, D = 1, interface
Dirichlet 0, borders Dirichlet 0 on the six recognised keys, Krylov solve to ||r|| <= 1e-10 ||b|| (src/solver/diffusion.jl:14-72 with the geometry of
BenchPhaseFlow/problems/scalar/Scalar_3D_Diffusion_Poisson_Dirichlet.jl:43-61 widened to many spheres).

    python tools/run_poisson3d.py --nx 256                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_poisson3d.py --gpus 8 --nx 1536   # z-slabs, 8 GPUs

A "step" of this workload is one Krylov iteration (the system has no V/dt shift: kappa = O(n^2), un-multigridded CG needs O(n) iterations), so the
line reports iterations, time-to-tolerance, DOF*iterations/s and the HBM fraction of the iteration.  Prints one JSON line (not the bench line:
bench.py measures configs[1])."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                     # noqa: E402  (Harness: process group, context, event timing on the library's stream)

SEED = 20261018


def random_spheres(K=64, L=4.0, seed=SEED, rmin=0.1, rmax=0.3, margin=0.5, gap=0.11):
    """K spheres, centres U[margin, L - margin]^3, radii U[rmin, rmax], rejection-sampled so that any two surfaces are at least `gap` apart
    (two cell diagonals at 128^3: no cell, and no staggered volume between two barycentres, meets two spheres)."""
    rng = np.random.default_rng(seed)
    cen, rad = [], []
    while len(cen) < K:
        c = rng.uniform(margin, L - margin, 3)
        r = rng.uniform(rmin, rmax)
        if all(np.linalg.norm(c - c2) >= r + r2 + gap for c2, r2 in zip(cen, rad)):
            cen.append(c)
            rad.append(r)
    return np.array(cen), np.array(rad)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--spheres", type=int, default=64)
    ap.add_argument("--rtol", type=float, default=1e-10)
    ap.add_argument("--maxit", type=int, default=40000)
    ap.add_argument("--check-every", type=int, default=16)
    ap.add_argument("--repeat", type=int, default=2, help="solves (a steady solve always starts from a zero guess, penguin_b200.cu fold step); the first one builds the folded system and captures the graphs, the last one is timed")
    ap.add_argument("--path", default="auto", choices=["auto", "folded", "generic"])
    ap.add_argument("--precond", default="default", choices=["default", "mg"], help="mg: geometric multigrid V-cycle as the CG preconditioner (csrc/mg.cuh, one GPU)")
    args = ap.parse_args()
    H = bench.Harness(args)
    pb, torch = H.pb, H.torch
    N = H.world
    peak, peak_src = bench.peaks()
    nx = args.nx
    mesh = pb.Mesh((nx, nx, nx), (4.0, 4.0, 4.0))
    cen, rad = random_spheres(args.spheres)
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
    kw = dict(reltol=args.rtol, maxiter=args.maxit, check_every=args.check_every, path=args.path, precond=args.precond)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = 0.0
    for k in range(max(1, args.repeat)):
        H.barrier()
        e0.record(H.ext)
        pb.solve_DiffusionSteadyMono_(s, method="cg", **kw)
        e1.record(H.ext)
        H.barrier()
        ms = H.allmax(e0.elapsed_time(e1))           # includes the D2H read of the state (solver.x), as the reference's solve does
    ch = s.ch[-1]
    it, dof = int(ch["iters"]), int(ch["dof_bulk"])
    solve_ms = H.allmax(float(ch["solve_ms"]))
    cu, cg = int(ch["apply_cells_uniform"]), int(ch["apply_cells_general"])
    cells = cu + cg
    # per iteration of the fused CG: apply 6 passes + N coefficient arrays on the general tiles, update 3 passes
    it_bytes = H.allsum(8.0 * (9 * cells + 3 * cg))
    agg = it_bytes * it / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 and args.precond != "mg" else 0.0      # (the byte model is the plain CG iteration's)
    T = s.x[:n]
    out = {"workload": "steady Poisson 3-D, union of random disjoint spheres (fluid outside), f = 1, Dirichlet 0 on interface and borders (BASELINE.json configs[4])",
           "grid": [nx, nx, nx], "spheres": int(args.spheres), "seed": SEED, "n_gpus": N, "dof": dof, "rtol": args.rtol,
           "krylov": "CG on the folded (block-Jacobi-scaled) system, " + ("multigrid V-cycle preconditioner (rediscretised levels, Chebyshev smoothers)" if args.precond == "mg" else "no multigrid"), "iterations": it, "converged": bool(ch["converged"]),
           "final_rel_residual": ch["rnorm"] / ch["bnorm"] if ch["bnorm"] else 0.0,
           "time_to_tolerance_ms": ms, "krylov_loop_ms": solve_ms, "prologue_ms": H.allmax(float(ch["setup_ms"])), "capacity_build_s": cap_s,
           "ms_per_iteration": solve_ms / max(it, 1), "dof_iterations_per_s": dof * it / (solve_ms * 1e-3) if solve_ms > 0 else 0.0,
           "algorithmic_bytes_per_iteration_all_ranks": it_bytes, "aggregate_gbs": agg, "frac_of_measured_hbm": agg / (N * peak), "peak_source": peak_src,
           "cells_constant_coef_tiles_rank0": cu, "cells_streamed_coef_tiles_rank0": cg, "launches": int(ch["launches"]),
           "max_T_rank0": float(T.max()), "min_T_rank0": float(T.min())}
    if H.rank == 0:
        print(json.dumps(out), flush=True)
    del s
    pb.finalize()
    if H.dist is not None:
        H.dist.destroy_process_group()


if __name__ == "__main__":
    main()
