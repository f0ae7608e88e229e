"""BASELINE.json configs[4]: steady cut-cell Poisson in 3-D, union of K random disjoint spheres (fluid outside), f = 1, D = 1, interface
Dirichlet 0, borders Dirichlet 0 on the six recognised keys, Krylov solve to ||r|| <= 1e-10 ||b|| (src/solver/diffusion.jl:14-72 with the geometry of
BenchPhaseFlow/problems/scalar/Scalar_3D_Diffusion_Poisson_Dirichlet.jl:43-61 widened to many spheres).

    python tools/run_poisson3d.py --nx 256                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/run_poisson3d.py --gpus 8 --nx 1536   # z-slabs, 8 GPUs: plain CG only, NOT yet run (no steady solve has been measured on several ranks)

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

SEED = bench.POISSON_SEED
random_spheres = bench.random_spheres


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--nx", type=int, default=256)
    ap.add_argument("--spheres", type=int, default=64)
    ap.add_argument("--rtol", type=float, default=1e-10)
    ap.add_argument("--maxit", type=int, default=40000)
    ap.add_argument("--check-every", type=int, default=16)
    ap.add_argument("--repeat", type=int, default=2, help="solves (a steady solve always starts from a zero guess); the first one builds the folded system and captures the graphs, the last one is timed")
    ap.add_argument("--path", default="auto", choices=["auto", "folded", "generic"])
    ap.add_argument("--precond", default="default", choices=["default", "mg"], help="mg: geometric multigrid V-cycle as the CG preconditioner (csrc/mg.cuh, one GPU)")
    args = ap.parse_args()
    H = bench.Harness(args)
    out = bench.run_poisson3d(H, args.nx, args.precond, args.spheres, args.rtol, args.maxit, args.check_every, args.repeat, args.path)
    if H.rank == 0:
        print(json.dumps(out), flush=True)
    H.pb.finalize()
    if H.dist is not None:
        H.dist.destroy_process_group()


if __name__ == "__main__":
    main()
