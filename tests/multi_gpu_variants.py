"""Slab-partitioned solves at sizes where the multi-rank machinery is really exercised -- interior-class AND ghost-class tiles, the
fused iteration with the halo exchange on the second stream, ghost planes of >= 16384 doubles (several blocks of the peer-memory
halo kernel) -- compared with the SAME problem solved on one GPU (itself oracle-checked at the sizes of the parity tests).

    python tests/multi_gpu_variants.py --single /tmp/ref.npz                       # one GPU: reference states
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
        tests/multi_gpu_variants.py --ref /tmp/ref.npz                              # N ranks: every kernel variant against it
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import penguin_b200 as pb                      # noqa: E402
from penguin_b200 import slab                  # noqa: E402

VARIANTS = {"fused_pipelined": {}, "fused_band_launches": {"PB200_NO_BANDFUSE": "1"}, "unfused_tma": {"PB200_NO_FUSED": "1"}, "register_kernels": {"PB200_NO_TMA": "1", "PB200_NO_REPITCH": "1"}}


def run_case(case, rank, world):
    cut = lambda a, dims: np.ascontiguousarray(slab.scatter_owned(a, dims, rank, world))
    if case == "diph3d":
        dims, L = (136, 128, 64), (4.25, 4.0, 2.0)
        body = pb.Sphere((2.1, 2.0, 1.0), 0.8)
    elif case == "mono3d":
        dims, L = (136, 128, 64), (4.25, 4.0, 2.0)
        body = -pb.Sphere((2.1, 2.0, 1.0), 0.7)
    elif case == "diph2d_heads":
        # one interface per slab of a 2-rank split, away from the slab faces: the band lies in interior-class tiles on every rank, so the band
        # heads (the band work inside the two streaming kernels) run with several ranks as well -- bench.py's weak-scaled configs[1] in small
        dims, L = (128, 512), (4.0, 16.0)
        body = pb.Balls([[2.0, 4.0], [2.0, 12.0]], [1.5, 1.5])
    else:
        dims, L = (512, 1024), (4.0, 8.0)
        body = pb.Balls([[2.0, 4.0]], [1.3])          # the interface crosses the slab boundary of a 2-rank split
    mesh = pb.Mesh(dims, L)
    n = int(np.prod([d + 1 for d in dims]))
    h = L[0] / dims[0]
    if case == "mono3d":
        cap = pb.Capacity(body, mesh, compute_centroids=False)
        ph = pb.Phase(cap, pb.DiffusionOps(cap), 0.0, 1.0)
        keys = ("left", "right", "top", "bottom", "forward", "backward")
        bc = pb.BorderConditions({k: pb.Dirichlet(1.0) for k in keys})
        dt = 0.75 * h * h
        u0 = np.concatenate([cut(np.zeros(n), dims)] * 2)
        s = pb.DiffusionUnsteadyMono(ph, bc, pb.Dirichlet(1.0), dt, u0, "BE")
        pb.solve_DiffusionUnsteadyMono_(s, ph, dt, 2.5 * dt, bc, pb.Dirichlet(1.0), "CN", reltol=1e-12, warm_start=2)
        nblk = 2
    else:
        c1, c2 = pb.Capacity(body, mesh, compute_centroids=False), pb.Capacity(-body, mesh, compute_centroids=False)
        p1, p2 = pb.Phase(c1, pb.DiffusionOps(c1), 0.0, 1.0), pb.Phase(c2, pb.DiffusionOps(c2), 0.0, 1.0)
        ic = pb.InterfaceConditions(pb.ScalarJump(1.0, 2.0, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
        dt = 0.5 * h * h
        u0 = np.concatenate([cut(np.ones(n), dims)] * 2 + [cut(np.zeros(n), dims)] * 2)
        s = pb.DiffusionUnsteadyDiph(p1, p2, pb.BorderConditions(), ic, dt, u0, "BE")
        pb.solve_DiffusionUnsteadyDiph_(s, p1, p2, dt, 2.5 * dt, pb.BorderConditions(), ic, "BE", reltol=1e-12, warm_start=2)
        nblk = 4
    assert all(c["converged"] for c in s.ch), [(c["iters"], c["rnorm"]) for c in s.ch]
    return dims, nblk, s


def assemble(per_rank_states, dims, nblk):
    out = []
    for k in range(len(per_rank_states[0])):
        blocks = []
        for b in range(nblk):
            parts = []
            for st in per_rank_states:
                nl = st[k].shape[0] // nblk
                parts.append(st[k][b * nl:(b + 1) * nl])
            blocks.append(slab.gather_owned(parts, dims))
        out.append(np.concatenate(blocks))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--single", default=None)
    ap.add_argument("--ref", default=None)
    ap.add_argument("--cases", default="diph3d,mono3d,diph2d,diph2d_heads")
    ap.add_argument("--team", type=int, default=0, help="ONE process driving this many GPUs (pb200_init_multi): global host arrays, compared with --ref")
    args = ap.parse_args()
    cases = args.cases.split(",")
    if args.team:
        import faulthandler
        faulthandler.enable()
        import torch  # noqa: F401  (loads the NCCL build that ships with torch before the library dlopens libnccl.so.2)
        print("init_multi ...", flush=True)
        pb.init_multi(list(range(args.team)))
        print("init_multi done", flush=True)
        ref = np.load(args.ref)
        worst = 0.0
        for case in cases:
            dims, nblk, s = run_case(case, 0, 1)          # the team handle takes and returns GLOBAL arrays
            for k, x in enumerate(s.states):
                r = ref[f"{case}_{k}"]
                err = np.linalg.norm(x - r) / np.linalg.norm(r)
                worst = max(worst, err)
                print(f"{case} [one process, {args.team} GPUs] state {k}: rel L2 vs one GPU = {err:.3e}  iterations {s.ch[k]['iters']} (one GPU {int(ref[case + '_iters'][k])})", flush=True)
                assert err < 1e-9, err
        print("SINGLE_PROCESS_MULTI_GPU_OK worst", worst)
        pb.finalize()
        return
    if args.single:
        pb.init(0)
        out = {}
        for case in cases:
            dims, nblk, s = run_case(case, 0, 1)
            for k, st in enumerate(s.states):
                out[f"{case}_{k}"] = st
            out[f"{case}_iters"] = np.array([c["iters"] for c in s.ch])
        np.savez(args.single, **out)
        print("SINGLE_GPU_REFERENCE_WRITTEN", args.single)
        pb.finalize()
        return
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def bcast(ident):
        obj = [ident]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]
    pb.init_distributed(rank, world, local, bcast)
    ref = np.load(args.ref)
    worst = 0.0
    for case in cases:
        for vname, env in VARIANTS.items():
            for k in ("PB200_NO_FUSED", "PB200_NO_TMA", "PB200_NO_REPITCH", "PB200_NO_PIPE", "PB200_NO_BANDFUSE"):
                os.environ.pop(k, None)
            os.environ.update(env)
            dims, nblk, s = run_case(case, rank, world)
            gathered = [None] * world
            dist.all_gather_object(gathered, [st for st in s.states])
            if rank == 0:
                states = assemble(gathered, dims, nblk)
                its = [c["iters"] for c in s.ch]
                for k, x in enumerate(states):
                    r = ref[f"{case}_{k}"]
                    err = np.linalg.norm(x - r) / np.linalg.norm(r)
                    worst = max(worst, err)
                    print(f"{case} [{vname}, p2p {'off' if os.environ.get('PB200_NO_P2P') else 'on'}] state {k}: rel L2 vs one GPU = {err:.3e}  iterations {its[k]} (one GPU {int(ref[case + '_iters'][k])})",
                          flush=True)
                    assert err < 1e-9, err
                    assert abs(its[k] - int(ref[f"{case}_iters"][k])) <= 1
            dist.barrier()
    if rank == 0:
        print("MULTI_GPU_VARIANTS_OK worst", worst)
    dist.destroy_process_group()
    pb.finalize()


if __name__ == "__main__":
    main()
