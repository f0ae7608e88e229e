"""Run under torchrun (one rank per GPU): slab-partitioned solve vs the oracle on the global problem.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_parity.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import penguin_b200 as pb                      # noqa: E402
from penguin_b200 import slab                  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def bcast(ident):
        obj = [ident]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]
    pb.init_distributed(rank, world, local, bcast)
    worst = 0.0
    # mono3d_dcn: Dirichlet interface (no interface unknowns -> polynomial preconditioner) and CN (explicit part = one folded apply)
    for case in ("diph2d", "mono3d", "mono3d_dcn"):
        if case == "diph2d":
            dims, L = (40, 36), (8.0, 7.2)
            body = pb.Balls([[4.03, 3.67]], [2.0])   # no tangency to a grid line
        else:
            dims, L = (14, 12, int(os.environ.get("PARITY_NZ", "13"))), (4.0, 4.0, 4.0)
            body = -pb.Sphere((2.01, 2.01, 2.01), 1.0)
        mesh = pb.Mesh(dims, L)
        n = int(np.prod([d + 1 for d in dims]))
        h = L[0] / dims[0]
        dt = 0.5 * h * h
        f = lambda x, y, z, t: 0.2 * x + 0.0 * y
        cut = lambda a: np.ascontiguousarray(slab.scatter_owned(a, dims, rank, world))
        if case == "diph2d":
            c1, c2 = pb.Capacity(body, mesh), pb.Capacity(-body, mesh)
            p1, p2 = pb.Phase(c1, pb.DiffusionOps(c1), f, 1.0), pb.Phase(c2, pb.DiffusionOps(c2), f, 2.0)
            ic = pb.InterfaceConditions(pb.ScalarJump(1.0, 2.0, 0.1), pb.FluxJump(1.0, 1.5, 0.05))
            u0g = [np.ones(n), np.ones(n), np.zeros(n), np.zeros(n)]
            u0 = np.concatenate([cut(a) for a in u0g])
            bc = pb.BorderConditions({"left": pb.Dirichlet(0.5), "top": pb.Dirichlet(0.25)})
            s = pb.DiffusionUnsteadyDiph(p1, p2, bc, ic, dt, u0, "BE")
            pb.solve_DiffusionUnsteadyDiph_(s, p1, p2, dt, 2.5 * dt, bc, ic, "CN", reltol=1e-13)
            nblk = 4
        else:
            c1 = pb.Capacity(body, mesh)
            p1 = pb.Phase(c1, pb.DiffusionOps(c1), f, 1.0)
            keys = ("left", "right", "top", "bottom", "forward", "backward")
            bc = pb.BorderConditions({k: pb.Dirichlet(1.0) for k in keys})
            u0 = np.concatenate([cut(np.zeros(n)), cut(np.zeros(n))])
            bci = pb.Dirichlet(0.7) if case == "mono3d_dcn" else pb.Robin(1.0, 0.5, 0.3)
            sch = "CN" if case == "mono3d_dcn" else "BE"
            s = pb.DiffusionUnsteadyMono(p1, bc, bci, dt, u0, "BE")
            pb.solve_DiffusionUnsteadyMono_(s, p1, dt, 2.5 * dt, bc, bci, sch, reltol=1e-13)
            nblk = 2
        nloc = c1.nloc
        if rank == 0:
            print(case, "solver log:", [(c["iters"], c["converged"], "%.1e" % c["rnorm"]) for c in s.ch], flush=True)
        gathered = [None] * world
        dist.all_gather_object(gathered, [st for st in s.states])
        caps = [None] * world
        dist.all_gather_object(caps, dict(V=c1.V, ct=c1.cell_types, W=[w for w in c1.W], A=[w for w in c1.A], B=[w for w in c1.B], Gam=c1.Γ))
        if rank == 0:
            from oracle import geom, penguin_oracle as po
            mo = po.Mesh(dims, L)
            ls = geom.LevelSet.balls(body.centers, body.radii, body.fluid_inside)
            o1 = geom.capacity(mo, ls)
            assert np.array_equal(slab.gather_owned([c["ct"] for c in caps], dims), o1.cell_types)
            bad = False
            def cmp(name, dev, ref, bound):
                nonlocal bad
                e = np.abs(dev - ref)
                i = int(np.argmax(e))
                pd = [d + 1 for d in dims]
                idx = np.unravel_index(i, pd[::-1])[::-1]
                ok = e[i] <= bound
                bad = bad or not ok
                print(f"{case} {name}: max err {e[i]:.3e} (bound {bound:.1e}) at cell {idx} dev {dev[i]!r} ref {ref[i]!r} {'ok' if ok else 'FAIL'}")
            vol = h ** len(dims)
            cmp("V", slab.gather_owned([c["V"] for c in caps], dims), o1.V, 1e-12 * vol)
            cmp("Gamma", slab.gather_owned([c["Gam"] for c in caps], dims), o1.Gamma, 1e-12 * vol / h)
            for d in range(len(dims)):
                cmp(f"A{d}", slab.gather_owned([c["A"][d] for c in caps], dims), o1.A[d], 1e-12 * vol / h)
                cmp(f"B{d}", slab.gather_owned([c["B"][d] for c in caps], dims), o1.B[d], 1e-8 * vol / h)
                cmp(f"W{d}", slab.gather_owned([c["W"][d] for c in caps], dims), o1.W[d], 1e-9 * vol)
            assert not bad
            if case == "diph2d":
                o2 = geom.capacity(mo, ls.flipped())
                q1, q2 = po.Phase(o1, po.DiffusionOps(o1), f, 1.0), po.Phase(o2, po.DiffusionOps(o2), f, 2.0)
                ico = po.InterfaceConditions(po.ScalarJump(1.0, 2.0, 0.1), po.FluxJump(1.0, 1.5, 0.05))
                bco = po.BorderConditions({"left": po.Dirichlet(0.5), "top": po.Dirichlet(0.25)})
                so = po.DiffusionUnsteadyDiph(q1, q2, bco, ico, dt, np.concatenate(u0g), "BE")
                po.solve_DiffusionUnsteadyDiph(so, q1, q2, dt, 2.5 * dt, bco, ico, "CN")
            else:
                q1 = po.Phase(o1, po.DiffusionOps(o1), f, 1.0)
                bco = po.BorderConditions({k: po.Dirichlet(1.0) for k in keys})
                bcio = po.Dirichlet(0.7) if case == "mono3d_dcn" else po.Robin(1.0, 0.5, 0.3)
                so = po.DiffusionUnsteadyMono(q1, bco, bcio, dt, np.zeros(2 * n), "BE")
                po.solve_DiffusionUnsteadyMono(so, q1, dt, 2.5 * dt, bco, bcio, sch)
            assert len(so.states) == len(gathered[0])
            for k, ref in enumerate(so.states):
                # each rank's state is [blk0_local; blk1_local; ...]: reassemble block by block
                blocks = []
                for b in range(nblk):
                    parts = []
                    for r in range(world):
                        st = gathered[r][k]
                        nl = st.shape[0] // nblk
                        parts.append(st[b * nl:(b + 1) * nl])
                    blocks.append(slab.gather_owned(parts, dims))
                x = np.concatenate(blocks)
                err = np.linalg.norm(x - ref) / np.linalg.norm(ref)
                worst = max(worst, err)
                print(f"{case} state {k}: rel L2 vs oracle = {err:.3e}")
                assert err < 1e-9, err
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_PARITY_OK worst", worst)
    dist.destroy_process_group()
    pb.finalize()


if __name__ == "__main__":
    main()
