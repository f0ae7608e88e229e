"""N > 1 host logic on CPU: world_size-2 gloo processes cut a global problem into y/z-slabs (penguin_b200.slab, the mirror of
make_grid in csrc/common.cuh), exchange one ghost plane per side the way halo_exchange does (send first/last owned plane,
receive into the ghost planes), apply the folded stencil on their slab and sum their dot-product partials with an
allreduce -- the result must equal the single-process apply on the global grid."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _stencil(x, off, dims):
    """y = x + sum_d off_d[l] x[l - s_d] + off_d[l + s_d] x[l + s_d] on a padded array with ghost planes (NumPy restatement of
    kf_apply_dense, csrc/fold.cuh) -- x, off_d include one ghost plane per side"""
    pd = [d + 1 for d in dims]
    strides = [1, pd[0], pd[0] * pd[1] if len(pd) > 2 else None][:len(pd)]
    p = 1
    for d in pd[:-1]:
        p *= d
    n = x.shape[0]
    y = x.copy()
    for d, s in enumerate(strides):
        lo = np.zeros(n); lo[s:] = x[:-s]
        hi = np.zeros(n); hi[:-s] = x[s:]
        offhi = np.zeros(n); offhi[:-s] = off[d][s:]
        y += off[d] * lo + offhi * hi
    return y[p:-p]


def _worker(rank, world, port, dims, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import penguin_b200.slab as slab
    rng = np.random.default_rng(7)
    pd = [d + 1 for d in dims]
    n = int(np.prod(pd))
    xg = rng.standard_normal(n)
    offg = [rng.standard_normal(n) * 0.1 for _ in dims]
    # global pad layer is zero, like every reference array
    idx = np.arange(n)
    c0 = idx % pd[0]
    mask = c0 < dims[0]
    if len(pd) > 1:
        mask &= (idx // pd[0]) % pd[1] < dims[1]
    if len(pd) > 2:
        mask &= idx // (pd[0] * pd[1]) < dims[2]
    xg *= mask
    offg = [o * mask for o in offg]
    p = slab.plane_size(dims)

    def exchange(owned):
        """halo_exchange (csrc/common.cuh): first owned plane -> rank-1's upper ghost, last owned plane -> rank+1's lower ghost"""
        lower = upper = None
        reqs = []
        t_lo, t_hi = torch.zeros(p, dtype=torch.float64), torch.zeros(p, dtype=torch.float64)
        if rank > 0:
            reqs.append(dist.isend(torch.from_numpy(owned[:p].copy()), rank - 1))
            reqs.append(dist.irecv(t_lo, rank - 1))
        if rank < world - 1:
            reqs.append(dist.isend(torch.from_numpy(owned[-p:].copy()), rank + 1))
            reqs.append(dist.irecv(t_hi, rank + 1))
        for r in reqs:
            r.wait()
        if rank > 0:
            lower = t_lo.numpy()
        if rank < world - 1:
            upper = t_hi.numpy()
        return slab.with_ghosts(owned, lower, upper, dims)

    xl = exchange(slab.scatter_owned(xg, dims, rank, world))
    offl = [exchange(slab.scatter_owned(o, dims, rank, world)) for o in offg]
    yl = _stencil(xl, offl, dims)
    part = torch.tensor([float(np.dot(xl[p:-p], yl))], dtype=torch.float64)
    dist.all_reduce(part)
    ys = [torch.zeros(1)] * world
    gathered = [None] * world
    dist.all_gather_object(gathered, yl)
    if rank == 0:
        yg = slab.gather_owned(gathered, dims)
        zero = np.zeros(p)
        yref = _stencil(np.concatenate([zero, xg, zero]), [np.concatenate([zero, o, zero]) for o in offg], dims)
        q.put((float(np.max(np.abs(yg - yref))), float(part.item()), float(np.dot(xg, yref))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("dims", [(9, 7), (5, 4, 6)])
def test_slab_halo_stencil_matches_global(dims):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, dims, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    err, dot_par, dot_ref = q.get(timeout=120)
    for pr in procs:
        pr.join(60)
        assert pr.exitcode == 0
    assert err == 0.0
    assert abs(dot_par - dot_ref) <= 1e-12 * abs(dot_ref)


def test_slab_ranges_cover_the_grid():
    import penguin_b200.slab as slab
    for nplanes in (2, 3, 17, 129, 1025):
        for nr in (1, 2, 3, 4, 8):
            if nplanes < nr:
                continue
            r = [slab.slab_range(nplanes, k, nr) for k in range(nr)]
            assert r[0][0] == 0 and r[-1][1] == nplanes
            assert all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
