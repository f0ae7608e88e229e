"""Slab partition of the padded grid over the ranks (host mirror of make_grid in csrc/common.cuh).

The slowest grid dimension is split into contiguous plane ranges; every per-cell host array handed to / returned by the C ABI
under pb200_init_dist holds the rank's OWNED planes only.  These helpers cut global arrays into slabs and reassemble them;
they are pure NumPy so the N > 1 host logic can be tested with the gloo backend on CPU.
"""
import numpy as np


def slab_range(nplanes, rank, nranks):
    """owned plane range [k0, k1) of `rank` -- same arithmetic as make_grid (csrc/common.cuh)"""
    if nplanes < nranks:
        raise ValueError("fewer planes than ranks")
    base, rem = divmod(nplanes, nranks)
    k0 = rank * base + min(rank, rem)
    return k0, k0 + base + (1 if rank < rem else 0)


def plane_size(dims):
    """cells per plane of the slowest dimension of a mesh with `dims` real cells per direction (padded: n_i + 1)"""
    p = 1
    for d in dims[:-1]:
        p *= d + 1
    return p


def scatter_owned(a, dims, rank, nranks):
    """global padded per-cell array (x fastest) -> the rank's owned slab (a view)"""
    k0, k1 = slab_range(dims[-1] + 1, rank, nranks)
    p = plane_size(dims)
    return np.asarray(a)[k0 * p:k1 * p]


def gather_owned(parts, dims):
    """list of owned slabs in rank order -> global padded array"""
    out = np.concatenate([np.asarray(p) for p in parts])
    n = plane_size(dims) * (dims[-1] + 1)
    if out.shape[0] != n:
        raise ValueError(f"slabs hold {out.shape[0]} cells, the grid has {n}")
    return out


def with_ghosts(owned, lower, upper, dims):
    """owned slab + one ghost plane per side (None -> zeros, i.e. outside the global domain): the device layout"""
    p = plane_size(dims)
    z = np.zeros(p, dtype=np.asarray(owned).dtype)
    return np.concatenate([z if lower is None else lower, owned, z if upper is None else upper])
