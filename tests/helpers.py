"""Shared helpers of the parity tests: build the same problem on the oracle and on the CUDA library."""
import numpy as np

from oracle import geom
from oracle import penguin_oracle as po


def rel_l2(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


def oracle_levelset(body):
    """product body descriptor -> oracle level-set descriptor (same numbers, independent code)."""
    if body.kind == 0:
        return geom.LevelSet.balls(body.centers, body.radii, body.fluid_inside)
    return geom.LevelSet.halfspace(body.dim, body.c, body.fluid_inside)


def import_capacity(pb, mesh, cap_o):
    """oracle capacity -> device capacity through pb200_capacity_import (stage-wise parity: solver checked alone)."""
    return pb.Capacity.from_arrays(mesh, cap_o.V, cap_o.Gamma, cap_o.cell_types, cap_o.A, cap_o.B, cap_o.W, cap_o.C_omega, cap_o.C_gamma)


def to_oracle_bc(pb, bc):
    if isinstance(bc, pb.Dirichlet):
        return po.Dirichlet(bc.value)
    if isinstance(bc, pb.Neumann):
        return po.Neumann(bc.value)
    if isinstance(bc, pb.Robin):
        return po.Robin(bc.α, bc.β, bc.value)
    raise TypeError(bc)


def to_oracle_borders(pb, bc_b):
    return po.BorderConditions({k: to_oracle_bc(pb, v) for k, v in bc_b.borders.items()})
