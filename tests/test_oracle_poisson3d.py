"""CPU checks behind BASELINE.json configs[4] (steady Poisson, union of random disjoint spheres): the sphere generator of
tools/run_poisson3d.py and the oracle's steady solve (src/solver/diffusion.jl:14-72) on a small instance of that geometry."""
import importlib.util
import os

import numpy as np

from oracle import geom
from oracle import penguin_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _runner():
    spec = importlib.util.spec_from_file_location("run_poisson3d", os.path.join(ROOT, "tools", "run_poisson3d.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_sphere_generator_is_seeded_and_disjoint():
    m = _runner()
    c, r = m.random_spheres()
    c2, r2 = m.random_spheres()
    assert c.shape == (64, 3) and np.array_equal(c, c2) and np.array_equal(r, r2)        # seed 20261018: the same geometry on every rank
    assert r.min() >= 0.1 and r.max() <= 0.3 and c.min() >= 0.5 and c.max() <= 3.5
    d = np.linalg.norm(c[:, None] - c[None], axis=2) - r[:, None] - r[None]
    np.fill_diagonal(d, np.inf)
    assert d.min() >= 0.11                                                            # two cell diagonals at 128^3
    assert (c - r[:, None]).min() > 0.0 and (c + r[:, None]).max() < 4.0              # no sphere touches the box


def small_case():
    """three disjoint spheres in [0,4]^3 on 16^3 cells: the configs[4] system at a size the sparse LU solves in a second"""
    cen = [[1.2, 1.3, 1.25], [2.9, 2.7, 1.4], [2.0, 1.6, 3.0]]
    rad = [0.55, 0.6, 0.5]
    return (16, 16, 16), (4.0, 4.0, 4.0), cen, rad


def oracle_solution(n, L, cen, rad):
    mesh = po.Mesh(n, L)
    cap = geom.capacity(mesh, geom.LevelSet.balls(cen, rad, False))
    ph = po.Phase(cap, po.DiffusionOps(cap), (lambda x, y, z: 1.0 + 0 * x), 1.0)
    keys = ("left", "right", "top", "bottom", "forward", "backward")
    s = po.DiffusionSteadyMono(ph, po.BorderConditions({k: po.Dirichlet(0.0) for k in keys}), po.Dirichlet(0.0))
    return mesh, cap, po.solve_DiffusionSteadyMono(s)


def test_oracle_multi_sphere_poisson_properties():
    n, L, cen, rad = small_case()
    mesh, cap, s = oracle_solution(n, L, cen, rad)
    nn = mesh.n
    T, Tg = s.x[:nn], s.x[nn:]
    V = cap.V
    # volume of the fluid = box minus the three balls (test/capacity_test.jl checks the same identity loosely; here the oracle's geometry is exact)
    assert abs(V.sum() - (64.0 - sum(4.0 / 3.0 * np.pi * r ** 3 for r in rad))) < 1e-9
    # -lap T = 1 with T = 0 on every boundary: maximum principle (T >= 0 in the fluid) and T below the box-only solution's maximum (~0.9 for L = 4)
    fluid = V > 0
    assert T[fluid].min() > -1e-12 and 0.05 < T.max() < 0.9
    assert np.all(T[~fluid] == 0.0)                                  # removed rows: exact zeros (src/solver.jl:59-78)
    cut = cap.Gamma > 0
    assert np.abs(Tg[cut]).max() < 1e-12 and np.all(Tg[~cut] == 0.0)   # interface Dirichlet 0
