"""BASELINE.json's full size (configs[1]: 2048 x 2048 diphasic heat) through size-independent properties -- the oracle's sparse LU does
not finish there in seconds, so the checks are the ones the domain offers: analytic area / perimeter of the interface, complementarity
of the two phases, the symmetry of the problem under x <-> y (NOT under reflections: the geometry grid is the h/2-shifted one of
src/mesh.jl:50, so the circle is off the centre of the cell grid by h/2), exact zeros on removed DOFs, and agreement of two different Krylov methods on the same step."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NX = 2048


@pytest.fixture(scope="module")
def problem():
    import penguin_b200 as pb
    pb.init()
    mesh = pb.Mesh((NX, NX), (8.0, 8.0))
    body = pb.Circle((4.0, 4.0), 2.0)
    c1, c2 = pb.Capacity(body, mesh, compute_centroids=False), pb.Capacity(-body, mesh, compute_centroids=False)
    return pb, mesh, c1, c2


def test_geometry_sums_and_complementarity(problem):
    pb, mesh, c1, c2 = problem
    h = 8.0 / NX
    assert abs(c1.V.sum() - np.pi * 4.0) < 1e-10                      # test/capacity_test.jl:26-36, here to round-off
    assert abs(c1.Γ.sum() - 4.0 * np.pi) < 1e-9
    px = NX + 1
    real = np.zeros((px, px), bool)
    real[:NX, :NX] = True
    real = real.ravel()
    assert np.max(np.abs((c1.V + c2.V)[real] - h * h)) < 1e-16        # the two phases tile every real cell
    assert np.all(c1.V[~real] == 0.0) and np.all(c2.V[~real] == 0.0)  # pad layer (src/capacity.jl:90-92 passes `zero`)
    t1, t2 = c1.cell_types.reshape(px, px), c2.cell_types.reshape(px, px)
    assert np.array_equal(t1 == -1.0, t2 == -1.0)                     # same cut cells
    assert np.array_equal((t1 == 1.0)[:NX, :NX], (t2 == 0.0)[:NX, :NX])
    assert np.array_equal(t1[:NX, :NX], t1[:NX, :NX].T)               # classification symmetric under x <-> y
    assert np.array_equal(c1.Γ > 0, c1.cell_types == -1.0)            # test/capacity_test.jl:256-257
    # divergence theorem per cell
    A0, A1, G = c1.A[0].reshape(px, px), c1.A[1].reshape(px, px), c1.Γ.reshape(px, px)
    assert np.all(np.hypot(A0[:-1, 1:] - A0[:-1, :-1], A1[1:, :-1] - A1[:-1, :-1]) <= G[:-1, :-1] + 1e-13)


def test_states_keep_the_symmetries_and_methods_agree(problem):
    pb, mesh, c1, c2 = problem
    p1, p2 = pb.Phase(c1, pb.DiffusionOps(c1), 0.0, 1.0), pb.Phase(c2, pb.DiffusionOps(c2), 0.0, 1.0)
    n = c1.nloc
    h = 8.0 / NX
    dt = 0.5 * h * h
    ic = pb.InterfaceConditions(pb.ScalarJump(1.0, 1.0, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
    u0 = np.concatenate([np.ones(2 * n), np.zeros(2 * n)])
    out = {}
    for method in ("cg", "bicgstab"):
        s = pb.DiffusionUnsteadyDiph(p1, p2, pb.BorderConditions(), ic, dt, u0, "BE")
        pb.solve_DiffusionUnsteadyDiph_(s, p1, p2, dt, 3.5 * dt, pb.BorderConditions(), ic, "BE", method=method, reltol=1e-12, warm_start=4,
                                        states_stride=100)
        assert all(c["converged"] for c in s.ch) and len(s.ch) == 5
        out[method] = s.x
    x = out["cg"]
    assert np.linalg.norm(x - out["bicgstab"]) <= 1e-9 * np.linalg.norm(x)
    assert np.all(np.isfinite(x))
    px = NX + 1
    for blk in range(4):
        f = x[blk * n:(blk + 1) * n].reshape(px, px)[:NX, :NX]
        sc = np.abs(f).max() + 1e-300
        assert np.max(np.abs(f - f.T)) <= 1e-8 * sc                   # x <-> y
    # removed DOFs are exactly zero (src/solver.jl:186-187): phase-1 bulk outside the circle, phase-2 bulk inside, pad layer
    T1, T2 = x[:n], x[2 * n:3 * n]
    assert np.all(T1[c1.cell_types == 0.0] == 0.0)
    assert np.all(T2[c2.cell_types == 0.0] == 0.0)
    # heat flows from phase 1 (u0 = 1) into phase 2 (u0 = 0): bulk values away from the cut cells stay within the initial bounds
    full1, full2 = c1.cell_types == 1.0, c2.cell_types == 1.0
    assert T1[full1].max() <= 1.0 + 1e-9 and T1[full1].min() >= -1e-9
    assert T2[full2].max() <= 1.0 + 1e-9 and T2[full2].min() >= -1e-9
