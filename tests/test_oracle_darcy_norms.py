"""CPU checks of the oracle's Darcy velocity and check_convergence restatements (src/solver/darcy.jl:26-40, src/convergence.jl:4-93)."""
import numpy as np

from oracle import geom
from oracle import penguin_oracle as po


def _setup(n=12):
    mesh = po.Mesh((n, n), (4.0, 4.0))
    cap = geom.capacity(mesh, geom.LevelSet.ball((2.0, 2.0), 1.0))
    return mesh, cap, po.DiffusionOps(cap)


def test_darcy_velocity_nan_pattern_is_structural():
    # dense restatement of W! (G p_w + H p_g) with IEEE arithmetic on every structurally present coefficient, (i, i) and (i, i-1)
    mesh, cap, op = _setup()
    n, pd = mesh.n, mesh.pdims
    x = np.random.default_rng(0).standard_normal(2 * n)
    u = po.solve_darcy_velocity(x, op, cap)
    pw, pg = x[:n].copy(), x[n:].copy()
    ct = cap.cell_types
    pw[ct == 0] = np.nan; pg[ct == 0] = np.nan; pg[ct == 1] = np.nan
    stride = [1, pd[0]]
    ref = np.zeros(2 * n)
    with np.errstate(invalid="ignore"):
        for d in range(2):
            A, B, Wd = cap.A[d], cap.B[d], op.Wdag_diag[d * n:(d + 1) * n]
            for l in range(n):
                c = (l % pd[0], l // pd[0])
                ei = 1.0 if c[d] < pd[d] - 1 else 0.0
                q = ei * (B[l] * pw[l] + (A[l] - B[l]) * pg[l])
                if c[d] > 0:
                    bm = B[l - stride[d]]
                    q -= bm * pw[l - stride[d]] + (A[l] - bm) * pg[l - stride[d]]
                ref[d * n + l] = -(Wd[l] * q)
    assert np.array_equal(np.isnan(u), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert 0 < ok.sum() < ok.size and np.allclose(u[ok], ref[ok], rtol=1e-12, atol=1e-14)


def test_velocity_of_constant_pressure_is_zero_where_defined():
    # test/operators_test.jl:13-16 (grad of ones = 0) carried over to the Darcy velocity
    mesh, cap, op = _setup()
    u = po.solve_darcy_velocity(np.ones(2 * mesh.n), op, cap)
    ok = ~np.isnan(u)
    assert ok.sum() > 0 and np.abs(u[ok]).max() < 1e-12


def test_check_convergence_norms():
    mesh, cap, op = _setup(16)
    ct, V = cap.cell_types, cap.V
    u = lambda x, y: 1.0 + x + 2.0 * y
    ua = u(cap.C_omega[:, 0], cap.C_omega[:, 1])
    num = ua - 0.5                                     # constant error 0.5 everywhere
    g, fu, cu, em = po.check_convergence(u, num, cap, 2)
    tot = V.sum()
    assert np.isclose(fu, np.sqrt(0.25 * V[ct == 1].sum() / tot)) and np.isclose(cu, np.sqrt(0.25 * V[ct == -1].sum() / tot))
    assert np.isclose(g, np.sqrt(0.25 * V[ct != 0].sum() / tot)) and em == 0.0       # empty cells carry no volume
    assert po.check_convergence(u, num, cap, np.inf)[:3] == (0.5, 0.5, 0.5)
    # same through the 2n state-vector form (solver.x[1:end/2])
    assert po.check_convergence(u, np.concatenate([num, np.zeros_like(num)]), cap, 2)[0] == g
    rel = po.check_convergence(u, num, cap, 1, relative=True)
    assert np.isclose(rel[1], (np.abs(0.5 / ua[ct == 1]) * V[ct == 1]).sum() / tot)


# ---- the reference's own asserts for this stage (test/solver/darcy_test.jl) ------------------------------------------------------
def _darcy_reference_case():
    mesh = po.Mesh((20, 20), (2.0, 2.0))
    cap = geom.capacity(mesh, geom.LevelSet.ball((0.5, 0.5), 0.5))       # LS < 0 inside the circle: darcy_test.jl:10
    op = po.DiffusionOps(cap)
    bcb = po.BorderConditions({"left": po.Dirichlet(10.0), "right": po.Dirichlet(20.0)})
    return mesh, cap, op, bcb


def test_reference_darcy_steady_assert():
    # darcy_test.jl:4-26: maximum(uo) ~ 20 +- 1e-2
    mesh, cap, op, bcb = _darcy_reference_case()
    ph = po.Phase(cap, op, (lambda x, y, z: 0.0 * x), (lambda x, y, z: 1.0 + 0.0 * x))
    s = po.solve_DiffusionSteadyMono(po.DiffusionSteadyMono(ph, bcb, po.Neumann(0.0)))
    assert abs(s.x[:mesh.n].max() - 20.0) < 1e-2


def test_reference_darcy_unsteady_assert():
    # darcy_test.jl:28-54: BE, dt = 0.1 (lx/nx)^2, Tend = 0.2 (shortened: the assert holds from the first state on, border rows carry 20)
    mesh, cap, op, bcb = _darcy_reference_case()
    ph = po.Phase(cap, op, (lambda x, y, z, t: 0.0 * x), (lambda x, y, z: 1.0 + 0.0 * x))
    dt = 0.1 * (2.0 / 20) ** 2
    u0 = np.full(2 * mesh.n, 10.0)
    s = po.DiffusionUnsteadyMono(ph, bcb, po.Neumann(0.0), dt, u0, "BE")
    po.solve_DiffusionUnsteadyMono(s, ph, dt, 0.2, bcb, po.Neumann(0.0), "BE", max_steps=12)
    assert len(s.states) == 13
    assert abs(s.states[-1][:mesh.n].max() - 20.0) < 1e-2


def test_reference_darcy_velocity_assert():
    # darcy_test.jl:56-76: the non-NaN velocities stay below 1e2
    mesh, cap, op, bcb = _darcy_reference_case()
    ph = po.Phase(cap, op, (lambda x, y, z: 0.0 * x), (lambda x, y, z: 1.0 + 0.0 * x))
    s = po.solve_DiffusionSteadyMono(po.DiffusionSteadyMono(ph, bcb, po.Neumann(0.0)))
    u = po.solve_darcy_velocity(s.x, op, cap)
    ok = ~np.isnan(u)
    assert 0 < ok.sum() < u.size and np.abs(u[ok]).max() < 1e2
