"""GPU parity of the "next" rows: Darcy flow + velocity (src/solver/darcy.jl:1-89) and check_convergence
(src/convergence.jl:4-93, reduced on the device state) against the CPU oracle on identical (imported) capacities."""
import numpy as np
import pytest

from oracle import geom
from oracle import penguin_oracle as po
from helpers import import_capacity, rel_l2, to_oracle_borders

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    import penguin_b200
    penguin_b200.init()
    return penguin_b200


def _problem(pb, n, L, ls, f, D):
    mo, mg = po.Mesh(n, L), pb.Mesh(n, L)
    cap_o = geom.capacity(mo, ls)
    op_o = po.DiffusionOps(cap_o)
    cap_g = import_capacity(pb, mg, cap_o)
    return mo, po.Phase(cap_o, op_o, f, D), pb.Phase(cap_g, pb.DiffusionOps(cap_g), f, D)


@pytest.mark.parametrize("dim", [2, 3])
def test_darcy_flow_and_velocity(pb, dim):
    # examples/2D/Darcy-style set-up: pressure 1 / 0 on two opposite borders, impermeable (Neumann 0) body
    n, L = ((20, 20), (2.0, 2.0)) if dim == 2 else ((10, 10, 10), (2.0, 2.0, 2.0))
    ls = geom.LevelSet.ball((1.01,) * dim, 0.4, False)
    f = (lambda x, y, z: 0.0 * x)
    mo, pho, phg = _problem(pb, n, L, ls, f, 1.0)
    bcb = pb.BorderConditions({"bottom": pb.Dirichlet(1.0), "top": pb.Dirichlet(0.0)})
    so = po.solve_DiffusionSteadyMono(po.DiffusionSteadyMono(pho, to_oracle_borders(pb, bcb), po.Robin(1e-3, 1.0, 0.0)))
    sg = pb.solve_DarcyFlow_(pb.DarcyFlow(phg, bcb, pb.Robin(1e-3, 1.0, 0.0)), reltol=1e-13, maxiter=50000)
    assert len(sg.states) >= 1 and rel_l2(sg.x, so.x) < 1e-9
    ug = pb.solve_darcy_velocity(sg, phg, state_i=len(sg.states))
    uo = po.solve_darcy_velocity(so.x, pho.operator, pho.capacity)
    assert ug.shape == uo.shape == (dim * mo.n,)
    assert np.array_equal(np.isnan(ug), np.isnan(uo)), "NaN pattern (cells without a pressure) differs"
    ok = ~np.isnan(uo)
    assert ok.sum() > 0 and rel_l2(ug[ok], uo[ok]) < 1e-8


def test_darcy_unsteady(pb):
    n, L = (16, 16), (2.0, 2.0)
    f = (lambda x, y, z, t: 1.0 + 0.0 * x)
    mo, pho, phg = _problem(pb, n, L, geom.LevelSet.ball((1.0, 1.0), 0.5, False), f, 1.0)
    keys = ("left", "right", "top", "bottom")
    bco = po.BorderConditions({k: po.Dirichlet(0.0) for k in keys})
    bcg = pb.BorderConditions({k: pb.Dirichlet(0.0) for k in keys})
    u0 = np.zeros(2 * mo.n)
    dt = 0.5 * (L[0] / n[0]) ** 2
    so = po.DiffusionUnsteadyMono(pho, bco, po.Dirichlet(0.0), dt, u0, "BE")
    po.solve_DiffusionUnsteadyMono(so, pho, dt, 2.5 * dt, bco, po.Dirichlet(0.0), "BE")
    sg = pb.DarcyFlowUnsteady(phg, bcg, pb.Dirichlet(0.0), dt, u0, "BE")
    pb.solve_DarcyFlowUnsteady_(sg, phg, dt, 2.5 * dt, bcg, pb.Dirichlet(0.0), "BE", reltol=1e-13, maxiter=50000)
    assert len(sg.states) == len(so.states)
    for a, b in zip(sg.states, so.states):
        assert rel_l2(a, b) < 1e-9


@pytest.mark.parametrize("p,relative", [(2, False), (1, False), (3.5, False), (np.inf, False), (2, True), (np.inf, True)])
def test_check_convergence_on_device(pb, p, relative):
    # test/convergence_test.jl:30-49: u = 1 - (x-2)^2 - (y-2)^2 inside the unit circle, f = 4
    mo, pho, phg = _problem(pb, (24, 24), (4.0, 4.0), geom.LevelSet.ball((2.0, 2.0), 1.0), (lambda x, y, z: 4.0 + 0 * x), 1.0)
    bcb = pb.BorderConditions({k: pb.Dirichlet(0.0) for k in ("left", "right", "top", "bottom")})
    so = po.solve_DiffusionSteadyMono(po.DiffusionSteadyMono(pho, to_oracle_borders(pb, bcb), po.Dirichlet(0.0)))
    sg = pb.solve_DiffusionSteadyMono_(pb.DiffusionSteadyMono(phg, bcb, pb.Dirichlet(0.0)), reltol=1e-13, maxiter=50000)
    u = lambda x, y: 1.3 - (x - 2.0) ** 2 - (y - 2.0) ** 2        # (offset: no zeros of u_ana at cell centroids for the relative norms)
    want = po.check_convergence(u, so.x, pho.capacity, p, relative)
    u_ana, u_num, *got = pb.check_convergence(u, sg, phg.capacity, p, relative)
    assert u_ana.shape == (mo.n,) and u_num is not None and rel_l2(u_num, so.x[:mo.n]) < 1e-9
    for g, w in zip(got, want):
        if np.isnan(w) or np.isinf(w):
            assert np.isnan(g) or np.isinf(g)
        else:
            assert abs(g - w) <= 1e-9 * max(abs(w), 1e-300) + 1e-13, (g, w)


def test_reference_darcy_test_jl(pb):
    # test/solver/darcy_test.jl:4-26,56-76 through the device library: circle (0.5, 0.5) r 0.5 in [0, 2]^2, pure Neumann(0) interface,
    # Dirichlet 10 / 20 on :left / :right; the reference's asserts (max p_omega = 20 +- 1e-2, |u| < 1e2) and parity with the oracle
    f = (lambda x, y, z: 0.0 * x)
    mo, pho, phg = _problem(pb, (20, 20), (2.0, 2.0), geom.LevelSet.ball((0.5, 0.5), 0.5), f, (lambda x, y, z: 1.0 + 0.0 * x))
    bcb = pb.BorderConditions({"left": pb.Dirichlet(10.0), "right": pb.Dirichlet(20.0)})
    so = po.solve_DiffusionSteadyMono(po.DiffusionSteadyMono(pho, to_oracle_borders(pb, bcb), po.Neumann(0.0)))
    sg = pb.solve_DarcyFlow_(pb.DarcyFlow(phg, bcb, pb.Neumann(0.0)), reltol=1e-13, maxiter=50000)
    assert abs(sg.x[:mo.n].max() - 20.0) < 1e-2
    assert rel_l2(sg.x, so.x) < 1e-9
    ug = pb.solve_darcy_velocity(sg, phg)
    uo = po.solve_darcy_velocity(so.x, pho.operator, pho.capacity)
    assert np.array_equal(np.isnan(ug), np.isnan(uo))
    ok = ~np.isnan(ug)
    assert np.abs(ug[ok]).max() < 1e2 and rel_l2(ug[ok], uo[ok]) < 1e-8
