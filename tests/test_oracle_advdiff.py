"""Oracle pins of the advection-diffusion restatement (oracle/penguin_oracle.py: ConvectionOps, A_/b_*_advdiff after
/root/reference/src/operators.jl:194-209 and src/solver/advectiondiffusion.jl:12-283).  The reference holds no test of this solver family (its
examples under examples/2D/AdvectionDiffusion only plot), so the pins are structural identities and a closed-form solution:
  * zero velocities: the advection-diffusion systems ARE the diffusion systems (diffusion.jl), matrix for matrix;
  * C_d of a constant velocity annihilates ... nothing in general, but sum_d C_d applied to a CONSTANT field is the discrete divergence of the
    face fluxes: zero in the interior of a body-free mesh with a uniform velocity;
  * 1-D steady advection-diffusion without a body, Dirichlet borders: T(x) = (exp(Pe x) - 1) / (exp(Pe) - 1) (second-order central scheme)."""
import numpy as np
import scipy.sparse as sp

from oracle import geom
from oracle import penguin_oracle as po


def test_zero_velocity_is_the_diffusion_system():
    m = po.Mesh((14, 12), (2.0, 1.5))
    cap = geom.capacity(m, geom.LevelSet.ball((1.0, 0.7), 0.45))
    n = m.n
    opc, opd = po.ConvectionOps(cap, [np.zeros(n), np.zeros(n)], np.zeros(2 * n)), po.DiffusionOps(cap)
    for bc in (po.Dirichlet(1.0), po.Robin(1.0, 0.5, 0.2)):
        for scheme in ("BE", "CN"):
            Ac = po.A_mono_unstead_advdiff(opc, cap, 1.3, bc, 0.01, scheme)
            Ad = po.A_mono_unstead_diff(opd, cap, 1.3, bc, 0.01, scheme)
            assert abs(Ac - Ad).max() <= 1e-15 * abs(Ad).max()
            Ti = np.random.default_rng(1).standard_normal(2 * n)
            f = lambda x, y, z, t: 1.0 + x
            bc_ = po.b_mono_unstead_advdiff(opc, f, cap, 1.3, bc, Ti, 0.01, 0.02, scheme)
            bd_ = po.b_mono_unstead_diff(opd, f, 1.3, cap, bc, Ti, 0.01, 0.02, scheme)
            assert np.abs(bc_ - bd_).max() <= 1e-15 * np.abs(bd_).max()
        As_ = po.A_mono_stead_diff(opd, cap, 1.3, bc)
        assert abs(po.A_mono_stead_advdiff(opc, cap, 1.3, bc) - As_).max() <= 1e-15 * abs(As_).max()


def test_uniform_velocity_transports_a_constant_without_change_in_the_interior():
    m = po.Mesh((10, 9), (1.0, 0.9))
    cap = po.nobody_capacity(m)
    n = m.n
    op = po.ConvectionOps(cap, [0.7 * np.ones(n), -0.3 * np.ones(n)], np.zeros(2 * n))
    r = (op.C[0] + op.C[1]) @ np.ones(n)
    px, py = m.pdims
    R = r.reshape(py, px)
    assert np.abs(R[1:py - 2, 1:px - 2]).max() < 1e-15          # away from the first / last rows and columns (padded layout)
    assert all(np.abs(K.diagonal()).max() == 0.0 for K in op.K)                # no interface velocity, no body: K = 0


def test_1d_steady_advection_diffusion_closed_form():
    nx, Pe = 200, 5.0
    m = po.Mesh((nx,), (1.0,))
    cap = po.nobody_capacity(m)
    n = m.n
    op = po.ConvectionOps(cap, [Pe * np.ones(n)], np.zeros(n))
    ph = po.Phase(cap, op, lambda x, y, z: 0.0 * x, 1.0)
    bcb = po.BorderConditions({"bottom": po.Dirichlet(0.0), "top": po.Dirichlet(1.0)})       # (the 1-D keys of the reference, test/convergence_test.jl:13)
    s = po.solve_AdvectionDiffusionSteadyMono(po.AdvectionDiffusionSteadyMono(ph, bcb, po.Dirichlet(0.0)))
    x = np.asarray(m.centers[0])
    T = s.x[:nx]
    exact = (np.exp(Pe * x) - 1.0) / (np.exp(Pe) - 1.0)
    # border rows pin the first and last cell CENTRES (src/solver.jl:379-456), not the domain ends: compare with the exact profile through them
    xa, xb = x[0], x[nx - 1]
    assert abs(T[0]) < 1e-12 and abs(T[nx - 1] - 1.0) < 1e-12
    exact = (np.exp(Pe * x) - np.exp(Pe * xa)) / (np.exp(Pe * xb) - np.exp(Pe * xa))
    assert np.abs(T - exact[:nx]).max() < 2e-4
