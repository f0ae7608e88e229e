"""Oracle pins on the known-answer asserts of the reference's OWN benchmark suite for the scalar diffusion path (BenchPhaseFlow/problems/scalar/*.jl, run by
.github/workflows/benchphaseflow.yml): every script ends in a @testset -- a fitted convergence order above 1 for the Poisson studies, a mass drift below 1e-10
for the Neumann heat problem.  The oracle (geometry restatement + block systems + direct solve) must satisfy the same asserts on the same inputs; the fit is
BenchPhaseFlow/utils/convergence.jl:compute_orders (least squares of log err against log h over the last three meshes, rounded to one digit)."""
import numpy as np

from oracle import geom
from oracle import penguin_oracle as po


def fitted_order(h, err, use_last=3):
    """BenchPhaseFlow/utils/convergence.jl:16-56 (`orders.all`)"""
    h, err = np.asarray(h, float), np.asarray(err, float)
    m = err > 0
    lh, le = np.log(h[m])[-use_last:], np.log(err[m])[-use_last:]
    return round(float(np.polyfit(lh, le, 1)[0]), 1)


def poisson_errors(nlist, L, center, radius, f, u_ana):
    """the study loop of Scalar_{2,3}D_Diffusion_Poisson_Dirichlet.jl: fluid inside the ball, Dirichlet 0 on the interface and on the border keys the script names"""
    N = len(L)
    hs, errs = [], []
    for n in nlist:
        mesh = po.Mesh((n,) * N, L)
        cap = geom.capacity(mesh, geom.LevelSet.ball(center, radius))
        ph = po.Phase(cap, po.DiffusionOps(cap), f, 1.0)
        # the 3-D script names :front / :back, which classify_boundary_cell_fast never returns (src/solver.jl:379-409): only the four 2-D keys act
        bc_b = po.BorderConditions({k: po.Dirichlet(0.0) for k in ("left", "right", "top", "bottom")})
        s = po.solve_DiffusionSteadyMono(po.DiffusionSteadyMono(ph, bc_b, po.Dirichlet(0.0)))
        all_err, full_err, cut_err, empty_err = po.check_convergence(u_ana, s.x, cap, 2, False)
        hs.append(min(L[d] / n for d in range(N)))
        errs.append(all_err)
    return hs, errs


def test_poisson_2d_dirichlet_order():
    # Scalar_2D_Diffusion_Poisson_Dirichlet.jl:168-172,207-213: radius 1, centre (2, 2), f = 4, u = 1 - r^2; meshes 4 .. 128; `@test vofi_orders.all > 1.0`
    c = (2.0, 2.0)
    hs, errs = poisson_errors([4, 8, 16, 32, 64, 128], (4.0, 4.0), c, 1.0, (lambda x, y, z: 4.0 + 0 * x), lambda x, y: 1.0 - (x - c[0]) ** 2 - (y - c[1]) ** 2)
    assert fitted_order(hs, errs) > 1.0
    assert hs[0] > hs[-1] and min(errs) < max(errs)


def test_poisson_3d_dirichlet_order():
    # Scalar_3D_Diffusion_Poisson_Dirichlet.jl:111-117,131-137: radius 0.5, centre (2, 2, 2), f = 1, u = (R^2 - r^2) / 6; meshes 8 .. 64 (here 8 .. 32:
    # the last three meshes of the fit are then 8, 16, 32 instead of 16, 32, 64 -- coarser, i.e. the harder side of the same assert)
    c, R = (2.0, 2.0, 2.0), 0.5
    u = lambda x, y, z: (R * R - ((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2)) / 6.0
    hs, errs = poisson_errors([8, 16, 32], (4.0, 4.0, 4.0), c, R, (lambda x, y, z: 1.0 + 0 * x), u)
    assert fitted_order(hs, errs) > 1.0
    assert min(errs) < max(errs)


def test_neumann_mass_conservation():
    # Scalar_2D_Diffusion_Heat_NeumannMass.jl:38-84,100-118: circle r = 0.25 in the unit box, homogeneous Neumann on the interface (and Neumann border keys:
    # no-ops in 2-D), u0 = 1, CN, dt = 0.25 h^2, T_end = 0.1; mass = V . T_omega per state; `@test results.data.drift < 1e-10`  (32^2 here, 64^2 in the script)
    nx = 32
    mesh = po.Mesh((nx, nx), (1.0, 1.0))
    cap = geom.capacity(mesh, geom.LevelSet.ball((0.5, 0.5), 0.25))
    ph = po.Phase(cap, po.DiffusionOps(cap), (lambda x, y, z, t: 0.0 * x), 1.0)
    bc_b = po.BorderConditions({k: po.Neumann(0.0) for k in ("left", "right", "top", "bottom")})
    n = mesh.n
    u0 = np.ones(2 * n)
    dt = 0.25 * (1.0 / nx) ** 2
    s = po.DiffusionUnsteadyMono(ph, bc_b, po.Neumann(0.0), dt, u0, "CN")
    po.solve_DiffusionUnsteadyMono(s, ph, dt, 0.1, bc_b, po.Neumann(0.0), "CN")
    masses = np.array([cap.V @ st[:n] for st in s.states])
    assert len(masses) == po.n_solves(dt, 0.1)
    assert np.max(np.abs(masses - masses[0])) < 1e-10
    assert abs(masses[0] - np.pi * 0.25 ** 2) < 1e-12            # V . 1 = area of the disc
