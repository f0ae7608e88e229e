"""Pin the CPU oracle (oracle/) on every numeric assertion the reference's own tests make for the hot path.

Each test cites the reference test it restates.  These are the only "golden" values the reference holds for this
path (no per-cell fixtures exist upstream).
"""
import numpy as np
import pytest

from oracle import geom
from oracle import penguin_oracle as po


def _circle(c, r, inside=True):
    return geom.LevelSet.ball(c, r, inside)


def test_mesh_centers_nodes_border():
    # test/mesh_test.jl:4-46,64 -- centers = x0 + j h, nodes = x0 + (j + 1/2) h, border-cell count
    m = po.Mesh((5,), (1.0,))
    assert np.allclose(m.nodes[0], [0.1, 0.3, 0.5, 0.7, 0.9, 1.1])
    assert np.allclose(m.centers[0], [0.0, 0.2, 0.4, 0.6, 0.8])
    m2 = po.Mesh((5, 5), (1.0, 1.0))
    assert len(m2.border_cells) == 16
    m3 = po.Mesh((4, 4, 4), (1.0, 1.0, 1.0))
    assert len(m3.border_cells) == 4 ** 3 - 2 ** 3


def test_capacity_circle_area_perimeter():
    # test/capacity_test.jl:6-37 (circle r=0.3 in the unit box, 20x20): area, perimeter vs analytic (rtol there 5-10 %)
    mesh = po.Mesh((20, 20), (1.0, 1.0))
    cap = geom.capacity(mesh, _circle((0.5, 0.5), 0.3))
    assert abs(cap.V.sum() - np.pi * 0.09) < 1e-13
    assert abs(cap.Gamma.sum() - 2 * np.pi * 0.3) < 1e-13


def test_cut_cells_are_gamma_positive_and_centroids_on_circle():
    # test/capacity_test.jl:228-258
    mesh = po.Mesh((30, 30), (1.0, 1.0))
    cap = geom.capacity(mesh, _circle((0.51, 0.51), 0.3))
    cut = np.nonzero(cap.cell_types == -1)[0]
    assert len(cut) > 0
    assert np.array_equal(cut, np.nonzero(cap.Gamma > 0)[0])
    d = np.hypot(cap.C_gamma[cut, 0] - 0.51, cap.C_gamma[cut, 1] - 0.51)
    assert np.all(np.abs(d - 0.3) < 0.05)


def test_capacity_sphere_3d():
    # test/capacity_test.jl:260-286 (sphere r=0.3, 10^3)
    mesh = po.Mesh((10, 10, 10), (1.0, 1.0, 1.0))
    cap = geom.capacity(mesh, geom.LevelSet.ball((0.5, 0.5, 0.5), 0.3))
    assert abs(cap.V.sum() - 4 / 3 * np.pi * 0.027) < 1e-13
    assert abs(cap.Gamma.sum() - 4 * np.pi * 0.09) < 1e-12
    cut = np.nonzero(cap.cell_types == -1)[0]
    d = np.linalg.norm(cap.C_gamma[cut] - 0.5, axis=1)
    assert np.all(np.abs(d - 0.3) < 0.1)


def test_grad_of_ones_vanishes():
    # test/operators_test.jl:4-17 -- grad(ones)[2] == 0 (and on every face with index >= 2)
    mesh = po.Mesh((10, 10), (2.0, 2.0))
    cap = geom.capacity(mesh, _circle((1.0, 1.0), 0.5))
    op = po.DiffusionOps(cap)
    g = po.grad(op, np.ones(2 * op.n))
    assert g[1] == 0.0
    gx = g[:op.n].reshape(mesh.pdims[::-1])
    assert np.max(np.abs(gx[:-1, 1:-1])) < 1e-12


def test_steady_mono_dirichlet():
    # test/solver/diffusion_test.jl:5-26
    mesh = po.Mesh((20, 20), (2.0, 2.0))
    cap = geom.capacity(mesh, _circle((0.5, 0.5), 0.5))
    op = po.DiffusionOps(cap)
    bc1 = po.Dirichlet(1.0)
    bc_b = po.BorderConditions({k: bc1 for k in ("left", "right", "top", "bottom")})
    ph = po.Phase(cap, op, lambda x, y, z: 0.0, lambda x, y, z: 1.0)
    s = po.solve_DiffusionSteadyMono(po.DiffusionSteadyMono(ph, bc_b, po.Dirichlet(1.0)))
    n = op.n
    assert abs(s.x[:n].max() - 1.0) < 1e-2
    assert abs(s.x[n:].max() - 1.0) < 1e-2


def test_steady_diph_max_u1():
    # test/solver/diffusion_test.jl:28-55 -- maximum(u1o) ~ 1.15 (atol 1e-2)
    mesh = po.Mesh((80, 80), (4.0, 4.0))
    ls = _circle((2.0, 2.0), 1.0)
    cap, capc = geom.capacity(mesh, ls), geom.capacity(mesh, ls.flipped())
    op, opc = po.DiffusionOps(cap), po.DiffusionOps(capc)
    bc_b = po.BorderConditions({k: po.Dirichlet(0.0) for k in ("left", "right", "top", "bottom")})
    ic = po.InterfaceConditions(po.ScalarJump(1.0, 1.0, 0.0), po.FluxJump(1.0, 1.0, 0.0))
    f = lambda x, y, z: 1.0
    D = lambda x, y, z: 1.0
    s = po.solve_DiffusionSteadyDiph(po.DiffusionSteadyDiph(po.Phase(cap, op, f, D), po.Phase(capc, opc, f, D), bc_b, ic))
    assert abs(s.x[:op.n].max() - 1.15) < 1e-2


def test_unsteady_mono_be():
    # test/solver/diffusion_test.jl:57-81 -- maximum(ug) ~ 1; loop count 1 + 1 solves
    nx = 20
    mesh = po.Mesh((nx, nx), (4.0, 4.0))
    cap = geom.capacity(mesh, _circle((2.0, 2.0), 1.0))
    op = po.DiffusionOps(cap)
    bc_b = po.BorderConditions({k: po.Dirichlet(0.0) for k in ("left", "right", "top", "bottom")})
    ph = po.Phase(cap, op, lambda x, y, z, t: 0.0, lambda x, y, z: 1.0)
    u0 = np.concatenate([np.zeros(op.n), np.ones(op.n)])
    dt = 0.25 * (4.0 / nx) ** 2
    s = po.DiffusionUnsteadyMono(ph, bc_b, po.Dirichlet(1.0), dt, u0, "BE")
    po.solve_DiffusionUnsteadyMono(s, ph, dt, 0.01, bc_b, po.Dirichlet(1.0), "BE")
    assert len(s.states) == 2 == po.n_solves(dt, 0.01)
    assert abs(s.x[op.n:].max() - 1.0) < 1e-2


def test_loop_counts():
    # SURVEY a19 (IEEE accumulation of `while t < Tend; t += dt`)
    assert po.n_solves(6.25e-4, 0.01) == 17
    assert po.n_solves(0.01, 0.1) == 12


def test_convergence_2d_manufactured():
    # test/convergence_test.jl:30-49 -- u = 1 - r^2, global L2 error < 1e-2
    mesh = po.Mesh((40, 40), (4.0, 4.0))
    cap = geom.capacity(mesh, _circle((2.0, 2.0), 1.0))
    op = po.DiffusionOps(cap)
    bc_b = po.BorderConditions({k: po.Dirichlet(1.0) for k in ("left", "right", "top", "bottom")})
    ph = po.Phase(cap, op, lambda x, y, z: 4.0, lambda x, y, z: 1.0)
    s = po.solve_DiffusionSteadyMono(po.DiffusionSteadyMono(ph, bc_b, po.Dirichlet(0.0)))
    g, _, _, _ = po.check_convergence(lambda x, y: 1.0 - (x - 2) ** 2 - (y - 2) ** 2, s.x, cap)
    assert g < 1e-2


def test_convergence_3d_manufactured():
    # test/convergence_test.jl:51-70 at 20^3 (the reference runs 40^3; same assert)
    mesh = po.Mesh((20, 20, 20), (4.0, 4.0, 4.0))
    cap = geom.capacity(mesh, geom.LevelSet.ball((2.0, 2.0, 2.0), 1.0))
    op = po.DiffusionOps(cap)
    keys = ("left", "right", "top", "bottom", "forward", "backward")
    bc_b = po.BorderConditions({k: po.Dirichlet(1.0) for k in keys})
    ph = po.Phase(cap, op, lambda x, y, z: 6.0, lambda x, y, z: 1.0)
    s = po.solve_DiffusionSteadyMono(po.DiffusionSteadyMono(ph, bc_b, po.Dirichlet(0.0)))
    g, _, _, _ = po.check_convergence(lambda x, y, z: 1.0 - (x - 2) ** 2 - (y - 2) ** 2 - (z - 2) ** 2, s.x, cap)
    assert g < 1e-2


def test_convergence_1d_steady():
    # test/convergence_test.jl:7-28
    mesh = po.Mesh((40,), (4.0,))
    c, r = 0.5, 0.1
    cap = geom.capacity(mesh, geom.LevelSet.ball((c,), r))
    op = po.DiffusionOps(cap)
    bc_b = po.BorderConditions({"top": po.Dirichlet(0.0), "bottom": po.Dirichlet(0.0)})
    ph = po.Phase(cap, op, lambda x, y, z: x, lambda x, y, z: 1.0)
    s = po.solve_DiffusionSteadyMono(po.DiffusionSteadyMono(ph, bc_b, po.Dirichlet(0.0)))
    ua = lambda x: -(x - c) ** 3 / 6 - (c * (x - c) ** 2) / 2 + r ** 2 / 6 * (x - c) + c * r ** 2 / 2
    g, _, _, _ = po.check_convergence(ua, s.x, cap)
    assert g < 1e-2


def test_nobody_known_answer():
    # SURVEY Appendix A.1 (examples/2D/Diffusion/Heat_Nobody.jl:13): full cells, H = 0 on interior rows,
    # G'W!G = h^(N-2) x 5-point Laplacian on interior cells
    mesh = po.Mesh((8, 8), (4.0, 4.0))
    cap = po.nobody_capacity(mesh)
    op = po.DiffusionOps(cap)
    L = (op.G.T @ op.Wdag @ op.G).toarray()
    i = mesh.lin((3, 4))
    assert np.isclose(L[i, i], 4.0)
    for nb in ((2, 4), (4, 4), (3, 3), (3, 5)):
        assert np.isclose(L[i, mesh.lin(nb)], -1.0)
    Hm = op.H.toarray()
    assert np.allclose(Hm[:, i], 0.0)
    # first real cell: diagonal carries h^(2N-2) (1 + h^-N) per direction instead of 2 h^(N-2)
    h = 0.5
    j = mesh.lin((0, 4))
    assert np.isclose(L[j, j], h ** 2 * (1 + h ** -2) + 2.0)


def test_unsteady_diph_1d_erfc():
    # test/convergence_test.jl:100-192 -- 1-D diphasic BE vs the erfc similarity solution
    # (asserts: err1, err2 < 1e-2 on all cells, < 5e-2 on cut cells)
    from scipy.special import erfc
    nx, lx, xint = 100, 8.0, 4.0
    mesh = po.Mesh((nx,), (lx,))
    ls = geom.LevelSet.halfspace(0, xint, True)            # body = x - xint
    cap, capc = geom.capacity(mesh, ls), geom.capacity(mesh, ls.flipped())
    op, opc = po.DiffusionOps(cap), po.DiffusionOps(capc)
    He, D1, D2 = 0.5, 1.0, 1.0
    bc_b = po.BorderConditions({"top": po.Dirichlet(1.0), "bottom": po.Dirichlet(0.0)})
    ic = po.InterfaceConditions(po.ScalarJump(1.0, He, 0.0), po.FluxJump(1.0, 1.0, 0.0))
    f = lambda x, y, z, t: 0.0
    ph1 = po.Phase(cap, op, f, lambda x, y, z: D1)
    ph2 = po.Phase(capc, opc, f, lambda x, y, z: D2)
    n = op.n
    u0 = np.concatenate([np.zeros(n), np.zeros(n), np.ones(n), np.ones(n)])
    dt = 0.5 * (lx / nx) ** 2
    Tend = 0.5
    s = po.DiffusionUnsteadyDiph(ph1, ph2, bc_b, ic, dt, u0, "BE")
    po.solve_DiffusionUnsteadyDiph(s, ph1, ph2, dt, Tend, bc_b, ic, "BE")
    T1 = lambda x: -He / (1 + He * np.sqrt(D1 / D2)) * (erfc((x - xint) / (2 * np.sqrt(D1 * Tend))) - 2)
    T2 = lambda x: -He / (1 + He * np.sqrt(D1 / D2)) * erfc((x - xint) / (2 * np.sqrt(D2 * Tend))) + 1
    u1, u2 = s.x[:n], s.x[2 * n:3 * n]
    g1, f1, c1, _ = po.check_convergence(T1, u1, cap)
    g2, f2, c2, _ = po.check_convergence(T2, u2, capc)
    assert g1 < 1e-2 and g2 < 1e-2 and f1 < 1e-2 and f2 < 1e-2
    assert c1 < 5e-2 and c2 < 5e-2
