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


def test_heat_2d_dirichlet_order():
    # Scalar_2D_Diffusion_Heat_Dirichlet.jl:22-52,70-100,135-147,158-160: disc r = 1 centre (2.01, 2.01), interface Dirichlet 1, borders Dirichlet 0, u0 = 0,
    # BE constructor then CN loop, dt = 0.5 h^2, T_end = 0.1, against the Bessel series at t = 0.1; meshes 4 .. 64; `@test orders.all > 1.0`
    from scipy.special import j0, j1, jn_zeros
    c, R, t_end = (2.01, 2.01), 1.0, 0.1
    al = jn_zeros(0, 200)

    def u_ana(x, y):
        r = np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2)
        s = np.sum(np.exp(-al[None, :] ** 2 * t_end) * j0(al[None, :] * (r[:, None] / R)) / (al[None, :] * j1(al[None, :])), axis=1)
        return np.where(r >= R, 0.0, 1.0 - 2.0 * s)
    hs, errs = [], []
    for nx in (4, 8, 16, 32, 64):
        mesh = po.Mesh((nx, nx), (4.0, 4.0))
        cap = geom.capacity(mesh, geom.LevelSet.ball(c, R))
        ph = po.Phase(cap, po.DiffusionOps(cap), (lambda x, y, z, t: 0.0 * x), 1.0)
        bc_b = po.BorderConditions({k: po.Dirichlet(0.0) for k in ("left", "right", "top", "bottom")})
        dt = 0.5 * (4.0 / nx) ** 2
        s = po.DiffusionUnsteadyMono(ph, bc_b, po.Dirichlet(1.0), dt, np.zeros(2 * mesh.n), "BE")
        po.solve_DiffusionUnsteadyMono(s, ph, dt, t_end, bc_b, po.Dirichlet(1.0), "CN")
        hs.append(4.0 / nx)
        errs.append(po.check_convergence(u_ana, s.x, cap, 2, False)[0])
    assert fitted_order(hs, errs) > 1.0
    assert min(errs) < max(errs)


def heat2ph2d_exact(center, R, t_end, Dg=1.0, Dl=1.0, He=1.0, cg0=1.0, nq=3000):
    """diphasic/Heat_2ph_2D.jl:39-93: the Bessel-integral solution of the circular two-phase heat problem at t_end (phase 1 inside, phase 2 outside), with the
    script's cut-off Umax = 5 / sqrt(Dg t_end); Gauss-Legendre in u for all points at once instead of one adaptive quadgk per point"""
    from scipy.special import j0, j1, y0, y1
    umax = 5.0 / np.sqrt(Dg * t_end)
    xg, wg = np.polynomial.legendre.leggauss(nq)
    u, w = 0.5 * umax * (xg + 1.0), 0.5 * umax * wg
    D = np.sqrt(Dg / Dl)
    phi = Dg * np.sqrt(Dl) * j1(u * R) * y0(D * u * R) - He * Dl * np.sqrt(Dg) * j0(u * R) * y1(D * u * R)
    psi = Dg * np.sqrt(Dl) * j1(u * R) * j0(D * u * R) - He * Dl * np.sqrt(Dg) * j0(u * R) * j1(D * u * R)
    den = phi ** 2 + psi ** 2
    e = np.exp(-Dg * u ** 2 * t_end) * j1(u * R)

    def u1(x, y):
        r = np.hypot(x - center[0], y - center[1])
        val = (j0(u[None, :] * r[:, None]) * (e / (u ** 2 * den))[None, :]) @ w
        return np.where(r >= R, 0.0, 4.0 * cg0 * Dg * Dl ** 2 * He / (np.pi ** 2 * R) * val)

    def u2(x, y):
        r = np.hypot(x - center[0], y - center[1])
        rr = np.maximum(r, 1e-12)
        contrib = j0(D * u[None, :] * rr[:, None]) * phi[None, :] - y0(D * u[None, :] * rr[:, None]) * psi[None, :]
        val = (contrib * (e / (u * den))[None, :]) @ w
        return np.where(r < R, 0.0, 2.0 * cg0 * Dg * np.sqrt(Dl) * He / np.pi * val)
    return u1, u2


def test_heat_2ph_2d_against_the_bessel_solution():
    # diphasic/Heat_2ph_2D.jl:94-160,216-231 -- the problem BASELINE.json configs[1] is quoted on (benchmark/Heat_2ph_2D.jl): [0, 8]^2, circle r = 2 centre (4, 4),
    # ScalarJump(1, He = 1, 0), FluxJump(1, 1, 0), u0 = [1, 1, 0, 0], no border conditions, BE constructor then CN, dt = 0.5 h^2, T_end = 0.1;
    # error = max over the phases of the volume-weighted L2 norm (src/convergence.jl:114-237).  The script asserts a fit that is not NaN and errors that
    # vary with the mesh; the oracle shows more: the errors FALL with h, at an order above 1 (meshes 8 .. 64 of the script's 4 .. 128)
    c, R, t_end = (4.0, 4.0), 2.0, 0.1
    u1, u2 = heat2ph2d_exact(c, R, t_end)
    hs, errs = [], []
    for nx in (8, 16, 32, 64):
        mesh = po.Mesh((nx, nx), (8.0, 8.0))
        ls = geom.LevelSet.ball(c, R)
        c1, c2 = geom.capacity(mesh, ls), geom.capacity(mesh, ls.flipped())
        f = lambda x, y, z, t: 0.0 * x
        p1, p2 = po.Phase(c1, po.DiffusionOps(c1), f, 1.0), po.Phase(c2, po.DiffusionOps(c2), f, 1.0)
        ic = po.InterfaceConditions(po.ScalarJump(1.0, 1.0, 0.0), po.FluxJump(1.0, 1.0, 0.0))
        n = mesh.n
        u0 = np.concatenate([np.ones(2 * n), np.zeros(2 * n)])
        dt = 0.5 * (8.0 / nx) ** 2
        s = po.DiffusionUnsteadyDiph(p1, p2, po.BorderConditions(), ic, dt, u0, "BE")
        po.solve_DiffusionUnsteadyDiph(s, p1, p2, dt, t_end, po.BorderConditions(), ic, "CN")
        x = s.states[-1]
        e1 = po.check_convergence(u1, x[:n], c1, 2, False)[0]
        e2 = po.check_convergence(u2, x[2 * n:3 * n], c2, 2, False)[0]
        hs.append(8.0 / nx)
        errs.append(max(e1, e2))
    order = fitted_order(hs, errs)
    assert not np.isnan(order) and min(errs) < max(errs)           # the script's asserts
    assert all(b < a for a, b in zip(errs, errs[1:])) and order > 1.0, (errs, order)


def test_poisson_1d_dirichlet_order():
    # Scalar_1D_Diffusion_Poisson_Dirichlet.jl:37-49,96-110,116-118: interval |x - 0.5| < 0.11 of [0, 1], f = x, Dirichlet 0 on the interface points and on
    # :top / :bottom, u = -(x-c)^3/6 - c (x-c)^2/2 + R^2 (x-c)/6 + c R^2/2; meshes 2 .. 256; `@test orders.all > 1.0`
    c, R = 0.5, 0.11
    u = lambda x: -(x - c) ** 3 / 6 - c * (x - c) ** 2 / 2 + R * R / 6 * (x - c) + c * R * R / 2
    hs, errs = [], []
    for nx in (2, 4, 8, 16, 32, 64, 128, 256):
        mesh = po.Mesh((nx,), (1.0,))
        cap = geom.capacity(mesh, geom.LevelSet.ball((c,), R))
        ph = po.Phase(cap, po.DiffusionOps(cap), (lambda x, y, z: x), 1.0)
        bc_b = po.BorderConditions({"top": po.Dirichlet(0.0), "bottom": po.Dirichlet(0.0)})
        s = po.solve_DiffusionSteadyMono(po.DiffusionSteadyMono(ph, bc_b, po.Dirichlet(0.0)))
        hs.append(1.0 / nx)
        errs.append(po.check_convergence(u, s.x, cap, 2, False)[0])
    assert fitted_order(hs, errs) > 1.0
    assert min(errs) < max(errs)


def test_heat_3d_dirichlet_order():
    # Scalar_3D_Diffusion_Heat_Dirichlet.jl:21-59,83-111,151-161,171-173: ball r = 1 centre (2, 2, 2) in [0, 4]^3, interface Dirichlet 1 (border keys: Dirichlet 1,
    # :front / :back never match), u0 = 0, BE throughout, dt = 0.25 h^2, T_end = 0.1, against the spherical sine series; meshes 8, 12, 16, 20; `@test orders.all > 1.0`
    c, R, t_end = (2.0, 2.0, 2.0), 1.0, 0.1
    nn = np.arange(1, 201)
    lam = nn * np.pi / R

    def u_ana(x, y, z):
        r = np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2)
        rr = np.maximum(r, 1e-12)
        s = np.sum(((-1.0) ** (nn + 1) / nn)[None, :] * np.sin(lam[None, :] * rr[:, None]) * np.exp(-lam ** 2 * t_end)[None, :], axis=1)
        return np.where(r >= R, 1.0, 1.0 - (2.0 * R / (np.pi * rr)) * s)
    hs, errs = [], []
    for nx in (8, 12, 16, 20):
        mesh = po.Mesh((nx,) * 3, (4.0,) * 3)
        cap = geom.capacity(mesh, geom.LevelSet.ball(c, R))
        ph = po.Phase(cap, po.DiffusionOps(cap), (lambda x, y, z, t: 0.0 * x), 1.0)
        bc_b = po.BorderConditions({k: po.Dirichlet(1.0) for k in ("left", "right", "top", "bottom")})
        dt = 0.25 * (4.0 / nx) ** 2
        s = po.DiffusionUnsteadyMono(ph, bc_b, po.Dirichlet(1.0), dt, np.zeros(2 * mesh.n), "BE")
        po.solve_DiffusionUnsteadyMono(s, ph, dt, t_end, bc_b, po.Dirichlet(1.0), "BE")
        hs.append(4.0 / nx)
        errs.append(po.check_convergence(u_ana, s.x, cap, 2, False)[0])
    assert fitted_order(hs, errs) > 1.0, (hs, errs)
    assert min(errs) < max(errs)


def test_heat_2d_robin_order():
    # Scalar_2D_Diffusion_Heat_Robin.jl:20-53,76-100,140-148,159-160: disc r = 1 centre (2.01, 2.01), interface Robin(1, 1, 1), borders Dirichlet 0, u0 = 0,
    # BE constructor then CN, dt = 0.25 h^2, T_end = 0.1, against the Robin Bessel series (roots of a J1(a) - k R J0(a)); meshes 4 .. 128 (here .. 64); order > 1
    from scipy.optimize import brentq
    from scipy.special import j0, j1
    c, R, t_end, k = (2.01, 2.01), 1.0, 0.1, 1.0
    eq = lambda a: a * j1(a) - k * R * j0(a)
    al = np.array([brentq(eq, max((m - 0.25 - 0.5) * np.pi, 1e-6), (m - 0.25 + 0.5) * np.pi) for m in range(1, 201)])
    An = 2.0 * k * R / ((k * k * R * R + al ** 2) * j0(al))

    def u_ana(x, y):
        r = np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2)
        s = np.sum(An[None, :] * np.exp(-al[None, :] ** 2 * t_end / R ** 2) * j0(al[None, :] * (r[:, None] / R)), axis=1)
        return np.where(r >= R, 0.0, 1.0 - s)
    hs, errs = [], []
    for nx in (4, 8, 16, 32, 64):
        mesh = po.Mesh((nx, nx), (4.0, 4.0))
        cap = geom.capacity(mesh, geom.LevelSet.ball(c, R))
        ph = po.Phase(cap, po.DiffusionOps(cap), (lambda x, y, z, t: 0.0 * x), 1.0)
        bc_b = po.BorderConditions({kk: po.Dirichlet(0.0) for kk in ("left", "right", "top", "bottom")})
        bci = po.Robin(1.0, 1.0, 1.0)
        dt = 0.25 * (4.0 / nx) ** 2
        s = po.DiffusionUnsteadyMono(ph, bc_b, bci, dt, np.zeros(2 * mesh.n), "BE")
        po.solve_DiffusionUnsteadyMono(s, ph, dt, t_end, bc_b, bci, "CN")
        hs.append(4.0 / nx)
        errs.append(po.check_convergence(u_ana, s.x, cap, 2, False)[0])
    assert fitted_order(hs, errs) > 1.0, (hs, errs)
    assert min(errs) < max(errs)


def test_heat_1d_dirichlet_order():
    # Scalar_1D_Diffusion_Heat_Dirichlet.jl:20-47,67-87,126-134,146-147: slab |x - 0.5| < 0.25 of [0, 1], interface Dirichlet 0, u0 = 1, CN constructor and CN loop,
    # dt = 0.5 h^2, T_end = 0.1, against the sine series; the script's border keys :left / :right never match a 1-D cell (src/solver.jl:379-409); meshes 2 .. 32
    c, R, t_end = 0.5, 0.25, 0.1
    nn = 2 * np.arange(400) + 1
    lam = nn * np.pi / (2 * R)

    def u_ana(x):
        xi = x - (c - R)
        th = 4.0 / np.pi * np.sum(np.sin(lam[None, :] * xi[:, None]) / nn[None, :] * np.exp(-lam ** 2 * t_end)[None, :], axis=1)
        return np.where((x < c - R) | (x > c + R), 1.0, th)
    hs, errs = [], []
    for nx in (2, 4, 8, 16, 32):
        mesh = po.Mesh((nx,), (1.0,))
        cap = geom.capacity(mesh, geom.LevelSet.ball((c,), R))
        ph = po.Phase(cap, po.DiffusionOps(cap), (lambda x, y, z, t: 0.0 * x), 1.0)
        bc_b = po.BorderConditions({"left": po.Dirichlet(0.0), "right": po.Dirichlet(0.0)})
        dt = 0.5 * (1.0 / nx) ** 2
        u0 = np.concatenate([np.ones(mesh.n), np.zeros(mesh.n)])
        s = po.DiffusionUnsteadyMono(ph, bc_b, po.Dirichlet(0.0), dt, u0, "CN")
        po.solve_DiffusionUnsteadyMono(s, ph, dt, t_end, bc_b, po.Dirichlet(0.0), "CN")
        hs.append(1.0 / nx)
        errs.append(po.check_convergence(u_ana, s.x, cap, 2, False)[0])
    order = fitted_order(hs, errs)
    assert not np.isnan(order) and order > 1.0, (hs, errs)
    assert min(errs) < max(errs)


def test_heat_2ph_1d_henry_100():
    # diphasic/Heat_2ph_1D.jl:26-45,58-93,174-187: [0, 8], interface at x = 4, ScalarJump(1, He = 100, 0), FluxJump(1, 1, 0), u0 = [0, 0, 1, 1], :bottom -> 0,
    # :top -> 1, CN constructor and CN loop, dt = 0.5 h^2, T_end = 0.1, erfc similarity solution; meshes 4 .. 256; asserts: the fit is not NaN, errors vary
    from scipy.special import erfc
    lx, xint, t_end, He = 8.0, 4.0, 0.1, 100.0
    pref = -He / (1.0 + He)
    den = 2.0 * np.sqrt(t_end)
    u1 = lambda x: pref * (erfc((x - xint) / den) - 2.0)
    u2 = lambda x: pref * erfc((x - xint) / den) + 1.0
    hs, errs = [], []
    for nx in (4, 8, 16, 32, 64, 128, 256):
        mesh = po.Mesh((nx,), (lx,))
        ls = geom.LevelSet.halfspace(0, xint, True)
        c1, c2 = geom.capacity(mesh, ls), geom.capacity(mesh, ls.flipped())
        f = lambda x, y, z, t: 0.0 * x
        p1, p2 = po.Phase(c1, po.DiffusionOps(c1), f, 1.0), po.Phase(c2, po.DiffusionOps(c2), f, 1.0)
        bc_b = po.BorderConditions({"bottom": po.Dirichlet(0.0), "top": po.Dirichlet(1.0)})
        ic = po.InterfaceConditions(po.ScalarJump(1.0, He, 0.0), po.FluxJump(1.0, 1.0, 0.0))
        n = mesh.n
        u0 = np.concatenate([np.zeros(2 * n), np.ones(2 * n)])
        dt = 0.5 * (lx / nx) ** 2
        s = po.DiffusionUnsteadyDiph(p1, p2, bc_b, ic, dt, u0, "CN")
        po.solve_DiffusionUnsteadyDiph(s, p1, p2, dt, t_end, bc_b, ic, "CN")
        x = s.states[-1]
        e1 = po.check_convergence(u1, x[:n], c1, 2, False)[0]
        e2 = po.check_convergence(u2, x[2 * n:3 * n], c2, 2, False)[0]
        hs.append(lx / nx)
        errs.append(max(e1, e2))
    order = fitted_order(hs, errs)
    assert not np.isnan(order) and min(errs) < max(errs), (errs, order)
    assert errs[-1] < 0.05 * errs[0], errs                 # (beyond the script's asserts: the oracle does converge to the similarity solution)


def test_johansen_colella_problem4():
    # johansenColella/Problem4_SchwartzColella_Poisson3D.jl:20-31,49-72,113-131: -lap(phi) = 14 sin x sin 2y sin 3z inside the sphere r = 0.392 centre (1/2, 1/2, 1/2) of
    # the unit cube, phi = sin x sin 2y sin 3z imposed at the interface CENTROIDS (a function-valued Dirichlet: build_g_g, src/solver.jl:309-323); meshes 8 .. 64 (here
    # .. 32); asserts: the fit is not NaN, errors vary with the mesh.  The oracle also shows second-order behaviour.
    fe = lambda x, y, z: np.sin(x) * np.sin(2 * y) * np.sin(3 * z)
    hs, errs = [], []
    for nx in (8, 16, 32):
        mesh = po.Mesh((nx,) * 3, (1.0,) * 3)
        cap = geom.capacity(mesh, geom.LevelSet.ball((0.5, 0.5, 0.5), 0.392))
        ph = po.Phase(cap, po.DiffusionOps(cap), (lambda x, y, z: 14.0 * fe(x, y, z)), 1.0)
        bc_b = po.BorderConditions({k: po.Dirichlet(0.0) for k in ("left", "right", "top", "bottom")})
        s = po.solve_DiffusionSteadyMono(po.DiffusionSteadyMono(ph, bc_b, po.Dirichlet(fe)))
        hs.append(1.0 / nx)
        errs.append(po.check_convergence(fe, s.x, cap, 2, False)[0])
    order = fitted_order(hs, errs)
    assert not np.isnan(order) and min(errs) < max(errs)
    assert order > 1.5, (errs, order)
