"""GPU parity of the operator / solver stages against the CPU oracle on identical (imported) capacities.

Bar (BASELINE.json north_star): per-step solutions agree with the reference solve at the same dt to relative
L2 <= 1e-9 in fp64.  Every call goes through the C ABI (penguin_b200 -> ctypes -> libpenguin_b200.so).
"""
import numpy as np
import pytest

from oracle import geom
from oracle import penguin_oracle as po
from helpers import import_capacity, rel_l2, to_oracle_bc, to_oracle_borders

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(params=["folded", "generic"])
def kw(request):
    """both implementations of the solve (pb200_krylov_opts.path) against the oracle"""
    return dict(reltol=1e-13, maxiter=50000, path=request.param)


@pytest.fixture(scope="module")
def pb():
    import penguin_b200
    penguin_b200.init()
    return penguin_b200


def _phases(pb, mesh_o, mesh_g, ls, f, D):
    cap_o = geom.capacity(mesh_o, ls)
    op_o = po.DiffusionOps(cap_o)
    cap_g = import_capacity(pb, mesh_g, cap_o)
    op_g = pb.DiffusionOps(cap_g)
    return po.Phase(cap_o, op_o, f, D), pb.Phase(cap_g, op_g, f, D)


def _meshes(pb, n, L, x0=None):
    return po.Mesh(n, L, x0), pb.Mesh(n, L, x0)


@pytest.mark.parametrize("n,L,c,r", [((12,), (4.0,), (2.1,), 1.0), ((16, 12), (4.0, 3.0), (2.05, 1.45), 0.9),
                                     ((9, 8, 7), (4.0, 4.0, 4.0), (2.0, 2.1, 1.9), 1.2)])
def test_grad_div(pb, n, L, c, r):
    mo, mg = _meshes(pb, n, L)
    f = lambda *a: 0.0
    pho, phg = _phases(pb, mo, mg, geom.LevelSet.ball(c, r), f, 1.0)
    rng = np.random.default_rng(0)
    nn = mo.n
    p = rng.standard_normal(2 * nn)
    assert rel_l2(pb.grad(phg.operator, p), po.grad(pho.operator, p)) < 1e-13
    N = len(n)
    qo, qg = rng.standard_normal(N * nn), rng.standard_normal(N * nn)
    assert rel_l2(pb.div(phg.operator, qo, qg), po.div(pho.operator, qo, qg)) < 1e-13
    assert np.array_equal(phg.operator.Wdag, pho.operator.Wdag_diag)


def test_steady_mono_2d(pb, kw):
    # test/solver/diffusion_test.jl:5-26 on both sides
    mo, mg = _meshes(pb, (20, 20), (2.0, 2.0))
    f, D = (lambda x, y, z: 0.0 * x), (lambda x, y, z: 1.0 + 0 * x)
    pho, phg = _phases(pb, mo, mg, geom.LevelSet.ball((0.5, 0.5), 0.5), f, D)
    bcb = pb.BorderConditions({k: pb.Dirichlet(1.0) for k in ("left", "right", "top", "bottom")})
    so = po.solve_DiffusionSteadyMono(po.DiffusionSteadyMono(pho, to_oracle_borders(pb, bcb), po.Dirichlet(1.0)))
    sg = pb.solve_DiffusionSteadyMono_(pb.DiffusionSteadyMono(phg, bcb, pb.Dirichlet(1.0)), **kw)
    assert rel_l2(sg.x, so.x) < TOL
    assert abs(sg.x[:mo.n].max() - 1.0) < 1e-2


@pytest.mark.parametrize("ifc", ["dirichlet", "robin", "neumann"])
def test_steady_mono_manufactured(pb, kw, ifc):
    # test/convergence_test.jl:30-49 geometry; f = 4, Dirichlet / Robin / Neumann interface rows
    mo, mg = _meshes(pb, (24, 24), (4.0, 4.0))
    f, D = (lambda x, y, z: 4.0 + 0 * x), 1.0
    pho, phg = _phases(pb, mo, mg, geom.LevelSet.ball((2.0, 2.0), 1.0), f, D)
    bcb = pb.BorderConditions({k: pb.Dirichlet(1.0) for k in ("left", "right", "top", "bottom")})
    bci = {"dirichlet": pb.Dirichlet(0.0), "robin": pb.Robin(1.0, 0.5, 0.25), "neumann": pb.Robin(1e-3, 1.0, 0.1)}[ifc]
    so = po.solve_DiffusionSteadyMono(po.DiffusionSteadyMono(pho, to_oracle_borders(pb, bcb), to_oracle_bc(pb, bci)))
    sg = pb.solve_DiffusionSteadyMono_(pb.DiffusionSteadyMono(phg, bcb, bci), **kw)
    assert rel_l2(sg.x, so.x) < TOL


def test_steady_diph_2d(pb, kw):
    # test/solver/diffusion_test.jl:28-55 (40^2 here): max u1 pinned on the oracle at 80^2
    mo, mg = _meshes(pb, (40, 40), (4.0, 4.0))
    f, D = (lambda x, y, z: 1.0 + 0 * x), 1.0
    ls = geom.LevelSet.ball((2.0, 2.0), 1.0)
    p1o, p1g = _phases(pb, mo, mg, ls, f, D)
    p2o, p2g = _phases(pb, mo, mg, ls.flipped(), f, D)
    bcb = pb.BorderConditions({k: pb.Dirichlet(0.0) for k in ("left", "right", "top", "bottom")})
    so = po.solve_DiffusionSteadyDiph(po.DiffusionSteadyDiph(p1o, p2o, to_oracle_borders(pb, bcb),
                                                             po.InterfaceConditions(po.ScalarJump(1.0, 1.0, 0.0), po.FluxJump(1.0, 1.0, 0.0))))
    ic = pb.InterfaceConditions(pb.ScalarJump(1.0, 1.0, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
    sg = pb.solve_DiffusionSteadyDiph_(pb.DiffusionSteadyDiph(p1g, p2g, bcb, ic), **kw)
    assert rel_l2(sg.x, so.x) < TOL


@pytest.mark.parametrize("scheme", ["BE", "CN"])
@pytest.mark.parametrize("ifc", ["dirichlet_fn", "robin"])
def test_unsteady_mono_2d(pb, kw, scheme, ifc):
    # README quick start (README.md:43-79) at 32^2: interface Dirichlet sin(pi x) sin(pi y), borders Dirichlet 0
    nx = 32
    mo, mg = _meshes(pb, (nx, nx), (4.0, 4.0))
    f = lambda x, y, z, t: 0.3 * np.sin(x) * (1 + t)
    D = lambda x, y, z: 1.0 + 0 * x
    pho, phg = _phases(pb, mo, mg, geom.LevelSet.ball((2.01, 2.01), 1.0), f, D)
    bcb = pb.BorderConditions({k: pb.Dirichlet(0.0) for k in ("left", "right", "top", "bottom")})
    if ifc == "dirichlet_fn":
        bci = pb.Dirichlet(lambda x, y, z, t: np.sin(np.pi * x) * np.sin(np.pi * y) * (1 + 10 * t))
    else:
        bci = pb.Robin(2.0, 1.0, 0.5)
    n = mo.n
    u0 = np.concatenate([np.zeros(n), np.ones(n)])
    dt = 0.25 * (4.0 / nx) ** 2
    Tend = 4.5 * dt
    so = po.DiffusionUnsteadyMono(pho, to_oracle_borders(pb, bcb), to_oracle_bc(pb, bci), dt, u0, "BE")
    po.solve_DiffusionUnsteadyMono(so, pho, dt, Tend, to_oracle_borders(pb, bcb), to_oracle_bc(pb, bci), scheme)
    sg = pb.DiffusionUnsteadyMono(phg, bcb, bci, dt, u0, "BE")
    pb.solve_DiffusionUnsteadyMono_(sg, phg, dt, Tend, bcb, bci, scheme, **kw)
    assert len(sg.states) == len(so.states) == 6
    for a, b in zip(sg.states, so.states):
        assert rel_l2(a, b) < TOL


@pytest.mark.parametrize("scheme", ["BE", "CN"])
def test_unsteady_diph_2d(pb, kw, scheme):
    # benchmark/Heat_2ph_2D.jl:64-111 at 32^2: empty BorderConditions, ScalarJump(1, He, 0), FluxJump(1, 1, 0)
    nx = 32
    mo, mg = _meshes(pb, (nx, nx), (8.0, 8.0))
    f = lambda x, y, z, t: 0.0 * x
    ls = geom.LevelSet.ball((4.0, 4.0), 2.0)
    p1o, p1g = _phases(pb, mo, mg, ls, f, lambda x, y, z: 1.0 + 0 * x)
    p2o, p2g = _phases(pb, mo, mg, ls.flipped(), f, lambda x, y, z: 2.0 + 0 * x)
    n = mo.n
    u0 = np.concatenate([np.ones(n), np.ones(n), np.zeros(n), np.zeros(n)])
    dt = 0.5 * (8.0 / nx) ** 2
    Tend = 3.5 * dt
    ico = po.InterfaceConditions(po.ScalarJump(1.0, 0.5, 0.0), po.FluxJump(1.0, 1.0, 0.0))
    icg = pb.InterfaceConditions(pb.ScalarJump(1.0, 0.5, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
    so = po.DiffusionUnsteadyDiph(p1o, p2o, po.BorderConditions(), ico, dt, u0, "BE")
    po.solve_DiffusionUnsteadyDiph(so, p1o, p2o, dt, Tend, po.BorderConditions(), ico, scheme)
    sg = pb.DiffusionUnsteadyDiph(p1g, p2g, pb.BorderConditions(), icg, dt, u0, "BE")
    pb.solve_DiffusionUnsteadyDiph_(sg, p1g, p2g, dt, Tend, pb.BorderConditions(), icg, scheme, **kw)
    assert len(sg.states) == len(so.states) == 5
    for a, b in zip(sg.states, so.states):
        assert rel_l2(a, b) < TOL


def test_unsteady_diph_borders_jump_values(pb, kw):
    # Dirichlet borders on a diphasic problem + non-zero jump data g, h
    nx = 24
    mo, mg = _meshes(pb, (nx, nx), (4.0, 4.0))
    f = lambda x, y, z, t: 1.0 + 0 * x
    ls = geom.LevelSet.ball((2.0, 2.0), 1.0)
    p1o, p1g = _phases(pb, mo, mg, ls, f, 1.0)
    p2o, p2g = _phases(pb, mo, mg, ls.flipped(), f, 3.0)
    n = mo.n
    u0 = np.zeros(4 * n)
    dt = 0.5 * (4.0 / nx) ** 2
    Tend = 2.5 * dt
    keys = ("left", "right", "top", "bottom")
    bco = po.BorderConditions({k: po.Dirichlet(0.5) for k in keys})
    bcg = pb.BorderConditions({k: pb.Dirichlet(0.5) for k in keys})
    ico = po.InterfaceConditions(po.ScalarJump(1.0, 2.0, 0.3), po.FluxJump(1.0, 1.5, 0.2))
    icg = pb.InterfaceConditions(pb.ScalarJump(1.0, 2.0, 0.3), pb.FluxJump(1.0, 1.5, 0.2))
    so = po.DiffusionUnsteadyDiph(p1o, p2o, bco, ico, dt, u0, "BE")
    po.solve_DiffusionUnsteadyDiph(so, p1o, p2o, dt, Tend, bco, ico, "BE")
    sg = pb.DiffusionUnsteadyDiph(p1g, p2g, bcg, icg, dt, u0, "BE")
    pb.solve_DiffusionUnsteadyDiph_(sg, p1g, p2g, dt, Tend, bcg, icg, "BE", **kw)
    for a, b in zip(sg.states, so.states):
        assert rel_l2(a, b) < TOL


def test_unsteady_mono_3d_cn(pb, kw):
    # benchmark/Heat3D.jl:53-74 at 14^3: sphere, interface Dirichlet 1, borders Dirichlet 1 (incl. forward/backward), BE then CN
    nx = 14
    mo, mg = _meshes(pb, (nx, nx, nx), (4.0, 4.0, 4.0))
    f = lambda x, y, z, t: 0.0 * x
    pho, phg = _phases(pb, mo, mg, geom.LevelSet.ball((2.01, 2.01, 2.01), 1.0, False), f, 1.0)
    keys = ("left", "right", "top", "bottom", "forward", "backward")
    bco = po.BorderConditions({k: po.Dirichlet(1.0) for k in keys})
    bcg = pb.BorderConditions({k: pb.Dirichlet(1.0) for k in keys})
    n = mo.n
    u0 = np.zeros(2 * n)
    dt = 0.75 * (4.0 / nx) ** 2
    Tend = 2.5 * dt
    so = po.DiffusionUnsteadyMono(pho, bco, po.Dirichlet(1.0), dt, u0, "BE")
    po.solve_DiffusionUnsteadyMono(so, pho, dt, Tend, bco, po.Dirichlet(1.0), "CN")
    sg = pb.DiffusionUnsteadyMono(phg, bcg, pb.Dirichlet(1.0), dt, u0, "BE")
    pb.solve_DiffusionUnsteadyMono_(sg, phg, dt, Tend, bcg, pb.Dirichlet(1.0), "CN", **kw)
    for a, b in zip(sg.states, so.states):
        assert rel_l2(a, b) < TOL


def test_unsteady_diph_1d_halfspace(pb, kw):
    # test/convergence_test.jl:100-192 (first steps)
    nx, lx, xint = 100, 8.0, 4.0
    mo, mg = _meshes(pb, (nx,), (lx,))
    f = lambda x, y, z, t: 0.0 * x
    ls = geom.LevelSet.halfspace(0, xint, True)
    p1o, p1g = _phases(pb, mo, mg, ls, f, 1.0)
    p2o, p2g = _phases(pb, mo, mg, ls.flipped(), f, 1.0)
    n = mo.n
    u0 = np.concatenate([np.zeros(n), np.zeros(n), np.ones(n), np.ones(n)])
    dt = 0.5 * (lx / nx) ** 2
    Tend = 5.5 * dt
    bco = po.BorderConditions({"top": po.Dirichlet(1.0), "bottom": po.Dirichlet(0.0)})
    bcg = pb.BorderConditions({"top": pb.Dirichlet(1.0), "bottom": pb.Dirichlet(0.0)})
    ico = po.InterfaceConditions(po.ScalarJump(1.0, 0.5, 0.0), po.FluxJump(1.0, 1.0, 0.0))
    icg = pb.InterfaceConditions(pb.ScalarJump(1.0, 0.5, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
    so = po.DiffusionUnsteadyDiph(p1o, p2o, bco, ico, dt, u0, "BE")
    po.solve_DiffusionUnsteadyDiph(so, p1o, p2o, dt, Tend, bco, ico, "BE")
    sg = pb.DiffusionUnsteadyDiph(p1g, p2g, bcg, icg, dt, u0, "BE")
    pb.solve_DiffusionUnsteadyDiph_(sg, p1g, p2g, dt, Tend, bcg, icg, "BE", **kw)
    for a, b in zip(sg.states, so.states):
        assert rel_l2(a, b) < TOL


def test_removed_dofs_are_exact_zero(pb, kw):
    # solve_system! scatters into zeros(n): removed DOFs are exactly 0.0 (src/solver.jl:186-187)
    mo, mg = _meshes(pb, (16, 16), (4.0, 4.0))
    f = lambda x, y, z: 1.0 + 0 * x
    pho, phg = _phases(pb, mo, mg, geom.LevelSet.ball((2.0, 2.0), 1.0), f, 1.0)
    so = po.solve_DiffusionSteadyMono(po.DiffusionSteadyMono(pho, po.BorderConditions(), po.Dirichlet(0.0)))
    sg = pb.solve_DiffusionSteadyMono_(pb.DiffusionSteadyMono(phg, pb.BorderConditions(), pb.Dirichlet(0.0)), **kw)
    # every DOF the reference removes is exactly 0.0 on the device too; a device zero on a kept DOF is legitimate only where
    # the exact value is 0 (T_gamma = g = 0 here -- the oracle's LU leaves round-off noise there)
    assert np.all(sg.x[so.x == 0.0] == 0.0)
    assert np.max(np.abs(so.x[sg.x == 0.0])) < 1e-14
    assert rel_l2(sg.x, so.x) < TOL


def test_diph_cg_on_folded_system(pb):
    # the symmetrised diphasic system is SPD: CG and BiCGSTAB reach the same solution (and the oracle's)
    nx = 40
    mo, mg = _meshes(pb, (nx, nx), (8.0, 8.0))
    f = lambda x, y, z, t: 0.1 * x
    ls = geom.LevelSet.ball((4.0, 4.0), 2.0)
    p1o, p1g = _phases(pb, mo, mg, ls, f, 1.0)
    p2o, p2g = _phases(pb, mo, mg, ls.flipped(), f, 2.5)
    n = mo.n
    u0 = np.concatenate([np.ones(n), np.ones(n), np.zeros(n), np.zeros(n)])
    dt = 0.5 * (8.0 / nx) ** 2
    Tend = 2.5 * dt
    ico = po.InterfaceConditions(po.ScalarJump(1.0, 2.0, 0.1), po.FluxJump(1.0, 3.0, 0.05))
    icg = pb.InterfaceConditions(pb.ScalarJump(1.0, 2.0, 0.1), pb.FluxJump(1.0, 3.0, 0.05))
    so = po.DiffusionUnsteadyDiph(p1o, p2o, po.BorderConditions(), ico, dt, u0, "BE")
    po.solve_DiffusionUnsteadyDiph(so, p1o, p2o, dt, Tend, po.BorderConditions(), ico, "CN")
    for method in ("cg", "bicgstab"):
        sg = pb.DiffusionUnsteadyDiph(p1g, p2g, pb.BorderConditions(), icg, dt, u0, "BE")
        pb.solve_DiffusionUnsteadyDiph_(sg, p1g, p2g, dt, Tend, pb.BorderConditions(), icg, "CN", method=method, reltol=1e-13, path="folded")
        for a, b in zip(sg.states, so.states):
            assert rel_l2(a, b) < TOL


def test_async_state_download(pb):
    # pb200_solver_get_state_async / wait_state deliver the same state as the blocking call, also when the next step overlaps the copy
    import ctypes as C
    from penguin_b200 import _lib as L
    nx = 24
    mo, mg = _meshes(pb, (nx, nx), (4.0, 4.0))
    f = lambda x, y, z, t: 1.0 + 0 * x
    pho, phg = _phases(pb, mo, mg, geom.LevelSet.ball((2.03, 1.98), 1.0), f, 1.0)
    n = mo.n
    dt = 0.25 * (4.0 / nx) ** 2
    sg = pb.DiffusionUnsteadyMono(phg, pb.BorderConditions(), pb.Dirichlet(0.5), dt, np.zeros(2 * n), "BE")
    pb.solve_DiffusionUnsteadyMono_(sg, phg, dt, 1.5 * dt, pb.BorderConditions(), pb.Dirichlet(0.5), "BE", reltol=1e-12)
    ref = sg.states[-1].copy()
    buf = np.full(2 * n, np.nan)
    lib = L.lib()
    L.check(lib.pb200_solver_get_state_async(sg._h, buf.ctypes.data_as(L.dp)), sg._ctx.h)
    # queue another step right away: its write-back must wait for the copy
    pb.api._step(sg, "BE", dt, 3 * dt, pb.Dirichlet(0.5), None, pb.api._krylov_opts("cg", dict(reltol=1e-12)))
    L.check(lib.pb200_solver_wait_state(sg._h), sg._ctx.h)
    assert np.array_equal(buf, ref)
    assert not np.array_equal(sg.x, ref)


@pytest.mark.parametrize("order", [2, 4])
def test_extrapolated_initial_guess_keeps_parity(pb, order):
    # warm_start = m: initial guess by polynomial extrapolation through the last m states (buffers rotate on the device);
    # the converged states must still match the oracle step by step, for the diphasic CN loop and with time-dependent data
    nx = 28
    mo, mg = _meshes(pb, (nx, nx), (8.0, 8.0))
    f = lambda x, y, z, t: 0.3 * np.cos(x) * (1 + 5 * t)
    ls = geom.LevelSet.ball((4.02, 3.97), 2.0)
    p1o, p1g = _phases(pb, mo, mg, ls, f, 1.0)
    p2o, p2g = _phases(pb, mo, mg, ls.flipped(), f, 2.0)
    n = mo.n
    u0 = np.concatenate([np.ones(n), np.ones(n), np.zeros(n), np.zeros(n)])
    dt = 0.5 * (8.0 / nx) ** 2
    Tend = 8.5 * dt
    ico = po.InterfaceConditions(po.ScalarJump(1.0, 0.5, 0.0), po.FluxJump(1.0, 1.0, 0.0))
    icg = pb.InterfaceConditions(pb.ScalarJump(1.0, 0.5, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
    so = po.DiffusionUnsteadyDiph(p1o, p2o, po.BorderConditions(), ico, dt, u0, "BE")
    po.solve_DiffusionUnsteadyDiph(so, p1o, p2o, dt, Tend, po.BorderConditions(), ico, "CN")
    sg = pb.DiffusionUnsteadyDiph(p1g, p2g, pb.BorderConditions(), icg, dt, u0, "BE")
    pb.solve_DiffusionUnsteadyDiph_(sg, p1g, p2g, dt, Tend, pb.BorderConditions(), icg, "CN", reltol=1e-13, warm_start=order)
    assert len(sg.states) == len(so.states) == 10
    for a, b in zip(sg.states, so.states):
        assert rel_l2(a, b) < TOL
    # later steps start closer to the solution than the first ones
    its = [c["iters"] for c in sg.ch]
    assert min(its[4:]) <= its[1]


def test_unsteady_diph_3d(pb, kw):
    # examples/3D/Diffusion/Heat_2ph.jl:13-30 at 12^3: sphere interface, ScalarJump(1, 2, 0), FluxJump(1, 1, 0), empty borders, BE
    nx = 12
    mo, mg = _meshes(pb, (nx, nx, nx), (4.0, 4.0, 4.0))
    f = lambda x, y, z, t: 0.0 * x
    ls = geom.LevelSet.ball((2.0, 2.0, 2.0), 1.0)
    p1o, p1g = _phases(pb, mo, mg, ls, f, 1.0)
    p2o, p2g = _phases(pb, mo, mg, ls.flipped(), f, 1.0)
    n = mo.n
    u0 = np.concatenate([np.ones(2 * n), np.zeros(2 * n)])
    dt = 0.5 * (4.0 / nx) ** 2
    Tend = 3.5 * dt
    ico = po.InterfaceConditions(po.ScalarJump(1.0, 2.0, 0.0), po.FluxJump(1.0, 1.0, 0.0))
    icg = pb.InterfaceConditions(pb.ScalarJump(1.0, 2.0, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
    so = po.DiffusionUnsteadyDiph(p1o, p2o, po.BorderConditions(), ico, dt, u0, "BE")
    po.solve_DiffusionUnsteadyDiph(so, p1o, p2o, dt, Tend, po.BorderConditions(), ico, "BE")
    sg = pb.DiffusionUnsteadyDiph(p1g, p2g, pb.BorderConditions(), icg, dt, u0, "BE")
    pb.solve_DiffusionUnsteadyDiph_(sg, p1g, p2g, dt, Tend, pb.BorderConditions(), icg, "BE", **kw)
    assert len(sg.states) == len(so.states) == 5
    for a, b in zip(sg.states, so.states):
        assert rel_l2(a, b) < TOL


@pytest.mark.parametrize("which", ["x", "y", "full"])
def test_periodic_borders(pb, kw, which):
    # test/solver_test.jl:78-171: Periodic rows (x_row - x_partner = 0; the low side's partner is the pad cell, which the trimming removes)
    nx = 20
    mo, mg = _meshes(pb, (nx, nx), (4.0, 4.0))
    f = lambda x, y, z: 1.0 + 0 * x
    pho, phg = _phases(pb, mo, mg, geom.LevelSet.ball((2.03, 1.98), 1.0, False), f, 1.0)
    if which == "x":
        bo = {"left": po.Periodic(), "right": po.Periodic(), "top": po.Dirichlet(0.0), "bottom": po.Dirichlet(0.0)}
        bg = {"left": pb.Periodic(), "right": pb.Periodic(), "top": pb.Dirichlet(0.0), "bottom": pb.Dirichlet(0.0)}
    elif which == "y":
        bo = {"left": po.Dirichlet(0.25), "right": po.Dirichlet(0.25), "top": po.Periodic(), "bottom": po.Periodic()}
        bg = {"left": pb.Dirichlet(0.25), "right": pb.Dirichlet(0.25), "top": pb.Periodic(), "bottom": pb.Periodic()}
    else:
        bo = {k: po.Periodic() for k in ("left", "right", "top", "bottom")}
        bg = {k: pb.Periodic() for k in ("left", "right", "top", "bottom")}
    so = po.solve_DiffusionSteadyMono(po.DiffusionSteadyMono(pho, po.BorderConditions(bo), po.Dirichlet(1.0)))
    sg = pb.solve_DiffusionSteadyMono_(pb.DiffusionSteadyMono(phg, pb.BorderConditions(bg), pb.Dirichlet(1.0)), **kw)
    assert rel_l2(sg.x, so.x) < TOL
    sol = sg.x[:mo.n].reshape(nx + 1, nx + 1)
    if which == "x":       # the reference's own assertion (:41): opposite columns agree
        assert np.allclose(sol[1:nx - 1, 0], sol[1:nx - 1, nx - 1], atol=1e-8)


@pytest.mark.parametrize("phase", ["mono", "diph"])
def test_neumann_border_1d(pb, phase):
    # src/solver.jl:471-493: in 1-D a Neumann border is a real row, (x_row - x_adj) / dx = g (in >= 2-D it is a warning no-op)
    nx, lx = 60, 6.0
    mo, mg = _meshes(pb, (nx,), (lx,))
    dt = 0.4 * (lx / nx) ** 2
    bo = po.BorderConditions({"bottom": po.Neumann(0.7), "top": po.Dirichlet(0.2)})
    bg = pb.BorderConditions({"bottom": pb.Neumann(0.7), "top": pb.Dirichlet(0.2)})
    n = mo.n
    if phase == "mono":
        f = lambda x, y, z, t: 1.0 + 0 * x
        pho, phg = _phases(pb, mo, mg, geom.LevelSet.halfspace(0, 4.13, True), f, 1.0)
        u0 = np.zeros(2 * n)
        so = po.DiffusionUnsteadyMono(pho, bo, po.Robin(1.0, 0.5, 0.3), dt, u0, "BE")
        po.solve_DiffusionUnsteadyMono(so, pho, dt, 3.5 * dt, bo, po.Robin(1.0, 0.5, 0.3), "CN")
        sg = pb.DiffusionUnsteadyMono(phg, bg, pb.Robin(1.0, 0.5, 0.3), dt, u0, "BE")
        pb.solve_DiffusionUnsteadyMono_(sg, phg, dt, 3.5 * dt, bg, pb.Robin(1.0, 0.5, 0.3), "CN", reltol=1e-13, maxiter=50000)
    else:
        f = lambda x, y, z, t: 0.0 * x
        ls = geom.LevelSet.halfspace(0, 3.07, True)
        p1o, p1g = _phases(pb, mo, mg, ls, f, 1.0)
        p2o, p2g = _phases(pb, mo, mg, ls.flipped(), f, 2.0)
        u0 = np.concatenate([np.zeros(2 * n), np.ones(2 * n)])
        bo = po.BorderConditions({"bottom": po.Dirichlet(0.0), "top": po.Neumann(-0.4)})
        bg = pb.BorderConditions({"bottom": pb.Dirichlet(0.0), "top": pb.Neumann(-0.4)})
        ico = po.InterfaceConditions(po.ScalarJump(1.0, 0.5, 0.0), po.FluxJump(1.0, 1.0, 0.0))
        icg = pb.InterfaceConditions(pb.ScalarJump(1.0, 0.5, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
        so = po.DiffusionUnsteadyDiph(p1o, p2o, bo, ico, dt, u0, "BE")
        po.solve_DiffusionUnsteadyDiph(so, p1o, p2o, dt, 3.5 * dt, bo, ico, "BE")
        sg = pb.DiffusionUnsteadyDiph(p1g, p2g, bg, icg, dt, u0, "BE")
        pb.solve_DiffusionUnsteadyDiph_(sg, p1g, p2g, dt, 3.5 * dt, bg, icg, "BE", reltol=1e-13, maxiter=50000)
    assert len(sg.states) == len(so.states)
    for a, b in zip(sg.states, so.states):
        assert rel_l2(a, b) < TOL
