"""GPU parity of BASELINE.json configs[4] at a size the oracle's sparse LU solves in a second: steady Poisson in 3-D around three disjoint
spheres (fluid outside), f = 1, Dirichlet 0 on the interface and on the six recognised border keys (src/solver/diffusion.jl:14-72; geometry of
BenchPhaseFlow/problems/scalar/Scalar_3D_Diffusion_Poisson_Dirichlet.jl:43-61 with several spheres).  The DEVICE builds the
capacities here (multi-ball level set), so the test covers geometry + operators + steady solve in one go; rel-L2 <= 1e-9 against the oracle.
tools/run_poisson3d.py is the full-size runner of the same problem."""
import numpy as np
import pytest

from helpers import rel_l2
from test_oracle_poisson3d import oracle_solution, small_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pb():
    import penguin_b200
    penguin_b200.init()
    return penguin_b200


@pytest.mark.parametrize("path", ["folded", "generic"])
def test_steady_poisson_multi_sphere_3d(pb, path):
    n, L, cen, rad = small_case()
    mesh_o, cap_o, so = oracle_solution(n, L, cen, rad)
    mesh = pb.Mesh(n, L)
    cap = pb.Capacity(pb.Balls(cen, rad, fluid_inside=False), mesh, compute_centroids=False)
    assert np.array_equal(cap.cell_types, cap_o.cell_types)
    assert np.linalg.norm(cap.V - cap_o.V) <= 1e-12 * np.linalg.norm(cap_o.V)
    ph = pb.Phase(cap, pb.DiffusionOps(cap), 1.0, 1.0)
    keys = ("left", "right", "top", "bottom", "forward", "backward")
    s = pb.DiffusionSteadyMono(ph, pb.BorderConditions({k: pb.Dirichlet(0.0) for k in keys}), pb.Dirichlet(0.0))
    pb.solve_DiffusionSteadyMono_(s, method="cg", reltol=1e-13, maxiter=50000, path=path)
    assert s.ch[-1]["converged"]
    assert rel_l2(s.x, so.x) < 1e-9
    nn = mesh_o.n
    assert np.all(s.x[:nn][cap_o.V == 0] == 0.0)                     # removed DOFs are exact zeros


def test_steady_poisson_multigrid_3d(pb, monkeypatch):
    """the same problem through the multigrid-preconditioned CG (csrc/mg.cuh: levels 16^3, 8^3, 4^3 rebuilt from the level set): same solution, fewer iterations"""
    monkeypatch.setenv("PB200_MG_RES", "0")          # coarsen all the way (by default a level whose cells outgrow the spheres is not built: 16^3 would stay alone)
    n, L, cen, rad = small_case()
    mesh_o, cap_o, so = oracle_solution(n, L, cen, rad)
    mesh = pb.Mesh(n, L)
    cap = pb.Capacity(pb.Balls(cen, rad, fluid_inside=False), mesh, compute_centroids=False)
    ph = pb.Phase(cap, pb.DiffusionOps(cap), 1.0, 1.0)
    keys = ("left", "right", "top", "bottom", "forward", "backward")
    bcb = pb.BorderConditions({k: pb.Dirichlet(0.0) for k in keys})
    s0 = pb.DiffusionSteadyMono(ph, bcb, pb.Dirichlet(0.0))
    pb.solve_DiffusionSteadyMono_(s0, method="cg", reltol=1e-13, maxiter=50000, path="folded")
    s = pb.DiffusionSteadyMono(ph, bcb, pb.Dirichlet(0.0))
    pb.solve_DiffusionSteadyMono_(s, method="cg", reltol=1e-13, maxiter=200, path="folded", precond="mg")
    assert s.ch[-1]["converged"]
    assert rel_l2(s.x, so.x) < 1e-9
    assert s.ch[-1]["iters"] < s0.ch[-1]["iters"]
    pb.solve_DiffusionSteadyMono_(s, method="cg", reltol=1e-13, maxiter=200, path="folded", precond="mg")     # hierarchy reused
    assert rel_l2(s.x, so.x) < 1e-9


def test_steady_poisson_multigrid_2d(pb):
    """2-D (slab dimension y): circle in a box, 64^2 -> 32^2 -> ... -> 4^2, against the plain folded CG"""
    mesh = pb.Mesh((64, 64), (4.0, 4.0))
    cap = pb.Capacity(-pb.Circle((2.05, 1.97), 0.8), mesh, compute_centroids=False)
    ph = pb.Phase(cap, pb.DiffusionOps(cap), (lambda x, y, z: 1.0 + x), 1.0)
    bcb = pb.BorderConditions({k: pb.Dirichlet(0.5) for k in ("left", "right", "top", "bottom")})
    s0 = pb.DiffusionSteadyMono(ph, bcb, pb.Dirichlet(1.0))
    pb.solve_DiffusionSteadyMono_(s0, method="cg", reltol=1e-13, maxiter=50000, path="folded")
    s = pb.DiffusionSteadyMono(ph, bcb, pb.Dirichlet(1.0))
    pb.solve_DiffusionSteadyMono_(s, method="cg", reltol=1e-13, maxiter=200, path="folded", precond="mg")
    assert s.ch[-1]["converged"]
    assert rel_l2(s.x, s0.x) < 1e-10
    assert s.ch[-1]["iters"] < s0.ch[-1]["iters"]


def test_multigrid_refuses_what_it_cannot_do(pb):
    """imported capacities carry no level set to rebuild on coarser meshes: PB200_EUNSUPPORTED, not a silent fallback"""
    from helpers import import_capacity
    n, L, cen, rad = small_case()
    mesh_o, cap_o, so = oracle_solution(n, L, cen, rad)
    mesh = pb.Mesh(n, L)
    cap = import_capacity(pb, mesh, cap_o)
    ph = pb.Phase(cap, pb.DiffusionOps(cap), 1.0, 1.0)
    keys = ("left", "right", "top", "bottom", "forward", "backward")
    s = pb.DiffusionSteadyMono(ph, pb.BorderConditions({k: pb.Dirichlet(0.0) for k in keys}), pb.Dirichlet(0.0))
    with pytest.raises(Exception):
        pb.solve_DiffusionSteadyMono_(s, method="cg", path="folded", precond="mg")
    with pytest.raises(Exception):                      # the generic path has no multigrid: refused, not ignored
        pb.solve_DiffusionSteadyMono_(s, method="cg", path="generic", precond="mg")


def test_diphasic_3d_on_device_built_capacities(pb):
    """geometry + operators + solve in one chain, 3-D: the diphasic BE heat problem of examples/3D/Diffusion/Heat_2ph.jl:13-30 at 20^3 with the capacities
    built ON THE DEVICE (near-empty cut cells included: the smallest cut-cell volume is 1.5e-5 h^3), against the oracle chain (oracle geometry -> oracle LU) at rel-L2 <= 1e-9
    per state.  This is the end-to-end answer to "B in near-empty cells agrees to 1e-10 only": what the states see of it is below 1e-11."""
    from oracle import geom, penguin_oracle as po
    nx = 20
    mesh_o, mesh_g = po.Mesh((nx,) * 3, (4.0,) * 3), pb.Mesh((nx,) * 3, (4.0,) * 3)
    body = pb.Sphere((2.0, 2.0, 2.0), 1.0)
    f = lambda x, y, z, t: 0.0 * x
    c1, c2 = pb.Capacity(body, mesh_g), pb.Capacity(-body, mesh_g)
    p1, p2 = pb.Phase(c1, pb.DiffusionOps(c1), f, 1.0), pb.Phase(c2, pb.DiffusionOps(c2), f, 1.0)
    n = (nx + 1) ** 3
    u0 = np.concatenate([np.ones(2 * n), np.zeros(2 * n)])
    dt = 0.5 * (4.0 / nx) ** 2
    ic = pb.InterfaceConditions(pb.ScalarJump(1.0, 2.0, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
    s = pb.DiffusionUnsteadyDiph(p1, p2, pb.BorderConditions(), ic, dt, u0, "BE")
    pb.solve_DiffusionUnsteadyDiph_(s, p1, p2, dt, 2.5 * dt, pb.BorderConditions(), ic, "BE", reltol=1e-13)
    ls = geom.LevelSet.ball((2.0, 2.0, 2.0), 1.0)
    o1, o2 = geom.capacity(mesh_o, ls), geom.capacity(mesh_o, ls.flipped())
    q1, q2 = po.Phase(o1, po.DiffusionOps(o1), f, 1.0), po.Phase(o2, po.DiffusionOps(o2), f, 1.0)
    ico = po.InterfaceConditions(po.ScalarJump(1.0, 2.0, 0.0), po.FluxJump(1.0, 1.0, 0.0))
    so = po.DiffusionUnsteadyDiph(q1, q2, po.BorderConditions(), ico, dt, u0, "BE")
    po.solve_DiffusionUnsteadyDiph(so, q1, q2, dt, 2.5 * dt, po.BorderConditions(), ico, "BE")
    assert np.array_equal(c1.cell_types, o1.cell_types) and np.array_equal(c2.cell_types, o2.cell_types)
    assert len(s.states) == len(so.states)
    for a, b in zip(s.states, so.states):
        assert rel_l2(a, b) < 1e-9
