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
