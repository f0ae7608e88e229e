"""CPU check of the ALGORITHM behind the multigrid preconditioner of csrc/mg.cuh (BASELINE.json configs[4]), on the oracle's matrices of the
steady system (src/solver/diffusion.jl:30-43): rediscretised coarse levels (the capacities recomputed on n/2, n/4), cell-aggregation transfers in the
Jacobi-scaled variables, degree-2 Chebyshev smoothers on [lambda_max / 3, lambda_max], a fixed polynomial as the coarsest solve.  The CUDA implementation is
compared with the oracle's direct solve in tests/test_gpu_zz_poisson3d.py; here the claims that make it a legitimate CG preconditioner are checked:
the V-cycle is a symmetric positive definite linear operator, and it cuts the iteration count without changing the solution."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "experiments"))
import mg_experiment as mg          # noqa: E402
from test_oracle_poisson3d import small_case   # noqa: E402


def test_vcycle_is_spd_and_cuts_the_iterations():
    _, _, cen, rad = small_case()                      # three disjoint spheres in [0, 4]^3
    levels = [mg.level(n, cen, rad) for n in (32, 16, 8)]
    H = mg.build(levels, m=2, alpha=3.0)
    L0 = H[0]
    nfree = L0["Mh"].shape[0]
    assert nfree > 20000 and H[1]["Mh"].shape[0] > 2000 and H[2]["Mh"].shape[0] > 200
    B = lambda r: mg.vcycle(H, 0, r, coarse_sweeps=12)
    rng = np.random.default_rng(3)
    u, v = rng.standard_normal(nfree), rng.standard_normal(nfree)
    Bu, Bv = B(u), B(v)
    assert abs(u @ Bv - v @ Bu) <= 1e-12 * (np.linalg.norm(u) * np.linalg.norm(Bv))      # symmetric: every piece is a polynomial in M^ or a P / P^T pair
    assert u @ Bu > 0 and v @ Bv > 0                                                      # positive
    assert np.allclose(B(2.0 * u - 3.0 * v), 2.0 * Bu - 3.0 * Bv, rtol=0, atol=1e-10 * np.linalg.norm(Bu))   # linear (a FIXED polynomial on the coarsest level)
    b = L0["s"] * levels[0]["cap"].V[L0["a"]]                                            # f = 1 in the scaled rows
    x0, it0 = mg.pcg(L0["Mh"], b, lambda r: r, rtol=1e-10)
    x1, it1 = mg.pcg(L0["Mh"], b, B, rtol=1e-10)
    assert it1 * 4 < it0, (it0, it1)
    assert np.linalg.norm(x1 - x0) <= 1e-8 * np.linalg.norm(x0)
    # the rediscretised coarse operator is the right one for cell aggregation: over-/under-weighting the coarse correction only loses
    it_half = mg.pcg(L0["Mh"], b, lambda r: mg.vcycle(H, 0, r, 0.5, 12), rtol=1e-10)[1]
    it_twice = mg.pcg(L0["Mh"], b, lambda r: mg.vcycle(H, 0, r, 2.0, 12), rtol=1e-10)[1]
    assert it1 <= it_half and it1 <= it_twice, (it1, it_half, it_twice)
