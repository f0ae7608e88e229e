"""Oracle parity of the kernel branches the BENCHMARK runs (round-1 verdict, weak #1 / #2).

The interior branch of the folded operator apply (tiles whose cells are all valid and whose coefficients are one constant per
direction: 88 % of the cells at 2048^2) only exists on grids that hold at least one whole 32 x 32 (2-D) / 32 x 8 x 4 (3-D) tile
of full cells away from the border ring and the interface.  The cases below are sized so that such tiles exist in the phase(s)
under test -- asserted through pb200_step_stats.apply_cells_fast -- and every state is compared with the oracle's direct solve
(sparse LU of the reference's assembled system, src/solver/diffusion.jl:212-454, src/solver.jl:158-188) at rel-L2 <= 1e-9, for
  * one chunk and several chunks of CUDA-graph replay (check_every),
  * the polynomial preconditioner off / degree 1 / degree 2 (PB200_POLY),
  * the bench's own settings (rtol = 1e-10 on the scaled residual, warm_start = 4: extrapolated initial guess).
"""
import numpy as np
import pytest

from oracle import geom
from oracle import penguin_oracle as po
from helpers import import_capacity, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def pb():
    import penguin_b200
    penguin_b200.init()
    return penguin_b200


def _phases(pb, mo, mg, ls, f, D):
    cap_o = geom.capacity(mo, ls)
    cap_g = import_capacity(pb, mg, cap_o)
    return po.Phase(cap_o, po.DiffusionOps(cap_o), f, D), pb.Phase(cap_g, pb.DiffusionOps(cap_g), f, D)


_ORACLE = {}


def _cached(key, fn):
    """oracle runs are shared between the kernel variants of one case (the oracle does not depend on them)"""
    if key not in _ORACLE:
        _ORACLE[key] = fn()
    return _ORACLE[key]


def _assert_fast_branch_ran(s):
    fast = [c["apply_cells_fast"] for c in s.ch]
    assert min(fast) > 0, "no tile took the interior constant-coefficient branch: the case does not test what the benchmark runs"
    assert all(c["converged"] for c in s.ch)


@pytest.fixture(scope="module")
def diph256(pb):
    """benchmark/Heat_2ph_2D.jl:64-111 at 256^2 (the size of bench.py's cpu_baseline leg): circle r = 2 in [0, 8]^2, 64 cells of radius"""
    nx = 256
    mo, mg = po.Mesh((nx, nx), (8.0, 8.0)), pb.Mesh((nx, nx), (8.0, 8.0))
    f = lambda x, y, z, t: 0.0 * x
    ls = geom.LevelSet.ball((4.0, 4.0), 2.0)
    p1o, p1g = _phases(pb, mo, mg, ls, f, 1.0)
    p2o, p2g = _phases(pb, mo, mg, ls.flipped(), f, 1.0)
    n = mo.n
    u0 = np.concatenate([np.ones(2 * n), np.zeros(2 * n)])
    dt = 0.5 * (8.0 / nx) ** 2
    ico = po.InterfaceConditions(po.ScalarJump(1.0, 1.0, 0.0), po.FluxJump(1.0, 1.0, 0.0))
    oracle = {}

    def states(scheme, nsteps):
        key = (scheme, nsteps)
        if key not in oracle:
            so = po.DiffusionUnsteadyDiph(p1o, p2o, po.BorderConditions(), ico, dt, u0, "BE")
            po.solve_DiffusionUnsteadyDiph(so, p1o, p2o, dt, (nsteps - 0.5) * dt, po.BorderConditions(), ico, scheme)
            oracle[key] = so.states
        return oracle[key]
    return dict(p1=p1g, p2=p2g, u0=u0, dt=dt, states=states)


# implementations of the Krylov iteration on the folded system (csrc/fold2.cuh): the default is the fused, TMA-staged one; the switches
# peel it back layer by layer so that a parity failure names its layer
# (fused_pipelined runs the band heads -- the interface-band work inside the two streaming kernels --, fused_band_launches the separate band kernels
#  that several ranks and large 3-D bands use)
VARIANTS = {"fused_pipelined": {}, "fused_band_launches": {"PB200_NO_BANDFUSE": "1"}, "fused_two_stage": {"PB200_NO_PIPE": "1"}, "unfused_tma": {"PB200_NO_FUSED": "1"}, "register_kernels": {"PB200_NO_TMA": "1"},
            "reference_pitch": {"PB200_NO_REPITCH": "1"}}


@pytest.fixture(params=list(VARIANTS))
def variant(request, monkeypatch):
    for k, v in VARIANTS[request.param].items():
        monkeypatch.setenv(k, v)
    return request.param


@pytest.mark.parametrize("scheme", ["BE", "CN"])
@pytest.mark.parametrize("check_every", [2, 8])
def test_diph_256_interior_tiles_vs_oracle(pb, diph256, scheme, check_every, variant):
    # check_every = 2 with a zero initial guess: the first solve is ~15 chunks of graph replay, later ones a long first chunk + short ones
    d = diph256
    ic = pb.InterfaceConditions(pb.ScalarJump(1.0, 1.0, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
    ref = d["states"](scheme, 4)
    s = pb.DiffusionUnsteadyDiph(d["p1"], d["p2"], pb.BorderConditions(), ic, d["dt"], d["u0"], "BE")
    pb.solve_DiffusionUnsteadyDiph_(s, d["p1"], d["p2"], d["dt"], 3.5 * d["dt"], pb.BorderConditions(), ic, scheme, reltol=1e-13, path="folded",
                                    warm_start=0, check_every=check_every)
    _assert_fast_branch_ran(s)
    assert len(s.states) == len(ref) == 5
    for a, b in zip(s.states, ref):
        assert rel_l2(a, b) < TOL
    if check_every == 2:
        assert s.ch[0]["iters"] > 4          # several chunks were replayed


def test_diph_256_at_the_bench_settings(pb, diph256):
    # bench.py: rtol = 1e-10 on the SCALED residual, warm_start = 4, check_every = 8.  Is ||r^|| <= 1e-10 ||b^|| enough for 1e-9 on the state?
    d = diph256
    ic = pb.InterfaceConditions(pb.ScalarJump(1.0, 1.0, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
    ref = d["states"]("BE", 12)
    s = pb.DiffusionUnsteadyDiph(d["p1"], d["p2"], pb.BorderConditions(), ic, d["dt"], d["u0"], "BE")
    pb.solve_DiffusionUnsteadyDiph_(s, d["p1"], d["p2"], d["dt"], 11.5 * d["dt"], pb.BorderConditions(), ic, "BE", reltol=1e-10, warm_start=4,
                                    check_every=8)
    _assert_fast_branch_ran(s)
    assert len(s.states) == len(ref) == 13
    worst = max(rel_l2(a, b) for a, b in zip(s.states, ref))
    assert worst < TOL, f"rel-L2 {worst:.3e} at the bench's stopping rule"
    its = [c["iters"] for c in s.ch]
    assert its[-1] < its[1]                  # the extrapolated guess pays off once it has a history


@pytest.mark.parametrize("poly", ["0", "1", "2"])
def test_mono_256_cn_polynomial_preconditioner_vs_oracle(pb, poly, monkeypatch, variant):
    # Heat-type monophasic problem, fluid OUTSIDE a circle (interface Dirichlet => no interface unknowns => the polynomial step, MODE 4,
    # runs on the interior tiles), Dirichlet borders, first step BE then CN (benchmark/Heat3D.jl:53-74 in 2-D)
    monkeypatch.setenv("PB200_POLY", poly)
    nx = 256
    mo, mg = po.Mesh((nx, nx), (4.0, 4.0)), pb.Mesh((nx, nx), (4.0, 4.0))
    f = lambda x, y, z, t: 0.25 * np.sin(x) * (1 + t)
    pho, phg = _phases(pb, mo, mg, geom.LevelSet.ball((2.01, 2.01), 0.5, False), f, 1.0)
    keys = ("left", "right", "top", "bottom")
    bco, bcg = po.BorderConditions({k: po.Dirichlet(1.0) for k in keys}), pb.BorderConditions({k: pb.Dirichlet(1.0) for k in keys})
    n = mo.n
    u0 = np.zeros(2 * n)
    dt = 0.75 * (4.0 / nx) ** 2
    def oracle():
        so = po.DiffusionUnsteadyMono(pho, bco, po.Dirichlet(1.0), dt, u0, "BE")
        po.solve_DiffusionUnsteadyMono(so, pho, dt, 3.5 * dt, bco, po.Dirichlet(1.0), "CN")
        return so.states
    ref = _cached("mono256", oracle)
    sg = pb.DiffusionUnsteadyMono(phg, bcg, pb.Dirichlet(1.0), dt, u0, "BE")
    pb.solve_DiffusionUnsteadyMono_(sg, phg, dt, 3.5 * dt, bcg, pb.Dirichlet(1.0), "CN", reltol=1e-13, path="folded", warm_start=2)
    _assert_fast_branch_ran(sg)
    assert len(sg.states) == len(ref) == 5
    for a, b in zip(sg.states, ref):
        assert rel_l2(a, b) < TOL


def test_mono_3d_interior_tiles_vs_oracle(pb, variant):
    # 72 x 20 x 12: x >= 64 so that a 32 x 8 x 4 tile of full cells exists away from the border ring; small sphere off to one side
    dims, L = (72, 20, 12), (4.0, 4.0, 4.0)
    mo, mg = po.Mesh(dims, L), pb.Mesh(dims, L)
    f = lambda x, y, z, t: 0.1 * y
    pho, phg = _phases(pb, mo, mg, geom.LevelSet.ball((0.9, 1.0, 2.0), 0.5, False), f, 1.0)
    keys = ("left", "right", "top", "bottom", "forward", "backward")
    bco, bcg = po.BorderConditions({k: po.Dirichlet(1.0) for k in keys}), pb.BorderConditions({k: pb.Dirichlet(1.0) for k in keys})
    n = mo.n
    u0 = np.zeros(2 * n)
    dt = 0.75 * (4.0 / 72) ** 2
    def oracle():
        so = po.DiffusionUnsteadyMono(pho, bco, po.Dirichlet(1.0), dt, u0, "BE")
        po.solve_DiffusionUnsteadyMono(so, pho, dt, 2.5 * dt, bco, po.Dirichlet(1.0), "CN")
        return so.states
    ref = _cached("mono3d", oracle)
    sg = pb.DiffusionUnsteadyMono(phg, bcg, pb.Dirichlet(1.0), dt, u0, "BE")
    pb.solve_DiffusionUnsteadyMono_(sg, phg, dt, 2.5 * dt, bcg, pb.Dirichlet(1.0), "CN", reltol=1e-13, path="folded")
    _assert_fast_branch_ran(sg)
    assert len(sg.states) == len(ref) == 4
    for a, b in zip(sg.states, ref):
        assert rel_l2(a, b) < TOL


def test_diph_3d_interior_tiles_vs_oracle(pb, variant):
    # examples/3D/Diffusion/Heat_2ph.jl:13-30 on a 72 x 24 x 12 slab of isotropic cells (h = 1/24): the sphere (r = 0.9) is thicker than the
    # slab, so the interface is two spherical caps across the whole y-z section and phase 1 holds a whole 32 x 8 x 4 tile of full cells
    dims, L = (72, 24, 12), (3.0, 1.0, 0.5)
    mo, mg = po.Mesh(dims, L), pb.Mesh(dims, L)
    f = lambda x, y, z, t: 0.0 * x
    ls = geom.LevelSet.ball((2.0, 0.5, 0.25), 0.9)
    p1o, p1g = _phases(pb, mo, mg, ls, f, 1.0)
    p2o, p2g = _phases(pb, mo, mg, ls.flipped(), f, 1.0)
    n = mo.n
    u0 = np.concatenate([np.ones(2 * n), np.zeros(2 * n)])
    dt = 0.5 * (3.0 / 72) ** 2
    ico = po.InterfaceConditions(po.ScalarJump(1.0, 2.0, 0.0), po.FluxJump(1.0, 1.0, 0.0))
    icg = pb.InterfaceConditions(pb.ScalarJump(1.0, 2.0, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
    def oracle():
        so = po.DiffusionUnsteadyDiph(p1o, p2o, po.BorderConditions(), ico, dt, u0, "BE")
        po.solve_DiffusionUnsteadyDiph(so, p1o, p2o, dt, 1.5 * dt, po.BorderConditions(), ico, "BE")
        return so.states
    ref = _cached("diph3d", oracle)
    sg = pb.DiffusionUnsteadyDiph(p1g, p2g, pb.BorderConditions(), icg, dt, u0, "BE")
    pb.solve_DiffusionUnsteadyDiph_(sg, p1g, p2g, dt, 1.5 * dt, pb.BorderConditions(), icg, "BE", reltol=1e-13, path="folded", warm_start=2)
    _assert_fast_branch_ran(sg)
    assert len(sg.states) == len(ref) == 3
    for a, b in zip(sg.states, ref):
        assert rel_l2(a, b) < TOL
