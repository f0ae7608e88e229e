"""Sizes at which every block of the tile kernels walks SEVERAL tiles (stage reuse of the TMA pipelines, empty / full barrier phases,
field changes between consecutive tiles of a block) -- beyond what the oracle's sparse LU does in seconds.  The check is the agreement
of the kernel variants with the register kernels on the reference pitch (the round-1 implementation, itself oracle-checked at the sizes
of test_gpu_solver_parity.py / test_gpu_fastpath_parity.py): same states to 1e-9, same iteration counts +- 1."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

VARIANTS = {"fused_pipelined": {}, "fused_band_launches": {"PB200_NO_BANDFUSE": "1"}, "fused_band_heads_all_blocks": {"PB200_BANDFUSE_HB": "0"},
            "fused_two_stage": {"PB200_NO_PIPE": "1"}, "unfused_tma": {"PB200_NO_FUSED": "1"}}
BASE = {"PB200_NO_TMA": "1", "PB200_NO_REPITCH": "1"}


@pytest.fixture(scope="module")
def pb():
    import penguin_b200
    penguin_b200.init()
    return penguin_b200


def _run_diph(pb, dims, L, center, radius, nsteps, scheme="BE", **kw):
    mesh = pb.Mesh(dims, L)
    body = pb.Balls([list(center)], [radius])
    c1, c2 = pb.Capacity(body, mesh, compute_centroids=False), pb.Capacity(-body, mesh, compute_centroids=False)
    p1, p2 = pb.Phase(c1, pb.DiffusionOps(c1), 0.0, 1.0), pb.Phase(c2, pb.DiffusionOps(c2), 0.0, 1.0)
    n = c1.nloc
    dt = 0.5 * (L[0] / dims[0]) ** 2
    ic = pb.InterfaceConditions(pb.ScalarJump(1.0, 2.0, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
    u0 = np.concatenate([np.ones(2 * n), np.zeros(2 * n)])
    s = pb.DiffusionUnsteadyDiph(p1, p2, pb.BorderConditions(), ic, dt, u0, "BE")
    pb.solve_DiffusionUnsteadyDiph_(s, p1, p2, dt, (nsteps - 0.5) * dt, pb.BorderConditions(), ic, scheme, reltol=1e-12, **kw)
    assert all(c["converged"] for c in s.ch), [(c["iters"], c["rnorm"] / max(c["bnorm"], 1e-300)) for c in s.ch]
    return s


def _run_mono(pb, dims, L, center, radius, nsteps, **kw):
    mesh = pb.Mesh(dims, L)
    cap = pb.Capacity(-pb.Balls([list(center)], [radius]), mesh, compute_centroids=False)
    ph = pb.Phase(cap, pb.DiffusionOps(cap), 0.0, 1.0)
    n = cap.nloc
    dt = 0.75 * (L[0] / dims[0]) ** 2
    keys = ("left", "right", "top", "bottom", "forward", "backward")[:2 * len(dims)]
    bc = pb.BorderConditions({k: pb.Dirichlet(1.0) for k in keys})
    s = pb.DiffusionUnsteadyMono(ph, bc, pb.Dirichlet(1.0), dt, np.zeros(2 * n), "BE")
    pb.solve_DiffusionUnsteadyMono_(s, ph, dt, (nsteps - 0.5) * dt, bc, pb.Dirichlet(1.0), "CN", reltol=1e-12, **kw)
    assert all(c["converged"] for c in s.ch)
    return s


CASES = {
    "diph2d_1024": lambda pb, **kw: _run_diph(pb, (1024, 1024), (8.0, 8.0), (4.0, 4.0), 2.0, 3, **kw),
    "diph2d_1024_cn_warm": lambda pb, **kw: _run_diph(pb, (1024, 1024), (8.0, 8.0), (4.0, 4.0), 2.0, 5, scheme="CN", warm_start=3, **kw),
    "diph3d_160x128x64": lambda pb, **kw: _run_diph(pb, (160, 128, 64), (5.0, 4.0, 2.0), (2.5, 2.0, 1.0), 0.8, 3, **kw),
    "mono2d_1024": lambda pb, **kw: _run_mono(pb, (1024, 1024), (4.0, 4.0), (2.01, 2.01), 0.5, 3, **kw),
    "mono3d_160x128x64": lambda pb, **kw: _run_mono(pb, (160, 128, 64), (5.0, 4.0, 2.0), (2.5, 2.0, 1.0), 0.6, 3, **kw),
}
_BASE = {}


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("variant", list(VARIANTS))
def test_variants_agree_with_the_register_kernels(pb, monkeypatch, case, variant):
    if case not in _BASE:
        with monkeypatch.context() as m:
            for k, v in BASE.items():
                m.setenv(k, v)
            b = CASES[case](pb)
            _BASE[case] = ([st.copy() for st in b.states], [c["iters"] for c in b.ch])
    ref_states, ref_iters = _BASE[case]
    for k, v in VARIANTS[variant].items():
        monkeypatch.setenv(k, v)
    s = CASES[case](pb)
    assert len(s.states) == len(ref_states)
    for a, b in zip(s.states, ref_states):
        assert np.linalg.norm(a - b) <= 1e-9 * np.linalg.norm(b)
    its = [c["iters"] for c in s.ch]
    assert max(abs(i - j) for i, j in zip(its, ref_iters)) <= 1, (its, ref_iters)
    assert min(c["apply_cells_fast"] for c in s.ch) > 0
