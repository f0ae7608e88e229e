"""Bisecting helper for the advection-diffusion parity (GPU): one BE / steady solve per case, rel-L2 vs the oracle printed."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import penguin_b200 as pb
from oracle import geom, penguin_oracle as po
from helpers import import_capacity, rel_l2, to_oracle_borders

pb.init()
KW = dict(reltol=1e-13, maxiter=50000)


def run(tag, dims, L, body, borders, ifc, uo_fn, ug_scale, steady):
    mo, mg = po.Mesh(dims, L), pb.Mesh(dims, L)
    N, n = mo.N, mo.n
    cap_o = po.nobody_capacity(mo) if body is None else geom.capacity(mo, body)
    cap_g = import_capacity(pb, mg, cap_o)
    uo = uo_fn(n, N)
    ug = ug_scale * np.random.default_rng(3).standard_normal(N * n)
    opo, opg = po.ConvectionOps(cap_o, uo, ug), pb.ConvectionOps(cap_g, uo, ug)
    f = (lambda x, y, z: 1.0 + 0 * x) if steady else (lambda x, y, z, t: 1.0 + 0 * x)
    pho, phg = po.Phase(cap_o, opo, f, 0.7), pb.Phase(cap_g, opg, f, 0.7)
    bcb = pb.BorderConditions({k: pb.Dirichlet(1.0) for k in borders})
    bci_g = pb.Dirichlet(0.5) if ifc == "d" else pb.Robin(1.0, 0.5, 0.25)
    bci_o = po.Dirichlet(0.5) if ifc == "d" else po.Robin(1.0, 0.5, 0.25)
    if steady:
        so = po.solve_AdvectionDiffusionSteadyMono(po.AdvectionDiffusionSteadyMono(pho, to_oracle_borders(pb, bcb), bci_o))
        sg = pb.solve_AdvectionDiffusionSteadyMono_(pb.AdvectionDiffusionSteadyMono(phg, bcb, bci_g), **KW)
        print(f"{tag:44s} steady   err {rel_l2(sg.x, so.x):.3e}  bulk {rel_l2(sg.x[:n], so.x[:n]):.3e}  iters {sg.ch[-1]['iters']}", flush=True)
    else:
        T0 = np.concatenate([np.random.default_rng(5).random(n), np.zeros(n)])
        so = po.AdvectionDiffusionUnsteadyMono(pho, to_oracle_borders(pb, bcb), bci_o, 0.02, T0, "BE")
        po.solve_AdvectionDiffusionUnsteadyMono(so, pho, 0.02, 0.01, to_oracle_borders(pb, bcb), bci_o, "BE")
        sg = pb.AdvectionDiffusionUnsteadyMono(phg, bcb, bci_g, 0.02, T0, "BE")
        pb.solve_AdvectionDiffusionUnsteadyMono_(sg, phg, 0.02, 0.01, bcb, bci_g, "BE", **KW)
        for k, (a, b) in enumerate(zip(sg.states, so.states)):
            print(f"{tag:44s} state {k}  err {rel_l2(a, b):.3e}  bulk {rel_l2(a[:n], b[:n]):.3e}", flush=True)


uni = lambda n, N: [np.full(n, v) for v in (0.8, -0.5, 0.3)[:N]]
zero = lambda n, N: [np.zeros(n) for _ in range(N)]
ball2 = geom.LevelSet.ball((2.0, 2.1), 1.0)
B4 = ("left", "right", "top", "bottom")
run("nobody, no borders, u uniform, unsteady", (12, 10), (4.0, 4.0), None, (), "d", uni, 0.0, False)
run("nobody, 4 borders, u uniform, unsteady", (12, 10), (4.0, 4.0), None, B4, "d", uni, 0.0, False)
run("nobody, 4 borders, u uniform, steady", (12, 10), (4.0, 4.0), None, B4, "d", uni, 0.0, True)
run("ball D, no borders, u uniform, ug 0, unsteady", (12, 10), (4.0, 4.0), ball2, (), "d", uni, 0.0, False)
run("ball D, no borders, u 0, ug 0.3, unsteady", (12, 10), (4.0, 4.0), ball2, (), "d", zero, 0.3, False)
run("ball R, no borders, u uniform, ug 0, unsteady", (12, 10), (4.0, 4.0), ball2, (), "r", uni, 0.0, False)
run("ball R, no borders, u 0, ug 0.3, unsteady", (12, 10), (4.0, 4.0), ball2, (), "r", zero, 0.3, False)
run("ball D, 4 borders, u uniform, ug 0, steady", (12, 10), (4.0, 4.0), ball2, B4, "d", uni, 0.0, True)
run("1-D nobody, u uniform, unsteady", (16,), (4.0,), None, (), "d", uni, 0.0, False)


def show(tag, dims, L, borders):
    mo, mg = po.Mesh(dims, L), pb.Mesh(dims, L)
    N, n = mo.N, mo.n
    cap_o = po.nobody_capacity(mo)
    cap_g = import_capacity(pb, mg, cap_o)
    uo = uni(n, N)
    ug = np.zeros(N * n)
    opo, opg = po.ConvectionOps(cap_o, uo, ug), pb.ConvectionOps(cap_g, uo, ug)
    f = lambda x, y, z, t: 0.0 * x
    pho, phg = po.Phase(cap_o, opo, f, 0.7), pb.Phase(cap_g, opg, f, 0.7)
    bcb = pb.BorderConditions({k: pb.Dirichlet(1.0) for k in borders})
    T0 = np.concatenate([np.linspace(0.0, 1.0, n) ** 2, np.zeros(n)])
    so = po.AdvectionDiffusionUnsteadyMono(pho, to_oracle_borders(pb, bcb), po.Dirichlet(0.5), 0.02, T0, "BE")
    po.solve_AdvectionDiffusionUnsteadyMono(so, pho, 0.02, 0.01, to_oracle_borders(pb, bcb), po.Dirichlet(0.5), "BE")
    sg = pb.AdvectionDiffusionUnsteadyMono(phg, bcb, pb.Dirichlet(0.5), 0.02, T0, "BE")
    pb.solve_AdvectionDiffusionUnsteadyMono_(sg, phg, 0.02, 0.01, bcb, pb.Dirichlet(0.5), "BE", **KW)
    np.set_printoptions(precision=5, linewidth=200)
    print(tag, "device", sg.states[1][:n])
    print(tag, "oracle", so.states[1][:n])
    print(tag, "diff  ", sg.states[1][:n] - so.states[1][:n])


show("1-D bottom", (8,), (4.0,), ("bottom",))
show("1-D top", (8,), (4.0,), ("top",))
