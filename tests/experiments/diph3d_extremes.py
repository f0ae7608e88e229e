"""3-D diphasic heat at nx^3 on device-built capacities: per-step parity with the oracle and the extremes of the state (the cut-cell
scheme does not bound the values of near-empty cut cells; this shows the device reproduces the oracle's extremes)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import penguin_b200 as pb
from oracle import geom, penguin_oracle as po
from helpers import rel_l2
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 24
pb.init()
mo, mg = po.Mesh((nx,) * 3, (4.0,) * 3), pb.Mesh((nx,) * 3, (4.0,) * 3)
body = pb.Sphere((2.0, 2.0, 2.0), 1.0)
c1, c2 = pb.Capacity(body, mg), pb.Capacity(-body, mg)
p1, p2 = pb.Phase(c1, pb.DiffusionOps(c1), 0.0, 1.0), pb.Phase(c2, pb.DiffusionOps(c2), 0.0, 1.0)
n = c1.nloc; h = 4.0 / nx; dt = 0.5 * h * h
u0 = np.concatenate([np.ones(2 * n), np.zeros(2 * n)])
ic = pb.InterfaceConditions(pb.ScalarJump(1.0, 2.0, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
s = pb.DiffusionUnsteadyDiph(p1, p2, pb.BorderConditions(), ic, dt, u0, "BE")
pb.solve_DiffusionUnsteadyDiph_(s, p1, p2, dt, 5.5 * dt, pb.BorderConditions(), ic, "BE", reltol=1e-12, warm_start=4)
ls = geom.LevelSet.ball((2.0, 2.0, 2.0), 1.0)
o1, o2 = geom.capacity(mo, ls), geom.capacity(mo, ls.flipped())
f = lambda x, y, z, t: 0.0 * x
q1, q2 = po.Phase(o1, po.DiffusionOps(o1), f, 1.0), po.Phase(o2, po.DiffusionOps(o2), f, 1.0)
ico = po.InterfaceConditions(po.ScalarJump(1.0, 2.0, 0.0), po.FluxJump(1.0, 1.0, 0.0))
so = po.DiffusionUnsteadyDiph(q1, q2, po.BorderConditions(), ico, dt, u0, "BE")
po.solve_DiffusionUnsteadyDiph(so, q1, q2, dt, 5.5 * dt, po.BorderConditions(), ico, "BE")
for k, (a, b) in enumerate(zip(s.states, so.states)):
    i = int(np.argmax(np.abs(b)))
    print(k, "rel L2", rel_l2(a, b), "| max|x| device", np.abs(a).max(), "oracle", np.abs(b).max(), "at block", i // n, "cell V1", o1.V[i % n], "V2", o2.V[i % n], flush=True)
pb.finalize()
