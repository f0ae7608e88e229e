import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import penguin_b200 as pb
from oracle import geom, penguin_oracle as po
from helpers import import_capacity
pb.init()
mo, mg = po.Mesh((16, 16), (4.0, 4.0)), pb.Mesh((16, 16), (4.0, 4.0))
f = lambda x, y, z: 1.0 + 0 * x
cap_o = geom.capacity(mo, geom.LevelSet.ball((2.0, 2.0), 1.0))
pho = po.Phase(cap_o, po.DiffusionOps(cap_o), f, 1.0)
cap_g = import_capacity(pb, mg, cap_o)
phg = pb.Phase(cap_g, pb.DiffusionOps(cap_g), f, 1.0)
so = po.solve_DiffusionSteadyMono(po.DiffusionSteadyMono(pho, po.BorderConditions(), po.Dirichlet(0.0)))
sg = pb.solve_DiffusionSteadyMono_(pb.DiffusionSteadyMono(phg, pb.BorderConditions(), pb.Dirichlet(0.0)), reltol=1e-13, maxiter=50000)
zg, zo = sg.x == 0.0, so.x == 0.0
bad = np.nonzero(zg != zo)[0]
n = mo.n
print("n", n, "mismatch idx", bad)
for i in bad:
    j = i % n
    print(i, "block", i // n, "cell", (j % 17, j // 17), "gpu", sg.x[i], "oracle", so.x[i], "V", cap_o.V[j], "ct", cap_o.cell_types[j], "Gam", cap_o.Gamma[j],
          "B", [b[j] for b in cap_o.B], "A", [a[j] for a in cap_o.A])
print(sg.ch)
