import sys, os, numpy as np
sys.path.insert(0, '/root/repo')
import penguin_b200 as pb
pb.init()
dims, L = (1024, 1024), (8.0, 8.0)
mesh = pb.Mesh(dims, L)
body = pb.Balls([[4.0, 4.0]], [2.0])
c1, c2 = pb.Capacity(body, mesh, compute_centroids=False), pb.Capacity(-body, mesh, compute_centroids=False)
p1, p2 = pb.Phase(c1, pb.DiffusionOps(c1), 0.0, 1.0), pb.Phase(c2, pb.DiffusionOps(c2), 0.0, 1.0)
n = c1.nloc
dt = 0.5 * (L[0] / dims[0]) ** 2
ic = pb.InterfaceConditions(pb.ScalarJump(1.0, 2.0, 0.0), pb.FluxJump(1.0, 1.0, 0.0))
u0 = np.concatenate([np.ones(2 * n), np.zeros(2 * n)])
s = pb.DiffusionUnsteadyDiph(p1, p2, pb.BorderConditions(), ic, dt, u0, "BE")
pb.solve_DiffusionUnsteadyDiph_(s, p1, p2, dt, 0.5 * dt, pb.BorderConditions(), ic, "BE", reltol=1e-12, maxiter=200)
print(os.environ.get("TAG"), [(c["iters"], c["converged"], c["rnorm"] / max(c["bnorm"], 1e-300)) for c in s.ch])
