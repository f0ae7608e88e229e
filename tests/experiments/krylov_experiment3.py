"""Where do the small eigenvalues of the block-Jacobi-scaled diphasic system live?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from oracle import geom, penguin_oracle as po
nx = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mesh = po.Mesh((nx, nx), (8.0, 8.0)); ls = geom.LevelSet.ball((4.0, 4.0), 2.0)
c1, c2 = geom.capacity(mesh, ls), geom.capacity(mesh, ls.flipped())
o1, o2 = po.DiffusionOps(c1), po.DiffusionOps(c2)
n = mesh.n; dt = 0.5 * (8.0 / nx) ** 2; one = np.ones(n)
GG1, GH1, HG1, HH1 = po._blocks(o1, one); GG2, GH2, HG2, HH2 = po._blocks(o2, one)
A = sp.bmat([[o1.V / dt + GG1, None, GH1], [None, o2.V / dt + GG2, GH2], [HG1, HG2, HH1 + HH2]], format="csr")
keep = np.nonzero(np.asarray(abs(A).sum(1)).ravel() != 0)[0]
Ar = A[keep][:, keep].tocsr()
pos = -np.ones(3 * n, int); pos[keep] = np.arange(len(keep))
# block-Jacobi Cholesky scaling
rows, cols, vals = [], [], []
for i in range(n):
    m = [q for q in (pos[i], pos[n + i], pos[2 * n + i]) if q >= 0]
    if not m: continue
    Li = np.linalg.inv(np.linalg.cholesky(Ar[m][:, m].toarray()))
    for a, qa in enumerate(m):
        for b, qb in enumerate(m):
            if Li[a, b] != 0: rows.append(qa); cols.append(qb); vals.append(Li[a, b])
Linv = sp.csr_matrix((vals, (rows, cols)), shape=Ar.shape)
Ah = (Linv @ Ar @ Linv.T).tocsc()
lam, vec = spla.eigsh(Ah, k=8, sigma=0, which="LM")
lmax = spla.eigsh(Ah, k=1, which="LA", return_eigenvectors=False)[0]
print("nx", nx, "lambda min", lam, "max", lmax, "cond", lmax / lam[0])
typ = keep // n; cell = keep % n
ct1 = c1.cell_types[cell]
for k in range(4):
    v = vec[:, k] ** 2
    print(f" eig {lam[k]:.4f}: energy u1 {v[typ==0].sum():.3f} u2 {v[typ==1].sum():.3f} w {v[typ==2].sum():.3f} | in cut cells {v[ct1==-1].sum():.3f} | participation {1.0/np.sum(v**2):.1f} unknowns")
# pure bulk reference: spectrum of the scaled operator restricted to non-band cells
