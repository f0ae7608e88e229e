"""CPU experiment: per-cell block-Jacobi (u1_i, u2_i, w_i) on the symmetrised diphasic system."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from oracle import geom, penguin_oracle as po

def run(nx, dtfac=0.5, ndim=2):
    if ndim == 2:
        mesh = po.Mesh((nx, nx), (8.0, 8.0)); ls = geom.LevelSet.ball((4.0, 4.0), 2.0)
    else:
        mesh = po.Mesh((nx, nx, nx), (4.0, 4.0, 4.0)); ls = geom.LevelSet.ball((2.0, 2.0, 2.0), 1.0)
    c1, c2 = geom.capacity(mesh, ls), geom.capacity(mesh, ls.flipped())
    o1, o2 = po.DiffusionOps(c1), po.DiffusionOps(c2)
    n = mesh.n
    dt = dtfac * (mesh.L[0] / nx) ** 2
    one = np.ones(n)
    GG1, GH1, HG1, HH1 = po._blocks(o1, one)
    GG2, GH2, HG2, HH2 = po._blocks(o2, one)
    # symmetrised: rows1 /dt, rows2 /dt (a=b=1)
    A = sp.bmat([[o1.V / dt + GG1, None, GH1], [None, o2.V / dt + GG2, GH2], [HG1, HG2, HH1 + HH2]], format="csr")
    rhs = np.concatenate([c1.V / dt, 0 * c2.V, np.zeros(n)])
    absA = abs(A)
    keepmask = (np.asarray(absA.sum(1)).ravel() != 0)
    keep = np.nonzero(keepmask)[0]
    Ar = A[keep][:, keep].tocsr()
    b = rhs[keep]
    xref = spla.splu(Ar.tocsc()).solve(b)
    # block Jacobi: per cell dense block over kept members of {i, n+i, 2n+i}
    pos = -np.ones(3 * n, int); pos[keep] = np.arange(len(keep))
    Acsc = Ar.tocsc()
    rows, cols, vals = [], [], []
    Ad = Ar.todok() if len(keep) < 20000 else None
    Alil = Ar.tolil() if Ad is None else None
    for i in range(n):
        m = [pos[i], pos[n + i], pos[2 * n + i]]
        m = [q for q in m if q >= 0]
        if not m: continue
        blk = Ar[m][:, m].toarray()
        inv = np.linalg.inv(blk)
        for a, qa in enumerate(m):
            for c_, qc in enumerate(m):
                rows.append(qa); cols.append(qc); vals.append(inv[a, c_])
    Minv = sp.csr_matrix((vals, (rows, cols)), shape=Ar.shape)
    cnt = [0]
    def cb(x): cnt[0] += 1
    M = spla.LinearOperator(Ar.shape, lambda v: Minv @ v)
    for name, fn in (("cg", spla.cg), ("bicgstab", spla.bicgstab)):
        cnt[0] = 0
        x, info = fn(Ar, b, rtol=1e-10, atol=0, maxiter=3000, M=M, callback=cb)
        print(ndim, nx, name, "block-jacobi: its", cnt[0], "info", info, "err", np.linalg.norm(x - xref) / np.linalg.norm(xref), "res", np.linalg.norm(Ar @ x - b) / np.linalg.norm(b))
    d = Ar.diagonal()
    Mj = spla.LinearOperator(Ar.shape, lambda v: v / d)
    cnt[0] = 0
    x, info = spla.cg(Ar, b, rtol=1e-10, atol=0, maxiter=3000, M=Mj, callback=cb)
    print(ndim, nx, "cg point-jacobi: its", cnt[0], "err", np.linalg.norm(x - xref) / np.linalg.norm(xref))

if __name__ == "__main__":
    nd = int(sys.argv[1])
    for nx in [int(a) for a in sys.argv[2:]]:
        run(nx, ndim=nd)
