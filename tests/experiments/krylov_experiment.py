"""CPU experiment (design aid, not product): Krylov behaviour of the eliminated diphasic system [u1, u2, w]."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from oracle import geom, penguin_oracle as po

def system(nx, c=None):
    mesh = po.Mesh((nx, nx), (8.0, 8.0))
    ls = geom.LevelSet.ball((4.0, 4.0), 2.0)
    c1, c2 = geom.capacity(mesh, ls), geom.capacity(mesh, ls.flipped())
    o1, o2 = po.DiffusionOps(c1), po.DiffusionOps(c2)
    n = mesh.n
    dt = 0.5 * (8.0 / nx) ** 2
    one = np.ones(n)
    GG1, GH1, HG1, HH1 = po._blocks(o1, one)
    GG2, GH2, HG2, HH2 = po._blocks(o2, one)
    a1 = a2 = b1 = b2 = 1.0
    k = a2 / a1
    A = sp.bmat([[o1.V + dt * GG1, None, dt * k * GH1],
                 [None, o2.V + dt * GG2, dt * GH2],
                 [b1 * HG1, b2 * HG2, b1 * k * HH1 + b2 * HH2]], format="csr")
    rhs = np.concatenate([c1.V * 1.0, c2.V * 0.0, np.zeros(n)])
    absA = abs(A)
    keep = np.nonzero((np.asarray(absA.sum(1)).ravel() != 0) & (np.asarray(absA.sum(0)).ravel() != 0))[0]
    A = A[keep][:, keep].tocsr()
    rhs = rhs[keep]
    # row scalings that symmetrise: rows1 * b1/(dt k), rows2 * b2/dt, rows w * 1
    s = np.concatenate([np.full(n, b1 / (dt * k)), np.full(n, b2 / dt), np.ones(n)])[keep]
    return A, rhs, s

def run(nx):
    A, b, s = system(nx)
    xref = spla.splu(A.tocsc()).solve(b)
    cnt = [0]
    def cb(x): cnt[0] += 1
    # (1) Jacobi right-preconditioned BiCGSTAB on the unsymmetrised system
    d = A.diagonal()
    M = spla.LinearOperator(A.shape, lambda v: v / d)
    cnt[0] = 0
    x, info = spla.bicgstab(A, b, rtol=1e-10, atol=0, maxiter=3000, M=M, callback=cb)
    print(nx, "bicgstab unsym jacobi: its", cnt[0], "info", info, "err", np.linalg.norm(x - xref) / np.linalg.norm(xref), "res", np.linalg.norm(A @ x - b) / np.linalg.norm(b))
    # (2) symmetrised + symmetric Jacobi scaling, CG and BiCGSTAB
    As = sp.diags(s) @ A
    asym = abs(As - As.T).max() / abs(As).max()
    ds = As.diagonal()
    S = sp.diags(1 / np.sqrt(ds))
    Ah = (S @ As @ S).tocsr()
    bh = S @ (s * b)
    for name, fn in (("cg", spla.cg), ("bicgstab", spla.bicgstab)):
        cnt[0] = 0
        xh, info = fn(Ah, bh, rtol=1e-10, atol=0, maxiter=3000, callback=cb)
        x = S @ xh
        print(nx, name, "sym scaled: its", cnt[0], "info", info, "asym", asym, "err", np.linalg.norm(x - xref) / np.linalg.norm(xref), "res(orig)", np.linalg.norm(A @ x - b) / np.linalg.norm(b))
    ev = spla.eigsh(Ah, k=1, which="LA", return_eigenvectors=False)[0]
    evs = spla.eigsh(Ah, k=1, sigma=0, which="LM", return_eigenvectors=False)[0]
    print(nx, "lambda max", ev, "lambda min", evs, "cond", ev / evs)

for nx in [int(a) for a in sys.argv[1:]] or [64, 128, 256]:
    run(nx)
