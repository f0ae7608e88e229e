"""CG on the block-Jacobi-scaled system with a polynomial (Chebyshev) preconditioner restricted to the interface band."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from oracle import geom, penguin_oracle as po

def build(nx):
    mesh = po.Mesh((nx, nx), (8.0, 8.0)); ls = geom.LevelSet.ball((4.0, 4.0), 2.0)
    c1, c2 = geom.capacity(mesh, ls), geom.capacity(mesh, ls.flipped())
    o1, o2 = po.DiffusionOps(c1), po.DiffusionOps(c2)
    n = mesh.n; dt = 0.5 * (8.0 / nx) ** 2; one = np.ones(n)
    GG1, GH1, HG1, HH1 = po._blocks(o1, one); GG2, GH2, HG2, HH2 = po._blocks(o2, one)
    A = sp.bmat([[o1.V / dt + GG1, None, GH1], [None, o2.V / dt + GG2, GH2], [HG1, HG2, HH1 + HH2]], format="csr")
    keep = np.nonzero(np.asarray(abs(A).sum(1)).ravel() != 0)[0]
    Ar = A[keep][:, keep].tocsr()
    pos = -np.ones(3 * n, int); pos[keep] = np.arange(len(keep))
    rows, cols, vals = [], [], []
    for i in range(n):
        m = [q for q in (pos[i], pos[n + i], pos[2 * n + i]) if q >= 0]
        if not m: continue
        Li = np.linalg.inv(np.linalg.cholesky(Ar[m][:, m].toarray()))
        for a, qa in enumerate(m):
            for b, qb in enumerate(m):
                if Li[a, b] != 0: rows.append(qa); cols.append(qb); vals.append(Li[a, b])
    Linv = sp.csr_matrix((vals, (rows, cols)), shape=Ar.shape)
    Ah = (Linv @ Ar @ Linv.T).tocsr()
    b = Linv @ np.concatenate([c1.V / dt, 0 * c2.V, np.zeros(n)])[keep]
    wact = pos[2 * n:3 * n] >= 0
    bandcells = np.nonzero(wact)[0]
    band = np.array(sorted(q for i in bandcells for q in (pos[i], pos[n + i], pos[2 * n + i]) if q >= 0))
    return Ah, b, band

def cheb_band(Ah, band, deg, lo, hi):
    ABB = Ah[band][:, band].tocsr()
    theta, delta = (hi + lo) / 2, (hi - lo) / 2
    def apply(r):
        z = r.copy()
        rb = r[band]
        # Chebyshev iteration for ABB y = rb, y0 = 0
        y = np.zeros_like(rb); res = rb.copy()
        sigma = theta / delta; rho = 1 / sigma
        d = res / theta
        for k in range(deg):
            y = y + d
            res = res - ABB @ d
            rho_n = 1 / (2 * sigma - rho)
            d = rho_n * rho * d + 2 * rho_n / delta * res
            rho = rho_n
        z[band] = y
        return z
    return apply

for nx in [int(a) for a in sys.argv[1:]] or [128, 256]:
    Ah, b, band = build(nx)
    cnt = [0]
    def cb(x): cnt[0] += 1
    x, info = spla.cg(Ah, b, rtol=1e-10, atol=0, maxiter=2000, callback=cb)
    print(nx, "plain CG on scaled system:", cnt[0], "band unknowns", len(band), "of", Ah.shape[0])
    ABB = Ah[band][:, band]
    lmin = spla.eigsh(ABB.tocsc(), k=1, sigma=0, which="LM", return_eigenvectors=False)[0]; lmax = spla.eigsh(ABB, k=1, which="LA", return_eigenvectors=False)[0]
    print("   band block spectrum", lmin, lmax)
    for deg in (1, 2, 3, 4, 6):
        M = spla.LinearOperator(Ah.shape, cheb_band(Ah, band, deg, 0.9 * lmin, 1.05 * lmax))
        cnt[0] = 0
        x, info = spla.cg(Ah, b, rtol=1e-10, atol=0, maxiter=2000, M=M, callback=cb)
        print("   cheb deg", deg, "its", cnt[0], "res", np.linalg.norm(Ah @ x - b) / np.linalg.norm(b))
    # exact band solve
    lu = spla.splu(ABB.tocsc())
    def ex(r):
        z = r.copy(); z[band] = lu.solve(r[band]); return z
    cnt[0] = 0
    x, info = spla.cg(Ah, b, rtol=1e-10, atol=0, maxiter=2000, M=spla.LinearOperator(Ah.shape, ex), callback=cb)
    print("   exact band solve its", cnt[0])
