"""CPU experiment behind the multigrid preconditioner of the steady Poisson config (BASELINE.json configs[4]): rediscretised cut-cell operators on
coarsened meshes (the capacity code run on n/2, n/4, ...), cell-aggregation transfers, Chebyshev-polynomial smoothers in the Jacobi-scaled variables
(what the folded path applies), V-cycle as the CG preconditioner.  Prints CG iteration counts with / without it."""
import sys, os, time
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import geom, penguin_oracle as po
import importlib.util
spec = importlib.util.spec_from_file_location("rp", os.path.join(os.path.dirname(__file__), "..", "..", "tools", "run_poisson3d.py"))


def level(n, cen, rad):
    mesh = po.Mesh((n,) * 3, (4.0,) * 3)
    cap = geom.capacity(mesh, geom.LevelSet.balls(cen, rad, False), compute_centroids=False)
    op = po.DiffusionOps(cap)
    GG = (op.G.T @ op.Wdag @ op.G).tocsr()
    pd = mesh.pdims
    idx = np.arange(mesh.n).reshape(pd[::-1])          # x fastest
    border = np.zeros(pd[::-1], bool)
    for ax in range(3):
        sl = [slice(None)] * 3; sl[ax] = 0; border[tuple(sl)] = True
        sl[ax] = pd[::-1][ax] - 2; border[tuple(sl)] = True      # last real cell
        sl[ax] = pd[::-1][ax] - 1; border[tuple(sl)] = True      # padding
    rs = np.asarray(abs(GG).sum(axis=1)).ravel()
    act = (rs != 0) & ~border.ravel() & (cap.V > 0)
    return dict(n=n, mesh=mesh, cap=cap, GG=GG, act=act, pd=pd)


def transfer(lf, lc):
    """P: coarse cell -> its 8 children (piecewise constant) on the padded grids"""
    pf, pc = lf["pd"], lc["pd"]
    nf = np.prod(pf)
    k, j, i = np.meshgrid(np.arange(pf[2]), np.arange(pf[1]), np.arange(pf[0]), indexing="ij")
    ci, cj, ck = np.minimum(i // 2, pc[0] - 1), np.minimum(j // 2, pc[1] - 1), np.minimum(k // 2, pc[2] - 1)
    col = (ck * pc[1] + cj) * pc[0] + ci
    P = sp.csr_matrix((np.ones(nf), (np.arange(nf), col.ravel())), shape=(nf, np.prod(pc)))
    return P


def cheb(M, lo, hi, m):
    """z = q_m(M) r by the Chebyshev iteration from a zero guess"""
    theta, delta = 0.5 * (hi + lo), 0.5 * (hi - lo)
    sigma = theta / delta
    def apply(r):
        rho = 1.0 / sigma
        d = r / theta
        z = d.copy()
        for _ in range(m):
            rho_n = 1.0 / (2 * sigma - rho)
            d = rho_n * rho * d + 2 * rho_n / delta * (r - M @ z)
            z = z + d
            rho = rho_n
        return z
    return apply


def build(levels, m=2, alpha=8.0, pconst=True):
    H = []
    for l in levels:
        a = np.nonzero(l["act"])[0]
        M = l["GG"][a][:, a].tocsr()
        d = M.diagonal()
        s = 1.0 / np.sqrt(d)
        Mh = sp.diags(s) @ M @ sp.diags(s)
        # lambda_max by power iteration
        x = np.random.default_rng(0).standard_normal(len(a))
        for _ in range(30):
            y = Mh @ x; lam = (y @ y) / (y @ x); x = y / np.linalg.norm(y)
        hi = 1.03 * lam
        H.append(dict(a=a, M=M, Mh=Mh.tocsr(), s=s, hi=hi, sm=cheb(Mh.tocsr(), hi / alpha, hi, m), n=l["n"]))
    for q in range(len(levels) - 1):
        P = transfer(levels[q], levels[q + 1])
        H[q]["P"] = P[H[q]["a"]][:, H[q + 1]["a"]].tocsr()
    return H


def vcycle(H, q, rh, omega_c=1.0, coarse_sweeps=None, alpha_c=40.0):
    """in hat variables of level q: returns zh ~ Mh^-1 rh.  coarse_sweeps = None: exact coarsest solve (sparse LU); an integer: what csrc/mg.cuh does --
    one degree-2 Chebyshev polynomial on [hi / alpha_c, hi] from a zero guess followed by that many residual-correction sweeps with the same polynomial"""
    L = H[q]
    if q == len(H) - 1:
        if coarse_sweeps is not None:
            sm = L.setdefault("smc", cheb(L["Mh"], L["hi"] / alpha_c, L["hi"], 2))
            z = sm(rh)
            for _ in range(coarse_sweeps):
                z = z + sm(rh - L["Mh"] @ z)
            return z
        import scipy.sparse.linalg as spla
        if "lu" not in L:
            L["lu"] = spla.splu(L["Mh"].tocsc())
        return L["lu"].solve(rh)
    z = L["sm"](rh)
    res = rh - L["Mh"] @ z
    C = H[q + 1]
    rc = C["s"] * (L["P"].T @ (res / L["s"]))                 # r = S^-1 r^ (true residual); coarse r^_c = S_c r_c
    ec = vcycle(H, q + 1, rc, omega_c, coarse_sweeps, alpha_c)
    z = z + omega_c * (L["P"] @ (C["s"] * ec)) / L["s"]         # x = S x^ ; x^ = x / S
    res = rh - L["Mh"] @ z
    return z + L["sm"](res)


def pcg(Mh, b, prec, rtol=1e-10, maxit=5000):
    x = np.zeros_like(b); r = b.copy(); z = prec(r); p = z.copy(); rz = r @ z; nb = np.linalg.norm(b)
    for it in range(1, maxit + 1):
        q = Mh @ p; a = rz / (p @ q); x += a * p; r -= a * q
        if np.linalg.norm(r) <= rtol * nb:
            return x, it
        z = prec(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return x, maxit


if __name__ == "__main__":
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    cen, rad = m.random_spheres()
    n0 = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    ns = [n0]
    while ns[-1] > 8:
        ns.append(ns[-1] // 2)
    t0 = time.time()
    levels = [level(n, cen, rad) for n in ns]
    print("levels", ns, "built in %.1fs" % (time.time() - t0), "active", [int(l["act"].sum()) for l in levels])
    for mdeg, alpha in [(1, 4.0), (2, 8.0), (2, 16.0), (3, 16.0)]:
        H = build(levels, m=mdeg, alpha=alpha)
        L0 = H[0]
        b = L0["s"] * (levels[0]["cap"].V[L0["a"]] * 1.0)
        if mdeg == 1:
            _, it0 = pcg(L0["Mh"], b, lambda r: r)
            print("plain CG (Jacobi-scaled):", it0)
        for oc in (1.0, 1.5, 2.0):
            _, it = pcg(L0["Mh"], b, lambda r: vcycle(H, 0, r, oc))
            print(f"MG-PCG  smoother degree {mdeg} on [hi/{alpha:g}, hi], coarse over-correction {oc}: {it} iterations")
