"""Backward-Euler time stepping with the cubic extrapolated initial guess on the oracle's diphasic system: CG iterations per step for
the preconditioner candidates none / band (device default) / polynomial only / sums of the two.  Usage: python krylov_experiment6.py [nx] [steps]
(see krylov_experiment5.py for the findings)"""
import sys, os, time
HERE=os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from oracle import geom, penguin_oracle as po
# build like krylov_experiment4.build but keep pieces to time-step: A x^{n+1} = (V/dt) x^n on bulk rows
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
    Vd = np.concatenate([c1.V / dt, c2.V / dt, np.zeros(n)])[keep]
    wact = pos[2 * n:3 * n] >= 0
    bandcells = np.nonzero(wact)[0]
    band = np.array(sorted(q for i in bandcells for q in (pos[i], pos[n + i], pos[2 * n + i]) if q >= 0))
    x0 = np.concatenate([np.ones(n), np.zeros(n), np.zeros(n)])[keep]   # T1 = 1, T2 = 0
    return Ah, Linv, Vd, band, x0
NX=int(sys.argv[1]) if len(sys.argv) > 1 else 96
NSTEPS=int(sys.argv[2]) if len(sys.argv) > 2 else 14
Ah,Linv,Vd,band,x0=build(NX)
n=Ah.shape[0]
ABB=Ah[band][:,band].tocsr()
lmin=spla.eigsh(ABB.tocsc(),k=1,sigma=0,which='LM',return_eigenvectors=False)[0]; lmax=spla.eigsh(ABB,k=1,which='LA',return_eigenvectors=False)[0]
gl=spla.eigsh(Ah,k=1,which='LA',return_eigenvectors=False)[0]
# device-like q_B: degree-1 Chebyshev on [0.9 lmin, 1.05 lmax]
lo_b,hi_b=0.9*lmin,1.05*lmax
th,de=(hi_b+lo_b)/2,(hi_b-lo_b)/2; sg=th/de; r0=1/sg; r1=1/(2*sg-r0)
pa0=(1+r1*r0)/th+2*r1/de; pa1=-2*r1/(de*th)
def qB(v):   # q_B(A_BB) v_B (band vector in, band vector out)
    return pa0*v+pa1*(ABB@v)
lo,hi=1/3,1.03*gl
th,de=(hi+lo)/2,(hi-lo)/2; sg=th/de; r0=1/sg; r1=1/(2*sg-r0)
cr=(1+r1*r0)/th+2*r1/de; cA=-2*r1/(de*th)
q=lambda r: cr*r+cA*(Ah@r)
def P_band(r):
    z=r.copy(); z[band]=qB(r[band]); return z
def P_sum_minusI(r):
    z=q(r); z[band]+=qB(r[band])-r[band]; return z
def P_sum(r, w=1.0):
    z=q(r); z[band]+=w*qB(r[band]); return z
cands={'none':None,'band (current)':P_band,'q only':q,'q + (qB - I)':P_sum_minusI,'q + qB':P_sum,'q + 0.5 qB':lambda r:P_sum(r,0.5)}
# time stepping with cubic extrapolated guess (LinvT scaling: xhat = L^T x; rhs bhat = Linv (Vd * x))
Lm=sp.linalg.inv(Linv.tocsc()).tocsr() if False else None
# unknown in scaled space: Ah xh = Linv (Vd .* x), x = Linv^T xh
hist=[]
x=x0.copy()
cnt=[0]
def cb(_): cnt[0]+=1
LinvT=Linv.T.tocsr()
# xh for initial x: solve LinvT xh = x  (L^T xh... ) use spsolve once per need
from scipy.sparse.linalg import splu
lu=splu(LinvT.tocsc())
res_tab={k:[] for k in cands}
for step in range(NSTEPS):
    bh=Linv@(Vd*x)
    # guess in scaled space from history of scaled solutions
    if len(hist)>=4: g=4*hist[-1]-6*hist[-2]+4*hist[-3]-hist[-4]
    elif len(hist)>=1: g=hist[-1]
    else: g=lu.solve(x)
    sol=None
    for name,P in cands.items():
        cnt[0]=0
        xs,info=spla.cg(Ah,bh,x0=g,rtol=1e-10,atol=0,maxiter=400,M=None if P is None else spla.LinearOperator(Ah.shape,P),callback=cb)
        res_tab[name].append(cnt[0])
        if name=='band (current)': sol=xs
    hist.append(sol); x=LinvT@sol
for k,v in res_tab.items(): print('%-16s'%k, v)
