"""CPU experiments behind the polynomial preconditioner of the folded CG, on the oracle's matrices.

Part 1 (zero initial guess): CG on the block-Jacobi-scaled diphasic system with P = q_m(BA) B, q_m(A) + (B - I), B q_m(A) B for the band
preconditioner B; counts matvecs and the bytes per unknown a fused implementation moves (outer CG iteration 80 B, polynomial step 32 B).
Part 2 (run with `steps` as second argument): backward-Euler time stepping with the cubic extrapolated initial guess and the candidates
none / band / q only / q + (q_B - I) / q + q_B.  Usage: python krylov_experiment5.py [nx] [steps]
Findings (96^2 .. 400^2): the polynomial alone needs about as many iterations as the band preconditioner alone; the sum with (q_B - I) is
indefinite (iterations grow with the grid); on the device at 2048^2 after the spin-up the band preconditioner wins (8.35 vs 12.8)."""
import sys, os, time
HERE=os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
src=open(os.path.join(HERE,'krylov_experiment4.py')).read().split("for nx in [int(a)")[0]
ns={'__file__':os.path.join(HERE,'krylov_experiment4.py')}; exec(compile(src,'k4','exec'),ns)
build, cheb_band = ns['build'], ns['cheb_band']
NX=int(sys.argv[1]) if len(sys.argv) > 1 else 96
t=time.time(); Ah,b,band=build(NX); print('build',time.time()-t, Ah.shape)
ABB=Ah[band][:,band]
lmin=spla.eigsh(ABB.tocsc(),k=1,sigma=0,which='LM',return_eigenvectors=False)[0]; lmax=spla.eigsh(ABB,k=1,which='LA',return_eigenvectors=False)[0]
B=cheb_band(Ah,band,1,0.9*lmin,1.05*lmax)   # current band preconditioner (degree 1)
n=Ah.shape[0]
cnt=[0]
def cb(x): cnt[0]+=1
x,info=spla.cg(Ah,b,rtol=1e-10,atol=0,maxiter=500,M=spla.LinearOperator(Ah.shape,B),callback=cb)
base=cnt[0]; print('current: CG + band prec: its',base,'bytes/unknown',base*80)
BA=spla.LinearOperator(Ah.shape, lambda v: B(Ah@v))
ev_hi=spla.eigs(BA,k=1,which='LM',return_eigenvectors=False)[0].real
# smallest: power iteration on (hi I - BA)
v=np.random.default_rng(0).standard_normal(n)
for _ in range(400):
    w=ev_hi*v-BA@v; v=w/np.linalg.norm(w)
ev_lo=ev_hi-np.linalg.norm(ev_hi*v-BA@v)
print('spec(BA) ~',ev_lo,ev_hi)
def chebprec(m,lo,hi):
    theta,delta=(hi+lo)/2,(hi-lo)/2
    def apply(r):
        z=np.zeros_like(r); res=r.copy()
        sigma=theta/delta; rho=1/sigma
        d=B(res)/theta
        for k in range(m):
            z=z+d
            res=res-Ah@d
            rho_n=1/(2*sigma-rho)
            d=rho_n*rho*d+2*rho_n/delta*B(res)
            rho=rho_n
        z=z+d
        return z
    return apply
for m in (1,2,3,4):
    for lo,hi in ((0.9*ev_lo,1.05*ev_hi),(0.3,1.75)):
        cnt[0]=0
        x,info=spla.cg(Ah,b,rtol=1e-10,atol=0,maxiter=500,M=spla.LinearOperator(Ah.shape,chebprec(m,lo,hi)),callback=cb)
        it=cnt[0]
        print('m',m,'interval %.2f-%.2f'%(lo,hi),'outer',it,'matvecs',it*(m+1),'bytes/unknown',it*(80+32*m),'vs',base*80,'ratio %.2f'%(it*(80+32*m)/(base*80)),'res %.1e'%(np.linalg.norm(Ah@x-b)/np.linalg.norm(b)))
print('--- variants with a pure-A inner polynomial (3-term recurrence form: 32 B per inner matvec)')
def polyA(m,lo,hi):
    theta,delta=(hi+lo)/2,(hi-lo)/2
    def apply(r):
        z=np.zeros_like(r); res=r.copy()
        sigma=theta/delta; rho=1/sigma
        d=res/theta
        for k in range(m):
            z=z+d
            res=res-Ah@d
            rho_n=1/(2*sigma-rho)
            d=rho_n*rho*d+2*rho_n/delta*res
            rho=rho_n
        return z+d
    return apply
def dzB(r):   # (B - I) r
    return B(r)-r
for m in (1,2,3):
    for lo,hi in ((0.3,1.8),(0.2,1.8)):
        q=polyA(m,lo,hi)
        for name,P in (('q+(B-I)',lambda r,q=q: q(r)+dzB(r)),('B q B',lambda r,q=q: B(q(B(r))))):
            cnt[0]=0
            try:
                x,info=spla.cg(Ah,b,rtol=1e-10,atol=0,maxiter=300,M=spla.LinearOperator(Ah.shape,P),callback=cb)
            except Exception as e:
                print(name,'fail',e); continue
            it=cnt[0]
            print('m',m,'[%.1f,%.1f]'%(lo,hi),'%-8s'%name,'outer',it,'matvecs',it*(m+1),'bytes',it*(80+32*m),'ratio %.2f'%(it*(80+32*m)/(base*80)),'res %.1e'%(np.linalg.norm(Ah@x-b)/np.linalg.norm(b)),'info',info)
