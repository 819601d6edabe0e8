import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import oracle
from icqsol_b200.mesh.generators import uvSphere
import scipy.linalg as sla

def pcg(M, b, apply_pre, tol, maxit):
    x = np.zeros_like(b); r = b.copy(); w = apply_pre(r); p = np.zeros_like(b)
    rho_old = 1.0; bmax = np.abs(b).max()
    for k in range(maxit):
        rho = r.dot(w)
        beta = rho/rho_old if k>0 else 0.0
        p = w + beta*p
        z = M.dot(p)
        alpha = rho/p.dot(z)
        x += alpha*p; r -= alpha*z
        w = apply_pre(r); rho_old = rho
        yield k+1, x, r

for nt in (48, 72, 100):
    xyz, tri = uvSphere(1.0, nt, nt//2)
    n = len(tri)
    t0=time.time(); G = oracle.green_matrix(xyz, tri, 5, threads=True); t1=time.time()
    A = oracle.area2(xyz, tri)
    M = A[:,None]*G   # symmetric negative definite
    M = 0.5*(M+M.T)
    cen = (xyz[tri[:,0]]+xyz[tri[:,1]]+xyz[tri[:,2]])/3
    v = 1/np.sqrt(cen[:,0]**2+(cen[:,1]-2)**2+(cen[:,2]-4)**2)
    b = -v   # G sigma = -v
    bs = A*b
    print('N',n,'assembly',t1-t0)
    def run(name, pre):
        for k,x,r in pcg(M, bs, pre, 0, 5000):
            # unscaled residual
            res = np.abs(r/A).max()/np.abs(b).max()
            if res <= 5e-13:
                true = np.abs(b-G.dot(x)).max()/np.abs(b).max()
                print('  %-22s iters %5d rec %.2e true %.2e'%(name,k,res,true)); return
        print('  %-22s no conv'%name, res)
    d = np.diag(M)
    run('jacobi', lambda r: r/d)
    for bs_ in (128, 256, 512):
        invs=[]
        for s in range(0,n,bs_):
            e=min(n,s+bs_); invs.append(np.linalg.inv(M[s:e,s:e]))
        def pre(r):
            out=np.empty_like(r)
            for q,s in enumerate(range(0,n,bs_)):
                e=min(n,s+bs_); out[s:e]=invs[q].dot(r[s:e])
            return out
        run('blockjacobi%d'%bs_, pre)
