import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import oracle
from icqsol_b200.mesh.generators import uvSphere

def pcg(M, b, pre, A, braw, tol=5e-13, maxit=3000):
    x = np.zeros_like(b); r = b.copy(); w = pre(r); p = np.zeros_like(b); rho_old=1
    for k in range(maxit):
        rho = r.dot(w); beta = rho/rho_old if k>0 else 0.0
        p = w + beta*p; z = M.dot(p); alpha = rho/p.dot(z)
        x += alpha*p; r -= alpha*z; w = pre(r); rho_old = rho
        if np.abs(r/A).max() <= tol*np.abs(braw).max(): return k+1
    return -1

nt=int(sys.argv[1]) if len(sys.argv)>1 else 100
xyz, tri = uvSphere(1.0, nt, nt//2); n=len(tri)
G = oracle.green_matrix(xyz, tri, 5, threads=True)
A = oracle.area2(xyz, tri)
M = A[:,None]*G; M=0.5*(M+M.T)
cen=(xyz[tri[:,0]]+xyz[tri[:,1]]+xyz[tri[:,2]])/3
v=1/np.sqrt(cen[:,0]**2+(cen[:,1]-2)**2+(cen[:,2]-4)**2)
braw=-v; b=A*braw
def blockpre(bs):
    invs=[np.linalg.inv(M[s:min(n,s+bs),s:min(n,s+bs)]) for s in range(0,n,bs)]
    def pre(r):
        out=np.empty_like(r)
        for q,s in enumerate(range(0,n,bs)):
            e=min(n,s+bs); out[s:e]=invs[q]@r[s:e]
        return out
    return pre
def coarse(agg):
    nc=(n+agg-1)//agg
    R=np.zeros((nc,n))
    for c in range(nc): R[c,c*agg:min(n,(c+1)*agg)]=1.0
    Ac=R@M@R.T
    Aci=np.linalg.inv(Ac)
    return lambda r: R.T@(Aci@(R@r))
print('N',n)
d=np.diag(M)
print('jacobi', pcg(M,b,lambda r:r/d,A,braw))
b128=blockpre(128)
print('block128', pcg(M,b,b128,A,braw))
for agg in (128,32,8):
    c=coarse(agg)
    print('block128 + coarse agg',agg, pcg(M,b,lambda r:b128(r)+c(r),A,braw))
    print('jacobi + coarse agg',agg, pcg(M,b,lambda r:r/d+c(r),A,braw))
