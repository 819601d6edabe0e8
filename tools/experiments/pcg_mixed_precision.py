import sys, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import oracle
from icqsol_b200.mesh.generators import uvSphere
nt=100
xyz, tri = uvSphere(1.0, nt, nt//2); n=len(tri)
G = oracle.green_matrix(xyz, tri, 5, threads=True)
A = oracle.area2(xyz, tri)
G32=G.astype(np.float32).astype(np.float64)
cen=(xyz[tri[:,0]]+xyz[tri[:,1]]+xyz[tri[:,2]])/3
v=1/np.sqrt(cen[:,0]**2+(cen[:,1]-2)**2+(cen[:,2]-4)**2)
braw=-v; b=A*braw
bs=128
invs=[]
for s in range(0,n,bs):
    e=min(n,s+bs); Mb=A[s:e,None]*G[s:e,s:e]; invs.append(np.linalg.inv(0.5*(Mb+Mb.T)))
def pre(r):
    out=np.empty_like(r)
    for q,s in enumerate(range(0,n,bs)):
        e=min(n,s+bs); out[s:e]=invs[q]@r[s:e]
    return out
thr=5e-13*np.abs(braw).max()
def run(switch):
    x=np.zeros(n); r=b.copy(); w=pre(r); p=np.zeros(n); rho_old=1; n64=n32=0
    for k in range(1000):
        rho=r.dot(w); beta=rho/rho_old if k>0 else 0.0
        p=w+beta*p
        rel=np.abs(r/A).max()/np.abs(braw).max()
        if rel<switch: z=A*(G32@p); n32+=1
        else: z=A*(G@p); n64+=1
        alpha=rho/p.dot(z); x+=alpha*p; r-=alpha*z; w=pre(r); rho_old=rho
        if np.abs(r/A).max()<=thr: break
    true=np.abs(braw-G@x).max()/np.abs(braw).max()
    print('switch %.0e: iters %d (fp64 %d, fp32 %d) true backward %.2e  -> traffic %.1f fp64-products'%(switch,k+1,n64,n32,true,n64+0.5*n32))
for sw in (0, 1e-8, 1e-6, 1e-4, 1e-2, 10):
    run(sw)
