import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import oracle
from icqsol_b200.mesh.generators import uvSphere
nt=int(sys.argv[1]); nph=int(sys.argv[2]); kind=sys.argv[3]
xyz, tri = uvSphere(1.0, nt, nph); n=len(tri)
G = oracle.green_matrix(xyz, tri, 5, threads=True)
A = oracle.area2(xyz, tri)
M=A[:,None]*G; Ms=0.5*(M+M.T)
ev=np.linalg.eigvalsh(Ms) if n<=13000 else None
if ev is not None: print('N',n,'eig max %.3e min %.3e, #positive %d'%(ev.max(),ev.min(),(ev>0).sum()), flush=True)
cen=(xyz[tri[:,0]]+xyz[tri[:,1]]+xyz[tri[:,2]])/3
v=1/np.sqrt(cen[:,0]**2+(cen[:,1]-2)**2+(cen[:,2]-4)**2)
braw=-v; b=A*braw
bs=128
if kind=='block':
    invs=[np.linalg.inv(Ms[s:min(n,s+bs),s:min(n,s+bs)]) for s in range(0,n,bs)]
    def pre(r):
        out=np.empty_like(r)
        for q,s in enumerate(range(0,n,bs)):
            e=min(n,s+bs); out[s:e]=invs[q]@r[s:e]
        return out
else:
    d=np.diag(Ms); pre=lambda r:r/d
x=np.zeros(n); r=b.copy(); w=pre(r); p=np.zeros(n); rho_old=1
thr=5e-13*np.abs(braw).max()
for k in range(6000):
    rho=r.dot(w); beta=rho/rho_old if k>0 else 0.0
    p=w+beta*p; z=A*(G@p); pz=p.dot(z); alpha=rho/pz
    x+=alpha*p; r-=alpha*z; w=pre(r); rho_old=rho
    rec=np.abs(r/A).max()
    if (k+1)%50==0 or rec<=thr:
        true=np.abs(braw-G@x).max()
        print(k+1,'rec %.3e true %.3e rho %.2e pz %.2e'%(rec,true,rho,pz), flush=True)
    if rec<=thr: break
