import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import oracle
from icqsol_b200.mesh.generators import uvSphere
nt=int(sys.argv[1])
xyz, tri = uvSphere(1.0, nt, nt//2); n=len(tri)
G = oracle.green_matrix(xyz, tri, 5, threads=True)
A = oracle.area2(xyz, tri)
cen=(xyz[tri[:,0]]+xyz[tri[:,1]]+xyz[tri[:,2]])/3
v=1/np.sqrt(cen[:,0]**2+(cen[:,1]-2)**2+(cen[:,2]-4)**2)
braw=-v; b=A*braw
print('N',n,flush=True)
for bs in (128,256,512,1024,2048):
    invs=[]
    for s in range(0,n,bs):
        e=min(n,s+bs); Mb=A[s:e,None]*G[s:e,s:e]; invs.append(np.linalg.inv(0.5*(Mb+Mb.T)))
    def pre(r):
        out=np.empty_like(r)
        for q,s in enumerate(range(0,n,bs)):
            e=min(n,s+bs); out[s:e]=invs[q]@r[s:e]
        return out
    x=np.zeros(n); r=b.copy(); w=pre(r); p=np.zeros(n); rho_old=1
    thr=5e-13*np.abs(braw).max()
    for k in range(2000):
        rho=r.dot(w); beta=rho/rho_old if k>0 else 0.0
        p=w+beta*p; z=A*(G@p); alpha=rho/p.dot(z)
        x+=alpha*p; r-=alpha*z; w=pre(r); rho_old=rho
        if np.abs(r/A).max()<=thr: break
    print('block',bs,'iterations',k+1,'extra traffic per iteration / product traffic = %.3f'%(n*bs*8/(4.0*n*n)),flush=True)
