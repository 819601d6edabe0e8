import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import oracle
from icqsol_b200.mesh.generators import uvSphere
import scipy.sparse.linalg as sla
nt=int(sys.argv[1]); nph=int(sys.argv[2]); restart=int(sys.argv[3])
xyz, tri = uvSphere(1.0, nt, nph); n=len(tri)
G = oracle.green_matrix(xyz, tri, 5, threads=True)
cen=(xyz[tri[:,0]]+xyz[tri[:,1]]+xyz[tri[:,2]])/3
v=1/np.sqrt(cen[:,0]**2+(cen[:,1]-2)**2+(cen[:,2]-4)**2)
b=-v
bs=128
invs=[np.linalg.inv(G[s:min(n,s+bs),s:min(n,s+bs)]) for s in range(0,n,bs)]
def pre(r):
    out=np.empty_like(r)
    for q,s in enumerate(range(0,n,bs)):
        e=min(n,s+bs); out[s:e]=invs[q]@r[s:e]
    return out
cnt=[0]
def mv(x): cnt[0]+=1; return G@x
Aop=sla.LinearOperator((n,n),matvec=mv)
Mop=sla.LinearOperator((n,n),matvec=pre)
for name,M in (('block',Mop),('none',None)):
    cnt[0]=0
    t0=time.time()
    x,info=sla.gmres(Aop,b,M=M,restart=restart,maxiter=40,rtol=1e-13,atol=0.0)
    print(name,'info',info,'matvecs',cnt[0],'backward max-norm %.3e'%(np.abs(b-G@x).max()/np.abs(b).max()),'time %.1f'%(time.time()-t0),flush=True)
