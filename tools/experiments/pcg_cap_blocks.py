import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import oracle
from icqsol_b200.mesh.generators import uvSphere
nt=int(sys.argv[1]); nph=int(sys.argv[2]); rings=int(sys.argv[3])
xyz, tri = uvSphere(1.0, nt, nph); n=len(tri)
G = oracle.green_matrix(xyz, tri, 5, threads=True)
A = oracle.area2(xyz, tri)
M=A[:,None]*G; Ms=0.5*(M+M.T)
cap = nt + 2*nt*(rings-1) if rings>0 else 0
cap = (cap+127)//128*128 if cap else 0
cuts=[0]
if cap: cuts.append(cap)
s=cuts[-1]
while s+128 < n-cap: s+=128; cuts.append(s)
if cap: cuts.append(n-cap)
cuts.append(n)
cuts=sorted(set(cuts))
print('N',n,'cap',cap,'blocks',len(cuts)-1, flush=True)
invs=[np.linalg.inv(Ms[a:b,a:b]) for a,b in zip(cuts[:-1],cuts[1:])]
def pre(r):
    out=np.empty_like(r)
    for q,(a,b) in enumerate(zip(cuts[:-1],cuts[1:])): out[a:b]=invs[q]@r[a:b]
    return out
cen=(xyz[tri[:,0]]+xyz[tri[:,1]]+xyz[tri[:,2]])/3
v=1/np.sqrt(cen[:,0]**2+(cen[:,1]-2)**2+(cen[:,2]-4)**2)
braw=-v; b=A*braw
x=np.zeros(n); r=b.copy(); w=pre(r); p=np.zeros(n); rho_old=1
thr=5e-13*np.abs(braw).max()
for k in range(3000):
    rho=r.dot(w); beta=rho/rho_old if k>0 else 0.0
    p=w+beta*p; z=Ms@p; pz=p.dot(z); alpha=rho/pz
    x+=alpha*p; r-=alpha*z; w=pre(r); rho_old=rho
    rec=np.abs(r/A).max()
    if (k+1)%50==0 or rec<=thr:
        true=np.abs(braw-G@x).max()
        print(k+1,'rec %.3e true %.3e rho %.2e pz %.2e'%(rec,true,rho,pz), flush=True)
    if rec<=thr: break
