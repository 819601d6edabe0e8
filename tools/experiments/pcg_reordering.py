import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from oracle import oracle
from icqsol_b200.mesh.generators import uvSphere
nt=int(sys.argv[1]); bs=int(sys.argv[2])
xyz, tri = uvSphere(1.0, nt, nt//2); n=len(tri)
t0=time.time()
G = oracle.green_matrix(xyz, tri, 5, threads=True)
A = oracle.area2(xyz, tri)
print('N',n,'matrix %.0fs'%(time.time()-t0),flush=True)
cen=(xyz[tri[:,0]]+xyz[tri[:,1]]+xyz[tri[:,2]])/3
v=1/np.sqrt(cen[:,0]**2+(cen[:,1]-2)**2+(cen[:,2]-4)**2)
M=A[:,None]*G; M=0.5*(M+M.T)
def run(perm,label):
    Mp=M[np.ix_(perm,perm)]; Ap=A[perm]; braw=-v[perm]; b=Ap*braw
    invs=[]
    for s in range(0,n,bs):
        e=min(n,s+bs); invs.append(np.linalg.inv(Mp[s:e,s:e]))
    def pre(r):
        out=np.empty_like(r)
        for q,s in enumerate(range(0,n,bs)):
            e=min(n,s+bs); out[s:e]=invs[q]@r[s:e]
        return out
    x=np.zeros(n); r=b.copy(); w=pre(r); p=np.zeros(n); rho_old=1
    thr=5e-13*np.abs(braw).max()
    for k in range(3000):
        rho=r.dot(w); beta=rho/rho_old if k>0 else 0.0
        p=w+beta*p; z=Mp@p; alpha=rho/p.dot(z)
        x+=alpha*p; r-=alpha*z; w=pre(r); rho_old=rho
        if np.abs(r/Ap).max()<=thr: break
    print(label,'block',bs,'iterations',k+1,flush=True)
run(np.arange(n),'natural')
# patch ordering: Morton code of (theta, phi) of the centroid on a grid whose cells are ~square patches
theta=np.arccos(np.clip(cen[:,2]/np.linalg.norm(cen,axis=1),-1,1)); phi=np.arctan2(cen[:,1],cen[:,0])%(2*np.pi)
def morton(ix,iy,bits=10):
    code=np.zeros(len(ix),np.int64)
    for b in range(bits):
        code|=((ix>>b)&1)<<(2*b+1); code|=((iy>>b)&1)<<(2*b)
    return code
ix=np.minimum((theta/np.pi*1024).astype(np.int64),1023); iy=np.minimum((phi/(2*np.pi)*1024).astype(np.int64),1023)
run(np.argsort(morton(ix,iy),kind='stable'),'morton(theta,phi)')
# 3-D Morton of the centroid
q=np.minimum(((cen+1.0)/2.0*1024).astype(np.int64),1023)
def morton3(a,b,c,bits=10):
    code=np.zeros(len(a),np.int64)
    for t in range(bits):
        code|=((a>>t)&1)<<(3*t+2); code|=((b>>t)&1)<<(3*t+1); code|=((c>>t)&1)<<(3*t)
    return code
run(np.argsort(morton3(q[:,0],q[:,1],q[:,2]),kind='stable'),'morton3(xyz)')
