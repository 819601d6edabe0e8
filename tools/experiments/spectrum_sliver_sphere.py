import sys, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import oracle
from icqsol_b200.mesh.generators import uvSphere
nt,nph=int(sys.argv[1]),int(sys.argv[2])
xyz, tri = uvSphere(1.0, nt, nph); n=len(tri)
G = oracle.green_matrix(xyz, tri, 5, threads=True)
A = oracle.area2(xyz, tri)
M=A[:,None]*G; Ms=0.5*(M+M.T)
# Jacobi-scaled
d=np.sqrt(-np.diag(Ms)); S=Ms/d[:,None]/d[None,:]
ev,V=np.linalg.eigh(S)
pos=ev>0
print('N',n,'scaled eig: min %.3e max %.3e, #pos %d'%(ev.min(),ev.max(),pos.sum()))
# ring index of each triangle: triangles ordered ring by ring: first ring nt tris (pole fan), then 2*nt per band ..., last ring nt
ring=np.empty(n,int); idx=0; r=0
ring[:nt]=0; idx=nt; r=1
while idx<n-nt:
    ring[idx:idx+2*nt]=r; idx+=2*nt; r+=1
ring[idx:]=r
nr=r+1
W=(V[:,pos]**2)
mass=np.array([W[ring==k].sum() for k in range(nr)])/pos.sum()
print('share of positive-eigenvector mass per ring:', np.round(mass,3))
# aspect ratios per ring
a,b,c=xyz[tri[:,0]],xyz[tri[:,1]],xyz[tri[:,2]]
e=np.stack([np.linalg.norm(b-a,axis=1),np.linalg.norm(c-b,axis=1),np.linalg.norm(a-c,axis=1)],1)
asp=e.max(1)**2/(A+1e-300)
print('aspect (Lmax^2/2A) per ring:', np.round([asp[ring==k].mean() for k in range(nr)],1))
