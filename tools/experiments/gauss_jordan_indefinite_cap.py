import sys, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import oracle
from icqsol_b200.mesh.generators import uvSphere
xyz,tri=uvSphere(1.0,318,20); n=len(tri)
A=oracle.area2(xyz,tri)
m=1024
Gb=oracle.block(xyz,tri,0,m,0,m,5)
M=A[:m,None]*Gb; M=0.5*(M+M.T)
ev=np.linalg.eigvalsh(M); print('cap eig min %.3e max %.3e #pos %d cond %.2e'%(ev.min(),ev.max(),(ev>0).sum(), np.abs(ev).max()/np.abs(ev).min()))
a=M.copy()
minpiv=1e300
for k in range(m):
    piv=a[k,k]; minpiv=min(minpiv,abs(piv))
    pivinv=1.0/piv
    colk=a[:,k].copy(); rowk=a[k,:]*pivinv; rowk[k]=pivinv
    a-=np.outer(colk,rowk); a[k,:]=rowk; a[:,k]=-colk*pivinv; a[k,k]=pivinv
inv=0.5*(a+a.T)
ref=np.linalg.inv(M)
print('min |pivot| %.3e ; |GJ inv - LAPACK inv| / |inv| = %.3e ; |inv M - I| = %.3e (LAPACK %.3e)'%(minpiv, np.abs(inv-ref).max()/np.abs(ref).max(), np.abs(inv@M-np.eye(m)).max(), np.abs(ref@M-np.eye(m)).max()))
# sign pattern of pivots
