import sys, numpy as np, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
from oracle import oracle
from icqsol_b200.mesh.generators import uvSphere
nt=int(sys.argv[1])
xyz, tri = uvSphere(1.0, nt, nt//2); n=len(tri)
A=oracle.area2(xyz,tri)
print('N',n,'area2 min/max',A.min(),A.max())
worst=[]
t0=time.time()
for T in range(0,(n+127)//128):
    s=T*128; e=min(n,s+128)
    Gb=oracle.block(xyz,tri,s,e,s,e,5)
    M=A[s:e,None]*Gb
    asym=np.abs(M-M.T).max()/np.abs(M).max()
    ev=np.linalg.eigvalsh(0.5*(M+M.T))
    worst.append((ev.max(), ev.min(), asym, T))
    if T<3 or ev.max()>0: print('block',T,'eig max %.3e min %.3e asym %.1e cond %.2e'%(ev.max(),ev.min(),asym, ev.min()/ev.max()))
w=np.array([(a,b) for a,b,_,_ in worst])
print('max eig over blocks %.3e ; cond range %.2e..%.2e; time %.1f'%(w[:,0].max(), (w[:,1]/w[:,0]).min(), (w[:,1]/w[:,0]).max(), time.time()-t0))
