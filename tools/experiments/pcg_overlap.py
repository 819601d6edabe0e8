"""Overlapping diagonal blocks in the block preconditioner (symmetric additive Schwarz with the
partition of unity split as square roots): M^-1 = sum_b R_b^T D_b A_b^-1 D_b R_b, block b = the
tile rows of the product's own partition (icqRowPartition.sliverBlockPartition) extended by
`overlap` tiles on either side, D_b = 1/sqrt(number of blocks that cover the row).
  pcg_overlap.py <n_theta | -k (bolt cut k x k)> <n_phi> <blockTiles:overlapTiles,...>
Prints the CG iterations to max|r/s| <= 5e-13 max|b|, the traffic of the preconditioner step and
the tile products of its set-up, both relative to plain blocks of 8 tiles."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle
from icqsol_b200.mesh.generators import uvSphere, subdivide
from icqsol_b200.bem import icqRowPartition as rp
nt = int(sys.argv[1]); nph = int(sys.argv[2])
combos = [tuple(int(u) for u in t.split(':')) for t in sys.argv[3].split(',')]
if nt < 0:
    m_ = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', 'tests', 'golden', 'meshes.npz'))
    xyz, tri = subdivide(m_['boltFine_xyz'], m_['boltFine_tri'], -nt)
else:
    xyz, tri = uvSphere(1.0, nt, nph)
n = len(tri); nT = (n + 127) // 128
G = oracle.green_matrix(xyz, tri, 5, threads=True)
A = oracle.area2(xyz, tri)
M = A[:, None] * G; M = 0.5 * (M + M.T)
cen = (xyz[tri[:, 0]] + xyz[tri[:, 1]] + xyz[tri[:, 2]]) / 3
v = 1 / np.sqrt(cen[:, 0] ** 2 + (cen[:, 1] - 2) ** 2 + (cen[:, 2] - 4) ** 2)
braw = -v; b = A * braw
print('N', n, 'tiles', nT, flush=True)
for combo in combos:
    bt, ov = combo[0], combo[1]
    mode = combo[2] if len(combo) > 2 else 0
    isolate = mode == 1    # blocks over sliver runs neither extend nor are extended into
    # mode 2: sliver blocks extend outwards, their neighbours stay out of them; mode 3: the reverse
    starts, runs = rp.sliverBlockPartition(xyz, tri, blockTiles=bt)
    starts = [int(s) for s in starts] + [nT]
    sliver = np.zeros(nT + 1, bool)
    for (r0, cnt) in runs:
        sliver[r0:r0 + cnt] = True
    blocks = []
    cover = np.zeros(n)
    for t0, t1 in zip(starts[:-1], starts[1:]):
        e0, e1 = max(0, t0 - ov), min(nT, t1 + ov)
        if mode in (1, 3) and sliver[t0]:
            e0, e1 = t0, t1
        if mode in (1, 2) and not sliver[t0]:
            while e0 < t0 and sliver[e0]: e0 += 1
            while e1 > t1 and sliver[e1 - 1]: e1 -= 1
        lo, hi = e0 * 128, min(n, e1 * 128)
        blocks.append((lo, hi)); cover[lo:hi] += 1
    wgt = 1. / np.sqrt(cover)
    invs = [np.linalg.inv(M[lo:hi, lo:hi]) for lo, hi in blocks]
    def pre(r):
        out = np.zeros_like(r)
        for (lo, hi), inv in zip(blocks, invs):
            out[lo:hi] += wgt[lo:hi] * (inv @ (wgt[lo:hi] * r[lo:hi]))
        return out
    x = np.zeros(n); r = b.copy(); w = pre(r); p = np.zeros(n); rho_old = 1.; flips = 0
    thr = 5e-13 * np.abs(braw).max()
    for k in range(1500):
        rho = r.dot(w); beta = rho / rho_old if k > 0 else 0.
        p = w + beta * p; z = M @ p; pz = p.dot(z); alpha = rho / pz
        flips += (rho > 0) or (pz > 0)
        x += alpha * p; r -= alpha * z; w = pre(r); rho_old = rho
        if np.abs(r / A).max() <= thr: break
    traffic = sum((hi - lo) ** 2 for lo, hi in blocks) / float(n * 1024)
    setup = sum(((hi - lo) / 128.) ** 3 for lo, hi in blocks) / (nT * 64.)
    print('blockTiles %d overlap %d%s: %d blocks (dense runs %s), iterations %d%s, wrong-sign rho / p.z %d, '
          'preconditioner traffic x%.2f, set-up x%.2f, true backward error %.2e' % (
              bt, ov, ' (mode %d)' % mode if mode else '', len(blocks), runs, k + 1, '' if k + 1 < 1500 else ' (NOT converged)', flips, traffic, setup,
              np.abs(braw - G @ x).max() / np.abs(braw).max()), flush=True)
