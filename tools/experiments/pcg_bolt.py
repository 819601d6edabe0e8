#!/usr/bin/env python
"""Does the block-preconditioned CG of csrc/cg.cu converge on the REFINED BOLT (config C3's
geometry: test_results/testBoltFine.vtk, every triangle cut k x k)?  The indefiniteness of
DESIGN.md section 4.6 comes from neighbouring slivers and is a property of the triangles' shapes,
which uniform subdivision keeps: usage  pcg_bolt.py <k> [block_tiles]"""
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from icqsol_b200.mesh.generators import subdivide  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 2
bt = int(sys.argv[2]) if len(sys.argv) > 2 else 1
m = numpy.load(os.path.join(ROOT, 'tests', 'golden', 'meshes.npz'))
xyz, tri = subdivide(m['boltFine_xyz'], m['boltFine_tri'], k)
n = tri.shape[0]
oracle.build()
t0 = time.time()
g = oracle.green_matrix(xyz, tri, 5, threads=True)
a2 = oracle.area2(xyz, tri)
print('bolt k={0}: N={1}, G in {2:.0f} s'.format(k, n, time.time() - t0), flush=True)
a, b_, c = xyz[tri[:, 0]], xyz[tri[:, 1]], xyz[tri[:, 2]]
e2 = numpy.maximum.reduce([((b_ - a) ** 2).sum(1), ((c - b_) ** 2).sum(1), ((a - c) ** 2).sum(1)])
print('aspect Lmax^2/|db x dc|: median {0:.1f}, 99% {1:.1f}, max {2:.1f}'.format(
    numpy.median(e2 / a2), numpy.percentile(e2 / a2, 99), (e2 / a2).max()))
ms = a2[:, None] * g
del g
ms = 0.5 * (ms + ms.T)
cen = (a + b_ + c) / 3.
rhs = -1. / numpy.sqrt(cen[:, 0] ** 2 + (cen[:, 1] - 2.) ** 2 + (cen[:, 2] - 4.) ** 2)
B = 128 * bt
cuts = list(range(0, n, B)) + [n]
invs, wrong = [], 0
for p, q in zip(cuts[:-1], cuts[1:]):
    blk = ms[p:q, p:q]
    wrong += int((numpy.linalg.eigvalsh(blk) > 0).sum())
    invs.append(numpy.linalg.inv(blk))
print('diagonal blocks of {0}: {1} eigenvalues of the wrong sign'.format(B, wrong), flush=True)


def pre(r):
    out = numpy.empty_like(r)
    for j, (p, q) in enumerate(zip(cuts[:-1], cuts[1:])):
        out[p:q] = invs[j].dot(r[p:q])
    return out


x, r, p, rho_old = numpy.zeros(n), a2 * rhs, numpy.zeros(n), 1.
w = pre(r)
t0 = time.time()
for it in range(3000):
    rho = r.dot(w)
    p = w + (rho / rho_old if it else 0.) * p
    z = ms.dot(p)
    pz = p.dot(z)
    alpha = rho / pz
    x += alpha * p
    r -= alpha * z
    w = pre(r)
    rho_old = rho
    err = numpy.abs(r / a2).max() / numpy.abs(rhs).max()
    if it % 50 == 0:
        print('  it {0}: max|r/s|/max|b| {1:.2e}  rho {2:.2e}  p.z {3:.2e}'.format(it, err, rho, pz), flush=True)
    if err <= 5e-13:
        print('converged in {0} iterations ({1:.0f} s)'.format(it + 1, time.time() - t0))
        break
else:
    print('NOT converged: {0:.2e} after 3000 iterations'.format(err))
