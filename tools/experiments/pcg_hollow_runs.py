#!/usr/bin/env python
"""Slivers in the MIDDLE of the triangle list: a hollow sphere (outer sphere, then the inner
one with doubled resolution and flipped orientation, as tests/testHollowSphere.py:31-43) puts
the inner sphere's first pole in the middle.  Block-preconditioned CG of csrc/cg.cu with
(a) 128 x 128 blocks, (b) the two end caps of sliverCapTiles, (c) one dense block per run of
sliver tiles (icqRowPartition.sliverBlockPartition).   usage: pcg_hollow_runs.py [n_theta_inner]"""
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from icqsol_b200.bem import icqRowPartition as rp  # noqa: E402
from icqsol_b200.mesh.generators import hollowSphere  # noqa: E402

nti = int(sys.argv[1]) if len(sys.argv) > 1 else 320
xyz, tri = hollowSphere(2.0, 1.0, nti // 2, 10)
n = tri.shape[0]
nT = (n + 127) // 128
oracle.build()
t0 = time.time()
g = oracle.green_matrix(xyz, tri, 5, threads=True)
a2 = oracle.area2(xyz, tri)
print('hollow sphere, inner n_theta {0}: N={1}, {2} tiles, G in {3:.0f} s'.format(nti, n, nT, time.time() - t0),
      flush=True)
ms = a2[:, None] * g
del g
ms = 0.5 * (ms + ms.T)
cen = (xyz[tri[:, 0]] + xyz[tri[:, 1]] + xyz[tri[:, 2]]) / 3.
rhs = -1. / numpy.sqrt(cen[:, 0] ** 2 + (cen[:, 1] - 2.) ** 2 + (cen[:, 2] - 4.) ** 2)


def pcg(starts, label, maxit=1500):
    cuts = [128 * t for t in starts] + [n]
    invs = [numpy.linalg.inv(ms[p:q, p:q]) for p, q in zip(cuts[:-1], cuts[1:])]

    def pre(r):
        out = numpy.empty_like(r)
        for j, (p, q) in enumerate(zip(cuts[:-1], cuts[1:])):
            out[p:q] = invs[j].dot(r[p:q])
        return out
    x, r, p, rho_old = numpy.zeros(n), a2 * rhs, numpy.zeros(n), 1.
    w = pre(r)
    flips = 0
    for it in range(maxit):
        rho = r.dot(w)
        p = w + (rho / rho_old if it else 0.) * p
        z = ms.dot(p)
        pz = p.dot(z)
        flips += int(pz > 0) + int(rho > 0)
        alpha = rho / pz
        x += alpha * p
        r -= alpha * z
        w = pre(r)
        rho_old = rho
        err = numpy.abs(r / a2).max() / numpy.abs(rhs).max()
        if err <= 5e-13:
            print('{0}: converged in {1} iterations, {2} sign changes of rho / p.z'.format(label, it + 1, flips),
                  flush=True)
            return
    print('{0}: NOT converged, {1:.2e} after {2} iterations, {3} sign changes'.format(label, err, maxit, flips),
          flush=True)


head, tail = rp.sliverCapTiles(xyz, tri)
starts, runs = rp.sliverBlockPartition(xyz, tri)
print('sliver tiles', numpy.nonzero(rp.sliverTiles(xyz, tri))[0].tolist(), 'caps', (head, tail), 'runs', runs)
pcg(list(range(nT)), '(a) 128-blocks')
pcg(rp.blockPartition(n, 1, head, tail), '(b) end caps {0}'.format((head, tail)))
pcg(starts, '(c) dense block per run')
