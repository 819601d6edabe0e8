"""Generate the golden vectors under tests/golden/ by running the REFERENCE's
own code (run in the build container only; /root/reference is absent on the
GPU box, which is why the outputs are committed).

  * C++  : oracle/_ref/libicqref.so = unmodified bem/icqFunctor.cpp,
           icqLaplaceFunctor.cpp, icqQuadrature.cpp, icqLaplaceMatrices.cpp
           (recipe: oracle/Makefile `ref`).
  * Python: icqsol.bem.icqPotentialIntegrals / icqQuadrature / icqQuadrature1D /
           icqReferenceTriangle and icqsol.solvers.icqConjugateGradient imported
           unchanged through a temporary `icqsol -> /root/reference` alias.  The
           ten-line diagonal loop of bem/icqBaseLaplaceSolver.py:42-66 (whose
           module needs `import vtk`) is driven from here over numpy arrays,
           calling the reference's PotentialIntegrals unchanged.

Usage:  python tests/golden/make_golden.py
"""
import ctypes
import os
import sys
import tempfile

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)

alias = tempfile.mkdtemp()
os.symlink(REF, os.path.join(alias, 'icqsol'))
sys.path.insert(0, alias)

from icqsol.bem.icqPotentialIntegrals import PotentialIntegrals  # noqa: E402
from icqsol.bem import icqQuadrature as refQuad  # noqa: E402
from icqsol.bem import icqQuadrature1D as refQuad1D  # noqa: E402
from icqsol.solvers.icqConjugateGradient import ConjugateGradient  # noqa: E402

from oracle import oracle  # noqa: E402  (ctypes loader for libicqref.so only)
from icqsol_b200.mesh.vtk_legacy_io import readVtkPolyData  # noqa: E402
from icqsol_b200.mesh.polydata import extract_mesh  # noqa: E402
from icqsol_b200.mesh.generators import fanSplit, uvSphere  # noqa: E402

oracle.build(with_reference=True)
reflib = oracle.reflib()
FOUR_PI = 4. * numpy.pi


def ref_diag(xyz, tri, order):
    """bem/icqBaseLaplaceSolver.py:34-66 driven over arrays."""
    gpws = refQuad.gaussPtsAndWeights[order]
    npts = gpws.shape[1]
    xsis, etas, weights = gpws[0, :], gpws[1, :], gpws[2, :]
    out = numpy.zeros(tri.shape[0])
    for jSrc in range(tri.shape[0]):
        ia, ib, ic = tri[jSrc]
        paSrc, pbSrc, pcSrc = xyz[ia].copy(), xyz[ib].copy(), xyz[ic].copy()
        dbSrc = pbSrc - paSrc
        dcSrc = pcSrc - paSrc
        g = 0
        for ipt in range(npts):
            xObs = paSrc + xsis[ipt] * dbSrc + etas[ipt] * dcSrc
            pot0ab = PotentialIntegrals(xObs, paSrc, pbSrc, order)
            pot0bc = PotentialIntegrals(xObs, pbSrc, pcSrc, order)
            pot0ca = PotentialIntegrals(xObs, pcSrc, paSrc, order)
            g += weights[ipt] * (pot0ab.getIntegralOneOverR() +
                                 pot0bc.getIntegralOneOverR() +
                                 pot0ca.getIntegralOneOverR())
        out[jSrc] = g / (-FOUR_PI)
    return out


def ref_green(xyz, tri, order=5):
    g = oracle.ref_offdiag(xyz, tri)
    g[numpy.diag_indices_from(g)] = ref_diag(xyz, tri, order)
    return g


def load(relpath):
    xyz, cells = extract_mesh(readVtkPolyData(os.path.join(REF, relpath)))
    return xyz, fanSplit(cells)


def main():
    # ---- tables -------------------------------------------------------------
    tabs = {}
    for o in range(1, 9):
        tabs['tri%d' % o] = numpy.asarray(refQuad.gaussPtsAndWeights[o], numpy.float64)
    for o in range(1, 6):
        tabs['line%d' % o] = numpy.asarray(refQuad1D.gaussPtsAndWeights[o], numpy.float64)
    numpy.savez(os.path.join(HERE, 'tables.npz'), **tabs)

    # ---- meshes -------------------------------------------------------------
    meshes = {
        'twotri': load('data/2triangles.vtk'),
        'tet': load('data/tet.vtk'),
        'sphere4x2': load('data/sphere4x2.vtk'),
        'createSphere': load('test_results/testCreateSphere.vtk'),
        'sphere': load('data/sphere.vtk'),
        'boltFine': load('test_results/testBoltFine.vtk'),
    }
    # two coplanar unit-square triangles (tests/testPotentialIntegrals.py:37-72)
    meshes['square'] = (numpy.array([[0., 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]]),
                        numpy.array([[0, 1, 3], [1, 2, 3]], numpy.int64))
    numpy.savez_compressed(os.path.join(HERE, 'meshes.npz'),
                           **{k + '_xyz': v[0] for k, v in meshes.items()},
                           **{k + '_tri': v[1] for k, v in meshes.items()})

    # ---- full G on the small meshes, every diagonal order ---------------------
    small = {}
    for name in ('twotri', 'tet', 'sphere4x2', 'createSphere', 'square'):
        xyz, tri = meshes[name]
        g = oracle.ref_offdiag(xyz, tri)
        small[name + '_offdiag'] = g
        for order in range(1, 6):
            small['%s_diag%d' % (name, order)] = ref_diag(xyz, tri, order)
    numpy.savez_compressed(os.path.join(HERE, 'green_small.npz'), **small)

    # ---- sampled G + solve on the two real meshes ------------------------------
    big = {}
    rng = numpy.random.default_rng(0)
    for name in ('sphere', 'boltFine'):
        xyz, tri = meshes[name]
        n = tri.shape[0]
        g = ref_green(xyz, tri, 5)
        rows = numpy.sort(rng.choice(n, 12, replace=False))
        rows[0], rows[-1] = 0, n - 1
        big[name + '_rows'] = rows
        big[name + '_G_rows'] = g[rows, :]
        big[name + '_G_cols'] = g[:, rows]
        big[name + '_diag5'] = numpy.diag(g).copy()
        big[name + '_rowsum'] = g.sum(axis=1)
        big[name + '_colsum'] = g.sum(axis=0)
        cen = (xyz[tri[:, 0]] + xyz[tri[:, 1]] + xyz[tri[:, 2]]) / 3.
        v = 1. / numpy.sqrt(cen[:, 0] ** 2 + (cen[:, 1] - 2.) ** 2 + (cen[:, 2] - 4.) ** 2)
        big[name + '_v'] = v
        big[name + '_laplace_rsp'] = -numpy.linalg.solve(g, v)   # icqLaplaceSolver.py:36
        big[name + '_poisson_rsp_one'] = numpy.dot(g, numpy.ones(n))  # icqPoissonSolver.py:36
        big[name + '_poisson_rsp_x'] = numpy.dot(g, cen[:, 0])
        big[name + '_cond'] = numpy.array(numpy.linalg.cond(g))
    numpy.savez_compressed(os.path.join(HERE, 'green_sampled.npz'), **big)

    # ---- scalar known-answer vectors -------------------------------------------
    kat = {}
    tris = rng.normal(size=(40, 3, 3))
    kat['corner_tris'] = tris
    kat['corner_vals'] = numpy.array(
        [[PotentialIntegrals(t[0], t[1], t[2], o).getIntegralOneOverR() for o in range(1, 6)]
         for t in tris])
    kat['corner_unit_right'] = numpy.array(
        [PotentialIntegrals([0., 0., 0.], [1., 0., 0.], [0., 1., 0.], o).getIntegralOneOverR()
         for o in range(1, 6)])
    h = ctypes.c_void_p()
    reflib.icqQuadratureInit(ctypes.byref(h))
    dp = oracle._dp
    obs = rng.normal(size=(40, 3)) + 3.
    single = numpy.zeros((40, 10))       # orders 0..9 (0 and 9 unknown -> 0)
    for i in range(40):
        reflib.icqQuadratureSetObserver(ctypes.byref(h), dp(numpy.ascontiguousarray(obs[i])))
        for o in range(10):
            t = numpy.ascontiguousarray(tris[i])
            single[i, o] = reflib.icqQuadratureEvaluate(ctypes.byref(h), o, dp(t[0]), dp(t[1]), dp(t[2]))
    kat['single_obs'] = obs
    kat['single_vals'] = single
    # tests/testQuadratureCpp.py:33-53: observer at origin, unit right triangle at z=... as in the test
    pairs = rng.normal(size=(40, 2, 3, 3))
    pairs[:, 1] += 2.5
    dbl = numpy.zeros((40, 10))
    for i in range(40):
        a = numpy.ascontiguousarray(pairs[i, 0])
        b = numpy.ascontiguousarray(pairs[i, 1])
        for o in range(10):
            dbl[i, o] = reflib.icqQuadratureEvaluateDouble(ctypes.byref(h), o, dp(a[0]), dp(a[1]), dp(a[2]),
                                                          dp(b[0]), dp(b[1]), dp(b[2]))
    kat['double_pairs'] = pairs
    kat['double_vals'] = dbl
    kat['max_order'] = numpy.array(reflib.icqQuadratureGetMaxOrder(ctypes.byref(h)))
    reflib.icqQuadratureDel(ctypes.byref(h))
    numpy.savez_compressed(os.path.join(HERE, 'known_answers.npz'), **kat)

    # ---- reference ConjugateGradient on the area-symmetrised system -------------
    cg = {}
    for name, (nt, npz) in (('uv16x8', (16, 8)), ('uv40x20', (40, 20))):
        xyz, tri = uvSphere(1.0, nt, npz)
        g = ref_green(xyz, tri, 5)
        a2 = oracle.area2(xyz, tri)
        s = a2[:, None] * g
        cen = (xyz[tri[:, 0]] + xyz[tri[:, 1]] + xyz[tri[:, 2]]) / 3.
        v = 1. / numpy.sqrt(cen[:, 0] ** 2 + (cen[:, 1] - 2.) ** 2 + (cen[:, 2] - 4.) ** 2)
        solver = ConjugateGradient(s, a2 * v)
        solver.setTolerance(1.e-13)
        x, err, k = solver.solve(numpy.zeros(len(v)))
        cg[name + '_nt_np'] = numpy.array([nt, npz])
        cg[name + '_x'] = x
        cg[name + '_err'] = numpy.array(err)
        cg[name + '_iters'] = numpy.array(k)
        cg[name + '_lu'] = numpy.linalg.solve(g, v)
        cg[name + '_G_diag'] = numpy.diag(g).copy()
        cg[name + '_G_row0'] = g[0].copy()
        # raw (non-symmetric) G: the reference CG does not converge
        solver2 = ConjugateGradient(g, v)
        solver2.setMaxNumberOfIterations(200)
        _, err2, k2 = solver2.solve(numpy.zeros(len(v)))
        cg[name + '_raw_err'] = numpy.array(err2)
        cg[name + '_raw_iters'] = numpy.array(k2)
    numpy.savez_compressed(os.path.join(HERE, 'cg.npz'), **cg)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == '__main__':
    main()
