"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

ctypes front-end to the CPU restatement (oracle/icq_oracle.c) and, when it was
built, to the unmodified reference C++ (oracle/_ref/libicqref.so, see
oracle/Makefile).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this module.

Parity pin: tests/test_oracle.py checks every function here against
tests/golden/*.npz (outputs of the reference's own C++ and Python, generated
by tests/golden/make_golden.py) and, when libicqref.so is present, against the
reference C++ directly.
"""
import ctypes
import os
import subprocess

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
_REFDIR = os.path.join(_HERE, '_ref')

c_dp = ctypes.POINTER(ctypes.c_double)
c_lp = ctypes.POINTER(ctypes.c_longlong)


def build(with_reference=None):
    """Compile the restatement (always) and the reference C++ (when its sources
    are present, i.e. in the build container)."""
    targets = ['all']
    if with_reference is None:
        with_reference = os.path.isdir('/root/reference/bem')
    if with_reference:
        targets.append('ref')
    subprocess.check_call(['make', '-s', '-C', _HERE] + targets)


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _lp(a):
    return a.ctypes.data_as(c_lp)


_libs = {}


def lib(threads=False):
    """The restatement library; ``threads=True`` selects the OpenMP build."""
    name = 'libicqoracle_omp.so' if threads else 'libicqoracle.so'
    if name not in _libs:
        path = os.path.join(_REFDIR, name)
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.icqo_quadrature_evaluate.restype = ctypes.c_double
        L.icqo_quadrature_evaluate.argtypes = [ctypes.c_int] + [c_dp] * 4
        L.icqo_quadrature_evaluate_double.restype = ctypes.c_double
        L.icqo_quadrature_evaluate_double.argtypes = [ctypes.c_int] + [c_dp] * 6
        L.icqo_offdiag.restype = None
        L.icqo_offdiag.argtypes = [c_dp, c_lp, ctypes.c_longlong, c_dp]
        L.icqo_offdiag_entry.restype = ctypes.c_double
        L.icqo_offdiag_entry.argtypes = [c_dp, c_lp, ctypes.c_longlong, ctypes.c_longlong]
        for f in (L.icqo_offdiag_block, L.icqo_k_block):
            f.restype = None
            f.argtypes = [c_dp, c_lp] + [ctypes.c_longlong] * 4 + [c_dp]
        L.icqo_corner_integral.restype = ctypes.c_double
        L.icqo_corner_integral.argtypes = [ctypes.c_int, c_dp, c_dp, c_dp]
        L.icqo_diag.restype = ctypes.c_int
        L.icqo_diag.argtypes = [c_dp, c_lp, ctypes.c_longlong, ctypes.c_int, c_dp]
        L.icqo_gemv.restype = None
        L.icqo_gemv.argtypes = [c_dp, ctypes.c_longlong, ctypes.c_longlong, c_dp, c_dp]
        L.icqo_area2.restype = ctypes.c_double
        L.icqo_area2.argtypes = [c_dp] * 3
        L.icqo_num_threads.restype = ctypes.c_int
        _libs[name] = L
    return _libs[name]


def tri_rule(order):
    """[(xi, eta, w)] of the triangle rule of ``order`` (bem/icqQuadrature.cpp:20-144)."""
    L = lib()
    L.icqo_tri_rule.restype = ctypes.c_int
    L.icqo_tri_rule.argtypes = [ctypes.c_int] + [ctypes.POINTER(c_dp)] * 3
    xi, eta, w = c_dp(), c_dp(), c_dp()
    n = L.icqo_tri_rule(int(order), ctypes.byref(xi), ctypes.byref(eta), ctypes.byref(w))
    return [(xi[k], eta[k], w[k]) for k in range(n)]


def have_reference():
    return os.path.exists(os.path.join(_REFDIR, 'libicqref.so'))


def reflib():
    """The unmodified reference C++ (bem/icqQuadrature.cpp, icqLaplaceMatrices.cpp ...)."""
    if 'ref' not in _libs:
        L = ctypes.CDLL(os.path.join(_REFDIR, 'libicqref.so'))
        L.icqref_offdiag.restype = None
        L.icqref_offdiag.argtypes = [c_dp, c_lp, ctypes.c_longlong, c_dp]
        L.icqQuadratureEvaluate.restype = ctypes.c_double
        L.icqQuadratureEvaluate.argtypes = [ctypes.c_void_p, ctypes.c_int] + [c_dp] * 3
        L.icqQuadratureEvaluateDouble.restype = ctypes.c_double
        L.icqQuadratureEvaluateDouble.argtypes = [ctypes.c_void_p, ctypes.c_int] + [c_dp] * 6
        L.icqQuadratureGetMaxOrder.restype = ctypes.c_int
        _libs['ref'] = L
    return _libs['ref']


def _mesh(xyz, tri):
    xyz = numpy.ascontiguousarray(xyz, numpy.float64)
    tri = numpy.ascontiguousarray(tri, numpy.int64)
    return xyz, tri


def offdiag(xyz, tri, gMat=None, threads=False):
    """bem/icqLaplaceMatrices.cpp:37-105 restated; fills the off-diagonal of gMat."""
    xyz, tri = _mesh(xyz, tri)
    n = tri.shape[0]
    if gMat is None:
        gMat = numpy.zeros((n, n), numpy.float64)
    lib(threads).icqo_offdiag(_dp(xyz), _lp(tri), n, _dp(gMat))
    return gMat


def ref_offdiag(xyz, tri, gMat=None):
    """The reference's own computeOffDiagonalTerms on (xyz, tri)."""
    xyz, tri = _mesh(xyz, tri)
    n = tri.shape[0]
    if gMat is None:
        gMat = numpy.zeros((n, n), numpy.float64)
    reflib().icqref_offdiag(_dp(xyz), _lp(tri), n, _dp(gMat))
    return gMat


def diag(xyz, tri, order=5, threads=False):
    """bem/icqBaseLaplaceSolver.py:34-66 restated.  KeyError for order outside 1..5."""
    xyz, tri = _mesh(xyz, tri)
    out = numpy.zeros(tri.shape[0], numpy.float64)
    if lib(threads).icqo_diag(_dp(xyz), _lp(tri), tri.shape[0], int(order), _dp(out)) != 0:
        raise KeyError(order)
    return out


def green_matrix(xyz, tri, order=5, threads=False):
    """Full G as BaseLaplaceSolver builds it (diagonal, then off-diagonal)."""
    g = offdiag(xyz, tri, threads=threads)
    g[numpy.diag_indices_from(g)] = diag(xyz, tri, order, threads=threads)
    return g


def block(xyz, tri, r0, r1, c0, c1, order=5, threads=True):
    """G[r0:r1, c0:c1] including any diagonal entries in range."""
    xyz, tri = _mesh(xyz, tri)
    out = numpy.zeros((r1 - r0, c1 - c0), numpy.float64)
    lib(threads).icqo_offdiag_block(_dp(xyz), _lp(tri), r0, r1, c0, c1, _dp(out))
    lo, hi = max(r0, c0), min(r1, c1)
    if hi > lo:
        d = diag(xyz, tri[lo:hi], order)
        for k in range(lo, hi):
            out[k - r0, k - c0] = d[k - lo]
    return out


def k_block(xyz, tri, r0, r1, c0, c1, threads=True):
    """K[r0:r1, c0:c1] -- double layer, PARITY UNPINNED (not in the reference)."""
    xyz, tri = _mesh(xyz, tri)
    out = numpy.zeros((r1 - r0, c1 - c0), numpy.float64)
    lib(threads).icqo_k_block(_dp(xyz), _lp(tri), r0, r1, c0, c1, _dp(out))
    return out


def corner_integral(order, x, p, q):
    """PotentialIntegrals(x, p, q, order).getIntegralOneOverR()
    (bem/icqPotentialIntegrals.py:15-40)."""
    a = [numpy.ascontiguousarray(v, numpy.float64) for v in (x, p, q)]
    return lib().icqo_corner_integral(int(order), _dp(a[0]), _dp(a[1]), _dp(a[2]))


def quadrature_evaluate(order, pobs, pa, pb, pc):
    a = [numpy.ascontiguousarray(v, numpy.float64) for v in (pobs, pa, pb, pc)]
    return lib().icqo_quadrature_evaluate(int(order), *[_dp(v) for v in a])


def quadrature_evaluate_double(order, pa0, pb0, pc0, pa1, pb1, pc1):
    a = [numpy.ascontiguousarray(v, numpy.float64) for v in (pa0, pb0, pc0, pa1, pb1, pc1)]
    return lib().icqo_quadrature_evaluate_double(int(order), *[_dp(v) for v in a])


def area2(xyz, tri):
    """|db x dc| per triangle (twice the area; bem/icqLaplaceMatrices.cpp:9-30)."""
    a, b, c = xyz[tri[:, 0]], xyz[tri[:, 1]], xyz[tri[:, 2]]
    return numpy.linalg.norm(numpy.cross(b - a, c - a), axis=1)


def laplace_response(gMat, src):
    """bem/icqLaplaceSolver.py:36"""
    return -numpy.linalg.solve(gMat, src)


def poisson_response(gMat, src):
    """bem/icqPoissonSolver.py:36"""
    return numpy.dot(gMat, src)


def conjugate_gradient(mat, b, x0, precond=None, tol=1.e-10, maxNumIters=None):
    """solvers/icqConjugateGradient.py:49-79 restated: Jacobi-preconditioned CG,
    true residual recomputed every iteration, absolute 2-norm tolerance."""
    n = len(b)
    if precond is None:
        precond = numpy.array([mat[i, i] for i in range(n)])
    if maxNumIters is None:
        maxNumIters = n
    x = x0.copy()
    r = b - mat.dot(x)
    w = r / precond
    p = numpy.zeros(x0.shape, numpy.float64)
    beta = 0.0
    rho = r.dot(w)
    err = numpy.linalg.norm(r)
    k = 0
    while abs(err) > tol and k < maxNumIters:
        p = w + beta * p
        z = mat.dot(p)
        alpha = rho / p.dot(z)
        r -= alpha * z
        w = r / precond
        rhoOld = rho
        rho = r.dot(w)
        x += alpha * p
        beta = rho / rhoOld
        err = numpy.linalg.norm(b - mat.dot(x))
        k += 1
    return x, err, k
