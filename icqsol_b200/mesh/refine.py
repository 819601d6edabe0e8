"""Uniform refinement on the GPU (kernel ``k_subdivide``, csrc/refine.cu).

The part of the reference's ``RefineSurface`` (shapes/icqRefineSurface.py:46-177) that does
not need pytriangle: every triangle is cut into k*k children, cell data are inherited
(:155-174).  ``icqsol_b200.mesh.generators.subdivide`` is the host version with the same
point order, child order and roundings.  There is no CPU fallback here: without the
compiled library and a CUDA device the calls raise.
"""
import numpy

from icqsol_b200 import _lib


def subdivideOnDevice(xyz, tri, k, float32_points=True, device=0, ctx=None):
    """(points float64[ntri*(k+1)(k+2)/2, 3], triangles int64[ntri*k*k, 3]); children of
    parent t are triangles[t*k*k:(t+1)*k*k]."""
    k = int(k)
    if k < 1:
        raise ValueError('k must be >= 1')
    xyz = numpy.ascontiguousarray(xyz, numpy.float64)
    tri = numpy.ascontiguousarray(tri, numpy.int64)
    ntri = tri.shape[0]
    npl = (k + 1) * (k + 2) // 2
    out_xyz = numpy.empty((ntri * npl, 3), numpy.float64)
    out_tri = numpy.empty((ntri * k * k, 3), numpy.int64)
    own = ctx is None
    if own:
        ctx = _lib.Context(device)
    try:
        ctx.check(ctx.lib.icqb_subdivide_host(
            ctx.handle, xyz.ctypes.data_as(_lib.c_double_p), xyz.shape[0],
            tri.ctypes.data_as(_lib.c_int64_p), ntri, k, 1 if float32_points else 0,
            out_xyz.ctypes.data_as(_lib.c_double_p), out_tri.ctypes.data_as(_lib.c_int64_p)))
    finally:
        if own:
            ctx.close()
    return out_xyz, out_tri


def inheritCellData(values, k):
    """Cell array of the refined mesh: every child takes its parent's tuple
    (shapes/icqRefineSurface.py:155-174)."""
    return numpy.repeat(numpy.asarray(values), int(k) * int(k), axis=0)
