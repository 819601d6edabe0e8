"""Deterministic synthetic meshes for the benchmark configurations
 and the polygon -> triangle step in front of the BEM path.

All generators return ``(xyz float64[Np,3], tri int64[N,3])``; coordinates are
rounded to float32 and promoted back by default, which is what a mesh that
went through ``vtkPoints()`` looks like to the reference
(shapes/icqShapeManager.py:794-798, bem/icqLaplaceMatrices.cpp:81-83).
"""
import numpy


def _round_f32(xyz, float32_points):
    xyz = numpy.ascontiguousarray(xyz, numpy.float64)
    if float32_points:
        xyz = xyz.astype(numpy.float32).astype(numpy.float64)
    return xyz


def uvSphere(radius=1.0, n_theta=16, n_phi=8, center=(0., 0., 0.), outward=True,
             float32_points=True):
    """UV sphere: rings at phi_i = pi*i/n_phi, theta_j = 2*pi*j/n_theta, poles as
    single points, quads split along (i,j)-(i+1,j+1).  N = 2*n_theta*(n_phi-1)."""
    nt, npz = int(n_theta), int(n_phi)
    phi = numpy.pi * numpy.arange(1, npz) / npz
    theta = 2. * numpy.pi * numpy.arange(nt) / nt
    ring = numpy.zeros((npz - 1, nt, 3))
    ring[:, :, 0] = numpy.sin(phi)[:, None] * numpy.cos(theta)[None, :]
    ring[:, :, 1] = numpy.sin(phi)[:, None] * numpy.sin(theta)[None, :]
    ring[:, :, 2] = numpy.cos(phi)[:, None]
    xyz = numpy.vstack([[0., 0., 1.], ring.reshape(-1, 3), [0., 0., -1.]]) * radius
    xyz += numpy.asarray(center, numpy.float64)[None, :]
    north, south = 0, 1 + (npz - 1) * nt

    def vid(i, j):  # ring i (0-based), column j
        return 1 + i * nt + (j % nt)

    j = numpy.arange(nt)
    tris = [numpy.stack([numpy.full(nt, north), vid(0, j), vid(0, j + 1)], 1)]
    for i in range(npz - 2):
        a, b, c, d = vid(i, j), vid(i + 1, j), vid(i + 1, j + 1), vid(i, j + 1)
        tris.append(numpy.stack([a, b, c], 1))
        tris.append(numpy.stack([a, c, d], 1))
    tris.append(numpy.stack([numpy.full(nt, south), vid(npz - 2, j + 1), vid(npz - 2, j)], 1))
    tri = numpy.vstack(tris).astype(numpy.int64)
    if not outward:
        tri = tri[:, ::-1].copy()
    return _round_f32(xyz, float32_points), tri


def hollowSphere(outer_radius=2.0, inner_radius=1.0, n_theta=100, n_phi=50,
                 float32_points=True):
    """Outer sphere plus an inner shell with doubled resolution and inverted
    orientation, mirroring the CSG subtraction of tests/testHollowSphere.py:31-43."""
    xo, to = uvSphere(outer_radius, n_theta, n_phi, float32_points=float32_points)
    xi, ti = uvSphere(inner_radius, 2 * n_theta, 2 * n_phi, outward=False,
                      float32_points=float32_points)
    return numpy.vstack([xo, xi]), numpy.vstack([to, ti + xo.shape[0]])


def fanSplit(cells, return_parent=False):
    """Polygons -> triangles by a fan from the first vertex:
    (v0,v1,v2),(v0,v2,v3),...  The reference triangulates with pytriangle
    (shapes/icqRefineSurface.py:204-306), which is not available here; for
    triangles this is the identity, for convex polygons it is a valid split.
    With ``return_parent`` also the polygon every triangle came from (the children
    inherit the parent's cell data, shapes/icqRefineSurface.py:155-174)."""
    if isinstance(cells, numpy.ndarray) and cells.ndim == 2:
        k = cells.shape[1]
        if k == 3:
            tri = numpy.ascontiguousarray(cells, numpy.int64)
            parent = numpy.arange(tri.shape[0], dtype=numpy.int64)
        else:
            tri = numpy.stack([cells[:, [0, j, j + 1]] for j in range(1, k - 1)], 1).reshape(-1, 3)
            tri = numpy.ascontiguousarray(tri, numpy.int64)
            parent = numpy.repeat(numpy.arange(cells.shape[0], dtype=numpy.int64), k - 2)
        return (tri, parent) if return_parent else tri
    out, parent = [], []
    for ic, c in enumerate(cells):
        for k in range(1, len(c) - 1):
            out.append((c[0], c[k], c[k + 1]))
            parent.append(ic)
    tri = numpy.asarray(out, numpy.int64).reshape(-1, 3)
    return (tri, numpy.asarray(parent, numpy.int64)) if return_parent else tri


def _lattice(k):
    """Local lattice of the k-per-edge subdivision: (i/k, j/k) of its points and the local
    connectivity of the k*k children."""
    ij = [(i, j) for i in range(k + 1) for j in range(k + 1 - i)]
    index = {p: n for n, p in enumerate(ij)}
    iu = numpy.array([p[0] for p in ij], numpy.float64) / k
    ju = numpy.array([p[1] for p in ij], numpy.float64) / k
    local = []
    for i in range(k):
        for j in range(k - i):
            local.append((index[(i, j)], index[(i + 1, j)], index[(i, j + 1)]))
            if i + j < k - 1:
                local.append((index[(i + 1, j)], index[(i + 1, j + 1)], index[(i, j + 1)]))
    return iu, ju, numpy.asarray(local, numpy.int64)


def subdivide(xyz, tri, k, float32_points=True):
    """Uniform subdivision of every triangle into k*k children (k per edge),
    children keep the parent's orientation.  Points on shared edges are
    duplicated per parent (G does not need a conforming mesh); the new
    coordinates are a + (i/k) db + (j/k) dc."""
    k = int(k)
    if k <= 1:
        return _round_f32(xyz, float32_points), numpy.asarray(tri, numpy.int64)
    tri = numpy.asarray(tri, numpy.int64)
    a, b, c = xyz[tri[:, 0]], xyz[tri[:, 1]], xyz[tri[:, 2]]
    # local lattice (i, j), i + j <= k
    iu, ju, local = _lattice(k)
    pts = (a[:, None, :] + iu[None, :, None] * (b - a)[:, None, :] +
           ju[None, :, None] * (c - a)[:, None, :])
    npl = len(iu)
    base = (numpy.arange(tri.shape[0], dtype=numpy.int64) * npl)[:, None, None]
    newtri = (local[None, :, :] + base).reshape(-1, 3)
    return _round_f32(pts.reshape(-1, 3), float32_points), newtri


def refineToMaxEdge(xyz, tri, max_edge_length, float32_points=True, return_maps=False):
    """Per-triangle uniform subdivision with k = ceil(longest edge / max_edge_length).
    Stand-in for the edge splitting of shapes/icqRefineSurface.py:46-177 (whose
    Delaunay step needs pytriangle); not triangle-for-triangle identical to it.
    The children of a triangle follow one another, in the order of their parents.
    With ``return_maps`` also (parent, (ids, weights)): the triangle every child came from
    (cell data is copied to the children, shapes/icqRefineSurface.py:155-174) and, for every
    output point, three input point ids with barycentric weights (point data is interpolated;
    the input points come first, unchanged)."""
    tri = numpy.asarray(tri, numpy.int64)
    npts = xyz.shape[0]

    def identity():
        ids = numpy.repeat(numpy.arange(npts, dtype=numpy.int64)[:, None], 3, 1)
        w = numpy.zeros((npts, 3))
        w[:, 0] = 1.
        return numpy.arange(tri.shape[0], dtype=numpy.int64), (ids, w)

    if not (max_edge_length > 0) or not numpy.isfinite(max_edge_length) or tri.shape[0] == 0:
        return (xyz, tri) + (identity() if return_maps else ())
    a, b, c = xyz[tri[:, 0]], xyz[tri[:, 1]], xyz[tri[:, 2]]
    longest = numpy.sqrt(numpy.maximum.reduce([((b - a) ** 2).sum(1), ((c - b) ** 2).sum(1),
                                               ((a - c) ** 2).sum(1)]))
    ks = numpy.maximum(1, numpy.ceil(longest / max_edge_length - 1e-12).astype(int))
    if ks.max() == 1:
        return (xyz, tri) + (identity() if return_maps else ())
    out_xyz, out_tri, parents, off = [xyz], [], [], npts
    ids0, w0 = identity()[1]
    out_ids, out_w = [ids0], [w0]
    for k in numpy.unique(ks):
        which = numpy.nonzero(ks == k)[0]
        sel = tri[which]
        if k == 1:
            out_tri.append(sel)
            parents.append(which)
            continue
        px, pt = subdivide(xyz, sel, k, float32_points)
        out_xyz.append(px)
        out_tri.append(pt + off)
        parents.append(numpy.repeat(which, k * k))
        off += px.shape[0]
        iu, ju, _ = _lattice(int(k))
        out_ids.append(numpy.repeat(sel, len(iu), axis=0))
        out_w.append(numpy.tile(numpy.stack([1. - iu - ju, iu, ju], 1), (sel.shape[0], 1)))
    parents = numpy.concatenate(parents)
    order = numpy.argsort(parents, kind='stable')
    newtri = numpy.vstack(out_tri)[order]
    if not return_maps:
        return numpy.vstack(out_xyz), newtri
    return numpy.vstack(out_xyz), newtri, parents[order], (numpy.vstack(out_ids), numpy.vstack(out_w))
