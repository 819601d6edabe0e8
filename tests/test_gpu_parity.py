"""Parity of the CUDA path against the oracle (CPU restatement of the reference)
and the committed golden vectors (outputs of the reference's own code).
All calls go through the C ABI (ctypes) or the Python classes on top of it.

Tolerances: north_star demands 1e-12 relative for every G entry and for GEMV
outputs in fp64; solution vectors are checked by backward error <= 1e-12 and
forward difference vs LU <= 50*cond(G)*eps.
"""
import ctypes

import numpy
import pytest

pytestmark = pytest.mark.gpu

TOL_ENTRY = 1e-12
# K is not in the reference (parity unpinned).  Its integrand n.(x-y)/r^3 is nearly
# singular for touching, nearly coplanar triangles: two correctly rounded fp64
# evaluation orders differ by ~1e-11 of max|K| there (measured: numpy vs the C oracle
# 8.8e-12 on the 40x20 UV sphere), hence a max-norm tolerance.
K_TOL = 1e-9
EPS = numpy.finfo(numpy.float64).eps


def dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def lp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


@pytest.fixture()
def ctx(gpu_lib):
    from icqsol_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def assemble(ctx, xyz, tri, order=5, with_k=False, rows=None):
    xyz = numpy.ascontiguousarray(xyz, numpy.float64)
    tri = numpy.ascontiguousarray(tri, numpy.int64)
    n = tri.shape[0]
    ctx.check(ctx.lib.icqb_set_mesh(ctx.handle, dp(xyz), xyz.shape[0], lp(tri), n))
    r0, r1 = (0, n) if rows is None else rows
    ctx.check(ctx.lib.icqb_set_rows(ctx.handle, r0, r1))
    g = numpy.full((r1 - r0, n), numpy.nan)
    k = numpy.full((r1 - r0, n), numpy.nan) if with_k else None
    ctx.check(ctx.lib.icqb_assemble_host(ctx.handle, order, dp(g), dp(k) if with_k else None))
    return (g, k) if with_k else g


def relerr(a, b):
    return numpy.abs(a - b).max() / numpy.abs(b).max() if b.size else 0.0


def max_entry_relerr(a, b):
    return numpy.abs(a / b - 1.0).max()


# ---- G: golden vectors produced by the reference's own C++ and Python ------------------

@pytest.mark.parametrize('name', ['twotri', 'tet', 'sphere4x2', 'createSphere', 'square'])
@pytest.mark.parametrize('order', [1, 2, 3, 4, 5])
def test_green_small_golden(ctx, golden, name, order):
    xyz, tri = golden.mesh(name)
    g = assemble(ctx, xyz, tri, order)
    ref = golden['green_small'][name + '_offdiag'].copy()
    ref[numpy.diag_indices_from(ref)] = golden['green_small']['%s_diag%d' % (name, order)]
    assert max_entry_relerr(g, ref) < TOL_ENTRY


def test_survey_known_values(ctx, golden):
    """Values printed in SURVEY.md section 8c (run of the reference during the survey)."""
    g = assemble(ctx, *golden.mesh('twotri'), 5)
    assert abs(g[0, 0] / -0.15959645874363654 - 1) < TOL_ENTRY
    assert abs(g[1, 0] / -0.03637297439435873 - 1) < TOL_ENTRY
    g = assemble(ctx, *golden.mesh('tet'), 5)
    assert abs(g[0, 1] / -0.08250452568311496 - 1) < TOL_ENTRY
    assert abs(g[0, 3] / -0.16292331240300845 - 1) < TOL_ENTRY
    assert abs(g[3, 0] / -0.09406381827314243 - 1) < TOL_ENTRY
    assert abs(g[3, 3] / -0.21436428380303485 - 1) < TOL_ENTRY
    g = assemble(ctx, *golden.mesh('square'), 5)
    assert abs(g[1, 0] / -0.08186618508970281 - 1) < TOL_ENTRY


@pytest.mark.parametrize('name', ['sphere', 'boltFine'])
def test_green_real_meshes_golden(ctx, golden, name):
    """data/sphere.vtk (960 triangles after the fan split) and
    test_results/testBoltFine.vtk (2350 triangles, areas spanning 1500x)."""
    xyz, tri = golden.mesh(name)
    g = assemble(ctx, xyz, tri, 5)
    big = golden['green_sampled']
    rows = big[name + '_rows']
    assert max_entry_relerr(g[rows, :], big[name + '_G_rows']) < TOL_ENTRY
    assert max_entry_relerr(g[:, rows], big[name + '_G_cols']) < TOL_ENTRY
    assert max_entry_relerr(numpy.diag(g), big[name + '_diag5']) < TOL_ENTRY
    assert max_entry_relerr(g.sum(axis=1), big[name + '_rowsum']) < TOL_ENTRY
    assert max_entry_relerr(g.sum(axis=0), big[name + '_colsum']) < TOL_ENTRY


def test_green_vs_oracle_every_entry(ctx, golden, oracle_mod):
    xyz, tri = golden.mesh('boltFine')
    g = assemble(ctx, xyz, tri, 5)
    ref = oracle_mod.green_matrix(xyz, tri, 5, threads=True)
    assert max_entry_relerr(g, ref) < TOL_ENTRY


@pytest.mark.parametrize('n_theta,n_phi', [(5, 3), (16, 9), (33, 17), (64, 30)])
def test_green_uv_spheres_vs_oracle(ctx, oracle_mod, n_theta, n_phi):
    """Sizes around the 128-triangle tile boundary: 20, 256, 1056, 3712 triangles."""
    from icqsol_b200.mesh.generators import uvSphere
    xyz, tri = uvSphere(1.0, n_theta, n_phi)
    g = assemble(ctx, xyz, tri, 5)
    ref = oracle_mod.green_matrix(xyz, tri, 5, threads=True)
    assert max_entry_relerr(g, ref) < TOL_ENTRY


@pytest.mark.parametrize('n', [1, 2, 127, 128, 129, 255, 257])
def test_green_ragged_sizes(ctx, golden, oracle_mod, n):
    xyz, tri = golden.mesh('boltFine')
    tri = tri[500:500 + n]
    g = assemble(ctx, xyz, tri, 3)
    ref = oracle_mod.green_matrix(xyz, tri, 3)
    assert max_entry_relerr(g, ref) < TOL_ENTRY


def test_empty_mesh(ctx):
    xyz = numpy.zeros((0, 3))
    tri = numpy.zeros((0, 3), numpy.int64)
    ctx.check(ctx.lib.icqb_set_mesh(ctx.handle, dp(xyz), 0, lp(tri), 0))
    assert ctx.lib.icqb_num_triangles(ctx.handle) == 0
    ctx.check(ctx.lib.icqb_assemble_host(ctx.handle, 5, dp(numpy.zeros(1)), None))


def test_bad_vertex_index_rejected(ctx):
    xyz = numpy.zeros((3, 3))
    tri = numpy.array([[0, 1, 3]], numpy.int64)
    st = ctx.lib.icqb_set_mesh(ctx.handle, dp(xyz), 3, lp(tri), 1)
    assert st == -2
    with pytest.raises(RuntimeError):
        ctx.check(st)


@pytest.mark.parametrize('order', [0 + 6, 7, 8, -1])
def test_diag_order_out_of_range_is_keyerror(ctx, golden, order):
    """bem/icqQuadrature1D.py:38 raises KeyError for order > 5 (tests/testLaplace.py:30)."""
    xyz, tri = golden.mesh('tet')
    with pytest.raises(KeyError):
        assemble(ctx, xyz, tri, order)


def test_offdiag_dropin_keeps_diagonal(gpu_lib, golden):
    """computeOffDiagonalTermsArrays: the contract of computeOffDiagonalTerms
    (bem/icqLaplaceMatrices.cpp:37-105) -- off-diagonal filled, diagonal untouched."""
    xyz, tri = golden.mesh('createSphere')
    n = tri.shape[0]
    g = numpy.zeros((n, n))
    g[numpy.diag_indices(n)] = 7.25
    st = gpu_lib.computeOffDiagonalTermsArrays(dp(numpy.ascontiguousarray(xyz)), xyz.shape[0],
                                               lp(numpy.ascontiguousarray(tri)), n, dp(g))
    assert st == 0
    ref = golden['green_small']['createSphere_offdiag']
    mask = ~numpy.eye(n, dtype=bool)
    assert (numpy.diag(g) == 7.25).all()
    assert numpy.abs(g[mask] / ref[mask] - 1).max() < TOL_ENTRY


def test_mirror_rule(ctx, golden):
    """G[j,i] = G[i,j] * A_i / A_j (bem/icqLaplaceMatrices.cpp:97-98)."""
    xyz, tri = golden.mesh('boltFine')
    g = assemble(ctx, xyz, tri, 5)
    a, b, c = xyz[tri[:, 0]], xyz[tri[:, 1]], xyz[tri[:, 2]]
    area = numpy.linalg.norm(numpy.cross(b - a, c - a), axis=1)
    s = area[:, None] * g
    off = ~numpy.eye(len(tri), dtype=bool)
    assert numpy.abs(s - s.T)[off].max() / numpy.abs(s).max() < 1e-14


def test_row_block_assembly_matches_full(ctx, golden):
    """Row-sharded assembly (what each rank does) reproduces the full matrix."""
    xyz, tri = golden.mesh('sphere')
    n = tri.shape[0]
    full = assemble(ctx, xyz, tri, 5)
    for (r0, r1) in [(0, 240), (240, 480), (480, 961 - 1), (100, 101), (130, 389), (960, 960)]:
        part = assemble(ctx, xyz, tri, 5, rows=(r0, r1))
        assert part.shape == (r1 - r0, n)
        if r1 > r0:
            assert max_entry_relerr(part, full[r0:r1]) < 1e-14


# ---- K (double layer; not in the reference: parity unpinned, oracle = definition) ----

def test_k_matches_definition_oracle(ctx, golden, oracle_mod):
    xyz, tri = golden.mesh('createSphere')
    n = tri.shape[0]
    g, k = assemble(ctx, xyz, tri, 5, with_k=True)
    ref = oracle_mod.k_block(xyz, tri, 0, n, 0, n)
    assert numpy.abs(k - ref).max() / numpy.abs(ref).max() < K_TOL
    assert (numpy.diag(k) == 0).all()
    gref = golden['green_small']['createSphere_offdiag'].copy()
    gref[numpy.diag_indices(n)] = golden['green_small']['createSphere_diag5']
    assert max_entry_relerr(g, gref) < TOL_ENTRY     # G unchanged by the fused K path


def test_k_row_sums_closed_surface(ctx):
    """sum_j K[i,j] -> 1/2 on a closed outward-oriented surface (solid angle 2 pi)."""
    from icqsol_b200.mesh.generators import uvSphere
    xyz, tri = uvSphere(1.0, 48, 24)
    g, k = assemble(ctx, xyz, tri, 5, with_k=True)
    rs = k.sum(axis=1)
    assert numpy.abs(rs - 0.5).max() < 0.02


def test_k_against_analytic_solid_angles(ctx, oracle_mod):
    """The device K against an anchor that is not this repository's code: the double-layer
    potential of a flat triangle is its solid angle, K[i,j] = sum_k w_k Omega_j(o_k) / (4 pi) up to
    the source-side quadrature error (tests/test_oracle.py pins the oracle the same way):
    well separated pairs to 5e-12 of the largest such entry, in both kernels (direct and
    far-field split), K and the transposed part of the lower storage."""
    from test_oracle import _solid_angle
    from icqsol_b200.mesh.generators import uvSphere
    xyz, tri = uvSphere(1.0, 32, 16)
    n = tri.shape[0]
    nodes = numpy.array(oracle_mod.tri_rule(8))
    a, b, c = xyz[tri[:, 0]], xyz[tri[:, 1]], xyz[tri[:, 2]]
    obs = a[:, None, :] + nodes[None, :, 0:1] * (b - a)[:, None, :] + nodes[None, :, 1:2] * (c - a)[:, None, :]
    exact = numpy.zeros((n, n))
    for j in range(n):
        om = _solid_angle(obs.reshape(-1, 3), a[j], b[j], c[j]).reshape(n, 16)
        exact[:, j] = (om * nodes[None, :, 2]).sum(1) / (4. * numpy.pi)
    cen = (a + b + c) / 3.
    dist = numpy.sqrt(((cen[:, None, :] - cen[None, :, :]) ** 2).sum(-1))
    size = numpy.sqrt(numpy.maximum.reduce([((b - a) ** 2).sum(1), ((c - b) ** 2).sum(1), ((a - c) ** 2).sum(1)]))
    far = dist > 3. * (size[:, None] + size[None, :])
    assert far.sum() > 0.4 * n * n
    for variant in (0, 1):
        ctx.check(ctx.lib.icqb_set_assembly_variant(ctx.handle, variant))
        g, k = assemble(ctx, xyz, tri, 5, with_k=True)
        assert numpy.abs(k[far] - exact[far]).max() < 5e-12 * numpy.abs(exact[far]).max()
        # closed outward surface: every row of the exact entries adds up to 1/2; the computed ones
        # miss it by the quadrature error of the touching pairs only
        assert numpy.abs(k.sum(1) - 0.5).max() < 0.05
    ctx.check(ctx.lib.icqb_set_assembly_variant(ctx.handle, 0))


def test_k_row_block_and_sampled_oracle(ctx, golden, oracle_mod):
    xyz, tri = golden.mesh('boltFine')
    g, k = assemble(ctx, xyz, tri, 5, with_k=True, rows=(1000, 1300))
    ref = oracle_mod.k_block(xyz, tri, 1000, 1300, 0, tri.shape[0])
    assert numpy.abs(k - ref).max() / numpy.abs(ref).max() < K_TOL
    gref = oracle_mod.block(xyz, tri, 1000, 1300, 0, tri.shape[0], 5)
    assert max_entry_relerr(g, gref) < TOL_ENTRY


# ---- quadrature handle API (tests/testQuadratureCpp.py) ----------------------------------

def test_quadrature_handle_api(gpu_lib, golden):
    kat = golden['known_answers']
    h = ctypes.c_void_p()
    gpu_lib.icqQuadratureInit(ctypes.byref(h))
    assert h.value
    assert gpu_lib.icqQuadratureGetMaxOrder(ctypes.byref(h)) == int(kat['max_order']) == 8
    tris, obs = kat['corner_tris'], kat['single_obs']
    for i in range(tris.shape[0]):
        gpu_lib.icqQuadratureSetObserver(ctypes.byref(h), dp(numpy.ascontiguousarray(obs[i])))
        t = numpy.ascontiguousarray(tris[i])
        for order in range(10):
            v = gpu_lib.icqQuadratureEvaluate(ctypes.byref(h), order, dp(t[0]), dp(t[1]), dp(t[2]))
            ref = kat['single_vals'][i, order]
            assert abs(v - ref) <= 1e-13 * abs(ref)      # unknown orders: exactly 0
    pairs = kat['double_pairs']
    for i in range(pairs.shape[0]):
        a = numpy.ascontiguousarray(pairs[i, 0])
        b = numpy.ascontiguousarray(pairs[i, 1])
        for order in range(10):
            v = gpu_lib.icqQuadratureEvaluateDouble(ctypes.byref(h), order, dp(a[0]), dp(a[1]),
                                                    dp(a[2]), dp(b[0]), dp(b[1]), dp(b[2]))
            ref = kat['double_vals'][i, order]
            assert abs(v - ref) <= 1e-13 * abs(ref)
    # zero-area triangle -> 0 (bem/icqQuadrature.cpp:185)
    z = numpy.zeros(3)
    assert gpu_lib.icqQuadratureEvaluate(ctypes.byref(h), 5, dp(z), dp(z), dp(z)) == 0.0
    gpu_lib.icqQuadratureDel(ctypes.byref(h))


def test_quadrature_cpp_style(gpu_lib, golden):
    """tests/testQuadratureCpp.py:33-53: device quadrature == triangleQuadrature formula
    (observer at the origin, -1/(4 pi r)) for orders 1..8, tolerance 1e-10 as in the reference."""
    tables = golden['tables']
    pa, pb, pc = numpy.array([0., 0., 1.]), numpy.array([1., 0., 1.]), numpy.array([0., 1., 1.])
    h = ctypes.c_void_p()
    gpu_lib.icqQuadratureInit(ctypes.byref(h))
    for order in range(1, 9):
        xi, eta, w = tables['tri%d' % order]
        p = pa[None, :] + xi[:, None] * (pb - pa)[None, :] + eta[:, None] * (pc - pa)[None, :]
        area = numpy.linalg.norm(numpy.cross(pb - pa, pc - pa))
        expected = 0.5 * area * (w * (-1. / (4. * numpy.pi * numpy.linalg.norm(p, axis=1)))).sum()
        v = gpu_lib.icqQuadratureEvaluate(ctypes.byref(h), order, dp(pa), dp(pb), dp(pc))
        assert abs(v - expected) < 1e-10
    gpu_lib.icqQuadratureDel(ctypes.byref(h))


# ---- GEMV / CG / solver classes -----------------------------------------------------------

@pytest.mark.parametrize('nrows,ncols', [(1, 1), (3, 5), (4, 512), (127, 1023), (960, 960),
                                         (1001, 2350)])
def test_gemv_vs_numpy(ctx, nrows, ncols):
    import torch
    rng = numpy.random.default_rng(nrows * 7919 + ncols)
    a = rng.normal(size=(nrows, ncols))
    x = rng.normal(size=ncols)
    for ld in (ncols, (ncols + 15) // 16 * 16):
        dev = torch.zeros((nrows, ld), dtype=torch.float64, device='cuda:0')
        dev[:, :ncols] = torch.from_numpy(a).cuda()
        y = numpy.zeros(nrows)
        ctx.check(ctx.lib.icqb_gemv_host(ctx.handle, dev.data_ptr(), nrows, ncols, ld, dp(x), dp(y)))
        ref = a.dot(x)
        scale = numpy.abs(a).dot(numpy.abs(x))
        assert (numpy.abs(y - ref) <= 1e-13 * scale).all()


@pytest.mark.parametrize('storage', ['lower', 'full'])
def test_poisson_response_golden(golden, storage):
    """PoissonSolver on the reference's meshes: v = G.charge vs numpy.dot on the
    reference G (bem/icqPoissonSolver.py:36), 1e-12 relative."""
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqPoissonSolver import PoissonSolver
    big = golden['green_sampled']
    for name in ('sphere', 'boltFine'):
        xyz, tri = golden.mesh(name)
        solver = PoissonSolver(PolyData.from_arrays(xyz, tri, dtype=numpy.float64), float('inf'),
                               storage=storage)
        solver.setSourceFromExpression('1.0')
        rsp = solver.computeResponseField()
        assert numpy.abs(rsp / big[name + '_poisson_rsp_one'] - 1).max() < 1e-12
        solver.setSourceFromExpression('x')
        rsp = solver.computeResponseField()
        ref = big[name + '_poisson_rsp_x']
        assert numpy.abs(rsp - ref).max() / numpy.abs(ref).max() < 1e-12
        arr = solver.getVtkPolyData().GetCellData().GetArray('v')
        assert arr is not None and arr.GetNumberOfTuples() == len(tri)


@pytest.mark.parametrize('storage', ['lower', 'full'])
def test_laplace_response_golden(golden, storage):
    """LaplaceSolver: sigma = -G^-1 v.  Reference = LU (bem/icqLaplaceSolver.py:36).
    Backward error <= 1e-12; forward difference vs LU bounded by the conditioning."""
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver
    big = golden['green_sampled']
    for name in ('sphere', 'boltFine'):
        xyz, tri = golden.mesh(name)
        solver = LaplaceSolver(PolyData.from_arrays(xyz, tri, dtype=numpy.float64), float('inf'), 5,
                               storage=storage)
        big_rows = big[name + '_rows']
        gfull = solver.getGreenMatrix()
        assert max_entry_relerr(gfull[big_rows, :], big[name + '_G_rows']) < TOL_ENTRY
        assert max_entry_relerr(gfull[:, big_rows], big[name + '_G_cols']) < TOL_ENTRY
        solver.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
        v = solver.getSourceArray(solver.getSourceArrayIndex())
        assert numpy.abs(v / big[name + '_v'] - 1).max() < 1e-14
        rsp = solver.computeResponseField()
        ref = big[name + '_laplace_rsp']
        g = solver.getGreenMatrix()
        backward = numpy.abs(v + g.dot(rsp)).max() / numpy.abs(v).max()
        assert backward < 1e-12
        # forward error of ANY solver is bounded by cond(G) x (relative residual); LU's own
        # is ~cond*eps.  2-norm bound with the measured residual, plus LU's share.
        cond = float(big[name + '_cond'])
        res2 = numpy.linalg.norm(v + g.dot(rsp)) / numpy.linalg.norm(v)
        fwd2 = numpy.linalg.norm(rsp - ref) / numpy.linalg.norm(ref)
        assert fwd2 < cond * (res2 + 50 * EPS)
        assert res2 < 1e-12
        assert solver.lastSolveInfo['iterations'] < len(tri)
        arr = solver.getVtkPolyData().GetCellData().GetArray('normal_electric_field_jump')
        assert arr is not None


def test_conjugate_gradient_class_golden(golden):
    """ConjugateGradient on the symmetrised system of a UV sphere vs the reference
    class's own run (tests/golden/cg.npz) and LU."""
    from icqsol_b200.mesh.generators import uvSphere
    from icqsol_b200.solvers.icqConjugateGradient import ConjugateGradient
    from oracle import oracle
    cg = golden['cg']
    for name in ('uv16x8', 'uv40x20'):
        nt, npz = cg[name + '_nt_np']
        xyz, tri = uvSphere(1.0, int(nt), int(npz))
        g = oracle.green_matrix(xyz, tri, 5, threads=True)
        a2 = oracle.area2(xyz, tri)
        cen = (xyz[tri[:, 0]] + xyz[tri[:, 1]] + xyz[tri[:, 2]]) / 3.
        v = 1. / numpy.sqrt(cen[:, 0] ** 2 + (cen[:, 1] - 2.) ** 2 + (cen[:, 2] - 4.) ** 2)
        # (1) exactly the reference's usage: explicit symmetrised matrix
        s = a2[:, None] * g
        solver = ConjugateGradient(s, a2 * v)
        solver.setTolerance(1.e-13)
        x, err, k = solver.solve(numpy.zeros(len(v)))
        assert err <= 1.e-13
        assert abs(k - int(cg[name + '_iters'])) <= 3
        assert numpy.abs(x - cg[name + '_x']).max() / numpy.abs(x).max() < 1e-9
        assert abs(solver.getSolutionError(x) - err) < 1e-13
        # (2) raw G with row scaling
        solver = ConjugateGradient(g, v)
        solver.setRowScaling(a2)
        solver.setTolerance(1.e-12)
        x2, err2, k2 = solver.solve(numpy.zeros(len(v)))
        assert err2 <= 1.e-12
        lu = cg[name + '_lu']
        assert numpy.abs(x2 - lu).max() / numpy.abs(lu).max() < 1e-8


def test_cg_nonzero_start_and_maxiter(golden):
    from icqsol_b200.solvers.icqConjugateGradient import ConjugateGradient
    rng = numpy.random.default_rng(3)
    q = rng.normal(size=(300, 300))
    a = q.dot(q.T) + 300 * numpy.eye(300)
    xt = rng.normal(size=300)
    b = a.dot(xt)
    solver = ConjugateGradient(a, b)
    x, err, k = solver.solve(rng.normal(size=300))
    assert err <= 1e-10 and numpy.abs(x - xt).max() < 1e-10 and 0 < k < 300
    solver.setMaxNumberOfIterations(3)
    x, err, k = solver.solve(numpy.zeros(300))
    assert k == 3 and err > 1e-10
    # already converged start: zero iterations (while-condition at icqConjugateGradient.py:64)
    solver.setMaxNumberOfIterations(300)
    x, err, k = solver.solve(xt)
    assert k == 0


def test_point_data_projection_and_missing_source(golden):
    from icqsol_b200.mesh.polydata import PolyData, DataArray
    from icqsol_b200.bem.icqPoissonSolver import PoissonSolver
    xyz, tri = golden.mesh('tet')
    pd = PolyData.from_arrays(xyz, tri)
    solver = PoissonSolver(pd, float('inf'))
    with pytest.raises(RuntimeError):
        solver.computeResponseField()
    pd.GetPointData().AddArray(DataArray('charge', [1., 2., 3., 4.]))
    rsp = solver.computeResponseField()
    cellCharge = numpy.array([(1 + 2 + 4) / 3., (1 + 3 + 2) / 3., (1 + 4 + 3) / 3., (2 + 3 + 4) / 3.])
    g = solver.getGreenMatrix()
    assert numpy.abs(rsp - g.dot(cellCharge)).max() < 1e-14


# ---- block-lower-triangular storage -----------------------------------------------------------

def lower_assemble(ctx, xyz, tri, blocks, order=5, with_k=False):
    import torch
    from icqsol_b200.bem import icqRowPartition as rp
    xyz = numpy.ascontiguousarray(xyz, numpy.float64)
    tri = numpy.ascontiguousarray(tri, numpy.int64)
    n = tri.shape[0]
    ctx.check(ctx.lib.icqb_set_mesh(ctx.handle, dp(xyz), xyz.shape[0], lp(tri), n))
    ctx.check(ctx.lib.icqb_set_row_blocks(ctx.handle, *blocks))
    rows = (blocks[1] - blocks[0]) + (blocks[3] - blocks[2])
    ntiles = int(ctx.lib.icqb_lower_num_tiles(ctx.handle))
    assert ntiles == rp.lowerStrips(blocks)[1]
    bufs = [torch.full((max(ntiles, 1), 128, 128), float('nan'), dtype=torch.float64, device='cuda:0')
            for _ in range(3 if with_k else 1)]
    ctx.check(ctx.lib.icqb_assemble_lower_dev(ctx.handle, order, bufs[0].data_ptr(),
                                              bufs[1].data_ptr() if with_k else 0,
                                              bufs[2].data_ptr() if with_k else 0))
    ctx.check(ctx.lib.icqb_synchronize(ctx.handle))
    out = []
    for b in bufs:
        t = b[:ntiles].cpu().numpy()
        written = numpy.where(numpy.isnan(t), 0.0, t)
        rowsArr = rp.lowerTilesToRows(written, blocks, n)
        nanRows = rp.lowerTilesToRows(numpy.isnan(t).astype(float), blocks, n)
        out.append((rowsArr, nanRows))
    return out, rows


@pytest.mark.parametrize('blocks', [(0, 960, 960, 960), (0, 256, 768, 960), (256, 512, 512, 768),
                                    (128, 128, 384, 500), (0, 0, 0, 0)])
def test_lower_storage_tiles_vs_oracle(ctx, golden, oracle_mod, blocks):
    """Every stored entry (tiles J <= I, tile-contiguous layout) of arbitrary row blocks
    equals the oracle; KL/KU hold K[i,j] / K[j,i]."""
    xyz, tri = golden.mesh('sphere')
    n = tri.shape[0]
    bufs, rows = lower_assemble(ctx, xyz, tri, blocks, 5, with_k=True)
    (g, gnan), (kl, _), (ku, _) = bufs
    ref = oracle_mod.green_matrix(xyz, tri, 5, threads=True)
    kref = oracle_mod.k_block(xyz, tri, 0, n, 0, n)
    glob = list(range(blocks[0], blocks[1])) + list(range(blocks[2], blocks[3]))
    assert len(glob) == rows
    kmax = numpy.abs(kref).max()
    for lrow, i in enumerate(glob):
        cend = min(n, (i // 128 + 1) * 128)
        assert (gnan[lrow, :cend] == 0).all()                 # every stored entry written
        assert numpy.abs(g[lrow, :cend] / ref[i, :cend] - 1).max() < TOL_ENTRY
        assert numpy.abs(kl[lrow, :cend] - kref[i, :cend]).max() < K_TOL * kmax
        assert numpy.abs(ku[lrow, :cend] - kref[:cend, i]).max() < K_TOL * kmax


@pytest.mark.parametrize('name', ['tet', 'sphere', 'boltFine'])
def test_lower_storage_solver_matrices(golden, oracle_mod, name):
    """getGreenMatrix()/getNormalDerivativeGreenMatrix() rebuilt from the lower storage
    (mirror rule on read) equal the oracle entry by entry."""
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver
    xyz, tri = golden.mesh(name)
    n = tri.shape[0]
    solver = LaplaceSolver(PolyData.from_arrays(xyz, tri, dtype=numpy.float64), float('inf'), 5,
                           computeK=True, storage='lower')
    ref = oracle_mod.green_matrix(xyz, tri, 5, threads=True)
    assert max_entry_relerr(solver.getGreenMatrix(), ref) < TOL_ENTRY
    kref = oracle_mod.k_block(xyz, tri, 0, n, 0, n)
    assert numpy.abs(solver.getNormalDerivativeGreenMatrix() - kref).max() < K_TOL * numpy.abs(kref).max()
    a2 = solver.getTriangleAreas2()
    assert (a2 == oracle_mod.area2(xyz, tri)).all() or numpy.abs(a2 / oracle_mod.area2(xyz, tri) - 1).max() < 1e-15


@pytest.mark.parametrize('nt,nph', [(7, 4), (16, 9), (33, 17), (64, 30)])
def test_symmetric_product_vs_numpy(oracle_mod, nt, nph):
    """icqb_symv_dev (streams the stored half once) == numpy.dot with the full oracle G."""
    from icqsol_b200.mesh.generators import uvSphere
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqPoissonSolver import PoissonSolver
    xyz, tri = uvSphere(1.0, nt, nph)
    n = tri.shape[0]
    solver = PoissonSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5, storage='lower')
    ref = oracle_mod.green_matrix(xyz, tri, 5, threads=True)
    rng = numpy.random.default_rng(n)
    for _ in range(2):
        x = rng.normal(size=n)
        y = solver.applyGreenMatrix(x)
        scale = numpy.abs(ref).dot(numpy.abs(x))
        assert (numpy.abs(y - ref.dot(x)) <= 1e-13 * scale).all()


@pytest.mark.parametrize('n,blocks', [(6000, None), (2500, (0, 640, 1920, 2500)), (130, None)])
def test_symmetric_product_synthetic_matrix(ctx, n, blocks):
    """k_symv on a synthetic matrix obeying the mirror rule G[j,i] = G[i,j] A_i/A_j, written
    straight into the tile-contiguous lower storage: sizes where tiles are split between
    warps (head slots), spans cover whole tiles, and ragged last tiles.  The unstored
    remainder of ragged tiles is NaN-filled: it must never reach the result."""
    import torch
    from icqsol_b200.bem import icqRowPartition as rp
    from icqsol_b200.mesh.generators import uvSphere
    nt = 2
    while 2 * nt * (nt // 2 - 1) < n:
        nt += 2
    xyz, tri = uvSphere(1.0, nt, nt // 2)
    tri = numpy.ascontiguousarray(tri[:n])
    ctx.check(ctx.lib.icqb_set_mesh(ctx.handle, dp(xyz), xyz.shape[0], lp(tri), n))
    if blocks is None:
        blocks = (0, n, n, n)
    ctx.check(ctx.lib.icqb_set_row_blocks(ctx.handle, *blocks))
    a2 = numpy.zeros(n)
    ctx.check(ctx.lib.icqb_area2_host(ctx.handle, dp(a2)))
    rng = numpy.random.default_rng(n)
    s = rng.normal(size=(n, n))
    s = s + s.T
    g = s * a2[None, :]
    strips, ntiles = rp.lowerStrips(blocks)
    assert ntiles == ctx.lib.icqb_lower_num_tiles(ctx.handle)
    tiles = numpy.full((ntiles, 128, 128), numpy.nan)
    for (row0, nrows, tileRow, base) in strips:
        for J in range(tileRow + 1):
            ce = min(n, (J + 1) * 128)
            tiles[base + J, :nrows, :ce - J * 128] = g[row0:row0 + nrows, J * 128:ce]
    dev = torch.device('cuda', 0)
    dA = torch.from_numpy(tiles).to(dev)
    glob = numpy.array(list(range(blocks[0], blocks[1])) + list(range(blocks[2], blocks[3])), int)
    # the rows this context stores define the operator: entries (i,j) and (j,i) with i stored, j<=tile(i)
    mask = numpy.zeros((n, n), bool)
    for i in glob:
        ce = min(n, (i // 128 + 1) * 128)
        mask[i, :ce] = True
    tI = numpy.arange(n) // 128
    lowerPart = mask & (tI[None, :] < tI[:, None])      # strictly below the diagonal tiles
    full = numpy.where(mask, g, 0.0) + numpy.where(lowerPart, g, 0.0).T * (a2[None, :] / a2[:, None])
    for scaled in (0, 1):
        x = rng.normal(size=n)
        dx = torch.from_numpy(x).to(dev)
        dy = torch.zeros(n, dtype=torch.float64, device=dev)
        ctx.check(ctx.lib.icqb_symv_dev(ctx.handle, dA.data_ptr(), dx.data_ptr(), dy.data_ptr(), scaled))
        y = dy.cpu().numpy()
        ref = full.dot(x)
        bound = numpy.abs(full).dot(numpy.abs(x))
        if scaled:
            ref = a2 * ref
            bound = a2 * bound
        assert numpy.isfinite(y).all()
        assert (numpy.abs(y - ref) <= 1e-13 * bound + 1e-300).all()


@pytest.mark.parametrize('kind,starts,expect', [(1, None, 'fallback'), (2, [0, 1, 2, 3], 'fallback'),
                                                (2, [0, 2], 'weak')])
def test_weak_pivot_in_a_preconditioner_block(ctx, kind, starts, expect):
    """A 128 x 128 diagonal block whose unpivoted inversion meets a pivot that has lost 14 digits
    ([[1, 1], [1, 1 + 1e-14]] inside tile 1 of a synthetic symmetric matrix, row-block storage):
    the block falls back to 1/diag and is counted (both solvers; a block of ONE tile of the
    partitioned solver behaves the same), inside a larger dense block the pivot is reported only
    (icqb_cg_last_blocks)."""
    import torch
    from icqsol_b200.mesh.generators import uvSphere
    xyz, tri = uvSphere(1.0, 20, 11)
    n = tri.shape[0]
    ctx.check(ctx.lib.icqb_set_mesh(ctx.handle, dp(xyz), xyz.shape[0], lp(numpy.ascontiguousarray(tri)), n))
    ctx.check(ctx.lib.icqb_set_rows(ctx.handle, 0, n))
    a2 = numpy.zeros(n)
    ctx.check(ctx.lib.icqb_area2_host(ctx.handle, dp(a2)))
    rng = numpy.random.default_rng(5)
    noise = rng.standard_normal((n, n)) * 1e-3
    sym = 4. * numpy.eye(n) + 0.5 * (noise + noise.T)
    sym[128:130, :] = 0.
    sym[:, 128:130] = 0.
    sym[128:130, 128:130] = [[1., 1.], [1., 1. + 1e-14]]
    ld = (n + 15) // 16 * 16
    g = numpy.zeros((n, ld))
    g[:, :n] = sym / a2[:, None]          # diag(a2) g is the symmetric matrix above
    dev = torch.device('cuda', 0)
    dA = torch.from_numpy(g).to(dev)
    db = torch.ones(n, dtype=torch.float64, device=dev)
    dx = torch.zeros(n, dtype=torch.float64, device=dev)
    ctx.check(ctx.lib.icqb_cg_set_preconditioner(ctx.handle, kind))
    st = numpy.asarray(starts or [], numpy.int64)
    ctx.check(ctx.lib.icqb_cg_set_block_partition(ctx.handle, lp(st) if len(st) else None, len(st)))
    it, r2, ri = ctypes.c_int64(0), ctypes.c_double(0.), ctypes.c_double(0.)
    ctx.check(ctx.lib.icqb_cg_solve_dev(ctx.handle, dA.data_ptr(), n, ld, 0, ctx.lib.icqb_area2_dev(ctx.handle), 0,
                                        db.data_ptr(), dx.data_ptr(), 5e-13, 2, 3, ctypes.byref(it),
                                        ctypes.byref(r2), ctypes.byref(ri)))
    bad, weak = ctypes.c_int64(-1), ctypes.c_int64(-1)
    ctx.check(ctx.lib.icqb_cg_last_blocks(ctx.handle, ctypes.byref(bad), ctypes.byref(weak)))
    if expect == 'fallback':
        assert bad.value == 1
    else:
        assert bad.value == 0 and weak.value >= 1
    assert numpy.isfinite(dx.cpu().numpy()).all()


def test_block_preconditioner_cuts_iterations():
    """LaplaceSolver with the inverse-diagonal-block preconditioner needs fewer CG
    iterations than 1/diag (the reference class's default) and gives the same response."""
    from icqsol_b200.mesh.generators import uvSphere
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver
    xyz, tri = uvSphere(1.0, 48, 24)
    out = {}
    for storage in ('lower', 'full'):
        for kind in ('diagonal', 'block'):
            solver = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5, storage=storage,
                                   preconditioner=kind)
            solver.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
            v = solver.getSourceArray(solver.getSourceArrayIndex())
            rsp = solver.computeResponseField()
            g = solver.getGreenMatrix()
            assert numpy.abs(v + g.dot(rsp)).max() / numpy.abs(v).max() < 1e-12
            out[storage, kind] = (rsp, solver.lastSolveInfo['iterations'])
        assert out[storage, 'block'][1] < out[storage, 'diagonal'][1]
        # two solutions with backward error <= 5e-13 differ by <= cond(G) * 1e-12 (cond ~ 1e3..1e4)
        assert relerr(out[storage, 'block'][0], out[storage, 'diagonal'][0]) < 1e-6
    # the same arithmetic in both storages up to summation order
    assert abs(out['lower', 'block'][1] - out['full', 'block'][1]) <= 3
    assert relerr(out['lower', 'block'][0], out['full', 'block'][0]) < 1e-6


@pytest.mark.parametrize('n', [300, 128, 517])
def test_conjugate_gradient_block_preconditioner(n):
    """ConjugateGradient.setBlockPreconditioner on a dense SPD matrix whose size is not a
    multiple of the block (identity padding of the last block)."""
    from icqsol_b200.solvers.icqConjugateGradient import ConjugateGradient
    rng = numpy.random.default_rng(n)
    q = rng.normal(size=(n, n))
    a = q.dot(q.T) + n * numpy.eye(n)
    xt = rng.normal(size=n)
    b = a.dot(xt)
    ref = ConjugateGradient(a, b)
    x0, err0, k0 = ref.solve(numpy.zeros(n))
    solver = ConjugateGradient(a, b)
    solver.setBlockPreconditioner()
    x, err, k = solver.solve(numpy.zeros(n))
    assert err <= 1e-10 and numpy.abs(x - xt).max() < 1e-10
    assert 0 < k <= k0
    if n <= 128:
        assert k <= 2     # one block: the preconditioner is the inverse


def test_example_solve_laplace_cli(tmp_path, oracle_mod, capsys):
    """examples/solveLaplace.py (the reference's CLI, config C1 style): legacy VTK in, binary
    legacy VTK out with the response as a named cell field; the field solves G sigma = -v for
    the oracle's G."""
    import importlib.util
    import os
    from icqsol_b200.mesh.generators import uvSphere
    from icqsol_b200.mesh.polydata import PolyData, extract_mesh
    from icqsol_b200.mesh.vtk_legacy_io import readVtkPolyData, writeVtkPolyData
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('solveLaplace',
                                                  os.path.join(root, 'examples', 'solveLaplace.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    xyz, tri = uvSphere(1.0, 24, 12)
    src = str(tmp_path / 'in.vtk')
    dst = str(tmp_path / 'out.vtk')
    writeVtkPolyData(PolyData.from_arrays(xyz, tri), src)
    assert mod.main(['--input', src, '--output', dst, '--verbose',
                     '--dirichlet', 'sin(pi*x)*cos(pi*y)*z']) == 0
    assert 'normal electric field jump min/avg/max' in capsys.readouterr().out
    assert mod.main([]) == 3
    back = readVtkPolyData(dst)
    xyz2, cells2 = extract_mesh(back)
    assert (xyz2 == xyz).all() and (numpy.array(cells2) == tri).all()
    sigma = back.GetCellData().GetArray('normal_electric_field').as_numpy()[:, 0]
    v = back.GetCellData().GetArray('voltage').as_numpy()[:, 0]
    cen = (xyz[tri[:, 0]] + xyz[tri[:, 1]] + xyz[tri[:, 2]]) / 3.
    vref = numpy.sin(numpy.pi * cen[:, 0]) * numpy.cos(numpy.pi * cen[:, 1]) * cen[:, 2]
    assert numpy.abs(v - vref).max() < 1e-14
    g = oracle_mod.green_matrix(xyz, tri, 5, threads=True)
    assert numpy.abs(v + g.dot(sigma)).max() / numpy.abs(v).max() < 1e-12


# ---- BASELINE.json configurations at (or near) full size: sampled oracle rows/columns and
# ---- size-independent properties (the oracle cannot build a whole matrix of this size) ----

def _sampled_checks(solver, xyz, tri, oracle_mod, rows, rsp=None, v=None):
    """Column j of G by the device product G e_j (stored and mirrored entries) against the
    oracle's column; rows of the oracle against the device solution."""
    n = tri.shape[0]
    for j in rows:
        e = numpy.zeros(n)
        e[j] = 1.0
        col = solver.applyGreenMatrix(e)
        ref = oracle_mod.block(xyz, tri, 0, n, j, j + 1, 5)[:, 0]
        assert numpy.abs(col / ref - 1.).max() < TOL_ENTRY, j
    if rsp is not None:
        vmax = numpy.abs(v).max()
        for i in rows:
            grow = oracle_mod.block(xyz, tri, i, i + 1, 0, n, 5)[0]
            assert abs(v[i] + grow.dot(rsp)) < 1e-12 * vmax, i


def test_config_c2_full_size(oracle_mod):
    """Config C2 (19,880-triangle UV sphere, BASELINE.json configs[1]): sampled columns of G
    equal the oracle to 1e-12; the response satisfies the oracle's rows of G sigma = -v to
    1e-12 max|v|; the product is linear; the device's own backward error agrees."""
    from icqsol_b200.mesh.generators import uvSphere
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver
    xyz, tri = uvSphere(1.0, 142, 71)
    n = tri.shape[0]
    assert n == 19880
    solver = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5)
    solver.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
    v = solver.getSourceArray(solver.getSourceArrayIndex())
    rsp = solver.computeResponseField()
    assert solver.lastSolveInfo['iterations'] < 400
    assert solver.lastSolveInfo['residualInf'] <= 5e-13 * numpy.abs(v).max()
    _sampled_checks(solver, xyz, tri, oracle_mod, [0, 141, 9999, 19751, 19879], rsp, v)
    back = solver.applyGreenMatrix(rsp)
    assert numpy.abs(v + back).max() < 1e-12 * numpy.abs(v).max()
    rng = numpy.random.default_rng(5)
    x, y = rng.normal(size=n), rng.normal(size=n)
    lin = solver.applyGreenMatrix(2.0 * x - 3.0 * y) - (2.0 * solver.applyGreenMatrix(x) -
                                                        3.0 * solver.applyGreenMatrix(y))
    assert numpy.abs(lin).max() < 1e-12 * numpy.abs(solver.applyGreenMatrix(numpy.abs(x) + numpy.abs(y))).max()


def test_config_c3_refined_bolt(golden, oracle_mod):
    """Config C3's geometry at a size one GPU test can afford: test_results/testBoltFine.vtk
    with every triangle split in 4 (9,400 triangles, areas down to 2e-7)."""
    from icqsol_b200.mesh.generators import subdivide
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver
    xyz0, tri0 = golden.mesh('boltFine')
    xyz, tri = subdivide(xyz0, tri0, 2)
    n = tri.shape[0]
    assert n == 4 * tri0.shape[0]
    solver = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5)
    solver.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
    v = solver.getSourceArray(solver.getSourceArrayIndex())
    rsp = solver.computeResponseField()
    assert solver.lastSolveInfo['iterations'] < n
    _sampled_checks(solver, xyz, tri, oracle_mod, [0, 4701, n - 1], rsp, v)


@pytest.mark.parametrize('mesh', ['bolt_k6', 'sphere_318x159'])
def test_config_c3_full_size_on_one_gpu(golden, oracle_mod, mesh):
    """Config C3 at FULL size on one B200 (lower storage: 28.6 / 40.4 GB): the reference's
    testBoltFine.vtk with every triangle cut into 36 (84,600 triangles, bench.py's multi-GPU
    workload) and the UV sphere of the same size class (100,488 triangles) whose pole slivers
    make diag(A) G indefinite -- the CG with the dense sliver blocks converges,
    the response satisfies the oracle's rows of G sigma = -v to 1e-12 max|v|, sampled columns of
    the device matrix equal the oracle's to 1e-12."""
    import torch
    from icqsol_b200.mesh.generators import subdivide, uvSphere
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip('needs 60 GB of HBM')
    if mesh == 'bolt_k6':
        xyz, tri = subdivide(*golden.mesh('boltFine'), 6)
        assert tri.shape[0] == 84600
    else:
        xyz, tri = uvSphere(1.0, 318, 159)
        assert tri.shape[0] == 100488
    n = tri.shape[0]
    solver = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5, directFallback=False)
    solver.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
    v = solver.getSourceArray(solver.getSourceArrayIndex())
    rsp = solver.computeResponseField()
    assert solver.lastSolveInfo['method'] == 'cg' and solver.lastSolveInfo['converged']
    assert solver.lastSolveInfo['iterations'] < 400
    _sampled_checks(solver, xyz, tri, oracle_mod, [0, n // 2 + 77, n - 1], rsp, v)
    del solver
    torch.cuda.empty_cache()


def test_config_c4_hollow_sphere_poisson(oracle_mod):
    """Config C4's geometry (tests/testHollowSphere.py: outer shell plus an inward-oriented
    inner shell of doubled resolution), PoissonSolver v = G.charge against the oracle."""
    from icqsol_b200.mesh.generators import hollowSphere
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqPoissonSolver import PoissonSolver
    xyz, tri = hollowSphere(2.0, 1.0, 20, 10)
    n = tri.shape[0]
    for storage in ('lower', 'full'):
        solver = PoissonSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5, storage=storage)
        solver.setSourceFromExpression('x + 2.0')
        rsp = solver.computeResponseField()
        g = oracle_mod.green_matrix(xyz, tri, 5, threads=True)
        cen = (xyz[tri[:, 0]] + xyz[tri[:, 1]] + xyz[tri[:, 2]]) / 3.
        ref = g.dot(cen[:, 0] + 2.0)
        assert numpy.abs(rsp - ref).max() < 1e-12 * numpy.abs(ref).max()
    xyz, tri = hollowSphere(2.0, 1.0, 100, 50)
    assert tri.shape[0] == 49400      # the full-size C4 mesh of SURVEY.md section 8(d)
    solver = PoissonSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5)
    _sampled_checks(solver, xyz, tri, oracle_mod, [0, 9800, 49399])


def test_subdivide_on_device(golden):
    """k_subdivide against the host version: bit-identical points, same children; config C3's
    mesh (testBoltFine, k = 6 -> 84,600 triangles) is produced on the device."""
    from icqsol_b200.mesh.generators import subdivide
    from icqsol_b200.mesh.refine import subdivideOnDevice, inheritCellData
    xyz, tri = golden.mesh('boltFine')
    for k, f32 in ((2, True), (3, False), (6, True)):
        want_xyz, want_tri = subdivide(xyz, tri, k, float32_points=f32)
        got_xyz, got_tri = subdivideOnDevice(xyz, tri, k, float32_points=f32)
        assert (got_xyz == want_xyz).all() and (got_tri == want_tri).all()
    assert got_tri.shape[0] == 84600
    vals = inheritCellData(numpy.arange(len(tri)) * 1.5, 6)
    assert vals.shape[0] == 84600 and vals[35] == 0. and vals[36] == 1.5
    # areas are conserved: the children tile their parent
    def areas(p, t):
        return numpy.linalg.norm(numpy.cross(p[t[:, 1]] - p[t[:, 0]], p[t[:, 2]] - p[t[:, 0]]), axis=1)
    x2, t2 = subdivideOnDevice(xyz, tri, 4, float32_points=False)
    assert numpy.abs(areas(x2, t2).reshape(-1, 16).sum(1) / areas(numpy.asarray(xyz, float), tri) - 1).max() < 1e-9


def test_example_solve_laplace_cli_ply_input(tmp_path, oracle_mod):
    """The command line's other input format (reference examples/solveLaplace.py:58-69): the same
    mesh as PLY gives the same response file content as the legacy-VTK input."""
    import importlib.util
    import os
    from icqsol_b200.mesh.generators import uvSphere
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.mesh.ply_io import writePly
    from icqsol_b200.mesh.vtk_legacy_io import readVtkPolyData, writeVtkPolyData
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('solveLaplacePly',
                                                  os.path.join(root, 'examples', 'solveLaplace.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    xyz, tri = uvSphere(1.0, 12, 7)                      # float32-representable points
    out = {}
    for kind in ('vtk', 'ply'):
        src, dst = str(tmp_path / ('in.' + kind)), str(tmp_path / ('out_' + kind + '.vtk'))
        if kind == 'ply':
            writePly(PolyData.from_arrays(xyz, tri), src, binary=True)
        else:
            writeVtkPolyData(PolyData.from_arrays(xyz, tri), src)
        assert mod.main(['--input', src, '--output', dst, '--dirichlet', 'x*y + z']) == 0
        out[kind] = readVtkPolyData(dst).GetCellData().GetArray('normal_electric_field').as_numpy()[:, 0]
    assert (out['ply'] == out['vtk']).all()
    g = oracle_mod.green_matrix(xyz, tri, 5)
    cen = (xyz[tri[:, 0]] + xyz[tri[:, 1]] + xyz[tri[:, 2]]) / 3.
    v = cen[:, 0] * cen[:, 1] + cen[:, 2]
    assert numpy.abs(v + g.dot(out['ply'])).max() / numpy.abs(v).max() < 1e-12


# ---- dense blocks over the sliver runs in the preconditioner (cg_caps.cu) --------------------
# (first run on a B200 in round 2: gpurun_out/r2a_*)


@pytest.mark.parametrize('storage', ['lower', 'full'])
def test_cap_blocks_on_sliver_sphere(oracle_mod, storage):
    """UV sphere with the pole slivers of the C3 size (n_theta = 318): diag(A) G is
    indefinite, the plain block preconditioner does not converge, blocks + caps does
    (109 iterations in the numpy model) and the response satisfies the oracle's system."""
    from icqsol_b200.mesh.generators import uvSphere
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver
    xyz, tri = uvSphere(1.0, 318, 20)
    solver = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5, storage=storage,
                           preconditioner='block+caps')
    assert solver.capTiles == (8, 9)
    solver.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
    v = solver.getSourceArray(solver.getSourceArrayIndex())
    rsp = solver.computeResponseField()
    assert solver.lastSolveInfo['converged']
    assert solver.lastSolveInfo['iterations'] < 300
    g = oracle_mod.green_matrix(xyz, tri, 5, threads=True)
    assert numpy.abs(v + g.dot(rsp)).max() / numpy.abs(v).max() < 1e-12


def test_cap_blocks_same_answer_on_regular_sphere():
    """On a mesh the plain blocks handle, explicit caps give the same response."""
    from icqsol_b200.mesh.generators import uvSphere
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver
    xyz, tri = uvSphere(1.0, 48, 24)
    out = []
    for caps in (None, (2, 3)):
        solver = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5,
                               preconditioner='block' if caps is None else 'block+caps')
        if caps:
            solver.ctx.check(solver.lib.icqb_cg_set_cap_tiles(solver.ctx.handle, *caps))
        solver.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
        v = solver.getSourceArray(solver.getSourceArrayIndex())
        rsp = solver.computeResponseField()
        g = solver.getGreenMatrix()
        assert numpy.abs(v + g.dot(rsp)).max() / numpy.abs(v).max() < 1e-12
        out.append((rsp, solver.lastSolveInfo['iterations']))
    assert relerr(out[1][0], out[0][0]) < 1e-6
    assert out[1][1] <= out[0][1] + 3


@pytest.mark.parametrize('storage', ['lower', 'full'])
def test_block_tiles_cut_iterations(storage):
    """Diagonal blocks of several tiles (batched block Gauss-Jordan on the device): fewer
    iterations than single tiles, same response."""
    from icqsol_b200.mesh.generators import uvSphere
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver
    xyz, tri = uvSphere(1.0, 64, 32)        # 3,968 triangles, 31 tile rows
    out = {}
    for k in (1, 4, 8):
        solver = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5, storage=storage,
                               preconditioner='block+caps', blockTiles=k)
        solver.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
        v = solver.getSourceArray(solver.getSourceArrayIndex())
        rsp = solver.computeResponseField()
        g = solver.getGreenMatrix()
        assert numpy.abs(v + g.dot(rsp)).max() / numpy.abs(v).max() < 1e-12
        out[k] = (rsp, solver.lastSolveInfo['iterations'])
    assert out[8][1] < out[4][1] < out[1][1]
    assert relerr(out[8][0], out[1][0]) < 1e-6


# ---- far-field split of the off-diagonal assembly (assemble_offdiag_ff.cu) -------------------

@pytest.mark.parametrize('variant', [1])
@pytest.mark.parametrize('nt,nph', [(16, 9), (64, 30)])
def test_far_field_split_every_entry(ctx, oracle_mod, variant, nt, nph):
    """Every G entry of a whole matrix within 1e-12 of the oracle and within a few ulp of the
    direct kernel, K against the definition, and most (warp, source triangle) pairs of the
    larger sphere on the far-field arithmetic."""
    from icqsol_b200.mesh.generators import uvSphere
    xyz, tri = uvSphere(1.0, nt, nph)
    n = tri.shape[0]
    g0, k0 = assemble(ctx, xyz, tri, 5, with_k=True)
    ctx.check(ctx.lib.icqb_set_assembly_variant(ctx.handle, variant))
    g, k = assemble(ctx, xyz, tri, 5, with_k=True)
    fast, allp = ctypes.c_int64(0), ctypes.c_int64(0)
    ctx.check(ctx.lib.icqb_assembly_stats(ctx.handle, ctypes.byref(fast), ctypes.byref(allp)))
    ref = oracle_mod.green_matrix(xyz, tri, 5, threads=True)
    assert max_entry_relerr(g, ref) < TOL_ENTRY
    assert max_entry_relerr(g, g0) < 5e-15
    assert numpy.abs(k - k0).max() < 1e-12 * numpy.abs(k0).max()
    assert (numpy.diag(k) == 0).all()
    assert allp.value > 0
    if n > 3000:
        assert fast.value > 0.7 * allp.value


@pytest.mark.parametrize('variant', [1])
def test_far_field_split_row_blocks_and_ragged(ctx, golden, oracle_mod, variant):
    """Row blocks (direct-only tiles left and right of the block) on the bolt, ragged sizes."""
    xyz, tri = golden.mesh('boltFine')
    n = tri.shape[0]
    ctx.check(ctx.lib.icqb_set_assembly_variant(ctx.handle, variant))
    for (r0, r1) in [(1000, 1300), (0, 129), (2349, 2350)]:
        g, k = assemble(ctx, xyz, tri, 5, with_k=True, rows=(r0, r1))
        gref = oracle_mod.block(xyz, tri, r0, r1, 0, n, 5)
        assert max_entry_relerr(g, gref) < TOL_ENTRY
        kref = oracle_mod.k_block(xyz, tri, r0, r1, 0, n)
        assert numpy.abs(k - kref).max() / numpy.abs(kref).max() < K_TOL
    xs, ts = golden.mesh('sphere')
    for m in (1, 2, 127, 129, 257):
        sub = numpy.ascontiguousarray(ts[:m])
        g = assemble(ctx, xs, sub, 5)
        assert max_entry_relerr(g, oracle_mod.green_matrix(xs, sub, 5)) < TOL_ENTRY


@pytest.mark.parametrize('variant', [1])
def test_far_field_split_config_c2(oracle_mod, variant):
    """Config C2 through the solver classes (lower storage): sampled columns of G against the
    oracle, the response against oracle rows, ~96 % of the pairs on the far-field arithmetic."""
    from icqsol_b200.mesh.generators import uvSphere
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver
    xyz, tri = uvSphere(1.0, 142, 71)
    solver = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5, computeK=True,
                           assemblyVariant=variant)
    fast, allp = ctypes.c_int64(0), ctypes.c_int64(0)
    solver.ctx.check(solver.lib.icqb_assembly_stats(solver.ctx.handle, ctypes.byref(fast),
                                                    ctypes.byref(allp)))
    assert 0.94 * allp.value < fast.value < 0.97 * allp.value
    solver.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
    v = solver.getSourceArray(solver.getSourceArrayIndex())
    rsp = solver.computeResponseField()
    assert solver.lastSolveInfo['converged']
    _sampled_checks(solver, xyz, tri, oracle_mod, [0, 141, 9999, 19751, 19879], rsp, v)


@pytest.mark.parametrize('storage', ['lower', 'full'])
def test_direct_solve_matches_reference_lu(golden, storage):
    """solveGreenSystemDirect = LU with partial pivoting on the device matrix, against the
    reference's own solve (numpy.linalg.solve on the reference G, golden LU response)."""
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver
    xyz, tri = golden.mesh('boltFine')
    solver = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5, storage=storage)
    solver.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
    v = solver.getSourceArray(solver.getSourceArrayIndex())
    x = solver.solveGreenSystemDirect(v)
    assert solver.lastSolveInfo['method'] == 'lu' and solver.lastSolveInfo['converged']
    g = solver.getGreenMatrix()
    assert numpy.abs(v - g.dot(x)).max() < 1e-11 * numpy.abs(v).max()
    ref = numpy.linalg.solve(g, v)
    assert relerr(x, ref) < 1e-8          # cond(G) ~ 2e4: two LUs agree to ~cond * eps
    # the lower storage is untouched by the temporary row-major assembly
    rsp = solver.computeResponseField()
    assert relerr(-rsp, ref) < 1e-8


def test_direct_fallback_on_sliver_sphere(oracle_mod):
    """The UV sphere with C3-size pole slivers: the block-preconditioned CG does not converge
    (DESIGN.md 4.6); with directFallback the solver returns the LU solution."""
    import warnings
    from icqsol_b200.mesh.generators import uvSphere
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver
    xyz, tri = uvSphere(1.0, 318, 20)
    solver = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5, preconditioner='block')
    solver.setSolverMaxNumberOfIterations(300)
    solver.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
    v = solver.getSourceArray(solver.getSourceArrayIndex())
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter('always')
        rsp = solver.computeResponseField()
    assert any('solving by LU' in str(w.message) for w in caught)
    assert solver.lastSolveInfo['method'] == 'lu'
    g = oracle_mod.green_matrix(xyz, tri, 5, threads=True)
    assert numpy.abs(v + g.dot(rsp)).max() / numpy.abs(v).max() < 1e-10


@pytest.mark.parametrize('nt,nph', [(16, 9), (64, 30)])
def test_float32_product_kernel(oracle_mod, nt, nph):
    """k_symv_f32 (icqb_symv_f32_dev): the product with the float32-rounded stored entries,
    sums in double."""
    import torch
    from icqsol_b200.mesh.generators import uvSphere
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqPoissonSolver import PoissonSolver
    xyz, tri = uvSphere(1.0, nt, nph)
    n = tri.shape[0]
    solver = PoissonSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5, storage='lower')
    g = solver.getGreenMatrix()
    a2 = solver.getTriangleAreas2()
    t = numpy.arange(n) // 128
    stored = t[None, :] <= t[:, None]
    g32 = numpy.where(stored, g, 0.).astype(numpy.float32).astype(numpy.float64)
    full = g32 + numpy.where(t[None, :] < t[:, None], g32, 0.).T * (a2[None, :] / a2[:, None])
    rng = numpy.random.default_rng(n)
    x = rng.normal(size=n)
    dx = torch.from_numpy(x).to(solver.gMatDevice.device)
    dy = torch.zeros(n, dtype=torch.float64, device=dx.device)
    solver.ctx.check(solver.lib.icqb_symv_f32_dev(solver.ctx.handle, solver.gMatDevice.data_ptr(),
                                                  dx.data_ptr(), dy.data_ptr(), 0))
    y = dy.cpu().numpy()
    assert (numpy.abs(y - full.dot(x)) <= 1e-13 * numpy.abs(full).dot(numpy.abs(x))).all()


def test_mixed_precision_cg():
    """Float32 late products (switch at 1e-6, off by default) in the solver of kind 2: same backward error,
    a good share of the products in float32, a few more iterations at most."""
    from icqsol_b200.mesh.generators import uvSphere
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver
    xyz, tri = uvSphere(1.0, 64, 32)
    out = {}
    for mixed in (0., 1e-6):
        solver = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5,
                               preconditioner='block+caps', blockTiles=4, mixedPrecision=mixed)
        solver.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
        v = solver.getSourceArray(solver.getSourceArrayIndex())
        rsp = solver.computeResponseField()
        g = solver.getGreenMatrix()
        assert numpy.abs(v + g.dot(rsp)).max() / numpy.abs(v).max() < 1e-12
        out[mixed] = (rsp, solver.lastSolveInfo)
    assert out[0.][1]['float32Products'] == 0
    assert out[1e-6][1]['float32Products'] > 0.3 * out[1e-6][1]['iterations']
    assert out[1e-6][1]['iterations'] <= out[0.][1]['iterations'] + 8
    assert relerr(out[1e-6][0], out[0.][0]) < 1e-8
