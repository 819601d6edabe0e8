"""The oracle (oracle/icq_oracle.c, oracle/oracle.py) pinned against golden
vectors produced by the reference's own C++ and Python (tests/golden/*.npz,
generator tests/golden/make_golden.py) and -- when oracle/_ref/libicqref.so was
built from /root/reference -- against the reference C++ directly."""
import os
import re

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ['twotri', 'tet', 'sphere4x2', 'createSphere', 'square']


def test_tables_match_reference_in_both_headers(golden):
    """The 14-digit truncated triangle tables (bem/icqQuadrature.cpp:20-144) and the
    Gauss-Legendre tables (bem/icqQuadrature1D.py:9-25), in the oracle's header
    and in the CUDA header."""
    t = golden['tables']
    ref_tri = numpy.concatenate([t['tri%d' % o].T for o in range(1, 9)])
    ref_line = numpy.concatenate([t['line%d' % o] for o in range(1, 6)])
    src = open(os.path.join(ROOT, 'icqsol_b200', 'csrc', 'quad_tables.cuh')).read()
    body = src[src.index('ICQB_TRI_NODES_INIT'):src.index('constexpr TriNode kTriNodes')]
    nodes = numpy.array(re.findall(r'\{(-?[0-9.]+), (-?[0-9.]+), (-?[0-9.]+)\}', body), float)
    assert nodes.shape == ref_tri.shape and (nodes == ref_tri).all()
    body = src[src.index('ICQB_LINE_NODES_INIT'):src.index('constexpr LineNode kLineNodes')]
    line = numpy.array(re.findall(r'\{([0-9.]+), ([0-9.]+)\}', body), float)
    assert line.shape == ref_line.shape and (line == ref_line).all()
    hdr = open(os.path.join(ROOT, 'oracle', 'icq_tables.h')).read()

    def arr(name):
        m = re.search(name + r'\[\d+\] = \{(.*?)\};', hdr, re.S)
        return numpy.array([float(x) for x in m.group(1).replace('\n', ' ').split(',')])
    assert (arr('ICQO_TRI_XI') == ref_tri[:, 0]).all()
    assert (arr('ICQO_TRI_ETA') == ref_tri[:, 1]).all()
    assert (arr('ICQO_TRI_W') == ref_tri[:, 2]).all()
    assert (arr('ICQO_LINE_U') == ref_line[:, 0]).all()
    assert (arr('ICQO_LINE_W') == ref_line[:, 1]).all()
    assert [len(t['tri%d' % o][0]) for o in range(1, 9)] == [1, 3, 4, 6, 7, 12, 13, 16]


@pytest.mark.parametrize('name', SMALL)
def test_offdiag_bit_exact_vs_reference_cpp(oracle_mod, golden, name):
    xyz, tri = golden.mesh(name)
    g = oracle_mod.offdiag(xyz, tri)
    assert (g == golden['green_small'][name + '_offdiag']).all()
    assert (numpy.diag(g) == 0).all()     # diagonal untouched


@pytest.mark.parametrize('name', SMALL)
@pytest.mark.parametrize('order', [1, 2, 3, 4, 5])
def test_diag_vs_reference_python(oracle_mod, golden, name, order):
    xyz, tri = golden.mesh(name)
    d = oracle_mod.diag(xyz, tri, order)
    assert numpy.abs(d / golden['green_small']['%s_diag%d' % (name, order)] - 1).max() < 1e-14


def test_survey_values(oracle_mod, golden):
    g = oracle_mod.green_matrix(*golden.mesh('tet'), 5)
    assert g[0, 1] == -0.08250452568311496
    assert g[0, 3] == -0.16292331240300845
    assert g[3, 0] == -0.09406381827314243
    assert abs(g[3, 3] / -0.21436428380303485 - 1) < 1e-15
    g = oracle_mod.green_matrix(*golden.mesh('twotri'), 5)
    assert abs(g[0, 0] / -0.15959645874363654 - 1) < 1e-15 and g[1, 0] == -0.03637297439435873
    assert oracle_mod.offdiag(*golden.mesh('square'))[1, 0] == -0.08186618508970281


@pytest.mark.parametrize('name', ['sphere', 'boltFine'])
def test_real_meshes_sampled(oracle_mod, golden, name):
    xyz, tri = golden.mesh(name)
    big = golden['green_sampled']
    rows = big[name + '_rows']
    n = tri.shape[0]
    for k, r in enumerate(rows[:4]):
        blk = oracle_mod.block(xyz, tri, int(r), int(r) + 1, 0, n, 5, threads=False)
        assert numpy.abs(blk[0] / big[name + '_G_rows'][k] - 1).max() < 1e-14
    d = oracle_mod.diag(xyz, tri, 5)
    assert numpy.abs(d / big[name + '_diag5'] - 1).max() < 1e-14
    # block() and offdiag_entry() agree with the mirror rule of the full assembly
    sub = tri[:60]
    full = oracle_mod.green_matrix(xyz, sub, 5)
    assert (oracle_mod.block(xyz, sub, 0, 60, 0, 60, 5) == full).all()


def test_corner_integrals(oracle_mod, golden):
    """PotentialIntegrals.getIntegralOneOverR (bem/icqPotentialIntegrals.py:37-40) and
    the exact values quoted in tests/testPotentialIntegrals.py:13,23,33."""
    kat = golden['known_answers']
    for i, t in enumerate(kat['corner_tris']):
        for order in range(1, 6):
            v = oracle_mod.corner_integral(order, t[0], t[1], t[2])
            assert abs(v / kat['corner_vals'][i, order - 1] - 1) < 1e-14
    unit = [oracle_mod.corner_integral(o, [0., 0, 0], [1., 0, 0], [0., 1, 0]) for o in range(1, 6)]
    assert numpy.abs(numpy.array(unit) / kat['corner_unit_right'] - 1).max() < 1e-15
    assert abs(unit[4] - 1.2464461713881911) < 1e-15
    exact = numpy.sqrt(2.) * numpy.arcsinh(1.)
    assert abs(unit[4] / exact - 1) < 1e-3          # order-5 angular rule, not the closed form
    assert numpy.isnan(oracle_mod.corner_integral(6, [0., 0, 0], [1., 0, 0], [0., 1, 0]))
    with pytest.raises(KeyError):
        oracle_mod.diag(*golden.mesh('tet'), 6)


def test_quadrature_single_and_double(oracle_mod, golden):
    kat = golden['known_answers']
    for i, t in enumerate(kat['corner_tris']):
        for order in range(10):
            v = oracle_mod.quadrature_evaluate(order, kat['single_obs'][i], t[0], t[1], t[2])
            assert v == kat['single_vals'][i, order]
    for i, p in enumerate(kat['double_pairs']):
        for order in range(10):
            v = oracle_mod.quadrature_evaluate_double(order, p[0, 0], p[0, 1], p[0, 2],
                                                      p[1, 0], p[1, 1], p[1, 2])
            assert v == kat['double_vals'][i, order]
    assert oracle_mod.quadrature_evaluate(5, [0., 0, 0], [1., 1, 1], [1., 1, 1], [1., 1, 1]) == 0.0


def test_reference_cpp_direct(oracle_mod, golden):
    """Only where oracle/_ref/libicqref.so exists (built from /root/reference)."""
    if not oracle_mod.have_reference():
        pytest.skip('reference C++ not built here')
    from icqsol_b200.mesh.generators import uvSphere
    xyz, tri = uvSphere(1.0, 12, 7)
    assert (oracle_mod.ref_offdiag(xyz, tri) == oracle_mod.offdiag(xyz, tri)).all()
    xyz, tri = golden.mesh('boltFine')
    sub = tri[1000:1150]
    assert (oracle_mod.ref_offdiag(xyz, sub) == oracle_mod.offdiag(xyz, sub, threads=True)).all()


def test_cg_restatement_vs_reference_class(oracle_mod, golden):
    """oracle.conjugate_gradient == solvers/icqConjugateGradient.py on the
    area-symmetrised system (iteration count and iterate), and its divergence /
    slow convergence on the raw non-symmetric G."""
    from icqsol_b200.mesh.generators import uvSphere
    cg = golden['cg']
    nt, npz = cg['uv16x8_nt_np']
    xyz, tri = uvSphere(1.0, int(nt), int(npz))
    g = oracle_mod.green_matrix(xyz, tri, 5)
    assert numpy.abs(numpy.diag(g) / cg['uv16x8_G_diag'] - 1).max() < 1e-14
    assert numpy.abs(g[0, 1:] / cg['uv16x8_G_row0'][1:] - 1).max() < 1e-15
    a2 = oracle_mod.area2(xyz, tri)
    cen = (xyz[tri[:, 0]] + xyz[tri[:, 1]] + xyz[tri[:, 2]]) / 3.
    v = 1. / numpy.sqrt(cen[:, 0] ** 2 + (cen[:, 1] - 2.) ** 2 + (cen[:, 2] - 4.) ** 2)
    x, err, k = oracle_mod.conjugate_gradient(a2[:, None] * g, a2 * v, numpy.zeros(len(v)),
                                              tol=1.e-13)
    assert k == int(cg['uv16x8_iters'])
    assert numpy.abs(x - cg['uv16x8_x']).max() / numpy.abs(x).max() < 1e-12
    lu = numpy.linalg.solve(g, v)
    assert numpy.abs(lu - cg['uv16x8_lu']).max() / numpy.abs(lu).max() < 1e-12
    assert int(cg['uv40x20_raw_iters']) == 200 and float(cg['uv40x20_raw_err']) > 1e-6


def test_k_definition_properties(oracle_mod):
    """K is unpinned (not in the reference): check the definition through identities --
    row sums -> 1/2 on a closed outward surface, sign flips with orientation."""
    from icqsol_b200.mesh.generators import uvSphere
    xyz, tri = uvSphere(1.0, 24, 12)
    n = tri.shape[0]
    k = oracle_mod.k_block(xyz, tri, 0, 16, 0, n)
    assert numpy.abs(k.sum(axis=1) - 0.5).max() < 0.05
    kin = oracle_mod.k_block(xyz, tri[:, ::-1].copy(), 0, 16, 0, n)
    assert numpy.abs(kin + k).max() < 1e-13   # reversed vertex order: different quadrature-node order


def _solid_angle(x, a, b, c):
    """Signed solid angle of triangle (a, b, c) seen from x (van Oosterom & Strackee 1983),
    positive when x lies on the side the normal (b - a) x (c - a) points away from."""
    ra, rb, rc = a - x, b - x, c - x
    la, lb, lc = (numpy.sqrt((v * v).sum(-1)) for v in (ra, rb, rc))
    num = (ra * numpy.cross(rb, rc)).sum(-1)
    den = la * lb * lc + (ra * rb).sum(-1) * lc + (ra * rc).sum(-1) * lb + (rb * rc).sum(-1) * la
    return 2. * numpy.arctan2(num, den)


def test_k_against_the_analytic_double_layer_of_a_flat_triangle(oracle_mod):
    """An anchor for K that does not come from the code under test (the reference snapshot has no
    K): the double-layer potential of a flat triangle is its solid angle,
    integral over T_j of n_j.(x - y)/|x - y|^3 dS_y = -Omega_j(x) (signed as in _solid_angle), so
        K[i,j] = sum_k w_k Omega_j(o_k) / (4 pi)     up to the source-side quadrature error,
    o_k the observer triangle's quadrature points.  Measured on this mesh: pairs with
    |c_i - c_j| > 3 (L_i + L_j) (45 % of them) agree to 7e-13 of the largest such entry, > 2 (L_i +
    L_j) to 5e-11, > (L_i + L_j) (93 %) to 7e-8; touching triangles do not (a 16-point rule on a
    nearly singular kernel -- the same is true of the reference's G).  Checks the sign, the
    normal's orientation, the 1/(4 pi), the area factor and the observer averaging."""
    from icqsol_b200.mesh.generators import uvSphere
    xyz, tri = uvSphere(1.0, 32, 16)
    n = tri.shape[0]
    k = oracle_mod.k_block(xyz, tri, 0, n, 0, n)
    nodes = numpy.array(oracle_mod.tri_rule(8))            # [16, 3]: xi, eta, w
    a, b, c = xyz[tri[:, 0]], xyz[tri[:, 1]], xyz[tri[:, 2]]
    obs = a[:, None, :] + nodes[None, :, 0:1] * (b - a)[:, None, :] + nodes[None, :, 1:2] * (c - a)[:, None, :]
    exact = numpy.zeros((n, n))
    for j in range(n):
        om = _solid_angle(obs.reshape(-1, 3), a[j], b[j], c[j]).reshape(n, 16)
        exact[:, j] = (om * nodes[None, :, 2]).sum(1) / (4. * numpy.pi)
    numpy.fill_diagonal(exact, 0.)
    cen = (a + b + c) / 3.
    dist = numpy.sqrt(((cen[:, None, :] - cen[None, :, :]) ** 2).sum(-1))
    size = numpy.sqrt(numpy.maximum.reduce([((b - a) ** 2).sum(1), ((c - b) ** 2).sum(1), ((a - c) ** 2).sum(1)]))
    for kappa, share, tol in ((3., 0.4, 5e-12), (2., 0.7, 5e-10), (1., 0.9, 1e-6)):
        far = dist > kappa * (size[:, None] + size[None, :])
        assert far.sum() > share * n * n
        assert numpy.abs(k[far] - exact[far]).max() < tol * numpy.abs(exact[far]).max()
    # closed outward surface: the exact entries of a row add up to the interior-limit value 1/2
    # of the double-layer potential minus the (flat, vanishing) self term
    assert numpy.abs(exact.sum(1) - 0.5).max() < 1e-12
