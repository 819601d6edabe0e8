"""The CUDA translation units, compiled for the CPU and run thread by thread
(tools/emu: one fiber per CUDA thread, blocks one after the other, bulk copies and
mbarriers modelled, "device" memory = host memory), driven through the same C ABI.

What this checks where there is no GPU: the kernels' indexing, tile/strip/chunk
partitions, head slots, shuffles, reductions, the device-side CG state machine and the
host orchestration -- against the oracle.  What it cannot check: races, memory ordering,
PTX, performance.  The product never loads this build (different file, only this module
asks for it); the `-m gpu` tests remain the parity tests proper.
"""
import ctypes
import os
import sys

import numpy
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools', 'emu'))


@pytest.fixture(scope='module')
def emu():
    import build as emu_build
    from icqsol_b200 import _lib
    # every "device" allocation ends at an inaccessible page: a kernel that runs past the end
    # of a buffer faults (the emulation's memcheck, tools/emu/emu_runtime.cpp)
    os.environ.setdefault('ICQB_EMU_GUARD', '1')
    lib = ctypes.CDLL(emu_build.build())
    for name, restype, argtypes in _lib.SIGNATURES:
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    lib.emu_alloc.restype = ctypes.c_void_p
    lib.emu_alloc.argtypes = [ctypes.c_size_t]
    return lib


def dev_array(lib, shape, fill=numpy.nan):
    """float64 array in a guarded "device" allocation (never freed: test-sized), for the
    caller-owned matrices and vectors the kernels write."""
    shape = tuple(int(v) for v in numpy.atleast_1d(shape))
    count = int(numpy.prod(shape))
    ptr = lib.emu_alloc(8 * max(count, 1))
    assert ptr
    arr = numpy.ctypeslib.as_array((ctypes.c_double * max(count, 1)).from_address(ptr))[:count]
    arr = arr.reshape(shape)
    arr[...] = fill
    return arr


def dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def lp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


class Ctx:
    """One emulated context with a UV-sphere mesh; sms = SM count the library sees (the
    persistent product kernel sizes its grid and its static partition from it)."""

    def __init__(self, lib, n_theta, n_phi, sms=4):
        from icqsol_b200.mesh.generators import uvSphere
        self.lib = lib
        os.environ['ICQB_EMU_SMS'] = str(sms)
        self.h = ctypes.c_void_p()
        assert lib.icqb_create(ctypes.byref(self.h), 0) == 0
        xyz, tri = uvSphere(1.0, n_theta, n_phi)
        self.xyz, self.tri = numpy.ascontiguousarray(xyz), numpy.ascontiguousarray(tri)
        self.n = self.tri.shape[0]
        self.check(lib.icqb_set_mesh(self.h, dp(self.xyz), self.xyz.shape[0], lp(self.tri), self.n))

    def check(self, status):
        if status != 0:
            raise RuntimeError(self.lib.icqb_last_error(self.h))

    def close(self):
        self.lib.icqb_destroy(self.h)

    def rhs(self):
        t, x = self.tri, self.xyz
        cen = (x[t[:, 0]] + x[t[:, 1]] + x[t[:, 2]]) / 3.
        return -1. / numpy.sqrt(cen[:, 0] ** 2 + (cen[:, 1] - 2.) ** 2 + (cen[:, 2] - 4.) ** 2)

    def assemble_lower(self, blocks=None, with_k=False):
        n = self.n
        blocks = blocks or (0, n, n, n)
        self.check(self.lib.icqb_set_row_blocks(self.h, *blocks))
        nt = self.lib.icqb_lower_num_tiles(self.h)
        bufs = [dev_array(self.lib, (max(nt, 1), 128, 128)) for _ in range(3 if with_k else 1)]
        ptr = [b.ctypes.data for b in bufs] + [0, 0]
        self.check(self.lib.icqb_assemble_lower_dev(self.h, 5, ptr[0], ptr[1], ptr[2]))
        return bufs

    def solve(self, g, storage, kind, caps=None, ld=0, maxit=500, starts=None, extents=None):
        n = self.n
        b = dev_array(self.lib, n)
        b[:] = self.rhs()
        x = dev_array(self.lib, n, 0.)
        self.check(self.lib.icqb_cg_set_preconditioner(self.h, kind))
        self.check(self.lib.icqb_cg_set_cap_tiles(self.h, *(caps or (0, 0))))
        st = numpy.asarray(starts or [], numpy.int64)
        self.check(self.lib.icqb_cg_set_block_partition(self.h, lp(st) if len(st) else None, len(st)))
        if extents is not None:
            e0 = numpy.asarray(extents[0], numpy.int64)
            e1 = numpy.asarray(extents[1], numpy.int64)
            self.check(self.lib.icqb_cg_set_block_extents(self.h, lp(e0), lp(e1), len(e0)))
        it, r2, ri = ctypes.c_int64(0), ctypes.c_double(0.), ctypes.c_double(0.)
        self.check(self.lib.icqb_cg_solve_dev(
            self.h, g.ctypes.data, n, ld, storage, self.lib.icqb_area2_dev(self.h), 0, b.ctypes.data,
            x.ctypes.data, 5e-13, 2, maxit, ctypes.byref(it), ctypes.byref(r2), ctypes.byref(ri)))
        return x, it.value, ri.value, b


def model_iterations(g, a2, b, cuts, extents=None):
    """numpy model of the block-preconditioned CG of csrc/cg.cu with diagonal blocks
    [cuts[k], cuts[k+1]): iterations to max|r/s| <= 5e-13 max|b|.  ``extents`` (first rows, end
    rows): overlapping blocks, weighted 1/sqrt(cover) on both sides (symmetric additive Schwarz)."""
    n = len(b)
    ms = a2[:, None] * g
    ms = 0.5 * (ms + ms.T)
    cuts = [c for c in cuts if c < n] + [n]
    spans = list(zip(cuts[:-1], cuts[1:]))
    if extents is not None:
        spans = [(int(p), min(int(q), n)) for p, q in zip(*extents)]
    cover = numpy.zeros(n)
    for p, q in spans:
        cover[p:q] += 1.
    d = 1. / numpy.sqrt(cover)
    invs = [numpy.linalg.inv(ms[p:q, p:q]) for p, q in spans]

    def pre(r):
        out = numpy.zeros_like(r)
        for k, (p, q) in enumerate(spans):
            out[p:q] += d[p:q] * invs[k].dot(d[p:q] * r[p:q])
        return out
    x, r, p, rho_old = numpy.zeros(n), a2 * b, numpy.zeros(n), 1.
    w = pre(r)
    for k in range(2000):
        rho = r.dot(w)
        p = w + (rho / rho_old if k else 0.) * p
        z = ms.dot(p)
        alpha = rho / p.dot(z)
        x += alpha * p
        r -= alpha * z
        w = pre(r)
        rho_old = rho
        if numpy.abs(r / a2).max() <= 5e-13 * numpy.abs(b).max():
            return k + 1
    return -1


def test_emulated_assembly_full_storage(emu, oracle_mod):
    """k_prepare, k_diag, k_offdiag<G+K> (full storage, host-buffer API) on 144 triangles."""
    c = Ctx(emu, 12, 7)
    n = c.n
    g, k = numpy.full((n, n), numpy.nan), numpy.full((n, n), numpy.nan)
    c.check(emu.icqb_assemble_host(c.h, 5, dp(g), dp(k)))
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
    assert numpy.abs(g / ref - 1.).max() < 1e-12
    kref = oracle_mod.k_block(c.xyz, c.tri, 0, n, 0, n)
    assert numpy.abs(k - kref).max() < 1e-9 * numpy.abs(kref).max()
    c.close()


@pytest.mark.parametrize('sms', [1, 3, 4])
def test_emulated_lower_storage_and_product(emu, oracle_mod, sms):
    """Tile-contiguous lower storage (every stored entry, K and its transpose) and the
    persistent product kernel for several static partitions of the chunk sequence: with 1 SM
    a warp owns several chunks of a tile, with 4 most spans start inside a tile (head slots)."""
    from icqsol_b200.bem import icqRowPartition as rp
    c = Ctx(emu, 16, 9, sms=sms)
    n = c.n                                            # 256 triangles: 2 strips, 3 tiles
    g, kl, ku = c.assemble_lower(with_k=True)
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
    kref = oracle_mod.k_block(c.xyz, c.tri, 0, n, 0, n)
    rows = rp.lowerTilesToRows(g, (0, n, n, n), n)
    krows = rp.lowerTilesToRows(kl, (0, n, n, n), n)
    kurows = rp.lowerTilesToRows(ku, (0, n, n, n), n)
    for i in range(n):
        ce = min(n, (i // 128 + 1) * 128)
        assert numpy.abs(rows[i, :ce] / ref[i, :ce] - 1.).max() < 1e-12
        assert numpy.abs(krows[i, :ce] - kref[i, :ce]).max() < 1e-9 * numpy.abs(kref).max()
        assert numpy.abs(kurows[i, :ce] - kref[:ce, i]).max() < 1e-9 * numpy.abs(kref).max()
    rng = numpy.random.default_rng(sms)
    a2 = oracle_mod.area2(c.xyz, c.tri)
    for scaled in (0, 1):
        x, y = rng.normal(size=n), dev_array(emu, n)
        c.check(emu.icqb_symv_dev(c.h, g.ctypes.data, x.ctypes.data, y.ctypes.data, scaled))
        want = ref.dot(x) * (a2 if scaled else 1.)
        bound = numpy.abs(ref).dot(numpy.abs(x)) * (a2 if scaled else 1.)
        assert (numpy.abs(y - want) <= 1e-13 * bound).all()
    c.close()


def test_emulated_product_of_partial_row_blocks(emu, oracle_mod):
    """A rank that stores two separated row blocks (the folded layout of a multi-GPU run):
    its product is the operator restricted to the stored entries and their mirrors."""
    c = Ctx(emu, 20, 11, sms=2)                        # 400 triangles, tile rows 0..3
    n = c.n
    blocks = (128, 256, 384, 400)
    (g,) = c.assemble_lower(blocks)
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
    a2 = oracle_mod.area2(c.xyz, c.tri)
    mask = numpy.zeros((n, n), bool)
    for i in list(range(128, 256)) + list(range(384, 400)):
        mask[i, :min(n, (i // 128 + 1) * 128)] = True
    tile = numpy.arange(n) // 128
    lower = mask & (tile[None, :] < tile[:, None])
    full = numpy.where(mask, ref, 0.) + numpy.where(lower, ref, 0.).T * (a2[None, :] / a2[:, None])
    x, y = numpy.random.default_rng(1).normal(size=n), dev_array(emu, n)
    c.check(emu.icqb_symv_dev(c.h, g.ctypes.data, x.ctypes.data, y.ctypes.data, 0))
    assert (numpy.abs(y - full.dot(x)) <= 1e-13 * numpy.abs(full).dot(numpy.abs(x)) + 1e-300).all()
    c.close()


def test_emulated_cg_matches_model(emu, oracle_mod):
    """The device-side CG state machine (three kernels per iteration, batches in flight,
    confirmation with the true residual): same iteration count as the numpy model of the
    same preconditioner, solution checked with the oracle's matrix."""
    c = Ctx(emu, 12, 7)                                # 144 triangles: blocks of 128 and 16
    n = c.n
    (g,) = c.assemble_lower()
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
    a2 = oracle_mod.area2(c.xyz, c.tri)
    for kind, cuts in ((1, [0, 128]), (0, list(range(n)))):
        x, iters, rinf, b = c.solve(g, 1, kind)
        assert numpy.abs(b - ref.dot(x)).max() <= 1e-12 * numpy.abs(b).max()
        assert abs(iters - model_iterations(ref, a2, b, cuts)) <= 1
        assert rinf <= 5e-13 * numpy.abs(b).max()
    # row-block storage: k_gemv, k_cg_dir, k_cg_dot
    ld = (n + 15) // 16 * 16
    gf = dev_array(emu, (n, ld), 0.)
    c.check(emu.icqb_set_rows(c.h, 0, n))
    c.check(emu.icqb_assemble_dev(c.h, 5, gf.ctypes.data, 0, ld))
    x, iters, rinf, b = c.solve(gf, 0, 1, ld=ld)
    assert numpy.abs(b - ref.dot(x)).max() <= 1e-12 * numpy.abs(b).max()
    assert abs(iters - model_iterations(ref, a2, b, [0, 128])) <= 1
    c.close()


@pytest.mark.parametrize('caps,cuts', [((1, 1), [0, 128]), ((2, 0), [0])])
def test_emulated_cap_blocks(emu, oracle_mod, caps, cuts):
    """csrc/cg_caps.cu (preconditioner kind 2): dense cap blocks --
    extraction with the symmetric fill and the identity padding of the ragged last tile,
    Gauss-Jordan in global memory, split residual kernels, cap GEMV -- against the numpy
    model: caps of one tile each are the plain blocks, one cap over everything is the
    exact inverse (one iteration)."""
    c = Ctx(emu, 12, 7)
    (g,) = c.assemble_lower()
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
    a2 = oracle_mod.area2(c.xyz, c.tri)
    x, iters, rinf, b = c.solve(g, 1, 2, caps)
    assert numpy.abs(b - ref.dot(x)).max() <= 1e-12 * numpy.abs(b).max()
    assert abs(iters - model_iterations(ref, a2, b, cuts)) <= 1
    c.close()


def test_emulated_gemv(emu):
    """k_gemv, aligned and unaligned paths, ragged sizes."""
    c = Ctx(emu, 6, 4)
    rng = numpy.random.default_rng(0)
    for nrows, ncols, ld in ((5, 130, 130), (37, 129, 131), (4, 16, 16), (9, 1, 3)):
        a = dev_array(emu, (nrows, ld))
        a[...] = rng.normal(size=(nrows, ld))
        x, y = rng.normal(size=ncols), dev_array(emu, nrows)
        c.check(emu.icqb_gemv_host(c.h, a.ctypes.data, nrows, ncols, ld, dp(x), dp(y)))
        want = a[:, :ncols].dot(x)
        assert (numpy.abs(y - want) <= 1e-13 * numpy.abs(a[:, :ncols]).dot(numpy.abs(x))).all()
    c.close()


@pytest.mark.parametrize('k,f32', [(1, 1), (2, 1), (3, 0), (6, 1)])
def test_emulated_subdivide(emu, golden, k, f32):
    """k_subdivide: bit-identical points and the same children as the host version
    (mesh.generators.subdivide) on the reference's bolt mesh."""
    from icqsol_b200.mesh.generators import subdivide
    xyz, tri = golden.mesh('boltFine')
    xyz = numpy.ascontiguousarray(xyz[:], numpy.float64)
    tri = numpy.ascontiguousarray(tri[:300], numpy.int64)
    want_xyz, want_tri = subdivide(xyz, tri, k, float32_points=bool(f32))
    c = Ctx(emu, 6, 4)
    npl = (k + 1) * (k + 2) // 2
    got_xyz = numpy.full((len(tri) * npl, 3), numpy.nan)
    got_tri = numpy.full((len(tri) * k * k, 3), -1, numpy.int64)
    c.check(emu.icqb_subdivide_host(c.h, dp(xyz), len(xyz), lp(tri), len(tri), k, f32,
                                    dp(got_xyz), lp(got_tri)))
    if k == 1:
        # the host version returns the mesh itself for k = 1; the kernel re-indexes per parent
        assert (got_xyz[got_tri] == want_xyz[want_tri]).all()
    else:
        assert (got_xyz == want_xyz).all()
        assert (got_tri == want_tri).all()
    c.close()


def _run_ranks(emu, nranks, n_theta, n_phi, storage, kind, caps=None, peer=False, starts=None,
               mixed=0.):
    """One rank per thread (ctypes releases the GIL), the in-process NCCL stand-in of
    tools/emu underneath: folded row blocks, sharded assembly, distributed CG.  peer=True
    also sets up the optional peer exchange (handles swapped through a shared list)."""
    import threading
    from icqsol_b200.bem import icqRowPartition as rp
    uid = ctypes.create_string_buffer(128)
    assert emu.icqb_comm_unique_id(uid) == 0
    out, errors = [None] * nranks, []
    handles = [None] * nranks
    meet = threading.Barrier(nranks, timeout=120)

    def work(rank):
        try:
            c = Ctx(emu, n_theta, n_phi, sms=2)
            c.check(emu.icqb_comm_init(c.h, uid, rank, nranks))
            c.check(emu.icqb_cg_set_mixed_precision(c.h, mixed))
            n = c.n
            if peer:
                h = ctypes.create_string_buffer(64)
                c.check(emu.icqb_peer_create(c.h, 1024, h))
                handles[rank] = h.raw
                meet.wait()
                c.check(emu.icqb_peer_attach(c.h, ctypes.create_string_buffer(b''.join(handles), 64 * nranks)))
                meet.wait()
            if storage == 1:
                (g,) = c.assemble_lower(rp.foldedBlocks(n, rank, nranks))
                ld = 0
            else:
                r0, r1 = rp.rowRange(n, rank, nranks)
                ld = (n + 15) // 16 * 16
                g = dev_array(emu, (max(r1 - r0, 1), ld), 0.)
                c.check(emu.icqb_set_rows(c.h, r0, r1))
                c.check(emu.icqb_assemble_dev(c.h, 5, g.ctypes.data, 0, ld))
            out[rank] = c.solve(g, storage, kind, caps, ld=ld, starts=starts) + (c,)
        except Exception as e:    # noqa: BLE001 -- reported by the caller
            errors.append((rank, repr(e)))
    threads = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not errors, errors
    assert all(o is not None for o in out), 'a rank did not finish'
    return out


@pytest.mark.parametrize('nranks,storage,kind,caps,peer', [
    (2, 1, 1, None, False), (3, 1, 0, None, False), (2, 0, 1, None, False), (2, 1, 2, (1, 2), False),
    (3, 1, 2, None, True)])
def test_emulated_multi_rank_cg(emu, oracle_mod, nranks, storage, kind, caps, peer):
    """Sharded assembly and the distributed CG (all-reduce of the partial products with p.z in
    the slot behind them, all-reduce of the preconditioner blocks / cap blocks, all-gather for
    row blocks): every rank returns the same solution, it solves the oracle's system, and
    the iteration count is the single-rank one.  The last case runs the optional peer
    exchange (stores into the peers' buffers + flags instead of the all-reduce)."""
    out = _run_ranks(emu, nranks, 20, 11, storage, kind, caps, peer)     # 400 triangles, 4 tile rows
    c = out[0][-1]
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
    a2 = oracle_mod.area2(c.xyz, c.tri)
    x0, iters0, _, b, _ = out[0]
    for x, iters, rinf, _, _ in out:
        assert (x == x0).all() and iters == iters0
        assert rinf <= 5e-13 * numpy.abs(b).max()
    assert numpy.abs(b - ref.dot(x0)).max() <= 1e-12 * numpy.abs(b).max()
    cuts = {0: list(range(c.n)), 1: [0, 128, 256, 384], 2: [0, 128, 256, 384]}[kind]
    if caps == (1, 2):
        cuts = [0, 128, 256]      # head cap = tile 0, tiles 1 plain, tail cap = tiles 2..3
    assert abs(iters0 - model_iterations(ref, a2, b, cuts)) <= 1
    for o in out:
        o[-1].close()


def test_emulated_quadrature_handle_api(emu, golden):
    """The quadrature handle API (bem/icqQuadrature.h:19-40) of csrc/compat_quadrature.cu on a
    sample of the reference's own known answers (tests/golden/known_answers.npz)."""
    kat = golden['known_answers']
    h = ctypes.c_void_p()
    emu.icqQuadratureInit(ctypes.byref(h))
    assert h.value
    assert emu.icqQuadratureGetMaxOrder(ctypes.byref(h)) == 8
    tris, obs = kat['corner_tris'], kat['single_obs']
    for i in range(0, tris.shape[0], 8):
        emu.icqQuadratureSetObserver(ctypes.byref(h), dp(numpy.ascontiguousarray(obs[i])))
        t = numpy.ascontiguousarray(tris[i])
        for order in range(10):
            v = emu.icqQuadratureEvaluate(ctypes.byref(h), order, dp(t[0]), dp(t[1]), dp(t[2]))
            ref = kat['single_vals'][i, order]
            assert abs(v - ref) <= 1e-13 * abs(ref)
    pairs = kat['double_pairs']
    for i in range(0, pairs.shape[0], 8):
        a, b = numpy.ascontiguousarray(pairs[i, 0]), numpy.ascontiguousarray(pairs[i, 1])
        for order in (0, 1, 5, 8, 9):
            v = emu.icqQuadratureEvaluateDouble(ctypes.byref(h), order, dp(a[0]), dp(a[1]), dp(a[2]),
                                                dp(b[0]), dp(b[1]), dp(b[2]))
            ref = kat['double_vals'][i, order]
            assert abs(v - ref) <= 1e-13 * abs(ref)
    emu.icqQuadratureDel(ctypes.byref(h))


def test_emulated_ranks_without_rows(emu, oracle_mod):
    """Four ranks, 144 triangles: the folded blocks leave ranks 2 and 3 without a single row
    (no tiles, no product kernel, null task tables) -- they still take part in every
    collective and return the same solution."""
    out = _run_ranks(emu, 4, 12, 7, 1, 1)
    c = out[0][-1]
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
    x0, iters0, _, b, _ = out[0]
    for x, iters, rinf, _, _ in out:
        assert (x == x0).all() and iters == iters0
    assert numpy.abs(b - ref.dot(x0)).max() <= 1e-12 * numpy.abs(b).max()
    for o in out:
        o[-1].close()


@pytest.mark.parametrize('kind,starts,expect', [(1, None, 'fallback'), (2, [0, 1, 2, 3], 'fallback'),
                                                (2, [0, 2], 'weak')])
def test_emulated_weak_pivot_in_a_preconditioner_block(emu, oracle_mod, kind, starts, expect):
    """A 128 x 128 diagonal block whose unpivoted inversion meets a pivot that has lost 14 digits
    ([[1, 1], [1, 1 + 1e-14]] inside tile 1 of a synthetic symmetric matrix): the block
    falls back to 1/diag and is counted (both solvers; a block of ONE tile of the partitioned
    solver behaves the same), inside a larger dense block the pivot is reported only
    (icqb_cg_last_blocks; ADVICE of round 1)."""
    c = Ctx(emu, 20, 11)
    n = c.n
    a2 = oracle_mod.area2(c.xyz, c.tri)
    rng = numpy.random.default_rng(5)
    noise = rng.standard_normal((n, n)) * 1e-3
    sym = 4. * numpy.eye(n) + 0.5 * (noise + noise.T)
    sym[128:130, :] = 0.
    sym[:, 128:130] = 0.
    sym[128:130, 128:130] = [[1., 1.], [1., 1. + 1e-14]]
    ld = (n + 15) // 16 * 16
    g = dev_array(emu, (n, ld), 0.)
    g[:, :n] = sym / a2[:, None]          # diag(a2) g is the symmetric matrix above
    c.check(emu.icqb_set_rows(c.h, 0, n))
    c.solve(g, 0, kind, ld=ld, maxit=3, starts=starts)
    bad, weak = ctypes.c_int64(-1), ctypes.c_int64(-1)
    emu.icqb_cg_last_blocks.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
    c.check(emu.icqb_cg_last_blocks(c.h, ctypes.byref(bad), ctypes.byref(weak)))
    if expect == 'fallback':
        assert bad.value == 1
    else:
        assert bad.value == 0 and weak.value >= 1
    c.close()


@pytest.mark.parametrize('nt,nph,starts,ext,storage', [
    (20, 11, [0, 1, 2, 3], ([0, 0, 2, 3], [2, 3, 4, 4]), 1),     # single tiles, one tile of overlap
    (20, 11, [0, 2], ([0, 1], [3, 4]), 1),                       # two blocks of two tiles
    (20, 11, [0, 2], ([0, 1], [3, 4]), 0),                       # the same from row-block storage
    (30, 15, [0, 2, 4, 6], ([0, 1, 3, 5], [3, 5, 7, 7]), 1),     # 840 triangles, 7 tile rows (last ragged)
    (30, 15, [0, 3, 5], ([0, 3, 4], [3, 6, 7]), 1),              # a block that is not extended (sliver-run rule)
])
def test_emulated_overlapping_blocks(emu, oracle_mod, nt, nph, starts, ext, storage):
    """Overlapping preconditioner blocks (icqb_cg_set_block_extents: symmetric additive Schwarz,
    k_cg_update_overlap): backward error against the oracle's matrix and the iteration count of
    the numpy model of the same operator."""
    c = Ctx(emu, nt, nph)
    n = c.n
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
    a2 = oracle_mod.area2(c.xyz, c.tri)
    if storage == 1:
        (g,) = c.assemble_lower()
        ld = 0
    else:
        ld = (n + 15) // 16 * 16
        g = dev_array(emu, (n, ld), 0.)
        c.check(emu.icqb_set_rows(c.h, 0, n))
        c.check(emu.icqb_assemble_dev(c.h, 5, g.ctypes.data, 0, ld))
    x, iters, rinf, b = c.solve(g, storage, 2, ld=ld, starts=starts, extents=ext)
    assert numpy.abs(b - ref.dot(x)).max() <= 1e-12 * numpy.abs(b).max()
    rows = ([128 * t for t in ext[0]], [128 * t for t in ext[1]])
    model = model_iterations(ref, a2, b, [128 * t for t in starts], extents=rows)
    assert abs(iters - model) <= 1
    # fewer iterations than the same partition without overlap
    x0, iters0, _, _ = c.solve(g, storage, 2, ld=ld, starts=starts)
    assert iters < iters0
    c.close()


def test_emulated_block_cache_between_contexts(emu, oracle_mod):
    """The contexts' device buffers come from a process-wide cache of freed blocks
    (csrc/context.cu): a second solver of the same size gets the first one's blocks back, stale
    contents included.  Two different meshes of the same size, one after the other, each checked
    against the oracle; then the same again after icqb_release_cached_memory."""
    from icqsol_b200.mesh.generators import uvSphere
    emu.icqb_release_cached_memory.argtypes = [ctypes.c_int]
    emu.icqb_release_cached_memory.restype = ctypes.c_int
    for sweep in range(2):
        for radius in (1.0, 1.7):
            c = Ctx(emu, 12, 7)
            xyz, tri = uvSphere(radius, 12, 7)
            c.xyz, c.tri = numpy.ascontiguousarray(xyz), numpy.ascontiguousarray(tri)
            c.check(emu.icqb_set_mesh(c.h, dp(c.xyz), c.xyz.shape[0], lp(c.tri), c.n))
            ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
            (g,) = c.assemble_lower()
            x, iters, rinf, b = c.solve(g, 1, 2, starts=[0])
            assert numpy.abs(b - ref.dot(x)).max() <= 1e-12 * numpy.abs(b).max()
            c.close()
        assert emu.icqb_release_cached_memory(-1) == 0


@pytest.mark.parametrize('starts,storage', [([0, 2], 1), ([0, 3], 1), ([0, 1, 3], 0), ([0], 1), ([0], 0)])
def test_emulated_block_partition(emu, oracle_mod, starts, storage):
    """An explicit partition of the tile rows into diagonal blocks -- batched
    extraction, block Gauss-Jordan with 128^3 tile products, batched apply -- against the
    numpy model of the same blocks (400 triangles = 4 tile rows, the last one ragged)."""
    c = Ctx(emu, 20, 11)
    n = c.n
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
    a2 = oracle_mod.area2(c.xyz, c.tri)
    if storage == 1:
        (g,) = c.assemble_lower()
        ld = 0
    else:
        ld = (n + 15) // 16 * 16
        g = dev_array(emu, (n, ld), 0.)
        c.check(emu.icqb_set_rows(c.h, 0, n))
        c.check(emu.icqb_assemble_dev(c.h, 5, g.ctypes.data, 0, ld))
    x, iters, rinf, b = c.solve(g, storage, 2, ld=ld, starts=starts)
    assert numpy.abs(b - ref.dot(x)).max() <= 1e-12 * numpy.abs(b).max()
    assert abs(iters - model_iterations(ref, a2, b, [128 * t for t in starts])) <= 1
    c.close()


# ---- far-field split of the off-diagonal assembly (assemble_offdiag_ff.cu) ---------------------

def _assembly_stats(emu, c):
    fast, allp = ctypes.c_int64(0), ctypes.c_int64(0)
    c.check(emu.icqb_assembly_stats(c.h, ctypes.byref(fast), ctypes.byref(allp)))
    return fast.value, allp.value


@pytest.mark.parametrize('variant', [1])
@pytest.mark.parametrize('with_k', [False, True])
def test_emulated_far_field_split_full_storage(emu, oracle_mod, variant, with_k):
    """k_offdiag_ff on 576 triangles, full storage (mirrored tiles inside the block): every G
    entry within 1e-12 of the oracle and within a few ulp of the direct kernel, K against
    the definition oracle; a third of the (warp, source) pairs take the far-field arithmetic."""
    c = Ctx(emu, 24, 13)
    n = c.n
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
    g0 = numpy.full((n, n), numpy.nan)
    c.check(emu.icqb_assemble_host(c.h, 5, dp(g0), None))
    assert _assembly_stats(emu, c) == (0, 0)
    c.check(emu.icqb_set_assembly_variant(c.h, variant))
    g, k = numpy.full((n, n), numpy.nan), numpy.full((n, n), numpy.nan)
    c.check(emu.icqb_assemble_host(c.h, 5, dp(g), dp(k) if with_k else None))
    fast, allp = _assembly_stats(emu, c)
    assert allp == 7424 and 0.3 * allp < fast < 0.5 * allp
    assert numpy.abs(g / ref - 1.).max() < 1e-12
    assert numpy.abs(g / g0 - 1.).max() < 2e-15
    if with_k:
        kref = oracle_mod.k_block(c.xyz, c.tri, 0, n, 0, n)
        assert numpy.abs(k - kref).max() < 1e-12 * numpy.abs(kref).max()
        assert (numpy.diag(k) == 0.).all()
    assert emu.icqb_set_assembly_variant(c.h, 2) != 0
    c.close()


@pytest.mark.parametrize('variant', [1])
def test_emulated_far_field_split_lower_storage(emu, oracle_mod, variant):
    """Lower storage with two separated row blocks, G, K and the transposed K entries."""
    from icqsol_b200.bem import icqRowPartition as rp
    c = Ctx(emu, 20, 11)                               # 400 triangles, tile rows 0..3
    n = c.n
    blocks = (128, 256, 384, 400)
    c.check(emu.icqb_set_assembly_variant(c.h, variant))
    g, kl, ku = c.assemble_lower(blocks, with_k=True)
    fast, allp = _assembly_stats(emu, c)
    assert allp > 0 and fast > 0
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
    kref = oracle_mod.k_block(c.xyz, c.tri, 0, n, 0, n)
    rows, krows, kurows = (rp.lowerTilesToRows(m, blocks, n) for m in (g, kl, ku))
    kmax = numpy.abs(kref).max()
    for l, i in enumerate(list(range(128, 256)) + list(range(384, 400))):   # local, global row
        ce = min(n, (i // 128 + 1) * 128)
        sel = numpy.arange(ce) != i
        assert numpy.abs(rows[l, :ce][sel] / ref[i, :ce][sel] - 1.).max() < 1e-12
        assert numpy.abs(krows[l, :ce] - kref[i, :ce]).max() < 1e-12 * kmax
        assert numpy.abs(kurows[l, :ce] - kref[:ce, i]).max() < 1e-12 * kmax
    c.close()


def test_emulated_far_field_guard_and_ragged(emu, oracle_mod):
    """A mesh far from the origin: |c_i - c_j| < max|coordinate| / 512 for every pair, so no pair
    may take the affine arithmetic (its points would differ from the reference's rounded ones
    by ~eps max|coordinate|) -- and the result is the validated kernel's.  Then ragged sizes."""
    from icqsol_b200.mesh.generators import uvSphere
    xyz, tri = uvSphere(1.0, 12, 7, center=(3000., 0., 0.))
    h = ctypes.c_void_p()
    assert emu.icqb_create(ctypes.byref(h), 0) == 0
    xyz, tri = numpy.ascontiguousarray(xyz), numpy.ascontiguousarray(tri)
    n = tri.shape[0]
    out = []
    for variant in (0, 1):
        assert emu.icqb_set_mesh(h, dp(xyz), xyz.shape[0], lp(tri), n) == 0
        assert emu.icqb_set_assembly_variant(h, variant) == 0
        g = numpy.full((n, n), numpy.nan)
        assert emu.icqb_assemble_host(h, 5, dp(g), None) == 0
        out.append(g)
    fast, allp = ctypes.c_int64(0), ctypes.c_int64(0)
    assert emu.icqb_assembly_stats(h, ctypes.byref(fast), ctypes.byref(allp)) == 0
    assert fast.value == 0 and allp.value > 0
    assert numpy.abs(out[1] / out[0] - 1.).max() < 1e-15
    # ragged: 1, 2, 33, 129 triangles (clamped lanes vote with a valid triangle's bounds)
    xyz, tri = uvSphere(1.0, 24, 13)
    xyz, tri = numpy.ascontiguousarray(xyz), numpy.ascontiguousarray(tri)
    for m in (1, 2, 33, 129):
        sub = numpy.ascontiguousarray(tri[::len(tri) // m][:m])
        assert emu.icqb_set_mesh(h, dp(xyz), xyz.shape[0], lp(sub), m) == 0
        g = numpy.full((m, m), numpy.nan)
        assert emu.icqb_assemble_host(h, 5, dp(g), None) == 0
        ref = oracle_mod.green_matrix(xyz, sub, 5)
        assert numpy.abs(g / ref - 1.).max() < 1e-12
    emu.icqb_destroy(h)


def test_emulated_row_major_assembly_between_lower_products(emu, oracle_mod):
    """The sequence behind BaseLaplaceSolver.solveGreenSystemDirect: a context that holds a lower
    storage assembles complete row-major rows in between (icqb_set_rows, icqb_assemble_dev) and
    gets its row blocks back; the lower matrix and its product plan are unaffected."""
    c = Ctx(emu, 16, 9, sms=3)                         # 256 triangles
    n = c.n
    blocks = (0, n, n, n)
    (g,) = c.assemble_lower(blocks)
    x = numpy.random.default_rng(3).normal(size=n)
    y0 = dev_array(emu, n)
    c.check(emu.icqb_symv_dev(c.h, g.ctypes.data, x.ctypes.data, y0.ctypes.data, 0))
    ld = (n + 15) // 16 * 16
    full = dev_array(emu, (n, ld))
    c.check(emu.icqb_set_rows(c.h, 0, n))
    c.check(emu.icqb_assemble_dev(c.h, 5, full.ctypes.data, 0, ld))
    c.check(emu.icqb_set_row_blocks(c.h, *blocks))
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
    assert numpy.abs(full[:, :n] / ref - 1.).max() < 1e-12
    y1 = dev_array(emu, n)
    c.check(emu.icqb_symv_dev(c.h, g.ctypes.data, x.ctypes.data, y1.ctypes.data, 0))
    assert (y1 == y0).all()
    assert numpy.abs(y1 - ref.dot(x)).max() <= 1e-13 * numpy.abs(ref).dot(numpy.abs(x)).max()
    # the LU the direct solve hands to cuSOLVER, here numpy's: same response as the reference's
    b = c.rhs()
    assert numpy.abs(numpy.linalg.solve(full[:, :n], b) - numpy.linalg.solve(ref, b)).max() < \
        1e-9 * numpy.abs(numpy.linalg.solve(ref, b)).max()
    c.close()


def test_emulated_guard_pages_catch_an_overrun():
    """The memcheck of the emulation is real: a product whose matrix buffer is one row too short
    dies with SIGSEGV at the guard page (in a child process), the correct call does not."""
    import subprocess
    script = r'''
import ctypes, os, sys, numpy
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tools', 'emu'))
os.environ['ICQB_EMU_GUARD'] = '1'
import build as emu_build
from icqsol_b200 import _lib
lib = ctypes.CDLL(emu_build.build())
for name, restype, argtypes in _lib.SIGNATURES:
    fn = getattr(lib, name); fn.restype = restype; fn.argtypes = argtypes
lib.emu_alloc.restype = ctypes.c_void_p; lib.emu_alloc.argtypes = [ctypes.c_size_t]
h = ctypes.c_void_p(); assert lib.icqb_create(ctypes.byref(h), 0) == 0
rows, cols, short = 8, 512, int(sys.argv[1])
a = lib.emu_alloc(8 * (rows - short) * cols)
ctypes.memset(a, 0, 8 * (rows - short) * cols)
x, y = numpy.ones(cols), numpy.zeros(rows)
dp = lambda v: v.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
assert lib.icqb_gemv_host(h, a, rows, cols, cols, dp(x), dp(y)) == 0
print('ok')
'''.format(root=ROOT)
    good = subprocess.run([sys.executable, '-c', script, '0'], capture_output=True, text=True, timeout=300)
    assert good.returncode == 0 and 'ok' in good.stdout, good.stderr[-2000:]
    bad = subprocess.run([sys.executable, '-c', script, '1'], capture_output=True, text=True, timeout=300)
    assert bad.returncode == -11, (bad.returncode, bad.stdout, bad.stderr[-500:])


@pytest.mark.parametrize('order', ['shuffle'])     # 'reverse' as well: ICQB_EMU_ORDER=reverse pytest ...
def test_emulated_kernels_under_other_thread_orders(order):
    """The threads of a block run in ascending order between two synchronisation points by
    default; a missing barrier can hide behind that order.  The assembly, product and CG tests
    again in a child process with the order reversed / freshly permuted every round."""
    import subprocess
    if os.environ.get('ICQB_EMU_ORDER'):
        pytest.skip('already running under ICQB_EMU_ORDER')
    env = dict(os.environ, ICQB_EMU_ORDER=order)
    out = subprocess.run(
        [sys.executable, '-m', 'pytest', os.path.abspath(__file__), '-x', '-q', '-k',
         'assembly_full_storage or lower_storage_and_product or cg_matches_model or cap_blocks or '
         'far_field_split_lower or block_partition or subdivide or gemv'],
        capture_output=True, text=True, env=env, timeout=1200, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]


@pytest.mark.parametrize('nranks,storage,caps,starts,peer,cuts', [
    (2, 1, (1, 2), None, True, [0, 128, 256]),        # caps + peer exchange
    (4, 1, (1, 1), None, True, [0, 128, 256, 384]),   # four ranks, two of them with one strip
    (2, 1, None, [0, 2], False, [0, 256]),            # explicit partition, all-reduced dense blocks
    (3, 0, None, [0, 1, 3], False, [0, 128, 384]),    # row blocks, three ranks
    (3, 1, None, [0, 3], True, [0, 384])])            # partition + peer exchange
def test_emulated_multi_rank_experimental_combinations(emu, oracle_mod, nranks, storage, caps, starts,
                                                       peer, cuts):
    """Solver of kind 2, the combinations tests/dist_worker.py runs on the GPUs (also with
    ICQB_PEER_EXCHANGE): dense blocks are extracted from the rows each
    rank stores and all-reduced, every rank applies them; same solution and iteration count on
    every rank, the count of the numpy model of the same blocks."""
    out = _run_ranks(emu, nranks, 20, 11, storage, 2, caps, peer, starts)     # 400 triangles
    c = out[0][-1]
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
    a2 = oracle_mod.area2(c.xyz, c.tri)
    x0, iters0, _, b, _ = out[0]
    for x, iters, rinf, _, _ in out:
        assert (x == x0).all() and iters == iters0
        assert rinf <= 5e-13 * numpy.abs(b).max()
    assert numpy.abs(b - ref.dot(x0)).max() <= 1e-12 * numpy.abs(b).max()
    assert abs(iters0 - model_iterations(ref, a2, b, cuts)) <= 1
    for o in out:
        o[-1].close()


# ---- optional: float32 late products of the CG (k_symv_f32, cg_caps.cu) -------------------------

@pytest.mark.parametrize('sms', [1, 3, 4])
def test_emulated_float32_product(emu, oracle_mod, sms):
    """k_symv_f32 on the float32 copy of the stored tiles: equals the product with the
    float32-rounded entries (sums in double) for several static partitions of the chunk
    sequence, two separated row blocks and a ragged last tile."""
    c = Ctx(emu, 20, 11, sms=sms)                      # 400 triangles, tile rows 0..3
    n = c.n
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
    a2 = oracle_mod.area2(c.xyz, c.tri)
    rng = numpy.random.default_rng(sms)
    for blocks in ((0, n, n, n), (128, 256, 384, 400)):
        (g,) = c.assemble_lower(blocks)
        rows = list(range(blocks[0], blocks[1])) + list(range(blocks[2], blocks[3]))
        mask = numpy.zeros((n, n), bool)
        for i in rows:
            mask[i, :min(n, (i // 128 + 1) * 128)] = True
        tile = numpy.arange(n) // 128
        lower = mask & (tile[None, :] < tile[:, None])
        r32 = ref.astype(numpy.float32).astype(numpy.float64)
        full = numpy.where(mask, r32, 0.) + numpy.where(lower, r32, 0.).T * (a2[None, :] / a2[:, None])
        for scaled in (0, 1):
            x, y = rng.normal(size=n), dev_array(emu, n)
            c.check(emu.icqb_symv_f32_dev(c.h, g.ctypes.data, x.ctypes.data, y.ctypes.data, scaled))
            want = full.dot(x) * (a2 if scaled else 1.)
            bound = numpy.abs(full).dot(numpy.abs(x)) * (a2 if scaled else 1.)
            assert (numpy.abs(y - want) <= 1e-13 * bound + 1e-300).all()
            # and it IS a float32 product: away from the double one by ~1e-8
            exact = (numpy.where(mask, ref, 0.) + numpy.where(lower, ref, 0.).T *
                     (a2[None, :] / a2[:, None])).dot(x) * (a2 if scaled else 1.)
            assert 1e-10 * bound.max() < numpy.abs(y - exact).max() < 1e-6 * bound.max()
    c.close()


def model_iterations_mixed(g, a2, b, cuts, switch):
    """numpy model of the mixed-precision CG: as model_iterations, the product reads the
    float32-rounded matrix from the first residual below switch * max|b| on.
    @return (iterations, float32 products)"""
    n = len(b)
    ms = a2[:, None] * g
    g32 = g.astype(numpy.float32).astype(numpy.float64)
    cuts = [c for c in cuts if c < n] + [n]
    msym = 0.5 * (ms + ms.T)
    invs = [numpy.linalg.inv(msym[p:q, p:q]) for p, q in zip(cuts[:-1], cuts[1:])]

    def pre(r):
        out = numpy.empty_like(r)
        for k, (p, q) in enumerate(zip(cuts[:-1], cuts[1:])):
            out[p:q] = invs[k].dot(r[p:q])
        return out
    x, r, p, rho_old = numpy.zeros(n), a2 * b, numpy.zeros(n), 1.
    w = pre(r)
    lowp, n32 = False, 0
    for k in range(2000):
        rho = r.dot(w)
        p = w + (rho / rho_old if k else 0.) * p
        z = a2 * ((g32 if lowp else g).dot(p))
        n32 += int(lowp)
        alpha = rho / p.dot(z)
        x += alpha * p
        r -= alpha * z
        w = pre(r)
        rho_old = rho
        res = numpy.abs(r / a2).max()
        if res <= switch * numpy.abs(b).max():
            lowp = True
        if res <= 5e-13 * numpy.abs(b).max():
            return k + 1, n32
    return -1, n32


@pytest.mark.parametrize('starts,cuts', [([0, 1, 2, 3], [0, 128, 256, 384]), ([0, 2], [0, 256])])
def test_emulated_mixed_precision_cg(emu, oracle_mod, starts, cuts):
    """The CG of cg_caps.cu with float32 late products (switch at 1e-6): same iteration and
    float32-product counts as the numpy model, the solution satisfies the oracle's system to
    1e-12, and switching the option off restores the double-precision solve."""
    c = Ctx(emu, 20, 11)
    n = c.n
    (g,) = c.assemble_lower()
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
    a2 = oracle_mod.area2(c.xyz, c.tri)
    assert emu.icqb_cg_set_mixed_precision(c.h, 1.5) != 0
    c.check(emu.icqb_cg_set_mixed_precision(c.h, 1e-6))
    x, iters, rinf, b = c.solve(g, 1, 2, starts=starts)
    n32 = ctypes.c_int64(-1)
    c.check(emu.icqb_cg_last_mixed(c.h, ctypes.byref(n32)))
    want_iters, want32 = model_iterations_mixed(ref, a2, numpy.array(b), cuts, 1e-6)
    assert numpy.abs(b - ref.dot(x)).max() <= 1e-12 * numpy.abs(b).max()
    assert rinf <= 1e-12 * numpy.abs(b).max()
    assert abs(iters - want_iters) <= 2 and abs(n32.value - want32) <= 2 and n32.value > 5
    c.check(emu.icqb_cg_set_mixed_precision(c.h, 0.))
    x2, iters2, rinf2, _ = c.solve(g, 1, 2, starts=starts)
    c.check(emu.icqb_cg_last_mixed(c.h, ctypes.byref(n32)))
    assert n32.value == 0 and abs(iters2 - model_iterations(ref, a2, numpy.array(b), cuts)) <= 1
    assert numpy.abs(x2 - x).max() <= 1e-9 * numpy.abs(x).max()
    c.close()


@pytest.mark.parametrize('nranks,peer', [(2, False), (3, True)])
def test_emulated_mixed_precision_multi_rank(emu, oracle_mod, nranks, peer):
    """Float32 late products with the partial products all-reduced (NCCL stand-in) or exchanged
    through the peer buffers: every rank takes the same switch decision (replicated state), the
    same solution and counts on every rank, the counts of the numpy model."""
    out = _run_ranks(emu, nranks, 20, 11, 1, 2, None, peer, [0, 2], mixed=1e-6)
    c = out[0][-1]
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
    a2 = oracle_mod.area2(c.xyz, c.tri)
    x0, iters0, _, b, _ = out[0]
    counts = []
    for x, iters, rinf, _, cc in out:
        assert (x == x0).all() and iters == iters0
        assert rinf <= 1e-12 * numpy.abs(b).max()
        n32 = ctypes.c_int64(-1)
        cc.check(emu.icqb_cg_last_mixed(cc.h, ctypes.byref(n32)))
        counts.append(n32.value)
    assert len(set(counts)) == 1 and counts[0] > 5
    assert numpy.abs(b - ref.dot(x0)).max() <= 1e-12 * numpy.abs(b).max()
    want_iters, want32 = model_iterations_mixed(ref, a2, numpy.array(b), [0, 256], 1e-6)
    assert abs(iters0 - want_iters) <= 2 and abs(counts[0] - want32) <= 2
    for o in out:
        o[-1].close()


def test_emulated_far_field_split_fuzz(emu):
    """Random triangle soups -- sizes over three decades, a fifth of the triangles slivers,
    random scale, the whole mesh displaced by up to a few hundred times its size -- through
    the far-field variants against the validated kernel: G within the bound of the guard
    (2e-13; it is a few 1e-14 when the displacement is large, 1e-15 otherwise), K to 1e-11."""
    rng = numpy.random.default_rng(7)
    h = ctypes.c_void_p()
    assert emu.icqb_create(ctypes.byref(h), 0) == 0
    for trial in range(4):
        n = int(rng.integers(40, 160))
        scale = 10.0 ** rng.uniform(-3, 3)
        offset = rng.normal(size=3) * scale * 10.0 ** rng.uniform(-1, 2.5)
        cen = rng.uniform(-1, 1, size=(n, 3))
        size = 10.0 ** rng.uniform(-3, -0.3, size=n)
        e1, e2 = rng.normal(size=(n, 3)), rng.normal(size=(n, 3))
        sliver = rng.uniform(size=n) < 0.2
        e2 = numpy.where(sliver[:, None], e1 * rng.uniform(0.5, 1.5, size=(n, 1)) + 0.02 * e2, e2)
        xyz = numpy.vstack([cen, cen + size[:, None] * e1, cen + size[:, None] * e2]) * scale + offset
        if trial % 2:
            xyz = xyz.astype(numpy.float32).astype(numpy.float64)
        xyz = numpy.ascontiguousarray(xyz)
        tri = numpy.ascontiguousarray(
            numpy.stack([numpy.arange(n), n + numpy.arange(n), 2 * n + numpy.arange(n)], 1).astype(numpy.int64))
        out = {}
        for variant in (0, 1):
            assert emu.icqb_set_mesh(h, dp(xyz), xyz.shape[0], lp(tri), n) == 0
            assert emu.icqb_set_assembly_variant(h, variant) == 0
            g, k = numpy.full((n, n), numpy.nan), numpy.full((n, n), numpy.nan)
            assert emu.icqb_assemble_host(h, 5, dp(g), dp(k)) == 0
            out[variant] = (g, k)
        off = ~numpy.eye(n, dtype=bool)
        g0, k0 = out[0]
        for variant in (1,):
            g, k = out[variant]
            assert numpy.abs(g[off] / g0[off] - 1.).max() < 2e-13
            assert numpy.abs(k - k0)[off].max() < 1e-11 * numpy.abs(k0[off]).max()
    emu.icqb_destroy(h)


# ---- LU with partial pivoting (csrc/lu.cu) ------------------------------------------------------

def _lu_local(n, nranks, rank):
    """global column indices of the block-cyclic local columns of `rank` (blocks of 128)"""
    nblk = (n + 127) // 128
    cols = []
    for kb in range(rank, nblk, nranks):
        cols += list(range(kb * 128, min(n, (kb + 1) * 128)))
    return numpy.asarray(cols, int)


@pytest.mark.parametrize('n', [1, 5, 128, 129, 300])
def test_emulated_lu_single_rank(emu, n):
    """icqb_lu_factor_dev / icqb_lu_solve_dev against numpy.linalg.solve (LAPACK dgesv, the
    reference's own solve) on matrices that need row exchanges: same pivot sequence as dgetrf
    (checked through the permutation parity-free way: P A = L U reproduces A), solution to
    ~cond * eps, with and without the column scaling, panels of 1..3 blocks and ragged sizes."""
    rng = numpy.random.default_rng(n)
    a = rng.standard_normal((n, n))
    a[numpy.arange(n), numpy.arange(n)] *= 0.01           # pivoting is needed
    b = rng.standard_normal(n)
    ld = (n + 15) // 16 * 16
    h = ctypes.c_void_p()
    assert emu.icqb_create(ctypes.byref(h), 0) == 0
    assert emu.icqb_lu_local_columns(h, n) == n
    for scale in (None, numpy.abs(rng.standard_normal(n)) + 0.1):
        w = dev_array(emu, (n, ld), 0.)
        w[:, :n] = a.T                                     # column c of A contiguous
        bb = dev_array(emu, n)
        bb[:] = b
        x = dev_array(emu, n)
        sc = None
        if scale is not None:
            sc = dev_array(emu, n)
            sc[:] = scale
        info = ctypes.c_int(-1)
        st = emu.icqb_lu_factor_dev(h, w.ctypes.data, n, ld, sc.ctypes.data if sc is not None else 0,
                                    ctypes.byref(info))
        assert st == 0, emu.icqb_last_error(h)
        assert info.value == 0
        st = emu.icqb_lu_solve_dev(h, w.ctypes.data, n, ld, sc.ctypes.data if sc is not None else 0,
                                   bb.ctypes.data, x.ctypes.data)
        assert st == 0, emu.icqb_last_error(h)
        m = a if scale is None else a * scale[None, :]     # columns scaled
        rhs = b if scale is None else scale * b
        ref = numpy.linalg.solve(m, rhs)
        assert numpy.abs(x - ref).max() <= 1e-9 * max(numpy.abs(ref).max(), 1e-300) * max(numpy.linalg.cond(m) / 1e4, 1.)
        assert numpy.abs(m.dot(x) - rhs).max() <= 1e-11 * (numpy.abs(m).sum(1).max() * numpy.abs(x).max() + 1e-300)
        # the factors: unit lower below the diagonal, upper on and above, |L| <= 1 (partial pivoting)
        f = w[:, :n].T
        assert numpy.abs(numpy.tril(f, -1)).max(initial=0.) <= 1. + 1e-15
    emu.icqb_destroy(h)


def test_emulated_lu_singular_matrix(emu):
    """An exactly singular matrix: info reports the first zero pivot (dgetrf's info > 0)."""
    n = 6
    a = numpy.arange(36.).reshape(6, 6)        # rank 2
    a[:, 3] = 0.
    ld = 16
    h = ctypes.c_void_p()
    assert emu.icqb_create(ctypes.byref(h), 0) == 0
    w = dev_array(emu, (n, ld), 0.)
    w[:, :n] = a.T
    info = ctypes.c_int(0)
    assert emu.icqb_lu_factor_dev(h, w.ctypes.data, n, ld, 0, ctypes.byref(info)) == 0
    assert info.value > 0
    emu.icqb_destroy(h)


@pytest.mark.parametrize('nranks', [2, 3])
def test_emulated_lu_multi_rank(emu, oracle_mod, nranks):
    """Column blocks dealt to the ranks (threads over the in-process NCCL stand-in): panels are
    factored by their owner and broadcast, every rank updates its own columns; the solve sweeps
    pass the vector from owner to owner.  The system is the reference's own: G of a 400-triangle
    sphere (rows assembled per 128-row block into the local columns, scaled by the areas: the
    columns of the symmetric diag(A) G), against numpy.linalg.solve on the oracle's G."""
    import threading
    uid = ctypes.create_string_buffer(128)
    assert emu.icqb_comm_unique_id(uid) == 0
    out, errors = [None] * nranks, []

    def work(rank):
        try:
            c = Ctx(emu, 20, 11, sms=2)
            c.check(emu.icqb_comm_init(c.h, uid, rank, nranks))
            n = c.n
            ld = (n + 15) // 16 * 16
            cols = _lu_local(n, nranks, rank)
            assert emu.icqb_lu_local_columns(c.h, n) == len(cols)
            w = dev_array(emu, (max(len(cols), 1), ld), 0.)
            for lb, kb in enumerate(range(rank, (n + 127) // 128, nranks)):
                r0, r1 = kb * 128, min(n, (kb + 1) * 128)
                c.check(emu.icqb_set_rows(c.h, r0, r1))
                c.check(emu.icqb_assemble_dev(c.h, 5, w[lb * 128:].ctypes.data, 0, ld))
            b = dev_array(emu, n)
            b[:] = c.rhs()
            x = dev_array(emu, n)
            info = ctypes.c_int(-1)
            a2 = emu.icqb_area2_dev(c.h)
            c.check(emu.icqb_lu_factor_dev(c.h, w.ctypes.data, n, ld, a2, ctypes.byref(info)))
            assert info.value == 0
            c.check(emu.icqb_lu_solve_dev(c.h, w.ctypes.data, n, ld, a2, b.ctypes.data, x.ctypes.data))
            out[rank] = (x.copy(), b.copy(), c)
        except Exception as e:    # noqa: BLE001 -- reported by the caller
            errors.append((rank, repr(e)))
    threads = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    assert not errors, errors
    c = out[0][2]
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5)
    b = out[0][1]
    lu = numpy.linalg.solve(ref, b)
    for x, _, _ in out:
        assert (x == out[0][0]).all()
        assert numpy.abs(b - ref.dot(x)).max() <= 1e-12 * numpy.abs(b).max()
        assert numpy.abs(x - lu).max() <= 1e-9 * numpy.abs(lu).max()
    for o in out:
        o[2].close()


@pytest.mark.parametrize('nranks,peer', [(2, False), (3, False), (2, True)])
def test_emulated_sharded_preconditioner_blocks(emu, oracle_mod, nranks, peer):
    """Diagonal blocks cut at the ranks' row-block boundaries (icqRowPartition.rankCuts): every
    block lives on one rank, which extracts, inverts and applies it alone; the pieces of w are
    all-gathered.  Same solution on every rank, the oracle's system solved, the iteration count
    of the numpy model of the same blocks."""
    from icqsol_b200.bem import icqRowPartition as rp
    nt, nph = 28, 15                                     # 784 triangles, 7 tile rows
    n = 2 * nt * (nph - 1)
    nT = (n + 127) // 128
    cuts = rp.rankCuts(n, nranks)
    starts, _ = rp.sliverBlockPartition(None, numpy.zeros((n, 3), int), blockTiles=2,
                                        flagged=numpy.zeros(nT, bool), cuts=cuts)
    # (every multi-tile block inside one rank's rows)
    rows = rp.foldedTileRows(n, nranks)
    for a, b in zip(starts, starts[1:] + [nT]):
        assert len({g for g, (a0, a1, b0, b1) in enumerate(rows) for t in range(a, b)
                    if a0 <= t < a1 or b0 <= t < b1}) == 1
    assert any(b - a > 1 for a, b in zip(starts, starts[1:] + [nT]))
    out = _run_ranks(emu, nranks, nt, nph, 1, 2, None, peer, starts)
    c = out[0][-1]
    ref = oracle_mod.green_matrix(c.xyz, c.tri, 5, threads=True)
    a2 = oracle_mod.area2(c.xyz, c.tri)
    x0, iters0, _, b, _ = out[0]
    for x, iters, rinf, _, _ in out:
        assert (x == x0).all() and iters == iters0
        assert rinf <= 5e-13 * numpy.abs(b).max()
    assert numpy.abs(b - ref.dot(x0)).max() <= 1e-12 * numpy.abs(b).max()
    assert abs(iters0 - model_iterations(ref, a2, b, [128 * s for s in starts])) <= 1
    for o in out:
        o[-1].close()
