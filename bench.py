#!/usr/bin/env python
"""Benchmark of the boundary-element hot path (BASELINE.json: "G+K entries/sec
(fp64) and CG GEMV GB/s at 1/2/4/8 B200 vs host-CPU reference").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path

One step = one pass of the hot path over one mesh: fused G+K assembly (diagonal +
off-diagonal kernels) into resident HBM buffers, then the dense solve of G sigma = -v to a
backward error max|v + G sigma| <= 5e-13 max|v| (inside the 1e-12 parity criterion).

Workloads:
  N = 1 (default)  config C2: UV sphere, 19,880 triangles.
  N > 1 (default)  config C3's own geometry: the reference's test_results/testBoltFine.vtk with
                   every triangle cut into 36 (84,600 triangles), the SAME problem on 2, 4 and
                   8 GPUs -> "strong" scaling.  It also fits ONE B200 (--gpus 1 --workload bolt
                   is the base of that curve; profiles/ holds the 1/2/4/8 runs).
  --workload c3|c5 UV spheres of 100,488 / 199,808 triangles (C3 size, config C5);
  --workload c4    config C4: 49,400-triangle hollow sphere, G+K assembly + the Poisson product
                   v = G q (bem/icqPoissonSolver.py:36) instead of the solve;
  --workload weak  C2 refined so that N^2 / GPU stays fixed (round 1's curve).

value   = G+K entries per second over the whole step = 2*N^2 / ms_per_step
          (N^2 entries per matrix are DEFINED by a step in either storage: in 'lower'
          storage the upper half is the mirror rule applied on read)
e2e     = same metric through the public Python API (LaplaceSolver / PoissonSolver on a host
          polydata -> host response), host<->device copies inside the timed region
roofline= dominant kernel (fused G+K off-diagonal assembly) against the FP64 FMA rate measured
          on this GPU in the same run; roofline_gemv = the CG product (+ its exchange between
          the ranks) against MEASURED_PEAKS.json's HBM copy bandwidth, heaviest rank
phases  = per-phase times: mean over the steps on rank 0, and max / min over the ranks
g_only  = G entries per second of a G-only pass (what the reference computes: it has no K)
cpu_baseline = the reference's own C++ for G (oracle/_ref/libicqref.so, unmodified sources)
          and the C restatement of K (the reference has no K), on a bounded sub-mesh sample,
          1 thread (the reference ships single-threaded).
--impl reference = the same step on the host cores: G by the reference's C++, K by the C
          restatement, self terms, and the reference's LU (numpy.linalg.solve), on a bounded
          sample; `value` is extrapolated to the config's N by the stated rule (assembly by
          pair count, LU by N^3).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'G+K entries/sec (fp64)'
UNIT = 'entries/s'
EXPR_V = '1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)'   # CMakeLists.txt:487 of the reference
FLOP_PER_EVAL_G = 12.0
FLOP_PER_EVAL_GK = 18.0    # + 1/r^3 (2 mul) and two weighted accumulations (2 fma), DESIGN.md
# diagonal blocks of the CG preconditioner, in 128-triangle tiles: (one rank, several ranks).  One
# rank: blocks of 4 tiles extended by one tile on either side (overlapping blocks are a one-rank
# feature, include/icqb.h icqb_cg_set_block_extents); several ranks: plain blocks of 8.
DEFAULT_BLOCK_TILES = (4, 8)
DEFAULT_BLOCK_OVERLAP = (1, 0)


def resolve_blocks(args, world):
    """(block tiles, overlap tiles) of this run."""
    many = 0 if world == 1 else 1
    overlap = DEFAULT_BLOCK_OVERLAP[many] if args.block_overlap is None else args.block_overlap
    if many:
        overlap = 0
    tiles = args.block_tiles
    if tiles is None:
        tiles = DEFAULT_BLOCK_TILES[many] if overlap > 0 else DEFAULT_BLOCK_TILES[1]
    return int(tiles), int(overlap)
REFERENCE_WORKERS = 16     # host processes of the reference arm (fewer if the box has fewer cores)


def resolve_workload(ngpus, name):
    if name != 'auto':
        return name
    return 'c2' if ngpus == 1 else 'bolt'


def make_mesh(ngpus, name):
    """(label, xyz, tri) of the workload."""
    name = resolve_workload(ngpus, name)
    if name == 'bolt':
        # config C3's geometry: the reference's test_results/testBoltFine.vtk (2,350 triangles,
        # kept as a golden fixture in tests/golden/meshes.npz) with every triangle cut into 36
        from icqsol_b200.mesh.generators import subdivide
        m = numpy.load(os.path.join(ROOT, 'tests', 'golden', 'meshes.npz'))
        xyz, tri = subdivide(m['boltFine_xyz'], m['boltFine_tri'], 6)
        return 'testBoltFine_k6_N{0} (config C3: refined bolt)'.format(tri.shape[0]), xyz, tri
    if name == 'c4':
        from icqsol_b200.mesh.generators import hollowSphere
        xyz, tri = hollowSphere(2.0, 1.0, 100, 50)
        return 'hollow_sphere_100x50+200x100_N{0} (config C4, Poisson product)'.format(tri.shape[0]), xyz, tri
    from icqsol_b200.mesh.generators import uvSphere
    label, nt, nph = sphere_workload(ngpus, name)
    xyz, tri = uvSphere(1.0, nt, nph)
    return label, xyz, tri


def sphere_workload(ngpus, name):
    """(label, n_theta, n_phi) of the UV sphere workloads."""
    if name == 'c2':
        return 'uv_sphere_142x71_N19880 (config C2)', 142, 71
    if name == 'c5':
        return 'uv_sphere_448x224_N199808 (config C5)', 448, 224
    if name == 'c3':
        return 'uv_sphere_318x159_N100488 (size of config C3)', 318, 159
    if name == 'small':
        return 'uv_sphere_48x24_N2208 (smoke size)', 48, 24
    if name == 'tiny':
        return 'uv_sphere_16x11_N320 (emulation size: tools/emu/run_emulated.py)', 16, 11
    # weak scaling from C2: N^2/P fixed  ->  n_theta ~ 142 * P^(1/4)
    nt = int(round(142 * ngpus ** 0.25 / 2.0)) * 2
    return 'uv_sphere_{0}x{1}_N{2} (C2 scaled to {3} GPUs, N^2/GPU fixed)'.format(
        nt, nt // 2, 2 * nt * (nt // 2 - 1), ngpus), nt, nt // 2


def scaling_label(ngpus, name):
    """'strong': the problem is fixed whatever the GPU count (every workload but 'weak')."""
    return 'weak' if resolve_workload(ngpus, name) == 'weak' else 'strong'


def bench_config(args, label, n, world):
    """`config` of the JSON line -- the SAME dictionary for the CUDA arm and the reference arm
    (nothing in it depends on a GPU being present)."""
    poisson = resolve_workload(world, args.workload) == 'c4'
    lower = args.storage == 'lower'
    return {
        'workload': label, 'triangles': int(n), 'diag_order': 5, 'offdiag_order': 8,
        'step': ('fused G+K assembly + Poisson product v = G q' if poisson else
                 'fused G+K assembly + dense solve of G sigma = -v (max|v+G sigma| <= 5e-13 max|v|)'),
        'storage': args.storage, 'preconditioner': args.preconditioner,
        'block_tiles': resolve_blocks(args, world)[0], 'block_overlap_tiles': resolve_blocks(args, world)[1],
        'assembly_variant': args.assembly_variant,
        'mixed_switch': args.mixed_switch,
        'parallelism': ('folded row blocks balanced by stored tiles, lower storage x{0}' if lower
                        else 'row-block x{0}').format(world),
        'l2': 'inputs larger than L2 ({0:.2f} GB of G streamed per product per GPU)'.format(
            (4.0 if lower else 8.0) * n * n / world / 1e9)}


def centroid_potential(xyz, tri):
    cen = (xyz[tri[:, 0]] + xyz[tri[:, 1]] + xyz[tri[:, 2]]) / 3.
    return 1. / numpy.sqrt(cen[:, 0] ** 2 + (cen[:, 1] - 2.) ** 2 + (cen[:, 2] - 4.) ** 2)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region: NVML polled every 20 ms
    (nvidia_ml_py), or -- when NVML cannot be opened -- nvidia-smi in a loop (the recipe's
    clocks line, ~10 samples per second)."""
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
    # nvmlClocksThrottleReason* bits
    REASON_BITS = (('sw_power_cap', 0x4), ('hw_slowdown', 0x8), ('sw_thermal_slowdown', 0x20),
                   ('hw_thermal_slowdown', 0x40))

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index = index
        self.sm = []
        self.max_mhz = None
        self.reasons = set()
        self.source = None
        self.stop_flag = False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(int(index))
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
            self.source = 'nvml'
        except Exception:
            self.nvml = None
            self.source = 'nvidia-smi'

    def _sample_nvml(self):
        nv = self.nvml
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        except Exception:
            mask = 0
        for name, bit in self.REASON_BITS:
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                              '--format=csv,noheader,nounits'], capture_output=True,
                             text=True, timeout=5).stdout.strip()
        if not out:
            return
        f = [t.strip() for t in out.split(',')]
        if f[0].replace('.', '').isdigit():
            self.sm.append(float(f[0]))
        if f[1].replace('.', '').isdigit():
            self.max_mhz = max(self.max_mhz or 0., float(f[1]))
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for k, nm in enumerate(names):
            if len(f) > 3 + k and f[3 + k].lower().startswith('active'):
                self.reasons.add(nm)

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                    time.sleep(0.02)
                else:
                    self._sample_smi()
                    time.sleep(0.05)
            except Exception:
                if self.nvml is not None:   # NVML went wrong mid-run: fall back
                    self.nvml = None
                    self.source = 'nvidia-smi'
                else:
                    time.sleep(0.05)

    def summary(self):
        return {'sm_mhz': float(numpy.median(self.sm)) if self.sm else None,
                'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(self.sm), 'source': self.source}


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured (MEASURED_PEAKS.json)'
    return {'hbm_gbs': 6650.0}, 'fallback (B200_PROFILING.md)'


def traffic_of(kernel, storage, n):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of `kernel` from
    the committed ncu --set full capture of this workload (profiles/traffic.json), or None."""
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    if not os.path.exists(path):
        return None
    with open(path) as f:
        table = json.load(f)
    return table.get('{0}|{1}|N{2}'.format(kernel, storage, n))


# ---------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ---------------------------------------------------------------------------------------

def _ref_worker(args):
    """G (reference C++ when built, else the C restatement), the self terms and K (C restatement:
    the reference has none) of one sub-mesh; returns the seconds spent on (G, K)."""
    xyz, tri, order, use_ref, with_k = args
    from oracle import oracle
    n = tri.shape[0]
    t0 = time.perf_counter()
    if use_ref:
        g = oracle.ref_offdiag(xyz, tri)
    else:
        g = oracle.offdiag(xyz, tri)
    g[numpy.diag_indices_from(g)] = oracle.diag(xyz, tri, order)
    t1 = time.perf_counter()
    if with_k:
        oracle.k_block(xyz, tri, 0, n, 0, n, threads=False)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def cpu_assembly_sample(xyz, tri, ns, workers, order=5, with_k=True):
    """Each worker assembles G (and K) for its own ns-triangle sub-mesh of the workload.
    Returns (seconds wall, seconds wall of the G part alone (estimate), workers*ns*ns, kind)."""
    from oracle import oracle
    oracle.build()
    use_ref = oracle.have_reference()
    n = tri.shape[0]
    ns = min(int(ns), n)
    jobs = []
    for w in range(workers):
        start = (w * ns) % max(1, n - ns)
        jobs.append((xyz, numpy.ascontiguousarray(tri[start:start + ns]), order, use_ref, with_k))
    t0 = time.perf_counter()
    if workers == 1:
        parts = [_ref_worker(jobs[0])]
    else:
        import multiprocessing
        with multiprocessing.get_context('fork').Pool(workers) as pool:
            parts = pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t0
    tg = max(p[0] for p in parts)
    tk = max(p[1] for p in parts)
    wall_g = wall * tg / max(tg + tk, 1e-30)
    return wall, wall_g, workers * ns * ns, 'reference' if use_ref else 'port'


def run_reference(args):
    """The step of the CUDA arm on the host cores, the way the reference does it: G by its own
    C++ (bem/icqLaplaceMatrices.cpp:37-105, one process per core on disjoint sub-meshes: the
    shipped code is single-threaded), self terms, K by the C restatement of its definition
    (the reference has none), and the reference's solve, numpy.linalg.solve
    (bem/icqLaplaceSolver.py:36).  Every step works on a bounded sample; `value` extrapolates
    the sample to the config's N: assembly time by entry count (cost per pair is
    size-independent), LU time by N^3."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    world = int(os.environ.get('WORLD_SIZE', str(args.gpus)))
    label, xyz, tri = make_mesh(world, args.workload)
    n = tri.shape[0]
    poisson = resolve_workload(world, args.workload) == 'c4'
    workers = max(1, min(REFERENCE_WORKERS, os.cpu_count() or 1))
    ns = min(args.ref_sample, n)
    nlu = min(args.ref_lu_size, n)
    from oracle import oracle
    oracle.build()
    rng = numpy.random.default_rng(0)
    lu_mat = None
    rec = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        wall, wall_g, entries, kind = cpu_assembly_sample(xyz, tri, ns, workers)
        t1 = time.perf_counter()
        if poisson:
            # the reference's apply: numpy.dot(gMat, src) (bem/icqPoissonSolver.py:36)
            if lu_mat is None:
                lu_mat = rng.standard_normal((nlu, nlu))
            tl0 = time.perf_counter()
            lu_mat.dot(numpy.ones(nlu))
            t_solve = time.perf_counter() - tl0
        else:
            # the reference's solve on a matrix of the sample size with G's structure
            # (LAPACK dgesv: the time does not depend on the entries)
            if lu_mat is None:
                lu_mat = oracle.green_matrix(xyz, numpy.ascontiguousarray(tri[:nlu]), 5, threads=True)
            v = centroid_potential(xyz, tri[:nlu])
            tl0 = time.perf_counter()
            oracle.laplace_response(lu_mat, v)
            t_solve = time.perf_counter() - tl0
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            rec.append((dt, wall, wall_g, t_solve, entries))
    t_step = float(numpy.mean([r[0] for r in rec]))
    t_asm = float(numpy.mean([r[1] for r in rec]))
    t_asm_g = float(numpy.mean([r[2] for r in rec]))
    t_solve = float(numpy.mean([r[3] for r in rec]))
    entries = rec[0][4]                      # G entries of the sample (workers * ns^2)
    # extrapolation to the config's N
    asm_rate_gk = 2.0 * entries / t_asm      # G+K entries/s of the host, size-independent
    asm_rate_g = entries / t_asm_g
    t_asm_full = 2.0 * n * n / asm_rate_gk
    t_solve_full = t_solve * ((n / float(nlu)) ** (2 if poisson else 3))
    value = 2.0 * n * n / (t_asm_full + t_solve_full)
    sample = ('per step: {0} processes x (G by the reference C++ + self terms + K by the C '
              'restatement) of a {1}-triangle sub-mesh of the workload, then one {2} of size {3}; '
              'extrapolated to N = {4}: assembly by entry count ({5:.1f} s), {2} by N^{6} ({7:.1f} s)'
              .format(workers, ns, 'numpy.dot' if poisson else 'numpy.linalg.solve', nlu, n,
                      t_asm_full, 2 if poisson else 3, t_solve_full))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t_step,
        'higher_is_better': True, 'scaling': scaling_label(world, args.workload),
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': bench_config(args, label, n, world),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': workers, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'g_only': {'value': asm_rate_g, 'unit': 'G entries/s (assembly only)', 'cores': workers},
        'extrapolation': {'assembly_GK_entries_per_s': asm_rate_gk, 'assembly_s_at_N': t_asm_full,
                          'solve_s_at_sample': t_solve, 'solve_sample_size': nlu,
                          'solve_s_at_N': t_solve_full, 'sample_step_s': t_step},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------

def local_work(blocks, n, lower, rows):
    """(pairs of triangles this rank evaluates, matrix entries it streams per product)."""
    if lower:
        pairs, stored = 0.0, 0.0
        for (b, e) in ((blocks[0], blocks[1]), (blocks[2], blocks[3])):
            for t0 in range(b, e, 128):
                h = min(128, e - t0)
                w_diag = min(128, n - t0)
                pairs += h * t0 + h * (w_diag - 1)      # below the diagonal tile + inside it
                stored += 128.0 * t0 + 128.0 * 128.0    # whole tiles are streamed
        return pairs, stored
    return rows * (rows - 1) / 2.0 + rows * (n - rows), float(rows) * n


def run_b200(args):
    import ctypes
    import torch
    import torch.distributed as dist
    from icqsol_b200 import _lib
    from icqsol_b200.bem import icqRowPartition as rowPartition
    from icqsol_b200.bem.icqBaseLaplaceSolver import leadingDimension
    from icqsol_b200.mesh.polydata import PolyData, DataArray
    from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver
    from icqsol_b200.bem.icqPoissonSolver import PoissonSolver

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py: no CUDA device; the CUDA path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    if world != args.gpus and rank == 0:
        print('warning: --gpus {0} but WORLD_SIZE {1}'.format(args.gpus, world), file=sys.stderr)

    wname = resolve_workload(world, args.workload)
    poisson = wname == 'c4'
    label, xyz, tri = make_mesh(world, args.workload)
    n = tri.shape[0]
    v = numpy.ones(n) if poisson else centroid_potential(xyz, tri)

    # ---- resident set-up (not timed): context, mesh in HBM, matrix buffers ----------------
    ctx = _lib.Context(local)
    lib = ctx.lib
    stream = torch.cuda.Stream(device=dev)
    ctx.check(lib.icqb_set_stream(ctx.handle, stream.cuda_stream))
    _lib.attachCommunicator(ctx, rank, world, rowPartition.broadcastBytes)
    kind = {'diagonal': 0, 'block': 1, 'block+caps': 2}[args.preconditioner]
    ctx.check(lib.icqb_cg_set_preconditioner(ctx.handle, kind))
    dense_runs = []
    if kind == 2:
        # one dense block per run of sliver tiles, blocks of --block-tiles tiles elsewhere
        cuts = rowPartition.rankCuts(n, world) if (args.storage == 'lower' and world > 1) else None
        block_tiles, block_overlap = resolve_blocks(args, world)
        starts, dense_runs = rowPartition.sliverBlockPartition(xyz, tri, blockTiles=block_tiles, cuts=cuts)
        start_list = [int(t) for t in starts]
        starts = numpy.ascontiguousarray(starts, numpy.int64)
        ctx.check(lib.icqb_cg_set_block_partition(ctx.handle, starts.ctypes.data_as(_lib.c_int64_p),
                                                  len(starts)))
        if block_overlap > 0:
            ext0, ext1 = rowPartition.blockExtents(start_list, (n + rowPartition.TILE - 1) // rowPartition.TILE,
                                                   dense_runs, overlapTiles=block_overlap)
            e0 = numpy.ascontiguousarray(ext0, numpy.int64)
            e1 = numpy.ascontiguousarray(ext1, numpy.int64)
            ctx.check(lib.icqb_cg_set_block_extents(ctx.handle, e0.ctypes.data_as(_lib.c_int64_p),
                                                    e1.ctypes.data_as(_lib.c_int64_p), len(e0)))
    ctx.check(lib.icqb_set_assembly_variant(ctx.handle, args.assembly_variant))
    ctx.check(lib.icqb_cg_set_mixed_precision(ctx.handle, args.mixed_switch))
    ctx.check(lib.icqb_set_mesh(ctx.handle, xyz.ctypes.data_as(_lib.c_double_p), xyz.shape[0],
                                tri.ctypes.data_as(_lib.c_int64_p), n))
    lower = args.storage == 'lower'
    if lower:
        blocks = rowPartition.foldedBlocks(n, rank, world)
        ctx.check(lib.icqb_set_row_blocks(ctx.handle, *blocks))
    else:
        r0, r1 = rowPartition.rowRange(n, rank, world)
        blocks = (r0, r1, r1, r1)
        ctx.check(lib.icqb_set_rows(ctx.handle, r0, r1))
    rows = (blocks[1] - blocks[0]) + (blocks[3] - blocks[2])
    ld = leadingDimension(n)
    if lower:
        ntiles = int(lib.icqb_lower_num_tiles(ctx.handle))
        shape = (max(ntiles, 1), 128, 128)
    else:
        shape = (max(rows, 1), ld)
    gbuf = torch.empty(shape, dtype=torch.float64, device=dev)
    kbuf = torch.empty(shape, dtype=torch.float64, device=dev)
    kubuf = torch.empty(shape, dtype=torch.float64, device=dev) if lower else None
    bdev = torch.from_numpy(v).to(dev)
    xdev = torch.zeros(n, dtype=torch.float64, device=dev)
    area2 = lib.icqb_area2_dev(ctx.handle)
    pairs_local, stored = local_work(blocks, n, lower, rows)
    torch.cuda.synchronize(dev)

    names = ('assembly_ms', 'offdiag_ms', 'diag_ms', 'solve_ms', 'gemv_ms')
    phase = {k: [] for k in names + ('iters', 'resid', 'residinf', 'replacements', 'products', 'products32')}

    def assemble(with_k=True):
        kp = kbuf.data_ptr() if with_k else 0
        if lower:
            ctx.check(lib.icqb_assemble_lower_dev(ctx.handle, 5, gbuf.data_ptr(), kp,
                                                  kubuf.data_ptr() if with_k else 0))
        else:
            ctx.check(lib.icqb_assemble_dev(ctx.handle, 5, gbuf.data_ptr(), kp, ld))

    def step(record):
        with torch.cuda.stream(stream):
            xdev.zero_()
        t0 = time.perf_counter()
        assemble()
        t1 = time.perf_counter()
        iters = ctypes.c_int64(0)
        r2 = ctypes.c_double(0.)
        rinf = ctypes.c_double(0.)
        if poisson:
            if lower:
                ctx.check(lib.icqb_symv_dev(ctx.handle, gbuf.data_ptr(), bdev.data_ptr(), xdev.data_ptr(), 0))
            else:
                ctx.check(lib.icqb_gemv_dev(ctx.handle, gbuf.data_ptr(), rows, n, ld, bdev.data_ptr(),
                                            xdev.data_ptr(), 0, stream.cuda_stream))
        else:
            ctx.check(lib.icqb_cg_solve_dev(ctx.handle, gbuf.data_ptr(), n, ld, 1 if lower else 0, area2, 0,
                                            bdev.data_ptr(), xdev.data_ptr(), 5e-13, 2, args.max_iters,
                                            ctypes.byref(iters), ctypes.byref(r2), ctypes.byref(rinf)))
        if record:
            phase['assembly_ms'].append(1e3 * (t1 - t0))
            phase['offdiag_ms'].append(ctx.last_ms(1))
            phase['diag_ms'].append(ctx.last_ms(2))
            phase['solve_ms'].append(0.0 if poisson else ctx.last_ms(4))
            phase['gemv_ms'].append(ctx.last_ms(3))
            phase['iters'].append(iters.value)
            phase['resid'].append(r2.value)
            phase['residinf'].append(rinf.value)
            rep, prod = (0, 1) if poisson else ctx.cg_stats()
            phase['replacements'].append(rep)
            phase['products'].append(prod)
            n32 = ctypes.c_int64(0)
            ctx.check(lib.icqb_cg_last_mixed(ctx.handle, ctypes.byref(n32)))
            phase['products32'].append(0 if poisson else n32.value)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(local)
    launches0 = ctx.launch_count()
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step(True)
    e1.record(stream)
    barrier()
    sampler.stop_flag = True
    ms_per_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches = ctx.launch_count() - launches0
    value = 2.0 * n * n / (ms_per_step * 1e-3)

    # ---- G-only pass (what the reference computes), same buffers, timed the same way ------
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    assemble(False)
    g0.record(stream)
    for _ in range(2):
        assemble(False)
    g1.record(stream)
    barrier()
    g_only_ms = max_over_ranks(g0.elapsed_time(g1)) / 2.0
    g_only_kernel_ms = ctx.last_ms(1)

    # ---- fp64 peak of this GPU, measured in-run -------------------------------------------
    peak = ctypes.c_double(0.)
    ctx.check(lib.icqb_measure_fp64_peak(ctx.handle, ctypes.byref(peak)))
    far_frac = None
    if args.assembly_variant:
        fast, allp = ctypes.c_int64(0), ctypes.c_int64(0)
        ctx.check(lib.icqb_assembly_stats(ctx.handle, ctypes.byref(fast), ctypes.byref(allp)))
        far_frac = fast.value / max(allp.value, 1)

    # ---- per-rank phase means -> max / min over the ranks ---------------------------------
    mine = {k: float(numpy.mean(phase[k])) for k in names}
    mine['stored_bytes'] = 8.0 * stored
    mine['pairs'] = pairs_local
    if world > 1:
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
    else:
        everyone = [mine]

    # ---- e2e through the public Python API (host polydata in, host response out) ----------
    # the resident buffers go first: on one GPU the 84,600-triangle matrices take 86 GB, and the
    # solver class holds its own
    matrix_bytes = 3.0 * 8.0 * stored if lower else 2.0 * 8.0 * stored
    tight = 2.0 * matrix_bytes > 0.8 * torch.cuda.mem_get_info(dev)[1]
    del gbuf, kbuf, kubuf
    if tight:
        torch.cuda.empty_cache()
    e2e_times = []
    e2e_phases = {}
    h2d = xyz.nbytes + tri.nbytes + v.nbytes
    d2h = 8 * n
    for it in range(1 + args.e2e_steps if args.e2e_steps > 0 else 0):
        pd = PolyData.from_arrays(xyz, tri, dtype=numpy.float64)
        pd.GetCellData().AddArray(DataArray('charge' if poisson else 'v', v))
        barrier()
        t0 = time.perf_counter()
        cls = PoissonSolver if poisson else LaplaceSolver
        solver = cls(pd, float('inf'), 5, computeK=True, device=local,
                     storage=args.storage, preconditioner=args.preconditioner,
                     blockTiles=resolve_blocks(args, world)[0], blockOverlap=resolve_blocks(args, world)[1],
                     assemblyVariant=args.assembly_variant,
                     mixedPrecision=args.mixed_switch)
        if not poisson:
            solver.setSolverMaxNumberOfIterations(args.max_iters)
        rsp = solver.computeResponseField()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if it > 0:
            e2e_times.append(dt)
        # host-side phases of this end-to-end step: cumulative milliseconds -> per phase
        e2e_phases = {}
        for group in ('constructor', 'solve'):
            last = 0.0
            for name, t in solver.hostTimings[group]:
                e2e_phases['{0}: {1}'.format(group, name)] = round(t - last, 3)
                last = t
        del solver, rsp
        if tight:       # (otherwise the caching allocator hands the next solver the same blocks)
            torch.cuda.empty_cache()
    e2e_s = None
    if e2e_times:
        e2e_s = max_over_ranks(float(numpy.mean(e2e_times)))
    if rank == 0:
        peaks, peak_src = measured_peaks()
        off_ms = mine['offdiag_ms']
        gemv_ms = mine['gemv_ms']
        iters_mean = float(numpy.mean(phase['iters']))
        bmax = float(numpy.abs(v).max())
        # the off-diagonal launch of the rank that evaluates the most pairs
        heavy = max(everyone, key=lambda d: d['pairs'])
        flop = heavy['pairs'] * 256.0 * FLOP_PER_EVAL_GK
        achieved = flop / (heavy['offdiag_ms'] * 1e-3) / 1e12
        # FP64-pipe instructions per evaluation of the kernel's inner loop (SASS counts,
        # tools/sass_loops.py; G+K): validated kernel 16; far-field arithmetic 13
        instr_per_eval = 16.0 if far_frac is None else 16.0 - far_frac * 3.0
        # bytes of matrix a rank streams per product (float32 late products: 4 bytes per entry
        # for their share of the products); heaviest rank, and its time (the exchange between
        # the ranks is inside the timed product)
        frac32 = float(numpy.mean(phase['products32'])) / max(iters_mean, 1.0)
        width = 8.0 - 4.0 * min(max(frac32, 0.0), 1.0)
        heavy_g = max(everyone, key=lambda d: d['stored_bytes'])
        gemv_bytes = heavy_g['stored_bytes'] / 8.0 * width
        gemv_ms_max = max(d['gemv_ms'] for d in everyone)
        gemv_gbs = gemv_bytes / (gemv_ms_max * 1e-3) / 1e9 if gemv_ms_max > 0 else 0.0
        products_mean = 1.0 if poisson else iters_mean + 2 + float(numpy.max(phase['replacements']))
        spread = {k: {'max': max(d[k] for d in everyone), 'min': min(d[k] for d in everyone)}
                  for k in names + ('stored_bytes', 'pairs')}
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': scaling_label(world, args.workload), 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': bench_config(args, label, n, world),
            'phases': {'assembly_ms': mine['assembly_ms'],
                       'offdiag_kernel_ms': off_ms,
                       'diag_kernel_ms': mine['diag_ms'],
                       'assembly_entries_per_s': 2.0 * n * n / (spread['offdiag_ms']['max'] * 1e-3),
                       'solve_ms': mine['solve_ms'],
                       'cg_iters': int(iters_mean),
                       'cg_residual_replacements': int(numpy.max(phase['replacements'])),
                       'cg_products_enqueued': int(numpy.mean(phase['products'])),
                       'cg_float32_products': int(numpy.mean(phase['products32'])),
                       'cg_converged': bool(poisson or numpy.max(phase['residinf']) <= 1e-12 * bmax),
                       'cg_residual2': float(numpy.max(phase['resid'])),
                       'cg_backward_error_max_norm': float(numpy.max(phase['residinf']) / bmax),
                       'dense_preconditioner_runs': [list(map(int, r)) for r in dense_runs],
                       'gemv_ms': gemv_ms, 'gemv_gbs_per_gpu': gemv_gbs,
                       'gemv_gbs_aggregate': sum(d['stored_bytes'] for d in everyone) / 8.0 * width /
                       (gemv_ms_max * 1e-3) / 1e9 if gemv_ms_max > 0 else 0.0,
                       'gemv_equivalent_dense_gbs_aggregate': 8.0 * n * n / (gemv_ms_max * 1e-3) / 1e9
                       if gemv_ms_max > 0 else 0.0,
                       'over_ranks': spread,
                       'stored_bytes_per_rank': [d['stored_bytes'] for d in everyone]},
            'roofline': None,
            'roofline_assembly': {
                'bound': 'fp64', 'kernel': 'k_offdiag<G+K>' if not args.assembly_variant else 'k_offdiag_ff<G+K> (variant {0})'.format(args.assembly_variant), 'achieved': achieved,
                'peak': peak.value, 'unit': 'TFLOP/s', 'frac': achieved / peak.value,
                'traffic': traffic_of('k_offdiag_ff' if args.assembly_variant else 'k_offdiag', args.storage, n),
                'peak_source': 'DFMA loop measured in this run (icqb_measure_fp64_peak)',
                'flop_per_evaluation': FLOP_PER_EVAL_GK, 'fp64_instr_per_evaluation': instr_per_eval,
                'far_field_fraction': far_frac,
                'issue_frac': (heavy['pairs'] * 256.0 * instr_per_eval * 2 / (heavy['offdiag_ms'] * 1e-3) / 1e12) / peak.value,
                'launches_per_step': 1, 'ms_per_launch': heavy['offdiag_ms'],
                'share_of_step': heavy['offdiag_ms'] / ms_per_step},
            'roofline_gemv': {
                'bound': 'hbm', 'kernel': 'k_symv+k_symv_reduce' if lower else 'k_gemv',
                'achieved': gemv_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                'frac': gemv_gbs / peaks['hbm_gbs'],
                'traffic': traffic_of('k_symv' if lower else 'k_gemv', args.storage, n),
                'peak_source': peak_src, 'algorithmic_bytes_per_launch': gemv_bytes,
                'launches_per_step': products_mean, 'ms_per_launch': gemv_ms_max,
                'share_of_step': min(1.0, products_mean * gemv_ms_max / ms_per_step),
                'rank': 'heaviest rank (bytes), slowest rank (time), exchange between the ranks inside'},
            'g_only': {'value': float(n) * n / (g_only_ms * 1e-3), 'unit': 'G entries/s (assembly only)',
                       'ms': g_only_ms, 'offdiag_kernel_ms': g_only_kernel_ms},
            'e2e': None if e2e_s is None else {
                'value': 2.0 * n * n / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': int(d2h), 'ms_per_step': 1e3 * e2e_s,
                'host_phases_ms': e2e_phases,
                'api': '{0}(pdata, inf, 5, computeK=True).computeResponseField()'.format(
                    'PoissonSolver' if poisson else 'LaplaceSolver')},
            'gpu_launches': int(launches),
            'clocks': sampler.summary(),
        }
        # strong scaling: the one-GPU run of the SAME workload (`--gpus 1` runs config C2, not this
        # mesh), as measured by this repository on a box of the same pool -- for the reader of the
        # line; nothing below depends on it
        if world > 1:
            try:
                ref_file = {'bolt': 'r2_bench_n1_bolt_final.json'}.get(wname)
                if ref_file:
                    with open(os.path.join(ROOT, 'profiles', ref_file)) as f:
                        one = json.loads([ln for ln in f if ln.startswith('{')][-1])
                    if one['config']['triangles'] == n:
                        line['one_gpu_same_workload'] = {
                            'ms_per_step': one['ms_per_step'], 'value': one['value'], 'unit': one['unit'],
                            'speedup_of_this_run': one['ms_per_step'] / ms_per_step,
                            'source': 'profiles/' + ref_file + ' (bench.py --workload bolt on one B200)'}
            except Exception:
                pass
        # `roofline` = the kernel with the larger share of the step
        dom = 'roofline_gemv' if line['roofline_gemv']['share_of_step'] > \
            line['roofline_assembly']['share_of_step'] else 'roofline_assembly'
        line['roofline'] = dict(line[dom])
        if not args.no_cpu_baseline:
            wall, wall_g, entries, kindc = cpu_assembly_sample(xyz, tri, args.cpu_sample, 1)
            line['cpu_baseline'] = {
                'value': 2.0 * entries / wall, 'unit': UNIT, 'cores': 1, 'kind': kindc,
                'g_only_value': entries / wall_g,
                'sample': 'G and K of the first {0} triangles of the workload, assembly only: reference '
                          'C++ off-diagonal + restated self term for G, C restatement for K (the '
                          'reference has none), {1:.1f} s, 1 thread'.format(min(args.cpu_sample, n), wall)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='auto',
                    choices=['auto', 'c2', 'c3', 'c4', 'c5', 'bolt', 'weak', 'small', 'tiny'],
                    help="auto = config C2 on one GPU, the refined bolt of config C3 (84,600 "
                         "triangles, strong scaling) on several; c3/c5 = UV spheres of the C3/C5 "
                         "sizes; c4 = hollow sphere, Poisson product; weak = C2 refined with "
                         "N^2/GPU fixed")
    ap.add_argument('--cpu-sample', type=int, default=2000,
                    help='triangles of the CPU baseline sub-mesh (2000 -> ~15 s single core, G and K)')
    ap.add_argument('--ref-sample', type=int, default=1000,
                    help='--impl reference: triangles of the sub-mesh each host process assembles per step')
    ap.add_argument('--ref-lu-size', type=int, default=5000,
                    help='--impl reference: size of the LU solve timed per step (extrapolated by N^3)')
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--storage', default='lower', choices=['lower', 'full'],
                    help="'lower': each pair entry stored once (mirror rule), symmetric product; "
                         "'full': complete row blocks, row-major GEMV")
    ap.add_argument('--preconditioner', default='block+caps', choices=['block', 'diagonal', 'block+caps'],
                    help="CG preconditioner: 'block+caps' (default) = inverse diagonal blocks of "
                         "--block-tiles tiles, with one dense block over every run of sliver tiles "
                         "(DESIGN.md 4.6); 'block' = inverse 128x128 blocks; 'diagonal' = 1/diag")
    ap.add_argument('--block-tiles', type=int, default=None,
                    help='with block+caps: 128-triangle tiles per diagonal block (default 4 on one '
                         'GPU, 8 on several)')
    ap.add_argument('--block-overlap', type=int, default=None,
                    help='with block+caps on one GPU: tiles by which every block is extended on either '
                         'side (overlapping blocks, default 1; 0 = plain blocks)')
    ap.add_argument('--assembly-variant', type=int, default=1, choices=[0, 1],
                    help='arithmetic of the off-diagonal kernel (include/icqb.h, '
                         'icqb_set_assembly_variant): 0 direct, 1 far-field split')
    ap.add_argument('--mixed-switch', type=float, default=0.,
                    help='with block+caps and lower storage: relative residual below '
                         'which the CG products stream a float32 copy of the matrix (1e-6); 0 = off')
    ap.add_argument('--max-iters', type=int, default=1500,
                    help='CG iteration cap (a solve that needs more is reported as not converged)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
