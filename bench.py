#!/usr/bin/env python
"""Benchmark of the boundary-element hot path (BASELINE.json: "G+K entries/sec
(fp64) and CG GEMV GB/s ... vs host-CPU reference").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path

One step = one pass of the hot path on a synthetic refined UV sphere
(SURVEY.md section 8d): fused G+K assembly (diagonal + off-diagonal kernels) into
resident HBM buffers, then the dense CG solve of G sigma = -v to a backward error
max|v + G sigma| <= 5e-13 max|v| (inside the 1e-12 parity criterion).  N=1 runs config C2 (19,880 triangles); for N>1 the mesh is
refined so that the N^2 matrix work per GPU stays fixed ("weak" scaling), rows
block-sharded over the ranks, NCCL only inside the CG.

value   = G+K entries per second over the whole step = 2*N^2 / ms_per_step
          (N^2 entries per matrix are DEFINED by a step in either storage: in 'lower'
          storage the upper half is the mirror rule applied on read)
e2e     = same metric through the public Python API (LaplaceSolver on a host
          polydata -> host response), host<->device copies inside the timed region
roofline= dominant kernel (fused G+K off-diagonal assembly) against the FP64 FMA
          rate measured on this GPU in the same run; roofline_gemv = CG GEMV against
          MEASURED_PEAKS.json's HBM copy bandwidth
cpu_baseline = the reference's own C++ (oracle/_ref/libicqref.so, unmodified
          sources) or, when that was not built, the C restatement, on a bounded
          sub-mesh sample, 1 thread (the reference ships single-threaded).
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


def make_mesh(ngpus, name):
    """(label, xyz, tri) of the workload: a UV sphere, or -- 'bolt' -- config C3's geometry:
    the reference's test_results/testBoltFine.vtk (2,350 triangles, kept as a golden fixture in
    tests/golden/meshes.npz) with every triangle cut into 36 (84,600 triangles)."""
    if name == 'bolt':
        from icqsol_b200.mesh.generators import subdivide
        m = numpy.load(os.path.join(ROOT, 'tests', 'golden', 'meshes.npz'))
        xyz, tri = subdivide(m['boltFine_xyz'], m['boltFine_tri'], 6)
        return 'testBoltFine_k6_N{0} (config C3: refined bolt)'.format(tri.shape[0]), xyz, tri
    from icqsol_b200.mesh.generators import uvSphere
    label, nt, nph = workload(ngpus, name)
    xyz, tri = uvSphere(1.0, nt, nph)
    return label, xyz, tri


def workload(ngpus, name):
    """(label, n_theta, n_phi) of the UV sphere for this GPU count."""
    if name == 'c2' or (name == 'auto' and ngpus == 1):
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
    xyz, tri, order, use_ref = args
    from oracle import oracle
    t0 = time.perf_counter()
    if use_ref:
        g = oracle.ref_offdiag(xyz, tri)
    else:
        g = oracle.offdiag(xyz, tri)
    g[numpy.diag_indices_from(g)] = oracle.diag(xyz, tri, order)
    return time.perf_counter() - t0, float(g[0, 1])


def cpu_assembly_sample(xyz, tri, ns, workers, order=5):
    """Each worker assembles G for its own ns-triangle sub-mesh of the workload with
    the reference's C++ (off-diagonal) and the restated self term.  Returns
    (seconds wall, entries, kind)."""
    from oracle import oracle
    oracle.build()
    use_ref = oracle.have_reference()
    n = tri.shape[0]
    jobs = []
    for w in range(workers):
        start = (w * ns) % max(1, n - ns)
        jobs.append((xyz, tri[start:start + ns], order, use_ref))
    t0 = time.perf_counter()
    if workers == 1:
        _ref_worker(jobs[0])
    else:
        import multiprocessing
        with multiprocessing.get_context('fork').Pool(workers) as pool:
            pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t0
    return wall, workers * ns * ns, 'reference' if use_ref else 'port'


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    label, xyz, tri = make_mesh(args.gpus, args.workload)
    cores = os.cpu_count() or 1
    ns = args.cpu_sample
    times = []
    kind = 'port'
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        wall, entries, kind = cpu_assembly_sample(xyz, tri, ns, cores)
        # the reference's solve: LU on the sample system (bem/icqLaplaceSolver.py:36)
        from oracle import oracle
        g = oracle.green_matrix(xyz, tri[:ns], 5, threads=True) if it == 0 else g
        v = centroid_potential(xyz, tri[:ns])
        oracle.laplace_response(g, v)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append((dt, entries))
    total_t = sum(t for t, _ in times)
    total_e = sum(e for _, e in times)
    value = total_e / total_t
    sample = ('{0} processes x G of a {1}-triangle sub-mesh of the workload (reference C++ '
              'off-diagonal + restated self term) + one numpy LU solve of that size; the reference '
              'has no K, entries counted are G entries'.format(cores, ns))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total_t / len(times),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': {'workload': label, 'sample': sample},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------------------

def run_b200(args):
    import ctypes
    import torch
    import torch.distributed as dist
    from icqsol_b200 import _lib
    from icqsol_b200.bem import icqRowPartition as rowPartition
    from icqsol_b200.bem.icqBaseLaplaceSolver import leadingDimension
    from icqsol_b200.mesh.polydata import PolyData, DataArray
    from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver

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

    label, xyz, tri = make_mesh(world, args.workload)
    n = tri.shape[0]
    v = centroid_potential(xyz, tri)

    # ---- resident set-up (not timed): context, mesh in HBM, matrix buffers ----------------
    ctx = _lib.Context(local)
    lib = ctx.lib
    stream = torch.cuda.Stream(device=dev)
    ctx.check(lib.icqb_set_stream(ctx.handle, stream.cuda_stream))
    _lib.attachCommunicator(ctx, rank, world, rowPartition.broadcastBytes)
    ctx.check(lib.icqb_cg_set_preconditioner(
        ctx.handle, {'diagonal': 0, 'block': 1, 'block+caps': 2}[args.preconditioner]))
    cap_tiles = (0, 0)
    if args.preconditioner == 'block+caps':    # experimental, see icqsol_b200/csrc/cg_caps.cu
        cap_tiles = rowPartition.sliverCapTiles(xyz, tri)
        ctx.check(lib.icqb_cg_set_cap_tiles(ctx.handle, *cap_tiles))
        starts, dense_runs = rowPartition.sliverBlockPartition(xyz, tri, blockTiles=args.block_tiles)
        if dense_runs or args.block_tiles > 1:
            starts = numpy.ascontiguousarray(starts, numpy.int64)
            ctx.check(lib.icqb_cg_set_block_partition(ctx.handle, starts.ctypes.data_as(_lib.c_int64_p),
                                                      len(starts)))
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
        gbuf = torch.zeros(shape, dtype=torch.float64, device=dev)
        kbuf = torch.zeros(shape, dtype=torch.float64, device=dev)
        kubuf = torch.zeros(shape, dtype=torch.float64, device=dev)
    else:
        gbuf = torch.empty((max(rows, 1), ld), dtype=torch.float64, device=dev)
        kbuf = torch.empty((max(rows, 1), ld), dtype=torch.float64, device=dev)
        kubuf = None
    bdev = torch.from_numpy(v).to(dev)
    xdev = torch.zeros(n, dtype=torch.float64, device=dev)
    area2 = lib.icqb_area2_dev(ctx.handle)
    # pairs of triangles this rank evaluates, stored entries it streams per product
    if lower:
        tiles = 0
        pairs_local = 0.0
        stored = 0.0
        for (b, e) in ((blocks[0], blocks[1]), (blocks[2], blocks[3])):
            for t0 in range(b, e, 128):
                h = min(128, e - t0)
                w_diag = min(128, n - t0)
                pairs_local += h * t0 + h * (w_diag - 1)      # below the diagonal tile + inside it
                stored += 128.0 * t0 + 128.0 * 128.0    # whole tiles are streamed
    else:
        pairs_local = rows * (rows - 1) / 2.0 + rows * (n - rows)
        stored = float(rows) * n
    torch.cuda.synchronize(dev)

    phase = {'assembly_ms': [], 'offdiag_ms': [], 'diag_ms': [], 'solve_ms': [], 'gemv_ms': [],
             'iters': [], 'resid': [], 'residinf': [], 'replacements': [], 'products': [], 'products32': []}

    def step(record):
        with torch.cuda.stream(stream):
            xdev.zero_()
        t0 = time.perf_counter()
        if lower:
            ctx.check(lib.icqb_assemble_lower_dev(ctx.handle, 5, gbuf.data_ptr(), kbuf.data_ptr(),
                                                  kubuf.data_ptr()))
        else:
            ctx.check(lib.icqb_assemble_dev(ctx.handle, 5, gbuf.data_ptr(), kbuf.data_ptr(), ld))
        t1 = time.perf_counter()
        iters = ctypes.c_int64(0)
        r2 = ctypes.c_double(0.)
        rinf = ctypes.c_double(0.)
        ctx.check(lib.icqb_cg_solve_dev(ctx.handle, gbuf.data_ptr(), n, ld, 1 if lower else 0, area2, 0,
                                        bdev.data_ptr(), xdev.data_ptr(), 5e-13, 2, args.max_iters,
                                        ctypes.byref(iters), ctypes.byref(r2), ctypes.byref(rinf)))
        if record:
            phase['assembly_ms'].append(1e3 * (t1 - t0))
            phase['offdiag_ms'].append(ctx.last_ms(1))
            phase['diag_ms'].append(ctx.last_ms(2))
            phase['solve_ms'].append(ctx.last_ms(4))
            phase['gemv_ms'].append(ctx.last_ms(3))
            phase['iters'].append(iters.value)
            phase['resid'].append(r2.value)
            phase['residinf'].append(rinf.value)
            rep, prod = ctx.cg_stats()
            phase['replacements'].append(rep)
            phase['products'].append(prod)
            n32 = ctypes.c_int64(0)
            ctx.check(lib.icqb_cg_last_mixed(ctx.handle, ctypes.byref(n32)))
            phase['products32'].append(n32.value)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

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
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = 2.0 * n * n / (ms_per_step * 1e-3)

    # ---- fp64 peak of this GPU, measured in-run -------------------------------------------
    peak = ctypes.c_double(0.)
    ctx.check(lib.icqb_measure_fp64_peak(ctx.handle, ctypes.byref(peak)))

    # ---- e2e through the public Python API (host polydata in, host response out) ----------
    e2e_times = []
    h2d = xyz.nbytes + tri.nbytes + v.nbytes
    d2h = 8 * n
    for it in range(1 + args.e2e_steps if args.e2e_steps > 0 else 0):
        pd = PolyData.from_arrays(xyz, tri, dtype=numpy.float64)
        pd.GetCellData().AddArray(DataArray('v', v))
        barrier()
        t0 = time.perf_counter()
        solver = LaplaceSolver(pd, float('inf'), 5, computeK=True, device=local,
                               storage=args.storage, preconditioner=args.preconditioner,
                               blockTiles=args.block_tiles, assemblyVariant=args.assembly_variant,
                               mixedPrecision=args.mixed_switch)
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
        del solver
    e2e_s = None
    if e2e_times:
        e2e_s = float(numpy.mean(e2e_times))
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
    if rank == 0:
        peaks, peak_src = measured_peaks()
        off_ms = float(numpy.mean(phase['offdiag_ms']))
        gemv_ms = float(numpy.mean(phase['gemv_ms']))
        # algorithmic flop of this rank's off-diagonal launch (pairs it evaluates x 256 x 18)
        flop = pairs_local * 256.0 * FLOP_PER_EVAL_GK
        achieved = flop / (off_ms * 1e-3) / 1e12
        # FP64-pipe instructions per evaluation of the kernel's inner loop (SASS counts,
        # tools/sass_loops.py; G+K): validated kernel 16; far-field arithmetic 13 (12 with the
        # Newton step) for the (warp, source triangle) pairs that took it
        far_frac = None
        instr_per_eval = 16.0
        if args.assembly_variant:
            fast, allp = ctypes.c_int64(0), ctypes.c_int64(0)
            ctx.check(lib.icqb_assembly_stats(ctx.handle, ctypes.byref(fast), ctypes.byref(allp)))
            far_frac = fast.value / max(allp.value, 1)
            instr_per_eval = 16.0 - far_frac * (3.0 if args.assembly_variant in (1, 3) else 4.0)
        # bytes of matrix this rank streams per product; with float32 late products
        # (--mixed-switch) the timed products are a mixture: 8 bytes per stored entry for the
        # double ones, 4 for the float32 ones, weighted by their share of the iterations
        frac32 = float(numpy.mean(phase['products32'])) / max(float(numpy.mean(phase['iters'])), 1.0)
        gemv_bytes = stored * (8.0 - 4.0 * min(max(frac32, 0.0), 1.0))
        gemv_gbs = gemv_bytes / (gemv_ms * 1e-3) / 1e9 if gemv_ms > 0 else 0.0
        # products that ran: iterations + the initial one + one per confirmation (those
        # enqueued behind the stop flag return at once and are not counted)
        products_mean = float(numpy.mean(phase['iters'])) + 2 + float(numpy.max(phase['replacements']))
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': label, 'triangles': n, 'diag_order': 5, 'offdiag_order': 8,
                       'step': 'fused G+K assembly + CG solve (max|v+G sigma| <= 5e-13 max|v|)',
                       'rows_per_gpu': rows, 'storage': args.storage,
                       'preconditioner': args.preconditioner, 'cap_tiles': list(cap_tiles), 'block_tiles': args.block_tiles,
                       'assembly_variant': args.assembly_variant, 'mixed_switch': args.mixed_switch,
                       'parallelism': ('folded row blocks, lower storage x{0}' if lower
                                       else 'row-block x{0}').format(world),
                       'l2': 'inputs larger than L2 ({0:.2f} GB of G streamed per product per GPU)'.format(
                           gemv_bytes / 1e9)},
            'phases': {'assembly_ms': float(numpy.mean(phase['assembly_ms'])),
                       'offdiag_kernel_ms': off_ms,
                       'diag_kernel_ms': float(numpy.mean(phase['diag_ms'])),
                       'assembly_entries_per_s': 2.0 * n * n / (numpy.mean(phase['assembly_ms']) * 1e-3),
                       'solve_ms': float(numpy.mean(phase['solve_ms'])),
                       'cg_iters': int(numpy.mean(phase['iters'])),
                       'cg_residual_replacements': int(numpy.max(phase['replacements'])),
                       'cg_products_enqueued': int(numpy.mean(phase['products'])),
                       'cg_float32_products': int(numpy.mean(phase['products32'])),
                       'cg_converged': bool(numpy.max(phase['residinf']) <= 1e-12 * numpy.abs(v).max()),
                       'cg_residual2': float(numpy.max(phase['resid'])),
                       'cg_backward_error_max_norm': float(numpy.max(phase['residinf']) / numpy.abs(v).max()),
                       'gemv_ms': gemv_ms, 'gemv_gbs_per_gpu': gemv_gbs,
                       'gemv_gbs_aggregate': gemv_gbs * world,
                       'gemv_equivalent_dense_gbs_aggregate': 8.0 * n * n / (gemv_ms * 1e-3) / 1e9
                       if gemv_ms > 0 else 0.0},
            'roofline': None,
            'roofline_assembly': {
                'bound': 'fp64', 'kernel': 'k_offdiag<G+K>' if not args.assembly_variant else 'k_offdiag_ff<G+K> (variant {0})'.format(args.assembly_variant), 'achieved': achieved,
                'peak': peak.value, 'unit': 'TFLOP/s', 'frac': achieved / peak.value,
                'traffic': traffic_of('k_offdiag', args.storage, n) if not args.assembly_variant else None,
                'peak_source': 'DFMA loop measured in this run (icqb_measure_fp64_peak)',
                'flop_per_evaluation': FLOP_PER_EVAL_GK, 'fp64_instr_per_evaluation': instr_per_eval,
                'far_field_fraction': far_frac,
                'issue_frac': (pairs_local * 256.0 * instr_per_eval * 2 / (off_ms * 1e-3) / 1e12) / peak.value,
                'launches_per_step': 1, 'ms_per_launch': off_ms,
                'share_of_step': off_ms / ms_per_step},
            'roofline_gemv': {
                'bound': 'hbm', 'kernel': 'k_symv+k_symv_reduce' if lower else 'k_gemv',
                'achieved': gemv_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                'frac': gemv_gbs / peaks['hbm_gbs'],
                'traffic': traffic_of('k_symv' if lower else 'k_gemv', args.storage, n),
                'peak_source': peak_src, 'algorithmic_bytes_per_launch': gemv_bytes,
                'launches_per_step': products_mean, 'ms_per_launch': gemv_ms,
                'share_of_step': min(1.0, products_mean * gemv_ms / ms_per_step)},
            'e2e': None if e2e_s is None else {
                'value': 2.0 * n * n / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': int(d2h), 'ms_per_step': 1e3 * e2e_s,
                'host_phases_ms': e2e_phases,
                'api': 'LaplaceSolver(pdata, inf, 5, computeK=True).computeResponseField()'},
            'gpu_launches': int(launches),
            'clocks': sampler.summary(),
        }
        # `roofline` = the kernel with the larger share of the step
        dom = 'roofline_gemv' if line['roofline_gemv']['share_of_step'] > \
            line['roofline_assembly']['share_of_step'] else 'roofline_assembly'
        line['roofline'] = dict(line[dom])
        if world == 1 and not args.no_cpu_baseline:
            wall, entries, kind = cpu_assembly_sample(xyz, tri, args.cpu_sample, 1)
            line['cpu_baseline'] = {
                'value': entries / wall, 'unit': UNIT, 'cores': 1, 'kind': kind,
                'sample': 'G (no K in the reference) of the first {0} triangles of the workload: '
                          'reference C++ off-diagonal + restated self term, {1:.1f} s'.format(
                              args.cpu_sample, wall)}
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
    ap.add_argument('--workload', default='auto', choices=['auto', 'c2', 'c3', 'c5', 'small', 'tiny', 'bolt'],
                    help="auto = config C2 on one GPU, C2 scaled with N^2/GPU fixed on several; c3/c5 = "
                         "UV spheres of the C3/C5 sizes; bolt = config C3's own geometry (refined "
                         "testBoltFine, 84,600 triangles)")
    ap.add_argument('--cpu-sample', type=int, default=3600,
                    help='triangles of the CPU baseline sub-mesh (3600 -> 10-20 s single core)')
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--storage', default='lower', choices=['lower', 'full'],
                    help="'lower': each pair entry stored once (mirror rule), symmetric product; "
                         "'full': complete row blocks, row-major GEMV")
    ap.add_argument('--preconditioner', default='auto', choices=['auto', 'block', 'diagonal', 'block+caps'],
                    help="CG preconditioner: inverse 128x128 diagonal blocks, 1/diag, or (experimental) "
                         "blocks with dense caps over the sliver tiles; auto = block, except for the "
                         "c3/c5 workloads whose pole slivers make the matrix indefinite (DESIGN.md 4.6): "
                         "block+caps there")
    ap.add_argument('--block-tiles', type=int, default=1,
                    help='experimental, with block+caps: 128-triangle tiles per diagonal block')
    ap.add_argument('--assembly-variant', type=int, default=0, choices=[0, 1, 2, 3, 4],
                    help='experimental arithmetic of the off-diagonal kernel (include/icqb.h, '
                         'icqb_set_assembly_variant): 0 validated kernel, 1 far-field split, '
                         '2 far-field split + Newton reciprocal square root, 3/4 = 1/2 at three CTAs per SM')
    ap.add_argument('--mixed-switch', type=float, default=0.,
                    help='experimental, with block+caps and lower storage: relative residual below '
                         'which the CG products stream a float32 copy of the matrix (1e-6); 0 = off')
    ap.add_argument('--max-iters', type=int, default=4000,
                    help='CG iteration cap (a solve that needs more is reported as not converged)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.preconditioner == 'auto':
        args.preconditioner = 'block+caps' if args.workload in ('c3', 'c5') else 'block'
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
