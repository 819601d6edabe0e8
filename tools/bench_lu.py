#!/usr/bin/env python
"""Direct solve (csrc/lu.cu: LU with partial pivoting, the reference's numpy.linalg.solve) on a
bench workload, one rank per GPU: times of the row assembly, the factorisation and the solve
sweeps, refinement history, and the backward error measured with the resident matrix.
    python tools/bench_lu.py [workload (c2|bolt|c3|...)]           (under torchrun for several GPUs)
Prints one JSON line (rank 0)."""
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import bench
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    name = sys.argv[1] if len(sys.argv) > 1 else 'c2'
    label, xyz, tri = bench.make_mesh(world, name)
    n = tri.shape[0]
    solver = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5, device=local)
    v = bench.centroid_potential(xyz, tri)
    t0 = time.perf_counter()
    x = solver.solveGreenSystemDirect(v)
    wall = time.perf_counter() - t0
    info = solver.lastSolveInfo
    t0 = time.perf_counter()
    xcg = solver.solveGreenSystem(v, tol=5e-13, maxNumIters=1500)
    wall_cg = time.perf_counter() - t0
    cg = solver.lastSolveInfo
    if rank == 0:
        flop = 2. / 3. * float(n) ** 3
        print(json.dumps({
            'workload': label, 'triangles': n, 'ranks': world, 'wall_s': wall,
            'lu_assembly_ms': info['luAssemblyMilliseconds'], 'lu_factor_ms': info['luFactorMilliseconds'],
            'lu_solve_ms': info['luSolveMilliseconds'],
            'factor_tflops_aggregate': flop / (info['luFactorMilliseconds'] * 1e-3) / 1e12,
            'refinements': info['refinements'],
            'backward_error_history': [r / numpy.abs(v).max() for r in info['residualHistory']],
            'converged': info['converged'],
            'cg_wall_s': wall_cg, 'cg_iterations': cg['iterations'], 'cg_method': cg['method'],
            'max_rel_diff_lu_vs_cg': float(numpy.abs(x - xcg).max() / numpy.abs(xcg).max())}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
