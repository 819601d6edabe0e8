#!/usr/bin/env python
"""Config C4 of BASELINE.json on one GPU (not the bench contract: a note-taking tool):
PoissonSolver on the 49,400-triangle hollow sphere -- G assembly (no K;
the reference has none and the Poisson response does not need it) + one product v = G charge
(bem/icqPoissonSolver.py:36), through the public classes, host mesh in / host response out.
Prints one JSON line.   usage: tools/bench_c4.py [steps] [assembly_variant] [n_theta n_phi (outer sphere; default 100 50)]"""
import json
import os
import sys
import time

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from icqsol_b200.mesh.generators import hollowSphere
    from icqsol_b200.mesh.polydata import PolyData, DataArray
    from icqsol_b200.bem.icqPoissonSolver import PoissonSolver
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    nt, nph = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (100, 50)
    xyz, tri = hollowSphere(2.0, 1.0, nt, nph)
    n = tri.shape[0]
    charge = numpy.ones(n)
    e2e, asm, prod = [], [], []
    for it in range(steps + 1):
        pd = PolyData.from_arrays(xyz, tri, dtype=numpy.float64)
        pd.GetCellData().AddArray(DataArray('charge', charge))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        solver = PoissonSolver(pd, float('inf'), 5, assemblyVariant=variant)
        v = solver.computeResponseField()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it > 0:
            e2e.append(dt)
            asm.append(solver.assemblyMilliseconds['offDiagonal'] + solver.assemblyMilliseconds['diagonal'])
            prod.append(solver.ctx.last_ms(3))
        del solver
    ntiles = sum(i + 1 for i in range((n + 127) // 128))
    out = {'config': 'C4: hollow sphere {1}x{2}, N={0}, PoissonSolver, diagonal order 5'.format(n, nt, nph),
           'assembly_variant': variant, 'e2e_ms': round(1e3 * float(numpy.mean(e2e)), 2),
           'assembly_ms': round(float(numpy.mean(asm)), 2),
           'G_entries_per_s_assembly': float(n) * n / (numpy.mean(asm) * 1e-3),
           'product_ms': round(float(numpy.mean(prod)), 4),
           'product_GB_per_s': round(8.0 * ntiles * 16384 / (numpy.mean(prod) * 1e-3) / 1e9, 1),
           'response_min_max': [float(v.min()), float(v.max())]}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
