#!/usr/bin/env python
"""Kernel-level timings on one GPU (not the bench contract: tools for the round's notes).
  tools/bench_kernels.py [n_theta n_phi]   -- default 142 71 (config C2)
Times, with CUDA events on the library's stream, after warm-up:
  * the off-diagonal assembly (G only and G+K) for assembly variants 0 and 1;
  * the symmetric product on the double matrix (k_symv + k_symv_reduce) and on its float32 copy
    (k_symv_f32 + k_symv_reduce), as GB/s of the bytes each one streams.
Prints one JSON line."""
import ctypes
import json
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from icqsol_b200 import _lib
    from icqsol_b200.bem import icqRowPartition as rp
    from icqsol_b200.mesh.generators import uvSphere
    nt, nph = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (142, 71)
    xyz, tri = uvSphere(1.0, nt, nph)
    n = tri.shape[0]
    dev = torch.device('cuda', 0)
    ctx = _lib.Context(0)
    lib = ctx.lib
    ctx.check(lib.icqb_set_mesh(ctx.handle, xyz.ctypes.data_as(_lib.c_double_p), xyz.shape[0],
                                tri.ctypes.data_as(_lib.c_int64_p), n))
    ctx.check(lib.icqb_set_row_blocks(ctx.handle, 0, n, n, n))
    ntiles = int(lib.icqb_lower_num_tiles(ctx.handle))
    g = torch.empty((ntiles, rp.TILE, rp.TILE), dtype=torch.float64, device=dev)
    k = torch.empty_like(g)
    ku = torch.empty_like(g)
    out = {'triangles': n, 'stored_tiles': ntiles, 'assembly_ms': {}, 'far_field_fraction': {}}
    pairs = 0.5 * n * n
    keep = {}
    for variant in (0, 1):
        ctx.check(lib.icqb_set_assembly_variant(ctx.handle, variant))
        for with_k in (False, True):
            times = []
            for it in range(4):
                ctx.check(lib.icqb_assemble_lower_dev(ctx.handle, 5, g.data_ptr(),
                                                      k.data_ptr() if with_k else 0,
                                                      ku.data_ptr() if with_k else 0))
                ctx.check(lib.icqb_synchronize(ctx.handle))
                if it > 0:
                    times.append(ctx.last_ms(1))
            key = 'variant{0}_{1}'.format(variant, 'G+K' if with_k else 'G')
            out['assembly_ms'][key] = round(float(numpy.mean(times)), 3)
        keep[variant] = (g.clone(), k.clone(), ku.clone())
        fast, allp = ctypes.c_int64(0), ctypes.c_int64(0)
        ctx.check(lib.icqb_assembly_stats(ctx.handle, ctypes.byref(fast), ctypes.byref(allp)))
        out['far_field_fraction'][str(variant)] = fast.value / max(allp.value, 1)
    out['max_rel_diff_variant1_vs_0_G'] = float(
        (torch.nan_to_num(keep[1][0] / keep[0][0], nan=1.0, posinf=1.0, neginf=1.0) - 1.0).abs().max())
    del keep
    out['assembly_Gevals_per_s'] = {key: round(pairs * 256 / (ms * 1e-3) / 1e9, 1)
                                    for key, ms in out['assembly_ms'].items()}
    ctx.check(lib.icqb_set_assembly_variant(ctx.handle, 0))
    ctx.check(lib.icqb_assemble_lower_dev(ctx.handle, 5, g.data_ptr(), 0, 0))
    x = torch.ones(n, dtype=torch.float64, device=dev)
    y = torch.empty(n, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > L2
    for name, fn, bytes_per in (('symv_f64', lib.icqb_symv_dev, 8), ('symv_f32', lib.icqb_symv_f32_dev, 4)):
        times = []
        for it in range(8):
            flush.zero_()
            torch.cuda.synchronize(dev)
            ctx.check(fn(ctx.handle, g.data_ptr(), x.data_ptr(), y.data_ptr(), 1))
            if it > 1:
                times.append(ctx.last_ms(3))
        ms = float(numpy.mean(times))
        out[name] = {'ms': round(ms, 4), 'GB_per_s': round(bytes_per * ntiles * 16384 / (ms * 1e-3) / 1e9, 1)}
    ctx.close()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
