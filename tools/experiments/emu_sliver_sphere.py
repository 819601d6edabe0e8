#!/usr/bin/env python
"""The experimental solver END TO END THROUGH THE CUDA SOURCES (CPU emulation, tools/emu) on a
sliver sphere: the oracle's G packed into the tile-contiguous lower storage, then
icqb_cg_solve_dev with preconditioner kind 2 and the partition of sliverBlockPartition --
k_dense_extract_lower, the unpivoted block Gauss-Jordan on the indefinite caps, k_dense_apply,
k_cg_vec / k_cg_close and the persistent product, exactly the code a GPU will run.
usage: emu_sliver_sphere.py <n_theta> <n_phi> [caps|plain] [block_tiles] [max_iters] [mixed_switch]
318 20 caps: 12,084 triangles, caps of 8 and 9 tiles, converged in 110 iterations (numpy model 109),
oracle backward error 3.2e-13; 16 minutes on one core."""
import ctypes, os, sys, time, numpy
ROOT='/root/repo'; sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/tools/emu')
os.environ['ICQB_EMU_SMS'] = '8'
import build as emu_build
from icqsol_b200 import _lib
from icqsol_b200.bem import icqRowPartition as rp
from oracle import oracle
oracle.build()
lib = ctypes.CDLL(emu_build.build())
for name, restype, argtypes in _lib.SIGNATURES:
    fn = getattr(lib, name); fn.restype = restype; fn.argtypes = argtypes
from icqsol_b200.mesh.generators import uvSphere
def dp(a): return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
def lp(a): return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
nt, nph = int(sys.argv[1]), int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else 'caps'
xyz, tri = uvSphere(1.0, nt, nph)
xyz = numpy.ascontiguousarray(xyz); tri = numpy.ascontiguousarray(tri); n = len(tri)
t0 = time.time()
g = oracle.green_matrix(xyz, tri, 5, threads=True)
print('N', n, 'oracle G %.0f s' % (time.time()-t0), flush=True)
# pack the lower block triangle into the tile-contiguous layout of one rank
nT = (n + 127)//128
strips, ntiles = rp.lowerStrips((0, n, n, n))
tiles = numpy.zeros((ntiles, 128, 128))
for (row0, nrows, I, base) in strips:
    for J in range(I + 1):
        c1 = min(n, (J+1)*128)
        tiles[base + J, :nrows, :c1 - J*128] = g[row0:row0+nrows, J*128:c1]
h = ctypes.c_void_p(); assert lib.icqb_create(ctypes.byref(h), 0) == 0
assert lib.icqb_set_mesh(h, dp(xyz), len(xyz), lp(tri), n) == 0
assert lib.icqb_set_row_blocks(h, 0, n, n, n) == 0
assert lib.icqb_lower_num_tiles(h) == ntiles
cen = (xyz[tri[:,0]]+xyz[tri[:,1]]+xyz[tri[:,2]])/3.
b = numpy.ascontiguousarray(-1./numpy.sqrt(cen[:,0]**2+(cen[:,1]-2.)**2+(cen[:,2]-4.)**2))
x = numpy.zeros(n)
starts, runs = rp.sliverBlockPartition(xyz, tri, blockTiles=int(sys.argv[4]) if len(sys.argv) > 4 else 1)
print('caps', rp.sliverCapTiles(xyz, tri), 'runs', runs, flush=True)
if mode == 'plain':
    assert lib.icqb_cg_set_preconditioner(h, 1) == 0
else:
    assert lib.icqb_cg_set_preconditioner(h, 2) == 0
    st = numpy.asarray(starts, numpy.int64)
    assert lib.icqb_cg_set_block_partition(h, lp(st), len(st)) == 0
mixed = float(sys.argv[6]) if len(sys.argv) > 6 else 0.
assert lib.icqb_cg_set_mixed_precision(h, mixed) == 0
it, r2, ri = ctypes.c_int64(0), ctypes.c_double(0.), ctypes.c_double(0.)
t0 = time.time()
st = lib.icqb_cg_solve_dev(h, tiles.ctypes.data, n, 0, 1, lib.icqb_area2_dev(h), 0, b.ctypes.data, x.ctypes.data,
                           5e-13, 2, int(sys.argv[5]) if len(sys.argv) > 5 else 400, ctypes.byref(it), ctypes.byref(r2), ctypes.byref(ri))
print('status', st, lib.icqb_last_error(h), 'iters', it.value, 'resid inf rel %.2e' % (ri.value/numpy.abs(b).max()), 'solve %.0f s' % (time.time()-t0), flush=True)
n32 = ctypes.c_int64(0)
lib.icqb_cg_last_mixed(h, ctypes.byref(n32))
print('float32 products', n32.value)
print('oracle backward error %.2e' % (numpy.abs(b - g.dot(x)).max()/numpy.abs(b).max()))
