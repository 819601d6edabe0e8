"""Multi-rank parity worker, launched by tests/test_gpu_multi.py under torchrun
(one rank per GPU, NCCL).  Row-sharded assembly + NCCL CG vs the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    local = int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    import datetime
    dist.init_process_group('nccl', device_id=torch.device('cuda', local),
                            timeout=datetime.timedelta(seconds=90))
    rank, world = dist.get_rank(), dist.get_world_size()
    from oracle import oracle
    from icqsol_b200.mesh.generators import uvSphere
    from icqsol_b200.mesh.polydata import PolyData
    from icqsol_b200.bem.icqLaplaceSolver import LaplaceSolver
    from icqsol_b200.bem.icqPoissonSolver import PoissonSolver

    # every case with both solvers: 128 x 128 blocks (cg.cu) and blocks + dense sliver blocks
    # (cg_caps.cu, the default; with ICQB_PEER_EXCHANGE=1 its product exchange goes through peer
    # memory instead of the NCCL all-reduce)
    kinds = ['block', 'block+caps']
    cases = [(nt, npz, storage, kind)
             for (nt, npz, storage) in ((40, 20, 'full'), (40, 20, 'lower'), (64, 30, 'lower'),
                                        (64, 30, 'full'), (7, 4, 'full'), (7, 4, 'lower'))
             for kind in kinds]
    refs = {}     # oracle matrices per mesh: built once, used by every storage / solver kind
    if os.environ.get('ICQB_WORKER_ONLY_FULL') == '1':
        cases = []
    for nt, npz, storage, kind in cases:
        xyz, tri = uvSphere(1.0, nt, npz)
        n = tri.shape[0]
        print('rank', rank, 'case', nt, npz, storage, kind, flush=True)
        if (nt, npz) not in refs:
            refs[(nt, npz)] = (oracle.green_matrix(xyz, tri, 5, threads=True),
                               oracle.k_block(xyz, tri, 0, n, 0, n))
        gref, kref = refs[(nt, npz)]
        solver = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5, computeK=True,
                               storage=storage, preconditioner=kind, blockTiles=2)
        assert solver.numRanks == world
        if storage == 'full':
            r0, r1 = solver.getRowRange()
            gloc = solver.getGreenMatrixDevice().cpu().numpy()
            if r1 > r0:
                assert numpy.abs(gloc / gref[r0:r1] - 1).max() < 1e-12, 'G rows'
                kloc = solver.getNormalDerivativeGreenMatrixDevice().cpu().numpy()
                assert numpy.abs(kloc - kref[r0:r1]).max() / numpy.abs(kref).max() < 1e-9, 'K rows'
        else:
            kfull = solver.getNormalDerivativeGreenMatrix()
            assert numpy.abs(kfull - kref).max() / numpy.abs(kref).max() < 1e-9, 'K lower'
        g = solver.getGreenMatrix()          # gathered on every rank
        assert numpy.abs(g / gref - 1).max() < 1e-12, 'G gathered'
        solver.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
        v = solver.getSourceArray(solver.getSourceArrayIndex())
        rsp = solver.computeResponseField()
        assert solver.lastSolveInfo['method'] == 'cg', solver.lastSolveInfo
        backward = numpy.abs(v + gref.dot(rsp)).max() / numpy.abs(v).max()
        assert backward < 1e-12, backward
        lu = oracle.laplace_response(gref, v)
        assert numpy.abs(rsp - lu).max() / numpy.abs(lu).max() < 1e-8
        info = solver.lastSolveInfo
        if kind == 'block+caps' and (nt, npz) != (40, 20):
            # the direct solve (csrc/lu.cu): column blocks dealt to the ranks, broadcast panels
            xlu = solver.solveGreenSystemDirect(v)
            assert solver.lastSolveInfo['method'] == 'lu' and solver.lastSolveInfo['converged']
            assert numpy.abs(v - gref.dot(xlu)).max() < 1e-12 * numpy.abs(v).max(), 'LU backward error'
            assert numpy.abs(-xlu - lu).max() / numpy.abs(lu).max() < 1e-8, 'LU vs numpy.linalg.solve'
        psolver = PoissonSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5, storage=storage)
        psolver.setSourceFromExpression('x + 1.0')
        prsp = psolver.computeResponseField()
        cen = (xyz[tri[:, 0]] + xyz[tri[:, 1]] + xyz[tri[:, 2]]) / 3.
        pref = gref.dot(cen[:, 0] + 1.0)
        assert numpy.abs(prsp - pref).max() / numpy.abs(pref).max() < 1e-12
        if rank == 0:
            print('N={0} ranks={1} {4} {5}: CG iters {2}, backward {3:.2e}'.format(
                n, world, info['iterations'], backward, storage, kind))
        del solver, psolver
    # full-size cases: the oracle cannot build these matrices, so columns are taken as device
    # products G e_j and compared with oracle columns, and the response is checked against
    # oracle rows (tests/test_gpu_parity.py, _sampled_checks); every rank checks different
    # columns and rows.  (1) the weak-scaling sphere (N^2 / GPU of config C2), (2) config C3's
    # own geometry: the refined bolt, 84,600 triangles -- bench.py's multi-GPU workload.
    from icqsol_b200.mesh.generators import subdivide
    nt = int(os.environ.get('ICQB_WORKER_NT', int(round(142 * world ** 0.25 / 2.0)) * 2))   # (override: dry runs)
    full = [('sphere', uvSphere(1.0, nt, nt // 2))]
    kbolt = int(os.environ.get('ICQB_WORKER_BOLT_K', 6))
    if kbolt > 0:
        m = numpy.load(os.path.join(ROOT, 'tests', 'golden', 'meshes.npz'))
        full.append(('bolt k={0}'.format(kbolt), subdivide(m['boltFine_xyz'], m['boltFine_tri'], kbolt)))
    # ICQB_WORKER_SPHERES="318x159,448x224": also the UV spheres of the C3 size (100,488
    # triangles) and of config C5 (199,808; eight GPUs) -- the meshes whose pole slivers make
    # diag(A) G indefinite
    for spec in filter(None, os.environ.get('ICQB_WORKER_SPHERES', '').split(',')):
        nth, nph = (int(v) for v in spec.split('x'))
        full.append(('sphere {0}x{1}'.format(nth, nph), uvSphere(1.0, nth, nph)))
    if os.environ.get('ICQB_WORKER_ONLY_FULL') == '1':
        full = full[2:]
    for name, (xyz, tri) in full:
        n = tri.shape[0]
        print('rank', rank, 'full-size case', name, n, flush=True)
        solver = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5)
        solver.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
        v = solver.getSourceArray(solver.getSourceArrayIndex())
        rsp = solver.computeResponseField()
        assert solver.lastSolveInfo['converged'] and solver.lastSolveInfo['method'] == 'cg', solver.lastSolveInfo
        cols = [0, n // 3 + 17, n - 1]
        for j in cols:                       # collective calls: same columns on every rank
            e = numpy.zeros(n)
            e[j] = 1.0
            col = solver.applyGreenMatrix(e)
            if cols.index(j) == rank % len(cols):
                ref = oracle.block(xyz, tri, 0, n, j, j + 1, 5)[:, 0]
                assert numpy.abs(col / ref - 1.).max() < 1e-12, ('column', j)
        for i in (rank * (n // world), min(n - 1, rank * (n // world) + 4321 % max(n // world, 1))):
            grow = oracle.block(xyz, tri, i, i + 1, 0, n, 5)[0]
            assert abs(v[i] + grow.dot(rsp)) < 1e-12 * numpy.abs(v).max(), ('row', i)
        if rank == 0:
            print('N={0} ranks={1} full size ({4}): CG iters {2}, backward {3:.2e}'.format(
                n, world, solver.lastSolveInfo['iterations'],
                solver.lastSolveInfo['residualInf'] / numpy.abs(v).max(), name))
        del solver
    dist.barrier()
    dist.destroy_process_group()
    print('rank {0} ok'.format(rank))


if __name__ == '__main__':
    main()
