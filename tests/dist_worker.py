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

    # ICQB_EXPERIMENTAL=1 runs every case with the experimental solver as well (dense cap blocks;
    # with ICQB_PEER_EXCHANGE=1 also the peer exchange instead of the NCCL all-reduce)
    kinds = ['block'] + (['block+caps'] if os.environ.get('ICQB_EXPERIMENTAL') == '1' else [])
    cases = [(nt, npz, storage, kind)
             for (nt, npz, storage) in ((40, 20, 'full'), (40, 20, 'lower'), (64, 30, 'lower'),
                                        (64, 30, 'full'), (7, 4, 'full'), (7, 4, 'lower'))
             for kind in kinds]
    for nt, npz, storage, kind in cases:
        xyz, tri = uvSphere(1.0, nt, npz)
        n = tri.shape[0]
        print('rank', rank, 'case', nt, npz, storage, flush=True)
        solver = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5, computeK=True,
                               storage=storage, preconditioner=kind)
        if kind == 'block+caps' and n > 512:
            # no slivers on these meshes: force caps so that the cap machinery runs
            solver.ctx.check(solver.lib.icqb_cg_set_cap_tiles(solver.ctx.handle, 2, 1))
        assert solver.numRanks == world
        if storage == 'full':
            r0, r1 = solver.getRowRange()
            gloc = solver.getGreenMatrixDevice().cpu().numpy()
            ref = oracle.block(xyz, tri, r0, r1, 0, n, 5)
            if r1 > r0:
                assert numpy.abs(gloc / ref - 1).max() < 1e-12, 'G rows'
                kref = oracle.k_block(xyz, tri, r0, r1, 0, n)
                kloc = solver.getNormalDerivativeGreenMatrixDevice().cpu().numpy()
                assert numpy.abs(kloc - kref).max() / numpy.abs(kref).max() < 1e-9, 'K rows'
        else:
            kfull = solver.getNormalDerivativeGreenMatrix()
            kref = oracle.k_block(xyz, tri, 0, n, 0, n)
            assert numpy.abs(kfull - kref).max() / numpy.abs(kref).max() < 1e-9, 'K lower'
        g = solver.getGreenMatrix()          # gathered on every rank
        gref = oracle.green_matrix(xyz, tri, 5, threads=True)
        assert numpy.abs(g / gref - 1).max() < 1e-12, 'G gathered'
        solver.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
        v = solver.getSourceArray(solver.getSourceArrayIndex())
        rsp = solver.computeResponseField()
        backward = numpy.abs(v + gref.dot(rsp)).max() / numpy.abs(v).max()
        assert backward < 1e-12, backward
        lu = oracle.laplace_response(gref, v)
        assert numpy.abs(rsp - lu).max() / numpy.abs(lu).max() < 1e-8
        info = solver.lastSolveInfo
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
    # the weak-scaling bench workload at full size (N^2 / GPU of config C2): the oracle cannot
    # build this matrix, so columns are taken as device products G e_j and compared with oracle
    # columns, and the response is checked against oracle rows (tests/test_gpu_parity.py,
    # _sampled_checks); every rank checks different columns
    nt = int(os.environ.get('ICQB_WORKER_NT', int(round(142 * world ** 0.25 / 2.0)) * 2))   # (override: dry runs)
    xyz, tri = uvSphere(1.0, nt, nt // 2)
    n = tri.shape[0]
    print('rank', rank, 'full-size case', nt, nt // 2, n, flush=True)
    solver = LaplaceSolver(PolyData.from_arrays(xyz, tri), float('inf'), 5)
    solver.setSourceFromExpression('1./sqrt(x**2 + (y-2.)**2 + (z-4.)**2)')
    v = solver.getSourceArray(solver.getSourceArrayIndex())
    rsp = solver.computeResponseField()
    assert solver.lastSolveInfo['converged'], solver.lastSolveInfo
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
        print('N={0} ranks={1} full size: CG iters {2}, backward {3:.2e}'.format(
            n, world, solver.lastSolveInfo['iterations'],
            solver.lastSolveInfo['residualInf'] / numpy.abs(v).max()))
    del solver
    dist.barrier()
    dist.destroy_process_group()
    print('rank {0} ok'.format(rank))


if __name__ == '__main__':
    main()
