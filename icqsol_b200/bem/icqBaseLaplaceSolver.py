"""Owner of the Green matrices, the drop-in boundary of the hot path.

Same role and accessors as the reference's ``BaseLaplaceSolver``
(bem/icqBaseLaplaceSolver.py:14-84): the constructor loads the shared library
(:26-27), allocates the response matrix (:29-30) and fills its diagonal and
off-diagonal terms (:74-77); ``getGreenMatrix`` returns it (:79-84).

Here the matrix lives in HBM (torch float64 buffers, sharded by observer rows over
the ranks when torch.distributed is initialised) and both kinds of terms are
computed by the CUDA library (include/icqb.h).  Two storages:
'lower' (default) keeps each pair entry once -- the upper part is the reference's
own mirror rule -- with row blocks g and 2P-1-g on rank g; 'full' keeps complete
rows [g*ceil(N/P), (g+1)*ceil(N/P)) on rank g.  ``getGreenMatrix()`` still
returns a ``numpy.float64[N,N]`` (materialised on demand); ``getGreenMatrixDevice()``
returns the resident local row block.  ``getNormalDerivativeGreenMatrix()`` is
the double-layer matrix K (not in the reference snapshot).
"""
import ctypes
import os
import time

import numpy

from icqsol_b200 import _lib
from icqsol_b200.bem.icqBaseSolver import BaseSolver
from icqsol_b200.bem import icqRowPartition as rowPartition

FOUR_PI = 4. * numpy.pi


def _requireTorchCuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('icqsol_b200 needs a CUDA device (B200, sm_100a): torch.cuda is not '
                           'available and there is no CPU fallback')
    return torch


def leadingDimension(n):
    """Row pitch of the device matrices: N rounded up to 16 doubles (128 B), so
    every row starts on a 128-byte boundary for the streaming GEMV loads."""
    return (int(n) + 15) // 16 * 16


class BaseLaplaceSolver(BaseSolver):

    # defaults of the solver options a caller leaves at None
    # blockTiles / blockOverlap: (one rank, several ranks) -- overlapping blocks are a
    # one-rank feature (include/icqb.h, icqb_cg_set_block_extents)
    DEFAULTS = {'preconditioner': 'block+caps', 'blockTiles': (4, 8), 'blockOverlap': (1, 0),
                'assemblyVariant': 1, 'mixedPrecision': 0.}

    def __init__(self, pdata, max_edge_length, order=5, computeK=False, device=None,
                 storage='lower', preconditioner=None, blockTiles=None, assemblyVariant=None,
                 directFallback=True, mixedPrecision=None, blockOverlap=None):
        """
        Constructor
        @param pdata polydata (vtkPolyData or stand-in)
        @param max_edge_length maximum edge length, used to turn
                               polygons into triangles
        @param order order of the self-term integration (1..5)
        @param computeK also assemble the double-layer matrix K
        @param device CUDA device index (default: LOCAL_RANK, else current device)
        @param storage 'lower': block-lower-triangular, each pair entry stored once, the
                       upper part given by the mirror rule G[j,i] = G[i,j] A_i/A_j
                       (bem/icqLaplaceMatrices.cpp:97-98); 'full': complete row blocks
        @param preconditioner of the CG solve: 'block+caps' (default) = inverse diagonal blocks
                       of ``blockTiles`` tiles of diag(A) G, with one dense block over every run
                       of tiles that hold sliver triangles (icqRowPartition.sliverBlockPartition;
                       on a UV sphere the two poles -- without them diag(A) G is indefinite
                       there and the iteration stalls); 'block' = inverse
                       128 x 128 blocks only; 'diagonal' = 1/diag (the reference CG class's
                       default, solvers/icqConjugateGradient.py:19); 'auto' = 'block+caps'
        @param blockTiles with 'block+caps': tiles (of 128 triangles) per diagonal block
                       between the sliver runs (config C2 on a B200: 165 iterations with 1,
                       91 with 8); default 4 on one rank (with blockOverlap 1), 8 on several
        @param blockOverlap with 'block+caps' on ONE rank: tiles by which every block is extended on
                       either side (symmetric additive Schwarz, icqRowPartition.blockExtents:
                       config C2 needs 45 iterations with blocks of 4 + 1 + 1 tiles instead of 91
                       with blocks of 8); default 1 on one rank; ignored on several ranks
        @param mixedPrecision with 'block+caps' and storage='lower' only: relative
                       residual below which the products of the CG stream a float32 copy of
                       the matrix (include/icqb.h, icqb_cg_set_mixed_precision); 0 = off
        @param directFallback when the CG stops short of the requested backward error, solve
                       by LU with partial pivoting instead (solveGreenSystemDirect) -- what the
                       reference does in the first place (bem/icqLaplaceSolver.py:36); without
                       it an unconverged solve raises RuntimeError
        @param assemblyVariant arithmetic of the off-diagonal kernel (include/icqb.h,
                       icqb_set_assembly_variant): 0 = direct (k_offdiag), 1 = far-field split
                       (k_offdiag_ff, the default: 9 % faster, entries within 5e-15 of 0)
        """
        # wall-clock marks of the host-side phases (milliseconds since construction began), for
        # the end-to-end breakdown of bench.py: hostTimings['constructor'] / ['solve']
        t_begin = time.perf_counter()
        self.hostTimings = {'constructor': [], 'solve': []}

        def mark(name):
            self.hostTimings['constructor'].append((name, 1e3 * (time.perf_counter() - t_begin)))

        defaults = dict(self.DEFAULTS)
        if preconditioner is None:
            preconditioner = defaults['preconditioner']
        rankInfo = rowPartition.worldInfo()
        single = 0 if rankInfo[1] == 1 else 1
        if blockOverlap is None:
            blockOverlap = defaults['blockOverlap'][single]
        if single == 1:
            blockOverlap = 0
        if blockTiles is None:
            blockTiles = defaults['blockTiles'][single] if blockOverlap > 0 else defaults['blockTiles'][1]
        if assemblyVariant is None:
            assemblyVariant = defaults['assemblyVariant']
        if mixedPrecision is None:
            mixedPrecision = defaults['mixedPrecision']
        if preconditioner == 'auto':
            preconditioner = 'block+caps'
        if preconditioner != 'block+caps' or storage != 'lower':
            mixedPrecision = 0.      # the float32 products exist in the block+caps solver only

        BaseSolver.__init__(self, pdata, max_edge_length, order)
        mark('mesh intake (host)')
        if storage not in ('lower', 'full'):
            raise ValueError("storage must be 'lower' or 'full'")
        if preconditioner not in ('block', 'diagonal', 'block+caps'):
            raise ValueError("preconditioner must be 'block', 'diagonal', 'block+caps' or 'auto'")
        self.storage = storage
        self.preconditioner = preconditioner
        self.directFallback = bool(directFallback)

        torch = _requireTorchCuda()
        if device is None:
            device = int(os.environ.get('LOCAL_RANK', torch.cuda.current_device()))
        self.device = int(device)
        torch.cuda.set_device(self.device)

        # the compiled library (raises if it was not built: no fallback)
        self.ctx = _lib.Context(self.device)
        self.lib = self.ctx.lib
        mark('context')

        # run the library on torch's current stream: buffer initialisation (torch) and the
        # kernels (library) are then ordered without extra synchronisation
        self.ctx.check(self.lib.icqb_set_stream(self.ctx.handle,
                                                torch.cuda.current_stream().cuda_stream or 1))

        self.assemblyVariant = int(assemblyVariant)
        self.ctx.check(self.lib.icqb_set_assembly_variant(self.ctx.handle, self.assemblyVariant))

        self.ctx.check(self.lib.icqb_cg_set_preconditioner(
            self.ctx.handle, {'diagonal': 0, 'block': 1, 'block+caps': 2}[preconditioner]))
        self.ctx.check(self.lib.icqb_cg_set_mixed_precision(self.ctx.handle, float(mixedPrecision)))
        self.rank, self.numRanks = rowPartition.worldInfo()
        self.capTiles = (0, 0)
        self.denseRuns = []
        if preconditioner == 'block+caps':
            # one dense block per run of tiles that hold slivers, wherever it lies in the list
            # (on a UV sphere: the head and tail caps of sliverCapTiles); the other blocks end
            # where the ranks' row blocks do, so that each of them lives on one rank
            flagged = rowPartition.sliverTiles(self._xyz, self._tri)
            self.capTiles = rowPartition.sliverCapTiles(self._xyz, self._tri, flagged=flagged)
            cuts = rowPartition.rankCuts(self.numTriangles, self.numRanks) \
                if (storage == 'lower' and self.numRanks > 1) else None
            starts, self.denseRuns = rowPartition.sliverBlockPartition(
                self._xyz, self._tri, blockTiles=int(blockTiles), cuts=cuts, flagged=flagged)
            self.blockStarts = [int(t) for t in starts]
            starts = numpy.ascontiguousarray(starts, numpy.int64)
            self.ctx.check(self.lib.icqb_cg_set_block_partition(
                self.ctx.handle, starts.ctypes.data_as(_lib.c_int64_p), len(starts)))
            self.blockExtents = None
            if int(blockOverlap) > 0 and self.numRanks == 1:
                numTiles = (self.numTriangles + rowPartition.TILE - 1) // rowPartition.TILE
                ext0, ext1 = rowPartition.blockExtents(self.blockStarts, numTiles, self.denseRuns,
                                                       overlapTiles=int(blockOverlap))
                self.blockExtents = (ext0, ext1)
                e0 = numpy.ascontiguousarray(ext0, numpy.int64)
                e1 = numpy.ascontiguousarray(ext1, numpy.int64)
                self.ctx.check(self.lib.icqb_cg_set_block_extents(
                    self.ctx.handle, e0.ctypes.data_as(_lib.c_int64_p),
                    e1.ctypes.data_as(_lib.c_int64_p), len(e0)))

        # one NCCL communicator per process, shared by every solver
        _lib.attachCommunicator(self.ctx, self.rank, self.numRanks, rowPartition.broadcastBytes)
        mark('options, communicator')

        n = self.numTriangles
        self.ctx.check(self.lib.icqb_set_mesh(
            self.ctx.handle, self._xyz.ctypes.data_as(_lib.c_double_p), self._xyz.shape[0],
            self._tri.ctypes.data_as(_lib.c_int64_p), n))
        mark('mesh upload, k_prepare')
        if storage == 'lower':
            self.rowBlocks = rowPartition.foldedBlocks(n, self.rank, self.numRanks)
            self.ctx.check(self.lib.icqb_set_row_blocks(self.ctx.handle, *self.rowBlocks))
        else:
            r0, r1 = rowPartition.rowRange(n, self.rank, self.numRanks)
            self.rowBlocks = (r0, r1, r1, r1)
            self.ctx.check(self.lib.icqb_set_rows(self.ctx.handle, r0, r1))
        self.rowBegin, self.rowEnd = self.rowBlocks[0], self.rowBlocks[1]
        self.numLocalRows = (self.rowBlocks[1] - self.rowBlocks[0]) + \
            (self.rowBlocks[3] - self.rowBlocks[2])
        mark('row blocks')

        self.ld = leadingDimension(n)
        dev = torch.device('cuda', self.device)
        self.kMatDevice = None
        self.kMatUpperDevice = None
        if storage == 'lower':
            # tile-contiguous lower triangle: [numTiles, 128, 128].  Not zero-filled (a 4N^2-byte
            # memset per matrix): the ragged edge tiles are only partly written, and every
            # reader masks rows and columns beyond N (k_symv, k_cg_blocks_lower,
            # k_dense_extract_lower, lowerTilesToRows) -- the parity tests run on NaN-filled tiles
            self.numLocalTiles = int(self.lib.icqb_lower_num_tiles(self.ctx.handle))
            shape = (max(self.numLocalTiles, 1), rowPartition.TILE, rowPartition.TILE)
            self.gMatDevice = torch.empty(shape, dtype=torch.float64, device=dev)
            if computeK:
                self.kMatDevice = torch.empty(shape, dtype=torch.float64, device=dev)
                self.kMatUpperDevice = torch.empty(shape, dtype=torch.float64, device=dev)
        else:
            rows = max(self.numLocalRows, 1)
            self.gMatDevice = torch.empty((rows, self.ld), dtype=torch.float64, device=dev)
            if computeK:
                self.kMatDevice = torch.empty((rows, self.ld), dtype=torch.float64, device=dev)
        self.gMat = None   # host copy, made on demand
        self.kMat = None
        mark('matrix buffers')

        self.__computeResponseMatrix()
        mark('assembly')

    def __computeResponseMatrix(self):
        # diagonal (self) terms and off-diagonal terms, one library call
        kPtr = self.kMatDevice.data_ptr() if self.kMatDevice is not None else 0
        if self.storage == 'lower':
            kuPtr = self.kMatUpperDevice.data_ptr() if self.kMatUpperDevice is not None else 0
            self.ctx.check(self.lib.icqb_assemble_lower_dev(
                self.ctx.handle, int(self.order), self.gMatDevice.data_ptr(), kPtr, kuPtr))
        else:
            self.ctx.check(self.lib.icqb_assemble_dev(self.ctx.handle, int(self.order),
                                                      self.gMatDevice.data_ptr(), kPtr, self.ld))
        self.ctx.check(self.lib.icqb_synchronize(self.ctx.handle))
        self.assemblyMilliseconds = {'diagonal': self.ctx.last_ms(2),
                                     'offDiagonal': self.ctx.last_ms(1),
                                     'prepare': self.ctx.last_ms(0)}

    def getTriangleAreas2(self):
        """|db x dc| of every triangle (twice the area), the weights of the mirror rule."""
        out = numpy.zeros(self.numTriangles, numpy.float64)
        self.ctx.check(self.lib.icqb_area2_host(self.ctx.handle, out.ctypes.data_as(_lib.c_double_p)))
        return out

    def __gatherStored(self, devMat):
        n = self.numTriangles
        if self.storage == 'lower':
            tiles = devMat[:self.numLocalTiles].cpu().numpy()
            local = rowPartition.lowerTilesToRows(tiles, self.rowBlocks, n)
            return rowPartition.gatherFoldedRows(local, n)
        local = devMat[:self.numLocalRows, :n].cpu().numpy()
        return rowPartition.allGatherMatrixRows(local, n)

    def __tileMask(self):
        """True where the lower storage holds an entry: tile(col) <= tile(row)."""
        t = numpy.arange(self.numTriangles) // rowPartition.TILE
        return t[None, :] <= t[:, None]

    def getGreenMatrix(self):
        """
        Return the Green function matrix
        @return matrix, numpy float64[N, N]
        """
        if self.gMat is None:
            g = self.__gatherStored(self.gMatDevice)
            if self.storage == 'lower':
                # upper part by the reference's mirror rule (icqLaplaceMatrices.cpp:97-98):
                # G[j,i] = G[i,j] * area_i / area_j
                stored = self.__tileMask()
                g = numpy.where(stored, g, 0.0)
                a2 = self.getTriangleAreas2()
                mirrored = (g * a2[:, None]).T / a2[:, None]
                g = numpy.where(stored, g, mirrored)
            self.gMat = g
        return self.gMat

    def getNormalDerivativeGreenMatrix(self):
        """
        Return the normal-derivative (double layer) Green matrix K
        @return matrix, numpy float64[N, N]
        """
        if self.kMatDevice is None:
            raise RuntimeError('ERROR: K was not assembled; construct the solver with computeK=True')
        if self.kMat is None:
            k = self.__gatherStored(self.kMatDevice)
            if self.storage == 'lower':
                stored = self.__tileMask()
                ku = self.__gatherStored(self.kMatUpperDevice)
                k = numpy.where(stored, k, numpy.where(stored, ku, 0.0).T)
            self.kMat = k
        return self.kMat

    def getGreenMatrixDevice(self):
        """The part of G resident in this rank's HBM (a torch float64 view).
        storage='full': rows [rowBegin, rowEnd) complete, float64[rows, N].
        storage='lower': float64[numLocalTiles, 128, 128], the tile-contiguous lower
        triangle of the rows in getRowBlocks(); entries of the edge tiles beyond row/column N
        are unspecified (icqRowPartition.lowerStrips gives the
        strip -> tile map; icqRowPartition.lowerTilesToRows rebuilds rows on the host)."""
        if self.storage == 'lower':
            return self.gMatDevice[:self.numLocalTiles]
        return self.gMatDevice[:self.numLocalRows, :self.numTriangles]

    def getNormalDerivativeGreenMatrixDevice(self):
        if self.kMatDevice is None:
            raise RuntimeError('ERROR: K was not assembled; construct the solver with computeK=True')
        if self.storage == 'lower':
            return self.kMatDevice[:self.numLocalTiles]
        return self.kMatDevice[:self.numLocalRows, :self.numTriangles]

    def getRowRange(self):
        """First block of observer rows [begin, end) owned by this rank."""
        return self.rowBegin, self.rowEnd

    def getRowBlocks(self):
        """(r0, r1, q0, q1): this rank stores observer rows [r0,r1) and [q0,q1)."""
        return self.rowBlocks

    # -- dense apply / solve on the resident matrix ---------------------------------

    def _toDevice(self, vec):
        torch = _requireTorchCuda()
        v = numpy.ascontiguousarray(vec, numpy.float64)
        return torch.from_numpy(v).to(self.gMatDevice.device)

    def applyGreenMatrix(self, src):
        """G . src (numpy.dot at bem/icqPoissonSolver.py:36) on the resident matrix.
        @param src float64[N] on the host
        @return float64[N] on the host"""
        torch = _requireTorchCuda()
        n = self.numTriangles
        x = self._toDevice(src)
        torch.cuda.current_stream().synchronize()
        if self.storage == 'lower':
            y = torch.empty(n, dtype=torch.float64, device=x.device)
            self.ctx.check(self.lib.icqb_symv_dev(self.ctx.handle, self.gMatDevice.data_ptr(),
                                                  x.data_ptr(), y.data_ptr(), 0))
            self.ctx.check(self.lib.icqb_synchronize(self.ctx.handle))
            return y.cpu().numpy()
        rows = self.rowEnd - self.rowBegin
        y = torch.empty(max(rows, 1), dtype=torch.float64, device=x.device)
        stream = self.lib.icqb_stream(self.ctx.handle)
        self.ctx.check(self.lib.icqb_gemv_dev(self.ctx.handle, self.gMatDevice.data_ptr(), rows, n,
                                              self.ld, x.data_ptr(), y.data_ptr(), 0, stream))
        self.ctx.check(self.lib.icqb_synchronize(self.ctx.handle))
        return rowPartition.allGatherRows(y[:rows].cpu().numpy(), n)

    def solveGreenSystem(self, rhs, tol=1.e-12, maxNumIters=None, x0=None):
        """Solve G x = rhs by CG on the area-symmetrised system
        diag(A) G x = diag(A) rhs  (G[j,i] = G[i,j] A_i/A_j, so diag(A) G is symmetric).
        Stops when max|rhs - G x| <= tol*max|rhs| (backward error in the max norm).
        @return x float64[N]; details in self.lastSolveInfo"""
        torch = _requireTorchCuda()
        t_begin = time.perf_counter()
        marks = self.hostTimings['solve'] = []
        n = self.numTriangles
        b = self._toDevice(rhs)
        x = self._toDevice(x0) if x0 is not None else torch.zeros(n, dtype=torch.float64,
                                                                  device=b.device)
        torch.cuda.current_stream().synchronize()
        marks.append(('vectors to device', 1e3 * (time.perf_counter() - t_begin)))
        iters = ctypes.c_int64(0)
        r2 = ctypes.c_double(0.)
        rinf = ctypes.c_double(0.)
        maxit = -1 if maxNumIters is None else int(maxNumIters)
        self.ctx.check(self.lib.icqb_cg_solve_dev(
            self.ctx.handle, self.gMatDevice.data_ptr(), n, self.ld,
            1 if self.storage == 'lower' else 0, self.lib.icqb_area2_dev(self.ctx.handle), 0,
            b.data_ptr(), x.data_ptr(), float(tol), 2, maxit, ctypes.byref(iters),
            ctypes.byref(r2), ctypes.byref(rinf)))
        marks.append(('icqb_cg_solve_dev', 1e3 * (time.perf_counter() - t_begin)))
        self.lastSolveInfo = {'iterations': iters.value, 'residual2': r2.value,
                              'residualInf': rinf.value, 'milliseconds': self.ctx.last_ms(4),
                              'gemvMilliseconds': self.ctx.last_ms(3)}
        self.lastSolveInfo['residualReplacements'], self.lastSolveInfo['products'] = \
            self.ctx.cg_stats()
        n32 = ctypes.c_int64(0)
        self.ctx.check(self.lib.icqb_cg_last_mixed(self.ctx.handle, ctypes.byref(n32)))
        self.lastSolveInfo['float32Products'] = n32.value
        nbad, nweak = ctypes.c_int64(0), ctypes.c_int64(0)
        self.ctx.check(self.lib.icqb_cg_last_blocks(self.ctx.handle, ctypes.byref(nbad), ctypes.byref(nweak)))
        self.lastSolveInfo['preconditionerBlocksReplaced'] = nbad.value
        self.lastSolveInfo['preconditionerWeakPivots'] = nweak.value
        # the reference's LU (bem/icqLaplaceSolver.py:36) cannot "not converge": when the
        # iteration stopped short of the requested backward error (iteration cap, attainable
        # accuracy, or a matrix the quadrature left too indefinite for CG) the system is solved
        # by LU instead, or the call raises -- an unconverged iterate is never returned
        bmax = float(numpy.abs(numpy.asarray(rhs, numpy.float64)).max()) if n else 0.
        reached = rinf.value <= float(tol) * bmax or bmax == 0.
        self.lastSolveInfo['converged'] = bool(reached)
        self.lastSolveInfo['method'] = 'cg'
        if not reached:
            import warnings
            msg = ('icqsol_b200: CG stopped after {0} iterations with max|rhs - G x| = {1:.3e} '
                   '(requested {2:.3e})'.format(iters.value, rinf.value, float(tol) * bmax))
            if not self.directFallback:
                raise RuntimeError(msg + '; directFallback=False, no solution returned')
            warnings.warn(msg + '; solving by LU instead (directFallback)', RuntimeWarning)
            cgInfo = dict(self.lastSolveInfo)
            xd = self.solveGreenSystemDirect(rhs)
            self.lastSolveInfo['cg'] = cgInfo
            return xd
        return x.cpu().numpy()

    def solveGreenSystemDirect(self, rhs, tol=5.e-13, maxRefinements=3):
        """Solve G x = rhs by LU with partial pivoting -- the reference's own solve
        (numpy.linalg.solve, bem/icqLaplaceSolver.py:36) -- on the GPU(s), for the matrices the
        reference's quadrature leaves too indefinite for a CG and as
        the fallback of solveGreenSystem.  The factorisation is this package's own
        (csrc/lu.cu: column blocks of 128 dealt to the ranks, panel factorisation, FP64 tile
        kernel for the updates).  What it factors is S = diag(A) G (symmetric): complete rows
        of G are assembled for the purpose by the same kernels (k_diag, k_offdiag) into a
        separate buffer -- 8 N^2 / P bytes per rank -- and are the columns of S once scaled.
        The solution is checked against the RESIDENT matrix with the product kernel and
        refined until max|rhs - G x| <= tol max|rhs| (at most ``maxRefinements`` times).
        @return x float64[N]; details in self.lastSolveInfo"""
        torch = _requireTorchCuda()
        n = self.numTriangles
        if n == 0:
            return numpy.zeros(0, numpy.float64)
        lib, ctx = self.lib, self.ctx
        b = self._toDevice(rhs)
        dev = b.device
        P, rank = self.numRanks, self.rank
        ncols = int(lib.icqb_lu_local_columns(ctx.handle, n))
        ld = self.ld
        t_wall = time.perf_counter()
        w = torch.empty((max(ncols, 1), ld), dtype=torch.float64, device=dev)
        # rows of G into the local column blocks: local block lb <-> global block lb * P + rank
        nblk = (n + rowPartition.TILE - 1) // rowPartition.TILE
        try:
            if P == 1:
                ctx.check(lib.icqb_set_rows(ctx.handle, 0, n))
                ctx.check(lib.icqb_assemble_dev(ctx.handle, int(self.order), w.data_ptr(), 0, ld))
            else:
                for lb, kb in enumerate(range(rank, nblk, P)):
                    r0 = kb * rowPartition.TILE
                    r1 = min(n, r0 + rowPartition.TILE)
                    ctx.check(lib.icqb_set_rows(ctx.handle, r0, r1))
                    ctx.check(lib.icqb_assemble_dev(ctx.handle, int(self.order),
                                                    w[lb * rowPartition.TILE].data_ptr(), 0, ld))
        finally:
            if self.storage == 'lower':
                ctx.check(lib.icqb_set_row_blocks(ctx.handle, *self.rowBlocks))
            else:
                ctx.check(lib.icqb_set_rows(ctx.handle, self.rowBlocks[0], self.rowBlocks[1]))
        assembly_ms = 1e3 * (time.perf_counter() - t_wall)
        area2 = lib.icqb_area2_dev(ctx.handle)
        info = ctypes.c_int(0)
        ctx.check(lib.icqb_lu_factor_dev(ctx.handle, w.data_ptr(), n, ld, area2, ctypes.byref(info)))
        factor_ms = ctx.last_ms(5)
        if info.value != 0:
            raise RuntimeError('solveGreenSystemDirect: the matrix is singular (zero pivot at row '
                               '{0})'.format(info.value))
        x = torch.empty(n, dtype=torch.float64, device=dev)
        ctx.check(lib.icqb_lu_solve_dev(ctx.handle, w.data_ptr(), n, ld, area2, b.data_ptr(), x.data_ptr()))
        solve_ms = ctx.last_ms(6)
        # residual with the resident matrix (the validated product), refinement on the factors
        bmax = float(b.abs().max().item())
        r = torch.empty(n, dtype=torch.float64, device=dev)
        dx = torch.empty(n, dtype=torch.float64, device=dev)
        history = []
        for sweep in range(int(maxRefinements) + 1):
            self._applyDevice(x, r)                       # r = G x
            torch.sub(b, r, out=r)
            rinf = float(r.abs().max().item())
            history.append(rinf)
            if rinf <= float(tol) * bmax or sweep == int(maxRefinements):
                break
            ctx.check(lib.icqb_lu_solve_dev(ctx.handle, w.data_ptr(), n, ld, area2, r.data_ptr(),
                                            dx.data_ptr()))
            solve_ms += ctx.last_ms(6)
            x += dx
        del w
        self.lastSolveInfo = {'iterations': 0, 'residual2': float('nan'), 'residualInf': history[-1],
                              'milliseconds': assembly_ms + factor_ms + solve_ms,
                              'luAssemblyMilliseconds': assembly_ms,
                              'luFactorMilliseconds': factor_ms, 'luSolveMilliseconds': solve_ms,
                              'gemvMilliseconds': 0., 'residualReplacements': 0, 'products': len(history),
                              'refinements': len(history) - 1, 'residualHistory': history,
                              'method': 'lu',
                              'converged': bool(history[-1] <= float(tol) * bmax or bmax == 0.)}
        return x.cpu().numpy()

    def _applyDevice(self, x, y):
        """y = G x on device vectors (replicated on every rank), resident matrix."""
        torch = _requireTorchCuda()
        n = self.numTriangles
        torch.cuda.current_stream().synchronize()
        if self.storage == 'lower':
            self.ctx.check(self.lib.icqb_symv_dev(self.ctx.handle, self.gMatDevice.data_ptr(),
                                                  x.data_ptr(), y.data_ptr(), 0))
            self.ctx.check(self.lib.icqb_synchronize(self.ctx.handle))
            return
        rows = self.rowEnd - self.rowBegin
        yl = torch.empty(max(rows, 1), dtype=torch.float64, device=x.device)
        stream = self.lib.icqb_stream(self.ctx.handle)
        self.ctx.check(self.lib.icqb_gemv_dev(self.ctx.handle, self.gMatDevice.data_ptr(), rows, n,
                                              self.ld, x.data_ptr(), yl.data_ptr(), 0, stream))
        self.ctx.check(self.lib.icqb_synchronize(self.ctx.handle))
        if self.numRanks == 1:
            y.copy_(yl[:n])
        else:
            y.copy_(torch.from_numpy(rowPartition.allGatherRows(yl[:rows].cpu().numpy(), n)))
