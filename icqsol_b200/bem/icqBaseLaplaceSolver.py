"""Owner of the Green matrices, the drop-in boundary of the hot path.

Same role and accessors as the reference's ``BaseLaplaceSolver``
(bem/icqBaseLaplaceSolver.py:14-84): the constructor loads the shared library
(:26-27), allocates the response matrix (:29-30) and fills its diagonal and
off-diagonal terms (:74-77); ``getGreenMatrix`` returns it (:79-84).

Here the matrix lives in HBM (a torch float64 buffer, row-block sharded when
torch.distributed is initialised: rank g owns observer rows
[g*ceil(N/P), (g+1)*ceil(N/P)) x all N columns) and both kinds of terms are
computed by the CUDA library (include/icqb.h).  ``getGreenMatrix()`` still
returns a ``numpy.float64[N,N]`` (materialised on demand); ``getGreenMatrixDevice()``
returns the resident local row block.  ``getNormalDerivativeGreenMatrix()`` is
the double-layer matrix K (not in the reference snapshot).
"""
import ctypes
import os

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

    def __init__(self, pdata, max_edge_length, order=5, computeK=False, device=None):
        """
        Constructor
        @param pdata polydata (vtkPolyData or stand-in)
        @param max_edge_length maximum edge length, used to turn
                               polygons into triangles
        @param order order of the self-term integration (1..5)
        @param computeK also assemble the double-layer matrix K
        @param device CUDA device index (default: LOCAL_RANK, else current device)
        """
        BaseSolver.__init__(self, pdata, max_edge_length, order)

        torch = _requireTorchCuda()
        if device is None:
            device = int(os.environ.get('LOCAL_RANK', torch.cuda.current_device()))
        self.device = int(device)
        torch.cuda.set_device(self.device)

        # the compiled library (raises if it was not built: no fallback)
        self.ctx = _lib.Context(self.device)
        self.lib = self.ctx.lib

        self.rank, self.numRanks = rowPartition.worldInfo()
        # one NCCL communicator per process, shared by every solver
        _lib.attachCommunicator(self.ctx, self.rank, self.numRanks, rowPartition.broadcastBytes)

        n = self.numTriangles
        self.ctx.check(self.lib.icqb_set_mesh(
            self.ctx.handle, self._xyz.ctypes.data_as(_lib.c_double_p), self._xyz.shape[0],
            self._tri.ctypes.data_as(_lib.c_int64_p), n))
        self.rowBegin, self.rowEnd = rowPartition.rowRange(n, self.rank, self.numRanks)
        self.ctx.check(self.lib.icqb_set_rows(self.ctx.handle, self.rowBegin, self.rowEnd))

        self.ld = leadingDimension(n)
        rows = max(self.rowEnd - self.rowBegin, 1)
        dev = torch.device('cuda', self.device)
        self.gMatDevice = torch.empty((rows, self.ld), dtype=torch.float64, device=dev)
        self.kMatDevice = torch.empty((rows, self.ld), dtype=torch.float64, device=dev) \
            if computeK else None
        self.gMat = None   # host copy, made on demand
        self.kMat = None

        self.__computeResponseMatrix()

    def __computeResponseMatrix(self):
        # diagonal (self) terms and off-diagonal terms, one library call
        kPtr = self.kMatDevice.data_ptr() if self.kMatDevice is not None else 0
        self.ctx.check(self.lib.icqb_assemble_dev(self.ctx.handle, int(self.order),
                                                  self.gMatDevice.data_ptr(), kPtr, self.ld))
        self.ctx.check(self.lib.icqb_synchronize(self.ctx.handle))
        self.assemblyMilliseconds = {'diagonal': self.ctx.last_ms(2),
                                     'offDiagonal': self.ctx.last_ms(1),
                                     'prepare': self.ctx.last_ms(0)}

    def __toHost(self, devMat):
        n = self.numTriangles
        rows = self.rowEnd - self.rowBegin
        local = devMat[:rows, :n].cpu().numpy()
        return rowPartition.allGatherMatrixRows(local, n)

    def getGreenMatrix(self):
        """
        Return the Green function matrix
        @return matrix, numpy float64[N, N]
        """
        if self.gMat is None:
            self.gMat = self.__toHost(self.gMatDevice)
        return self.gMat

    def getNormalDerivativeGreenMatrix(self):
        """
        Return the normal-derivative (double layer) Green matrix K
        @return matrix, numpy float64[N, N]
        """
        if self.kMatDevice is None:
            raise RuntimeError('ERROR: K was not assembled; construct the solver with computeK=True')
        if self.kMat is None:
            self.kMat = self.__toHost(self.kMatDevice)
        return self.kMat

    def getGreenMatrixDevice(self):
        """Local row block of G in HBM: torch float64[rowEnd-rowBegin, N] (a view)."""
        return self.gMatDevice[:self.rowEnd - self.rowBegin, :self.numTriangles]

    def getNormalDerivativeGreenMatrixDevice(self):
        if self.kMatDevice is None:
            raise RuntimeError('ERROR: K was not assembled; construct the solver with computeK=True')
        return self.kMatDevice[:self.rowEnd - self.rowBegin, :self.numTriangles]

    def getRowRange(self):
        """Observer rows [begin, end) owned by this rank."""
        return self.rowBegin, self.rowEnd

    # -- dense apply / solve on the resident matrix ---------------------------------

    def _localSlice(self, vec):
        torch = _requireTorchCuda()
        v = numpy.ascontiguousarray(vec, numpy.float64)
        return torch.from_numpy(v[self.rowBegin:self.rowEnd].copy()).to(self.gMatDevice.device)

    def applyGreenMatrix(self, src):
        """G . src with the GEMV kernel (numpy.dot at bem/icqPoissonSolver.py:36).
        @param src float64[N] on the host
        @return float64[N] on the host"""
        torch = _requireTorchCuda()
        n = self.numTriangles
        rows = self.rowEnd - self.rowBegin
        x = torch.from_numpy(numpy.ascontiguousarray(src, numpy.float64)).to(self.gMatDevice.device)
        y = torch.empty(max(rows, 1), dtype=torch.float64, device=self.gMatDevice.device)
        stream = self.lib.icqb_stream(self.ctx.handle)
        self.ctx.check(self.lib.icqb_synchronize(self.ctx.handle))
        torch.cuda.current_stream().synchronize()
        self.ctx.check(self.lib.icqb_gemv_dev(self.ctx.handle, self.gMatDevice.data_ptr(), rows, n,
                                              self.ld, x.data_ptr(), y.data_ptr(), 0, stream))
        self.ctx.check(self.lib.icqb_synchronize(self.ctx.handle))
        return rowPartition.allGatherRows(y[:rows].cpu().numpy(), n)

    def solveGreenSystem(self, rhs, tol=1.e-13, maxNumIters=None, x0=None):
        """Solve G x = rhs by CG on the area-symmetrised system
        diag(A) G x = diag(A) rhs  (G[j,i] = G[i,j] A_i/A_j, so diag(A) G is symmetric).
        Stops when ||rhs - G x||_2 <= tol*||rhs||_2.
        @return x float64[N]; details in self.lastSolveInfo"""
        torch = _requireTorchCuda()
        n = self.numTriangles
        rows = self.rowEnd - self.rowBegin
        dev = self.gMatDevice.device
        b = self._localSlice(rhs)
        x = self._localSlice(x0) if x0 is not None else torch.zeros(rows, dtype=torch.float64,
                                                                    device=dev)
        if rows == 0:
            b = torch.zeros(1, dtype=torch.float64, device=dev)
            x = torch.zeros(1, dtype=torch.float64, device=dev)
        torch.cuda.current_stream().synchronize()
        area2 = self.lib.icqb_area2_dev(self.ctx.handle) + 8 * self.rowBegin
        iters = ctypes.c_int64(0)
        r2 = ctypes.c_double(0.)
        rinf = ctypes.c_double(0.)
        maxit = -1 if maxNumIters is None else int(maxNumIters)
        self.ctx.check(self.lib.icqb_cg_solve_dev(
            self.ctx.handle, self.gMatDevice.data_ptr(), n, self.ld, area2, 0, b.data_ptr(),
            x.data_ptr(), float(tol), 1, maxit, ctypes.byref(iters), ctypes.byref(r2),
            ctypes.byref(rinf)))
        self.lastSolveInfo = {'iterations': iters.value, 'residual2': r2.value,
                              'residualInf': rinf.value, 'milliseconds': self.ctx.last_ms(4),
                              'gemvMilliseconds': self.ctx.last_ms(3)}
        return rowPartition.allGatherRows(x[:rows].cpu().numpy(), n)
