"""Dense conjugate gradient on the GPU.

Same class and methods as the reference's ``ConjugateGradient``
(solvers/icqConjugateGradient.py:5-87): ``ConjugateGradient(mat, b)``,
``setTolerance``, ``setMaxNumberOfIterations``, ``setVerbosity``,
``setDiagonalPreconditioner``, ``solve(x0) -> (x, err, k)``,
``getSolutionError(vec)``; defaults tol=1e-10 (absolute 2-norm of b - A x),
maxNumIters=n, preconditioner diag(mat).

``mat`` is a dense square matrix: a numpy array (uploaded once) or a torch
float64 CUDA tensor (used in place).  ``setRowScaling(s)`` makes the iteration
run on diag(s) A x = diag(s) b; the Green matrix needs s = triangle areas to be
symmetric (see icqsol_b200.bem.icqBaseLaplaceSolver.solveGreenSystem).
Single rank; the sharded solve lives in BaseLaplaceSolver.
"""
import ctypes

import numpy

from icqsol_b200 import _lib


class ConjugateGradient:

    def __init__(self, mat, b, device=None):
        """
        Constructor
        @param mat dense, square matrix
        @param b right hand side vector
        """
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('icqsol_b200 needs a CUDA device: there is no CPU fallback')
        if device is None:
            device = mat.device.index if isinstance(mat, torch.Tensor) and mat.is_cuda \
                else torch.cuda.current_device()
        self.device = torch.device('cuda', int(device))
        self.ctx = _lib.Context(int(device))
        self.lib = self.ctx.lib
        torch.cuda.set_device(self.device)
        self.ctx.check(self.lib.icqb_set_stream(self.ctx.handle,
                                                torch.cuda.current_stream().cuda_stream or 1))
        n = len(b)
        self.n = n
        if isinstance(mat, torch.Tensor) and mat.is_cuda:
            assert mat.dtype == torch.float64 and mat.shape[0] == n and mat.stride(1) == 1
            self.mat = mat
            self.ld = mat.stride(0)
        else:
            host = numpy.ascontiguousarray(mat, numpy.float64)
            assert host.shape == (n, n)
            self.ld = (n + 15) // 16 * 16
            self.mat = torch.zeros((n, self.ld), dtype=torch.float64, device=self.device)
            self.mat[:, :n] = torch.from_numpy(host).to(self.device)
        self.b = torch.from_numpy(numpy.ascontiguousarray(b, numpy.float64)).to(self.device)
        self.maxNumIters = n
        self.tol = 1.e-10
        self.verbose = False
        self.precond = None     # None = diag(mat), extracted on the device
        self.rowScaling = None

    def setTolerance(self, tol):
        self.tol = tol

    def setMaxNumberOfIterations(self, maxNumIters):
        self.maxNumIters = maxNumIters

    def setVerbosity(self, verbose):
        self.verbose = verbose

    def setDiagonalPreconditioner(self, precond):
        import torch
        self.precond = torch.from_numpy(numpy.ascontiguousarray(precond, numpy.float64)).to(self.device)

    def setBlockPreconditioner(self, on=True):
        """Extension: precondition with the inverse 128 x 128 diagonal blocks of the
        (row-scaled) matrix instead of 1/diag; ignored once setDiagonalPreconditioner
        has been called."""
        self.ctx.check(self.lib.icqb_cg_set_preconditioner(self.ctx.handle, 1 if on else 0))

    def setRowScaling(self, scale):
        import torch
        self.rowScaling = torch.from_numpy(numpy.ascontiguousarray(scale, numpy.float64)).to(self.device)

    def solve(self, x0):
        """
        Solve linear system
        @param x0 initial guess for solution
        @return solution, error, and number of iterations
        """
        import torch
        x = torch.from_numpy(numpy.ascontiguousarray(x0, numpy.float64)).to(self.device)
        torch.cuda.synchronize(self.device)
        iters = ctypes.c_int64(0)
        r2 = ctypes.c_double(0.)
        rinf = ctypes.c_double(0.)
        self.ctx.check(self.lib.icqb_cg_solve_dev(
            self.ctx.handle, self.mat.data_ptr(), self.n, self.ld, 0,
            self.rowScaling.data_ptr() if self.rowScaling is not None else 0,
            self.precond.data_ptr() if self.precond is not None else 0,
            self.b.data_ptr(), x.data_ptr(), float(self.tol), 0, int(self.maxNumIters),
            ctypes.byref(iters), ctypes.byref(r2), ctypes.byref(rinf)))
        if self.verbose:
            print('iterations {0} error = {1}'.format(iters.value, r2.value))
        self.lastSolveInfo = {'iterations': iters.value, 'residual2': r2.value,
                              'residualInf': rinf.value, 'milliseconds': self.ctx.last_ms(4),
                              'gemvMilliseconds': self.ctx.last_ms(3)}
        return x.cpu().numpy(), r2.value, iters.value

    def getSolutionError(self, vec):
        """
        Get the solution error
        @param vec solution
        @return error ||b - A vec||_2
        """
        y = numpy.zeros(self.n, numpy.float64)
        v = numpy.ascontiguousarray(vec, numpy.float64)
        self.ctx.check(self.lib.icqb_gemv_host(self.ctx.handle, self.mat.data_ptr(), self.n, self.n,
                                               self.ld, v.ctypes.data_as(_lib.c_double_p),
                                               y.ctypes.data_as(_lib.c_double_p)))
        return float(numpy.linalg.norm(self.b.cpu().numpy() - y))
