"""LaplaceSolver: normal electric field jump from a surface potential.

Same class, constructor, default field names and method as the reference
(bem/icqLaplaceSolver.py:8-40).  The reference computes
``rsp = -numpy.linalg.solve(gMat, src)`` (:36); here the dense system is solved
on the GPU(s) by conjugate gradients on the area-symmetrised matrix.
"""
from icqsol_b200.bem.icqBaseLaplaceSolver import BaseLaplaceSolver


class LaplaceSolver(BaseLaplaceSolver):

    def __init__(self, pdata, max_edge_length, order=5, **kwargs):
        """
        Constructor
        @param pdata polydata (vtkPolyData or stand-in)
        @param max_edge_length maximum edge length, used to turn
                               polygons into triangles
        @param order order of the self-term integration (1..5)
        """
        BaseLaplaceSolver.__init__(self, pdata, max_edge_length, order, **kwargs)
        self.responseName = 'normal_electric_field_jump'
        self.sourceName = 'v'
        self.solverTolerance = 5.e-13
        self.solverMaxNumIters = None

    def setSolverTolerance(self, tol):
        """Backward error max|v + G sigma| / max|v| at which the CG stops (default 5e-13,
        inside the 1e-12 parity criterion)."""
        self.solverTolerance = tol

    def setSolverMaxNumberOfIterations(self, maxNumIters):
        self.solverMaxNumIters = maxNumIters

    def computeResponseField(self):
        """
        Compute the response field, in this case the jump of the normal electric field
        due to a potential source
        @return response
        """
        srcIndex = self.getSourceArrayIndex()
        src = self.getSourceArray(srcIndex)

        # Compute the response: sigma = - G^{-1} v
        rsp = -self.solveGreenSystem(src, tol=self.solverTolerance,
                                     maxNumIters=self.solverMaxNumIters)

        self.addResponseField(rsp)

        return rsp
