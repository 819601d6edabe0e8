"""PoissonSolver: surface potential due to a surface charge, v = G . charge.

Same class, constructor, default field names and method as the reference
(bem/icqPoissonSolver.py:9-40); the ``numpy.dot`` at :36 is the GEMV kernel.
"""
from icqsol_b200.bem.icqBaseLaplaceSolver import BaseLaplaceSolver


class PoissonSolver(BaseLaplaceSolver):

    def __init__(self, pdata, max_edge_length, order=5, **kwargs):
        """
        Constructor
        @param pdata polydata (vtkPolyData or stand-in)
        @param max_edge_length maximum edge length, used to turn
                               polygons into triangles
        @param order order of the self-term integration (1..5)
        """
        BaseLaplaceSolver.__init__(self, pdata, max_edge_length, order, **kwargs)
        self.responseName = 'v'
        self.sourceName = 'charge'

    def computeResponseField(self):
        """
        Compute the response field, in this case the potential due to a charge source
        @return response
        """
        srcIndex = self.getSourceArrayIndex()
        src = self.getSourceArray(srcIndex)

        # Compute the response.
        rsp = self.applyGreenMatrix(src)

        self.addResponseField(rsp)

        return rsp
