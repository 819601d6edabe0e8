"""Mesh intake and field plumbing shared by the BEM solvers.

Same public surface as the reference's ``BaseSolver`` (bem/icqBaseSolver.py:9-157):
constructor ``(pdata, max_edge_length, order=5)``, ``getVtkPolyData``,
``getPoints``, ``getCells``, ``setResponseFieldName``, ``setSourceFieldName``,
``getSourceArrayIndex``, ``getSourceArray``, ``addResponseField``,
``setSourceFromExpression``.  ``pdata`` is duck-typed: a real ``vtkPolyData`` or
the in-repo stand-in (icqsol_b200.mesh.polydata.PolyData).

Differences, all forced by the image (no VTK, no pytriangle):
  * polygons are split into triangles by a fan and edges longer than
    ``max_edge_length`` by uniform per-triangle subdivision, instead of
    RefineSurface's constrained Delaunay step (shapes/icqRefineSurface.py:46-177);
    a mesh that is already triangulated and fine enough passes through untouched;
  * ``getCells`` returns int64 (the reference's ``numpy.int`` no longer exists).
"""
import numpy

from icqsol_b200.mesh import generators
from icqsol_b200.mesh.polydata import (PolyData, DataArray, extract_mesh, new_double_array,
                                       new_id_list)


def _findArray(data, name):
    # bem ... util/icqDataFetcher.py:7-21
    for i in range(data.GetNumberOfArrays()):
        if data.GetArray(i).GetName() == name:
            return i
    return -1


def _arrayValues(arr, ntuples):
    """float64[ntuples, components] of a (vtk or stand-in) data array."""
    if hasattr(arr, 'as_numpy'):
        return numpy.asarray(arr.as_numpy(), numpy.float64)[:ntuples].reshape(ntuples, -1)
    nc = arr.GetNumberOfComponents()
    return numpy.array([[arr.GetComponent(k, j) for j in range(nc)] for k in range(ntuples)],
                       numpy.float64).reshape(ntuples, nc)


class BaseSolver:

    def __init__(self, pdata, max_edge_length, order=5):
        """
        Constructor
        @param pdata polydata (vtkPolyData or stand-in)
        @param max_edge_length maximum edge length, used to turn
                               polygons into triangles
        @param order order of the self-term integration (1..5)
        """
        xyz, cells = extract_mesh(pdata)
        # every cell must be a polygon that can be turned into triangles
        # (the reference asserts 3 ids per refined cell, bem/icqBaseSolver.py:31)
        if isinstance(cells, numpy.ndarray):
            assert cells.shape[1] >= 3
            isTriangulated = cells.shape[1] == 3
        else:
            assert all(len(c) >= 3 for c in cells)
            isTriangulated = all(len(c) == 3 for c in cells)
        tri, fanParent = generators.fanSplit(cells, return_parent=True)
        refXyz, refTri, triParent, (ptIds, ptW) = generators.refineToMaxEdge(
            xyz, tri, max_edge_length, return_maps=True)
        if isTriangulated and refTri.shape[0] == tri.shape[0]:
            self.pdata = pdata
        else:
            # new polydata, as RefineSurface.getVtkPolyData() hands back: every child inherits
            # its parent's cell data (shapes/icqRefineSurface.py:155-174), point data is
            # interpolated onto the points the refinement added
            self.pdata = PolyData.from_arrays(refXyz, refTri)
            parentCell = fanParent[triParent]
            src = pdata.GetCellData()
            for i in range(src.GetNumberOfArrays()):
                arr = src.GetArray(i)
                vals = _arrayValues(arr, len(fanParent) and int(fanParent.max()) + 1)
                self.pdata.GetCellData().AddArray(DataArray(arr.GetName(), vals[parentCell]))
            src = pdata.GetPointData()
            for i in range(src.GetNumberOfArrays()):
                arr = src.GetArray(i)
                vals = _arrayValues(arr, xyz.shape[0])
                interp = (vals[ptIds[:, 0]] * ptW[:, 0:1] + vals[ptIds[:, 1]] * ptW[:, 1:2] +
                          vals[ptIds[:, 2]] * ptW[:, 2:3])
                self.pdata.GetPointData().AddArray(DataArray(arr.GetName(), interp))
            xyz, tri = self.pdata.GetPoints().as_float64(), refTri

        # point ids of each cell
        assert tri.ndim == 2 and tri.shape[1] == 3
        self._ptIdList = None
        self._xyz = numpy.ascontiguousarray(xyz, numpy.float64)
        self._tri = numpy.ascontiguousarray(tri, numpy.int64)

        self.points = self.pdata.GetPoints()
        self.polys = self.pdata.GetPolys()
        self.numTriangles = self.polys.GetNumberOfCells()
        assert self.numTriangles == self._tri.shape[0]

        # order of the integration, method dependent
        self.order = order

        # set in the derived classes
        self.responseName = 'NO-SET'
        self.sourceName = 'NOT-SET'

    @property
    def ptIdList(self):
        """Point ids of each triangle as nested lists (bem/icqBaseSolver.py:24-35)."""
        if self._ptIdList is None:
            self._ptIdList = self._tri.tolist()
        return self._ptIdList

    def getVtkPolyData(self):
        """
        Get the (modified) polydata object
        @return object
        """
        return self.pdata

    def getPoints(self):
        """
        Get grid points
        @return float64[numPoints, 3]
        """
        return self._xyz.copy()

    def getCells(self):
        """
        Get cell connectivity
        @return int64[numTriangles, 3]
        """
        return self._tri.copy()

    def setResponseFieldName(self, name):
        self.responseName = name

    def setSourceFieldName(self, name):
        self.sourceName = name

    def getSourceArrayIndex(self):
        """
        Index of the source cell array; a point array of that name is first
        projected onto the cells by averaging the three vertex values
        (util/icqDataFetcher.py:23-62).
        """
        cellData = self.pdata.GetCellData()
        srcIndex = _findArray(cellData, self.sourceName)
        if srcIndex < 0:
            pointData = self.pdata.GetPointData()
            index2 = _findArray(pointData, self.sourceName)
            if index2 >= 0:
                pointArr = pointData.GetArray(index2)
                cellArr = new_double_array(self.pdata)
                cellArr.SetName(self.sourceName)
                cellArr.SetNumberOfComponents(1)
                cellArr.SetNumberOfTuples(self.numTriangles)
                for i, (ia, ib, ic) in enumerate(self.ptIdList):
                    va = pointArr.GetComponent(ia, 0)
                    vb = pointArr.GetComponent(ib, 0)
                    vc = pointArr.GetComponent(ic, 0)
                    cellArr.SetComponent(i, 0, (va + vb + vc) / 3.)
                cellData.AddArray(cellArr)
                srcIndex = _findArray(cellData, self.sourceName)
        if srcIndex < 0:
            msg = 'ERROR: could not find any cell field named {0}!'.format(self.sourceName)
            raise RuntimeError(msg)
        return srcIndex

    def getSourceArray(self, srcIndex):
        """
        Source values, one per triangle
        @param srcIndex source array index
        @return float64[numTriangles]
        """
        srcArray = self.pdata.GetCellData().GetArray(srcIndex)
        n = self.pdata.GetNumberOfPolys()
        if hasattr(srcArray, 'as_numpy'):
            return srcArray.as_numpy()[:n, 0].astype(numpy.float64)
        src = numpy.zeros((n,), numpy.float64)
        for i in range(n):
            src[i] = srcArray.GetComponent(i, 0)
        return src

    def addResponseField(self, rsp):
        """
        Attach the response to the polydata as a named cell array
        @param rsp response numpy array
        """
        n = len(rsp)
        rspData = new_double_array(self.pdata)
        rspData.SetNumberOfComponents(1)
        rspData.SetNumberOfTuples(n)
        rspData.SetName(self.responseName)
        if hasattr(rspData, 'as_numpy'):
            rspData.as_numpy()[:, 0] = rsp
        else:
            for i in range(n):
                rspData.SetTuple(i, [rsp[i]])
        self.pdata.GetCellData().AddArray(rspData)

    def setSourceFromExpression(self, expression):
        """
        Set the source from expression, evaluated at each triangle's centroid
        @param expression expression of x, y, and z
        """
        from math import sqrt, pi, sin, cos, tan, log, exp  # noqa: F401 (names for eval)

        n = self.pdata.GetNumberOfPolys()
        sourceData = new_double_array(self.pdata)
        sourceData.SetNumberOfComponents(1)
        sourceData.SetNumberOfTuples(n)
        sourceData.SetName(self.sourceName)
        a, b, c = (self._xyz[self._tri[:, k]] for k in range(3))
        mid = a.copy()
        mid += b
        mid += c
        mid /= 3.0
        # same evaluation as the reference (eval per centroid, bem/icqBaseSolver.py:147-156),
        # with the expression compiled once instead of parsed n times
        code = compile(expression, '<source expression>', 'eval')
        for i, (x, y, z) in enumerate(mid.tolist()):
            v = eval(code)
            sourceData.SetTuple(i, [v])
        self.pdata.GetCellData().AddArray(sourceData)
