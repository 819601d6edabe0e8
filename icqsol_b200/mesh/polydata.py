"""In-repo stand-in for the slice of the vtkPolyData API the BEM path touches.

The reference solvers take a ``vtkPolyData`` (bem/icqBaseSolver.py:11-45) and
talk to it through a handful of methods.  VTK is not
installed in this image, so this module provides duck-typed classes with the
same method names backed by numpy arrays.  The host layer accepts either these
or real VTK objects (see :func:`extract_mesh`).

Points are stored the way VTK stores them by default -- float32 -- and are
promoted to float64 on ``GetPoint`` (shapes/icqShapeManager.py:794-798 builds
``vtkPoints()`` with the default type), unless ``dtype=float64`` is requested.
"""
import numpy


class IdList:
    """vtkIdList subset."""

    def __init__(self):
        self._ids = []

    @staticmethod
    def New():
        return IdList()

    def Delete(self):
        self._ids = []

    def GetNumberOfIds(self):
        return len(self._ids)

    def SetNumberOfIds(self, n):
        self._ids = [0] * n

    def GetId(self, i):
        return self._ids[i]

    def SetId(self, i, v):
        self._ids[i] = int(v)

    def InsertNextId(self, v):
        self._ids.append(int(v))
        return len(self._ids) - 1

    def _assign(self, ids):
        self._ids = [int(v) for v in ids]


class Points:
    """vtkPoints subset (float32 storage by default, like VTK)."""

    def __init__(self, array=None, dtype=numpy.float32):
        if array is None:
            self._xyz = numpy.zeros((0, 3), dtype)
        else:
            self._xyz = numpy.ascontiguousarray(array, dtype=dtype).reshape(-1, 3)

    def GetNumberOfPoints(self):
        return self._xyz.shape[0]

    def SetNumberOfPoints(self, n):
        self._xyz = numpy.zeros((n, 3), self._xyz.dtype)

    def GetPoint(self, i, out=None):
        p = self._xyz[i]
        if out is not None:
            out[0], out[1], out[2] = float(p[0]), float(p[1]), float(p[2])
            return None
        return (float(p[0]), float(p[1]), float(p[2]))

    def SetPoint(self, i, *xyz):
        if len(xyz) == 1:
            xyz = xyz[0]
        self._xyz[i, :] = xyz

    def InsertNextPoint(self, *xyz):
        if len(xyz) == 1:
            xyz = xyz[0]
        self._xyz = numpy.vstack([self._xyz, numpy.asarray(xyz, self._xyz.dtype)[None, :]])
        return self._xyz.shape[0] - 1

    def as_float64(self):
        """All points as a float64[Np,3] array (float32 values promoted)."""
        return self._xyz.astype(numpy.float64)


class CellArray:
    """vtkCellArray subset.  Cells of one size (the usual all-triangles case) are kept
    as an int64[ncells, k] array; mixed polygons as a list of id tuples."""

    def __init__(self, cells=None):
        self._array = None
        self._list = None
        if cells is None:
            self._list = []
        elif isinstance(cells, numpy.ndarray) and cells.ndim == 2:
            self._array = numpy.ascontiguousarray(cells, numpy.int64)
        else:
            self._list = [tuple(int(v) for v in c) for c in cells]
        self._cur = 0

    @property
    def _cells(self):
        """Cells as a list of id tuples (materialised on demand)."""
        if self._list is None:
            self._list = [tuple(row) for row in self._array.tolist()]
            self._array = None
        return self._list

    def as_array(self):
        """int64[ncells, k] when every cell has k vertices, else None."""
        if self._array is not None:
            return self._array
        if self._list and len(set(len(c) for c in self._list)) == 1:
            return numpy.asarray(self._list, numpy.int64)
        return None

    def GetNumberOfCells(self):
        return self._array.shape[0] if self._array is not None else len(self._list)

    def InitTraversal(self):
        self._cur = 0

    def GetNextCell(self, idList):
        if self._cur >= self.GetNumberOfCells():
            return 0
        if self._array is not None:
            idList._assign(self._array[self._cur])
        else:
            idList._assign(self._list[self._cur])
        self._cur += 1
        return 1

    def InsertNextCell(self, arg, ids=None):
        if ids is None:
            if isinstance(arg, IdList):
                ids = [arg.GetId(i) for i in range(arg.GetNumberOfIds())]
            else:
                ids = arg
        self._cells.append(tuple(int(v) for v in ids))
        return len(self._list) - 1


class DataArray:
    """vtkDoubleArray / vtkFloatArray subset."""

    def __init__(self, name=None, values=None):
        self._name = name
        if values is None:
            self._data = numpy.zeros((0, 1), numpy.float64)
        else:
            v = numpy.asarray(values, numpy.float64)
            self._data = v.reshape(v.shape[0], -1).copy()

    def GetName(self):
        return self._name

    def SetName(self, name):
        self._name = name

    def GetNumberOfComponents(self):
        return self._data.shape[1]

    def SetNumberOfComponents(self, n):
        self._data = numpy.zeros((self._data.shape[0], n), numpy.float64)

    def GetNumberOfTuples(self):
        return self._data.shape[0]

    def SetNumberOfTuples(self, n):
        self._data = numpy.zeros((n, self._data.shape[1]), numpy.float64)

    def GetComponent(self, i, j):
        return float(self._data[i, j])

    def SetComponent(self, i, j, v):
        self._data[i, j] = v

    def GetTuple(self, i):
        return tuple(float(v) for v in self._data[i])

    def SetTuple(self, i, values):
        self._data[i, :] = values

    def as_numpy(self):
        return self._data


class FieldData:
    """vtkCellData / vtkPointData subset."""

    def __init__(self):
        self._arrays = []

    def GetNumberOfArrays(self):
        return len(self._arrays)

    def GetArray(self, key):
        if isinstance(key, str):
            for a in self._arrays:
                if a.GetName() == key:
                    return a
            return None
        return self._arrays[key]

    def AddArray(self, arr):
        # VTK replaces an existing array of the same name
        for i, a in enumerate(self._arrays):
            if a.GetName() == arr.GetName():
                self._arrays[i] = arr
                return i
        self._arrays.append(arr)
        return len(self._arrays) - 1


class PolyData:
    """vtkPolyData subset: points, polygons, cell/point data."""

    def __init__(self, points=None, polys=None):
        self._points = points if points is not None else Points()
        self._polys = polys if polys is not None else CellArray()
        self._cellData = FieldData()
        self._pointData = FieldData()

    @classmethod
    def from_arrays(cls, xyz, cells, dtype=numpy.float32):
        return cls(Points(xyz, dtype=dtype), CellArray(cells))

    def GetPoints(self):
        return self._points

    def SetPoints(self, points):
        self._points = points

    def GetPolys(self):
        return self._polys

    def SetPolys(self, polys):
        self._polys = polys

    def GetNumberOfPolys(self):
        return self._polys.GetNumberOfCells()

    def GetNumberOfPoints(self):
        return self._points.GetNumberOfPoints()

    def GetCellData(self):
        return self._cellData

    def GetPointData(self):
        return self._pointData


def new_id_list(pdata):
    """An id list compatible with ``pdata`` (real VTK or the stand-in)."""
    if isinstance(pdata, PolyData):
        return IdList()
    import vtk  # real VTK polydata: needs a real vtkIdList
    return vtk.vtkIdList()


def new_double_array(pdata):
    """A named-double-array object compatible with ``pdata``."""
    if isinstance(pdata, PolyData):
        return DataArray()
    import vtk
    return vtk.vtkDoubleArray()


def extract_mesh(pdata):
    """Points float64[Np,3] and polygons of a duck-typed polydata: an int64[ncells,k]
    array when every cell has k vertices (stand-in only), else a list of id tuples.

    Mirrors the traversal at bem/icqBaseSolver.py:25-38 and
    bem/icqLaplaceMatrices.cpp:54-71.
    """
    if isinstance(pdata, PolyData):
        arr = pdata._polys.as_array()
        return pdata._points.as_float64(), (arr if arr is not None else list(pdata._polys._cells))
    points = pdata.GetPoints()
    npts = points.GetNumberOfPoints()
    xyz = numpy.zeros((npts, 3), numpy.float64)
    for i in range(npts):
        xyz[i, :] = points.GetPoint(i)
    polys = pdata.GetPolys()
    ids = new_id_list(pdata)
    polys.InitTraversal()
    cells = []
    for _ in range(polys.GetNumberOfCells()):
        polys.GetNextCell(ids)
        cells.append(tuple(ids.GetId(k) for k in range(ids.GetNumberOfIds())))
    return xyz, cells
