"""Legacy ``.vtk`` POLYDATA reader/writer (no VTK dependency).

Replaces, for the BEM path only, what the reference does through
``vtkPolyDataReader``/``vtkPolyDataWriter`` in
shapes/icqShapeManager.py:626-641 (load) and :725-751 (save): ASCII and
big-endian BINARY files with POINTS, POLYGONS and POINT_DATA / CELL_DATA
sections (SCALARS, NORMALS, VECTORS, FIELD).  Other dataset types raise.
"""
import numpy

from icqsol_b200.mesh.polydata import PolyData, Points, CellArray, DataArray

_DTYPES = {
    'float': numpy.dtype('>f4'), 'double': numpy.dtype('>f8'),
    'int': numpy.dtype('>i4'), 'unsigned_int': numpy.dtype('>u4'),
    'long': numpy.dtype('>i8'), 'unsigned_long': numpy.dtype('>u8'),
    'vtkIdType': numpy.dtype('>i8'), 'vtktypeint64': numpy.dtype('>i8'),
    'short': numpy.dtype('>i2'), 'unsigned_short': numpy.dtype('>u2'),
    'char': numpy.dtype('>i1'), 'unsigned_char': numpy.dtype('>u1'),
    'bit': numpy.dtype('>u1'),
}


class _Stream:
    """Byte stream with line/token access for the ASCII parts and raw reads
    for the BINARY payloads."""

    def __init__(self, buf):
        self.buf = buf
        self.pos = 0

    def eof(self):
        return self.pos >= len(self.buf)

    def readline(self):
        end = self.buf.find(b'\n', self.pos)
        if end < 0:
            end = len(self.buf)
        line = self.buf[self.pos:end]
        self.pos = end + 1
        return line.decode('latin-1')

    def next_keyword_line(self):
        """Next non-blank line split into tokens, or None at EOF."""
        while not self.eof():
            toks = self.readline().split()
            if toks:
                return toks
        return None

    def read_values(self, count, dtype, binary):
        if binary:
            nbytes = count * dtype.itemsize
            arr = numpy.frombuffer(self.buf, dtype=dtype, count=count, offset=self.pos)
            self.pos += nbytes
            # payload is followed by a newline
            if self.pos < len(self.buf) and self.buf[self.pos:self.pos + 1] == b'\n':
                self.pos += 1
            return arr.astype(dtype.newbyteorder('='))
        vals = []
        while len(vals) < count:
            if self.eof():
                raise ValueError('unexpected end of file in ASCII payload')
            vals.extend(self.readline().split())
        if len(vals) != count:
            raise ValueError('ASCII payload does not end on a line boundary')
        native = dtype.newbyteorder('=')
        if native.kind == 'f':
            # parse in the stored precision (float32 text -> float32 value), like VTK
            return numpy.array(vals, dtype=numpy.float64).astype(native)
        return numpy.array(vals, dtype=numpy.int64).astype(native)


def _read_attributes(st, first, n, binary, target):
    """Parse the arrays of one POINT_DATA / CELL_DATA section into ``target``.
    Returns the keyword line that ended the section (or None)."""
    toks = first
    while True:
        toks = st.next_keyword_line() if toks is None else toks
        if toks is None:
            return None
        kw = toks[0].upper()
        if kw in ('POINT_DATA', 'CELL_DATA'):
            return toks
        if kw == 'SCALARS':
            name, typ = toks[1], toks[2]
            ncomp = int(toks[3]) if len(toks) > 3 else 1
            nxt = st.next_keyword_line()
            if nxt[0].upper() != 'LOOKUP_TABLE':
                raise ValueError('SCALARS without LOOKUP_TABLE')
            vals = st.read_values(n * ncomp, _DTYPES[typ], binary)
            target.AddArray(DataArray(name, vals.reshape(n, ncomp)))
        elif kw in ('NORMALS', 'VECTORS'):
            vals = st.read_values(n * 3, _DTYPES[toks[2]], binary)
            target.AddArray(DataArray(toks[1], vals.reshape(n, 3)))
        elif kw == 'COLOR_SCALARS':
            ncomp = int(toks[2])
            dt = _DTYPES['unsigned_char'] if binary else _DTYPES['float']
            vals = st.read_values(n * ncomp, dt, binary)
            target.AddArray(DataArray(toks[1], vals.reshape(n, ncomp)))
        elif kw == 'TEXTURE_COORDINATES':
            dim = int(toks[2])
            vals = st.read_values(n * dim, _DTYPES[toks[3]], binary)
            target.AddArray(DataArray(toks[1], vals.reshape(n, dim)))
        elif kw == 'FIELD':
            for _ in range(int(toks[2])):
                h = st.next_keyword_line()
                ncomp, ntup = int(h[1]), int(h[2])
                vals = st.read_values(ncomp * ntup, _DTYPES[h[3]], binary)
                target.AddArray(DataArray(h[0], vals.reshape(ntup, ncomp)))
        elif kw == 'LOOKUP_TABLE':
            size = int(toks[2])
            dt = _DTYPES['unsigned_char'] if binary else _DTYPES['float']
            st.read_values(size * 4, dt, binary)
        elif kw == 'METADATA':
            while not st.eof() and st.readline().strip():
                pass
        else:
            raise ValueError('unsupported attribute keyword {0}'.format(toks[0]))
        toks = None


def readVtkPolyData(filename):
    """Read a legacy-format POLYDATA file into a :class:`PolyData` stand-in."""
    with open(filename, 'rb') as f:
        st = _Stream(f.read())
    header = st.readline()
    if not header.startswith('# vtk DataFile'):
        raise ValueError('{0}: not a legacy VTK file'.format(filename))
    st.readline()  # title
    fmt = st.readline().strip().upper()
    if fmt not in ('ASCII', 'BINARY'):
        raise ValueError('{0}: unknown format {1}'.format(filename, fmt))
    binary = fmt == 'BINARY'
    ds = st.next_keyword_line()
    if ds[0].upper() != 'DATASET' or ds[1].upper() != 'POLYDATA':
        raise ValueError('{0}: only DATASET POLYDATA is supported'.format(filename))

    pdata = PolyData()
    toks = None
    while True:
        toks = st.next_keyword_line() if toks is None else toks
        if toks is None:
            break
        kw = toks[0].upper()
        if kw == 'POINTS':
            n, dt = int(toks[1]), _DTYPES[toks[2]]
            vals = st.read_values(3 * n, dt, binary).reshape(n, 3)
            store = numpy.float64 if dt.itemsize == 8 else numpy.float32
            pdata.SetPoints(Points(vals, dtype=store))
        elif kw in ('POLYGONS', 'VERTICES', 'LINES', 'TRIANGLE_STRIPS'):
            ncells, size = int(toks[1]), int(toks[2])
            vals = st.read_values(size, _DTYPES['int'], binary)
            if kw == 'POLYGONS':
                cells, p = [], 0
                for _ in range(ncells):
                    k = int(vals[p])
                    cells.append(tuple(int(v) for v in vals[p + 1:p + 1 + k]))
                    p += 1 + k
                pdata.SetPolys(CellArray(cells))
        elif kw == 'POINT_DATA':
            toks = _read_attributes(st, None, int(toks[1]), binary, pdata.GetPointData())
            continue
        elif kw == 'CELL_DATA':
            toks = _read_attributes(st, None, int(toks[1]), binary, pdata.GetCellData())
            continue
        elif kw == 'METADATA':
            while not st.eof() and st.readline().strip():
                pass
        else:
            raise ValueError('{0}: unsupported section {1}'.format(filename, toks[0]))
        toks = None
    return pdata


def _write_arrays(f, fieldData, binary):
    n = fieldData.GetNumberOfArrays()
    if n == 0:
        return
    f.write('FIELD FieldData {0}\n'.format(n).encode('latin-1'))
    for i in range(n):
        arr = fieldData.GetArray(i)
        data = arr.as_numpy()
        f.write('{0} {1} {2} double\n'.format(arr.GetName(), data.shape[1],
                                              data.shape[0]).encode('latin-1'))
        flat = data.reshape(-1)
        if binary:
            f.write(flat.astype('>f8').tobytes() + b'\n')
        else:
            for k in range(0, flat.size, 9):
                f.write((' '.join(repr(float(v)) for v in flat[k:k + 9]) + '\n').encode('latin-1'))


def writeVtkPolyData(pdata, filename, binary=False):
    """Write a legacy POLYDATA file (points, polygons, cell/point arrays), ASCII or -- like
    the reference's default, shapes/icqShapeManager.py:725-751 -- big-endian BINARY."""
    xyz = pdata.GetPoints().as_float64()
    cells = pdata.GetPolys()._cells
    with open(filename, 'wb') as f:
        f.write('# vtk DataFile Version 3.0\nvtk output\n{0}\nDATASET POLYDATA\n'.format(
            'BINARY' if binary else 'ASCII').encode('latin-1'))
        isDouble = pdata.GetPoints()._xyz.dtype == numpy.float64
        f.write('POINTS {0} {1}\n'.format(xyz.shape[0], 'double' if isDouble else 'float').encode('latin-1'))
        if binary:
            f.write(xyz.astype('>f8' if isDouble else '>f4').tobytes() + b'\n')
        else:
            for p in xyz:
                f.write('{0!r} {1!r} {2!r}\n'.format(float(p[0]), float(p[1]), float(p[2])).encode('latin-1'))
        size = sum(len(c) + 1 for c in cells)
        f.write('POLYGONS {0} {1}\n'.format(len(cells), size).encode('latin-1'))
        if binary:
            flat = []
            for c in cells:
                flat.append(len(c))
                flat.extend(c)
            f.write(numpy.asarray(flat, '>i4').tobytes() + b'\n')
        else:
            for c in cells:
                f.write('{0} {1}\n'.format(len(c), ' '.join(str(v) for v in c)).encode('latin-1'))
        if pdata.GetCellData().GetNumberOfArrays():
            f.write('CELL_DATA {0}\n'.format(len(cells)).encode('latin-1'))
            _write_arrays(f, pdata.GetCellData(), binary)
        if pdata.GetPointData().GetNumberOfArrays():
            f.write('POINT_DATA {0}\n'.format(xyz.shape[0]).encode('latin-1'))
            _write_arrays(f, pdata.GetPointData(), binary)
