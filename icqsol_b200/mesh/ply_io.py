"""PLY polygon files without VTK: the other input format of the reference's BEM command line
(examples/solveLaplace.py:58-69 picks vtkPLYReader for ``*.ply`` through
shapes/icqShapeManager.py:995-997).  Reads ``ascii``, ``binary_little_endian`` and
``binary_big_endian`` files with a ``vertex`` element (x, y, z; other vertex properties become
point arrays) and a ``face`` element whose list property holds the vertex indices
(``vertex_indices`` / ``vertex_index``; other face properties are skipped); other elements
are skipped.  ``writePly`` writes what vtkPLYWriter writes for a plain mesh: float x, y, z and
``list uchar int vertex_indices``.
"""
import numpy

from icqsol_b200.mesh.polydata import PolyData, DataArray

_TYPES = {
    'char': 'i1', 'int8': 'i1', 'uchar': 'u1', 'uint8': 'u1',
    'short': 'i2', 'int16': 'i2', 'ushort': 'u2', 'uint16': 'u2',
    'int': 'i4', 'int32': 'i4', 'uint': 'u4', 'uint32': 'u4',
    'float': 'f4', 'float32': 'f4', 'double': 'f8', 'float64': 'f8',
}


def _header(buf):
    """(format, [(element name, count, [property ...])], offset of the data); a property is
    ('scalar', name, type) or ('list', name, count type, item type)."""
    if not buf.startswith(b'ply'):
        raise ValueError('not a PLY file')
    end = buf.find(b'end_header')
    if end < 0:
        raise ValueError('PLY header without end_header')
    nl = buf.find(b'\n', end)
    offset = (nl + 1) if nl >= 0 else len(buf)
    fmt, elements = None, []
    for line in buf[:end].decode('latin-1').splitlines()[1:]:
        t = line.split()
        if not t or t[0] in ('comment', 'obj_info'):
            continue
        if t[0] == 'format':
            fmt = t[1]
        elif t[0] == 'element':
            elements.append((t[1], int(t[2]), []))
        elif t[0] == 'property':
            if not elements:
                raise ValueError('PLY property before any element')
            if t[1] == 'list':
                elements[-1][2].append(('list', t[4], _TYPES[t[2]], _TYPES[t[3]]))
            else:
                elements[-1][2].append(('scalar', t[2], _TYPES[t[1]]))
    if fmt not in ('ascii', 'binary_little_endian', 'binary_big_endian'):
        raise ValueError('unsupported PLY format: {0}'.format(fmt))
    return fmt, elements, offset


def readPly(fileName):
    """Read a PLY file into the in-repo PolyData stand-in (points float64; vertex properties
    other than x, y, z as one-component point arrays)."""
    with open(fileName, 'rb') as f:
        buf = f.read()
    fmt, elements, pos = _header(buf)
    order = {'binary_little_endian': '<', 'binary_big_endian': '>'}.get(fmt)
    tokens, tpos = None, 0
    if fmt == 'ascii':
        tokens = buf[pos:].split()
    vertexColumns, faces = {}, []
    for name, count, props in elements:
        scalarOnly = all(p[0] == 'scalar' for p in props)
        if fmt != 'ascii' and scalarOnly:
            dt = numpy.dtype([(p[1], order + p[2]) for p in props])
            table = numpy.frombuffer(buf, dtype=dt, count=count, offset=pos)
            pos += count * dt.itemsize
            if name == 'vertex':
                vertexColumns = {p[1]: table[p[1]].astype(numpy.float64) for p in props}
            continue
        rows = []
        for _ in range(count):
            row = {}
            for p in props:
                if p[0] == 'scalar':
                    if fmt == 'ascii':
                        row[p[1]] = float(tokens[tpos])
                        tpos += 1
                    else:
                        dt = numpy.dtype(order + p[2])
                        row[p[1]] = float(numpy.frombuffer(buf, dt, 1, pos)[0])
                        pos += dt.itemsize
                else:
                    if fmt == 'ascii':
                        k = int(tokens[tpos])
                        row[p[1]] = [int(v) for v in tokens[tpos + 1:tpos + 1 + k]]
                        tpos += 1 + k
                    else:
                        dc, di = numpy.dtype(order + p[2]), numpy.dtype(order + p[3])
                        k = int(numpy.frombuffer(buf, dc, 1, pos)[0])
                        pos += dc.itemsize
                        row[p[1]] = numpy.frombuffer(buf, di, k, pos).astype(numpy.int64).tolist()
                        pos += k * di.itemsize
            rows.append(row)
        if name == 'vertex':
            vertexColumns = {p[1]: numpy.array([r[p[1]] for r in rows], numpy.float64)
                             for p in props if p[0] == 'scalar'}
        elif name == 'face':
            key = next((p[1] for p in props if p[0] == 'list' and
                        p[1] in ('vertex_indices', 'vertex_index')), None)
            if key is None:
                raise ValueError('PLY face element without a vertex_indices list')
            faces = [r[key] for r in rows]
    for c in ('x', 'y', 'z'):
        if c not in vertexColumns:
            raise ValueError('PLY vertex element without x, y, z')
    xyz = numpy.stack([vertexColumns['x'], vertexColumns['y'], vertexColumns['z']], 1)
    npts = xyz.shape[0]
    for f in faces:
        if len(f) < 3 or min(f) < 0 or max(f) >= npts:
            raise ValueError('PLY face with fewer than 3 vertices or an index out of range')
    sizes = set(len(f) for f in faces)
    cells = numpy.asarray(faces, numpy.int64) if len(sizes) == 1 else faces
    pdata = PolyData.from_arrays(xyz, cells, dtype=numpy.float64)
    for name, values in vertexColumns.items():
        if name not in ('x', 'y', 'z'):
            pdata.GetPointData().AddArray(DataArray(name, values))
    return pdata


def writePly(pdata, fileName, binary=True):
    """Write points and polygons (no data arrays): float x, y, z and
    ``list uchar int vertex_indices``, binary little endian or ASCII."""
    from icqsol_b200.mesh.polydata import extract_mesh
    xyz, cells = extract_mesh(pdata)
    xyz = numpy.asarray(xyz, numpy.float64)
    faces = cells.tolist() if isinstance(cells, numpy.ndarray) else [list(c) for c in cells]
    head = ['ply', 'format {0} 1.0'.format('binary_little_endian' if binary else 'ascii'),
            'comment icqsol_b200 PLY file', 'element vertex {0}'.format(xyz.shape[0]),
            'property float x', 'property float y', 'property float z',
            'element face {0}'.format(len(faces)), 'property list uchar int vertex_indices',
            'end_header']
    with open(fileName, 'wb') as f:
        f.write(('\n'.join(head) + '\n').encode('ascii'))
        if binary:
            f.write(xyz.astype('<f4').tobytes())
            for c in faces:
                f.write(numpy.uint8(len(c)).tobytes())
                f.write(numpy.asarray(c, '<i4').tobytes())
        else:
            for p in xyz.astype(numpy.float32):
                f.write('{0!r} {1!r} {2!r}\n'.format(float(p[0]), float(p[1]), float(p[2])).encode('ascii'))
            for c in faces:
                f.write((' '.join([str(len(c))] + [str(int(v)) for v in c]) + '\n').encode('ascii'))
