"""Row-block partition of the dense matrices over the ranks of one box, and the
few host-side collectives the solvers need around the CUDA library
(SURVEY.md section 8e).  The reference is single-process; this is new.

torch.distributed is plumbing only: it carries the NCCL unique id to the ranks
(the library then talks NCCL itself) and gathers result vectors.  Works with
the ``nccl`` backend on GPUs and with ``gloo`` on CPU (tests).
"""
import numpy


def worldInfo():
    """(rank, world_size) of the default process group, (0, 1) without one."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def rowChunk(n, nranks):
    """Rows per rank: ceil(n / nranks) (the slot size of the all-gathered vectors)."""
    return (int(n) + int(nranks) - 1) // int(nranks)


def rowRange(n, rank, nranks):
    """Observer rows [r0, r1) owned by ``rank``: consecutive blocks of
    ceil(n/nranks) rows; trailing ranks may own fewer (or no) rows.  Must match
    the partition icqb_cg_solve_dev assumes."""
    chunk = rowChunk(n, nranks)
    r0 = min(int(n), int(rank) * chunk)
    r1 = min(int(n), r0 + chunk)
    return r0, r1


TILE = 128


def _tilesIn(a, b):
    """Tiles stored by tile rows [a, b): tile row I holds the tiles J = 0..I."""
    return (b * (b + 1) - a * (a + 1)) // 2


_foldedCache = {}


def foldedTileRows(n, nranks):
    """[(a0, a1, b0, b1)] per rank, in tile rows: rank g owns tile rows [a0, a1) from the short
    end of the lower triangle and [b0, b1) from the long end ("folded"), the last rank the
    stretch left in the middle.  The cut points are chosen so that every rank STORES the same
    number of 128 x 128 tiles (tile row I holds I + 1 of them) -- that is what a rank assembles
    and what it streams per product; equal row counts would leave the ranks up to 2P/nT apart
    (4 % at 56,168 triangles on 8 ranks), because a tile row at the long end weighs nT tiles.
    Ranks 0..P-2 are cut one after the other against the cumulative target (g+1)/P of all
    tiles: for every count of long rows the matching count of short rows is solved for, and
    the pair closest to the target wins (two granule sizes -> resolution far below one long
    row: max/ideal <= 1.001 from 20,000 triangles on).  Deterministic, identical on every rank."""
    n, nranks = int(n), int(nranks)
    key = (n, nranks)
    if key in _foldedCache:
        return _foldedCache[key]
    nT = (n + TILE - 1) // TILE
    total = _tilesIn(0, nT)
    lo, hi, cum = 0, nT, 0
    out = []
    for g in range(nranks):
        if g == nranks - 1:
            out.append((lo, hi, hi, hi))
            break
        left = nranks - g - 1                    # ranks still to be served after this one
        target = (g + 1) * total / float(nranks) - cum
        rows = hi - lo
        best = None
        for khi in range(0, max(rows - left, 0) + 1):
            th = _tilesIn(hi - khi, hi)
            if th > target + nT:
                break
            rem = max(target - th, 0.)
            k0 = int(round((-(2 * lo + 1) + numpy.sqrt((2 * lo + 1) ** 2 + 8. * rem)) / 2.))
            for k in (k0 - 1, k0, k0 + 1):
                if k < 0 or k + khi > rows - left:
                    continue
                t = th + _tilesIn(lo, lo + k)
                cand = (abs(t - target), abs(k - khi), k, khi, t)
                if best is None or cand < best:
                    best = cand
        if best is None:                         # fewer tile rows than ranks: hand out what is left
            k, khi = 0, (1 if rows > 0 else 0)
            t = _tilesIn(hi - khi, hi)
        else:
            _, _, k, khi, t = best
        out.append((lo, lo + k, hi - khi, hi))
        lo += k
        hi -= khi
        cum += t
    _foldedCache[key] = out
    return out


def foldedBlocks(n, rank, nranks):
    """Row blocks (r0, r1, q0, q1) of ``rank`` for the block-lower-triangular storage
    (foldedTileRows in rows; an empty first block is replaced by the second, so that
    0 <= r0 <= r1 <= q0 <= q1 <= n always holds)."""
    n = int(n)
    a0, a1, b0, b1 = foldedTileRows(n, nranks)[int(rank)]
    r0, r1 = min(n, a0 * TILE), min(n, a1 * TILE)
    q0, q1 = min(n, b0 * TILE), min(n, b1 * TILE)
    if r0 == r1:
        r0, r1, q0, q1 = q0, q1, q1, q1
    if q0 == q1:
        q0 = q1 = r1
    return r0, r1, q0, q1


def storedTiles(blocks):
    """128 x 128 tiles held by row blocks (r0, r1, q0, q1)."""
    return lowerStrips(blocks)[1]


def lowerStrips(blocks):
    """[(first global row, rows in the strip, global tile row I, index of its tile J=0)]
    for the 128-row strips of row blocks (r0, r1, q0, q1), in storage order -- the
    tile-contiguous layout of include/icqb.h (icqb_assemble_lower_dev)."""
    strips, tile = [], 0
    for (b, e) in ((blocks[0], blocks[1]), (blocks[2], blocks[3])):
        for row0 in range(b, e, TILE):
            tileRow = row0 // TILE
            strips.append((row0, min(TILE, e - row0), tileRow, tile))
            tile += tileRow + 1
    return strips, tile


def lowerTilesToRows(tiles, blocks, n):
    """Local tile sequence float64[ntiles, 128, 128] -> stored rows float64[rows, n]
    (columns beyond each row's diagonal tile are left zero)."""
    strips, ntiles = lowerStrips(blocks)
    rows = (blocks[1] - blocks[0]) + (blocks[3] - blocks[2])
    out = numpy.zeros((rows, n), numpy.float64)
    lrow = 0
    for (row0, nrows, tileRow, base) in strips:
        t = tiles[base:base + tileRow + 1]                     # [I+1, 128, 128]
        strip = t.transpose(1, 0, 2).reshape(TILE, (tileRow + 1) * TILE)
        ncols = min(n, (tileRow + 1) * TILE)
        out[lrow:lrow + nrows, :ncols] = strip[:nrows, :ncols]
        lrow += nrows
    return out


def sliverCapTiles(xyz, tri, threshold=30., maxTiles=32, flagged=None):
    """(head_tiles, tail_tiles): how many leading (trailing) 128-triangle tiles hold slivers --
    triangles whose aspect Lmax^2 / |db x dc| exceeds ``threshold`` -- among the first (last)
    ``maxTiles`` tiles.  Neighbouring slivers make diag(area) G indefinite (DESIGN.md section
    4.6); on the UV spheres they are the fan at each pole and every other triangle of the next
    band(s), i.e. they sit at the two ends of the triangle list (aspect 51 at n_theta = 318, 26
    for the rest of the first band, below 30 from the second band on), and a dense block has to
    span the whole band to restore convergence.  Informational: the preconditioner's blocks
    come from sliverBlockPartition, which also covers slivers in the middle of the list.
    ``flagged``: the result of sliverTiles, if already at hand."""
    if flagged is None:
        flagged = sliverTiles(xyz, tri, threshold)
    flagged = numpy.asarray(flagged, bool)
    nT = flagged.shape[0]
    if nT == 0:
        return 0, 0
    w = min(int(maxTiles), nT)
    lead = numpy.nonzero(flagged[:w])[0]
    head = int(lead[-1]) + 1 if lead.size else 0
    w = min(int(maxTiles), nT - head)
    trail = numpy.nonzero(flagged[nT - w:][::-1])[0] if w > 0 else numpy.zeros(0, int)
    tail = int(trail[-1]) + 1 if trail.size else 0
    return head, tail


def sliverTiles(xyz, tri, threshold=30.):
    """bool[number of 128-triangle tiles]: the tile holds a sliver, a triangle whose aspect
    Lmax^2 / |db x dc| exceeds ``threshold``."""
    xyz = numpy.asarray(xyz, numpy.float64)
    tri = numpy.asarray(tri, numpy.int64)
    n = tri.shape[0]
    nT = (n + TILE - 1) // TILE
    if n == 0:
        return numpy.zeros(nT, bool)
    # component arrays (contiguous), no numpy.cross: this runs inside every solver constructor
    x, y, z = (numpy.ascontiguousarray(xyz[:, k]) for k in range(3))
    i0, i1, i2 = (numpy.ascontiguousarray(tri[:, k]) for k in range(3))
    bx, by, bz = x[i1] - x[i0], y[i1] - y[i0], z[i1] - z[i0]
    cx, cy, cz = x[i2] - x[i0], y[i2] - y[i0], z[i2] - z[i0]
    e2 = numpy.maximum(bx * bx + by * by + bz * bz, cx * cx + cy * cy + cz * cz)
    dx, dy, dz = cx - bx, cy - by, cz - bz
    e2 = numpy.maximum(e2, dx * dx + dy * dy + dz * dz)
    nx, ny, nz = by * cz - bz * cy, bz * cx - bx * cz, bx * cy - by * cx
    sliver = e2 * e2 > (threshold * threshold) * (nx * nx + ny * ny + nz * nz)
    pad = numpy.zeros(nT * TILE, bool)
    pad[:n] = sliver
    return pad.reshape(nT, TILE).any(axis=1)


def rankCuts(n, nranks):
    """Tile rows at which the folded row blocks of the ranks begin or end (foldedTileRows): a
    diagonal block of the preconditioner that does not cross one of them lives on ONE rank,
    which then extracts, inverts and applies it alone (csrc/cg_caps.cu)."""
    cuts = set()
    for (a0, a1, b0, b1) in foldedTileRows(n, nranks):
        cuts.update((a0, a1, b0, b1))
    return sorted(cuts)


def sliverBlockPartition(xyz, tri, blockTiles=1, threshold=30., gapTiles=4, maxTiles=64,
                         budgetFraction=0.125, budgetFloor=6.4e7, flagged=None, cuts=None):
    """First tile row of every diagonal block of the block preconditioner (kind 2, the default)
    (icqb_cg_set_block_partition) for a mesh whose slivers may sit ANYWHERE in the triangle
    list -- e.g. the poles of the inner sphere of a hollow sphere: every run of tiles that hold
    slivers (runs separated by at most ``gapTiles`` clean tiles are merged: on a UV sphere
    every other triangle of a band is a sliver and whole tiles in between are clean) becomes
    ONE dense block, cut into pieces of at most ``maxTiles`` tiles; the clean stretches get
    blocks of ``blockTiles`` tiles.  On a UV sphere this is sliverCapTiles' head and tail cap.
    If the dense blocks would take more than ``budgetFraction`` of the lower storage
    (N^2 / 2 entries; at least ``budgetFloor`` entries), the gaps are no longer merged; if that is still too much, only runs
    that reach the two ends of the list are kept.
    ``cuts`` (tile rows, e.g. rankCuts): the blocks of the clean stretches also end there (the
    sliver runs are never cut).
    @return (starts, denseRuns): list of first tile rows; [(first tile, tiles)] of the runs"""
    if flagged is None:
        flagged = sliverTiles(xyz, tri, threshold)
    flagged = numpy.asarray(flagged, bool)     # (tests pass the tile flags directly)
    nT = flagged.shape[0]
    k = max(int(blockTiles), 1)
    maxTiles = max(int(maxTiles), 1)
    if nT == 0:
        return [0], []
    idx = numpy.nonzero(flagged)[0]

    def runsFor(gap):
        runs = []
        for t in idx.tolist():
            if runs and t - runs[-1][1] <= gap + 1:
                runs[-1][1] = t
            else:
                runs.append([t, t])
        return [(a, b - a + 1) for a, b in runs]

    def cost(runs):
        total = 0
        for a, m in runs:
            while m > 0:
                piece = min(m, maxTiles)
                if piece > 1:
                    total += (piece * TILE) ** 2
                m -= piece
        return total

    budget = max(budgetFraction * 0.5 * float(len(tri)) ** 2, float(budgetFloor))
    runs = runsFor(int(gapTiles))
    if cost(runs) > budget:
        runs = runsFor(0)
    if cost(runs) > budget:
        runs = [r for r in runs if r[0] == 0 or r[0] + r[1] == nT]
    if cost(runs) > budget:
        runs = []
    cutList = sorted(int(c) for c in (cuts or []) if 0 < int(c) < nT)

    def step(t, limit):
        """end of the clean block that starts at t: k tiles, the next cut or `limit`"""
        end = min(t + k, limit)
        for c in cutList:
            if t < c < end:
                end = c
                break
        return end

    starts, t = [], 0
    for a, m in runs:
        while t < a:
            starts.append(t)
            t = step(t, a)
        end = a + m
        while t < end:
            starts.append(t)
            t = min(t + maxTiles, end)
    while t < nT:
        starts.append(t)
        t = step(t, nT)
    return (starts or [0]), [r for r in runs if r[1] > 1]


def blockExtents(starts, numTiles, denseRuns, overlapTiles=1, maxTiles=64):
    """Overlapping extents of the diagonal blocks of a partition (icqb_cg_set_block_extents):
    block q = tile rows [starts[q], starts[q+1]) grows by ``overlapTiles`` tiles on either side
    (symmetric additive Schwarz: a row in two extents is weighted 1/sqrt(2) in each).  The runs of
    sliver tiles (``denseRuns`` of sliverBlockPartition) must never be split between blocks --
    the wrong-sign eigenvectors of diag(A) G live on them and a block that holds
    only a piece of such a run makes the iteration stall (tools/experiments/pcg_overlap.py) --
    so a sliver block grows OUTWARDS like the others, but no other block grows into it.
    An extent never exceeds ``maxTiles`` tiles, never reaches past its neighbours' far ends
    (a row lies in two extents at most).
    @return (begins, ends): tile rows, one pair per block"""
    starts = [int(t) for t in starts]
    nT = int(numTiles)
    ov = max(int(overlapTiles), 0)
    nb = len(starts)
    ends = starts[1:] + [nT]
    sliver = [False] * (nT + 1)
    for (a, m) in denseRuns:
        for t in range(int(a), min(int(a) + int(m), nT)):
            sliver[t] = True
    ext0, ext1 = [], []
    for q in range(nb):
        t0, t1 = starts[q], ends[q]
        # not past the far end of the neighbouring blocks: at most two extents per row
        lo = starts[q - 1] if q > 0 else 0
        hi = ends[q + 1] if q + 1 < nb else nT
        e0, e1 = max(t0 - ov, lo, 0), min(t1 + ov, hi, nT)
        if not sliver[t0]:
            while e0 < t0 and sliver[e0]:
                e0 += 1
            while e1 > t1 and sliver[e1 - 1]:
                e1 -= 1
        while e1 - e0 > int(maxTiles):     # shrink the overlap, never the block
            if e1 > t1 and (e1 - t1 >= t0 - e0):
                e1 -= 1
            elif e0 < t0:
                e0 += 1
            else:
                break
        ext0.append(e0)
        ext1.append(e1)
    # a row may lie in the extents of two CONSECUTIVE blocks only
    for q in range(2, nb):
        if ext0[q] < ext1[q - 2]:
            ext0[q] = min(ext1[q - 2], starts[q])
    return ext0, ext1


def blockPartition(n, blockTiles, headTiles=0, tailTiles=0):
    """First tile row of every diagonal block of the block preconditioner (kind 2): a cap of
    headTiles, blocks of blockTiles, a cap of tailTiles (icqb_cg_set_block_partition)."""
    nT = (int(n) + TILE - 1) // TILE
    head = min(max(int(headTiles), 0), nT)
    tail = min(max(int(tailTiles), 0), nT - head)
    k = max(int(blockTiles), 1)
    starts = [0] if head > 0 or nT - tail > 0 else []
    t = head
    while t < nT - tail:
        if t > 0:
            starts.append(t)
        t += k
    if tail > 0 and nT - tail > 0:
        starts.append(nT - tail)
    return starts or [0]


def gatherFoldedRows(localRows, n):
    """Stored rows float64[len0+len1, ncols] of every rank -> float64[n, ncols] in
    global row order, on every rank."""
    rank, nranks = worldInfo()
    localRows = numpy.ascontiguousarray(localRows, numpy.float64)
    if nranks == 1:
        return localRows[:n]
    import torch
    import torch.distributed as dist
    allBlocks = [foldedBlocks(n, g, nranks) for g in range(nranks)]
    slot = max((b[1] - b[0]) + (b[3] - b[2]) for b in allBlocks)     # rows of the largest rank
    ncols = localRows.shape[1]
    mine = allBlocks[rank]
    nmine = (mine[1] - mine[0]) + (mine[3] - mine[2])
    dev = _backendDevice()
    send = torch.zeros((slot, ncols), dtype=torch.float64)
    send[:nmine] = torch.from_numpy(localRows[:nmine])
    send = send.to(dev)
    recv = torch.empty((slot * nranks, ncols), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.cpu().numpy()
    out = numpy.zeros((n, ncols), numpy.float64)
    for g, (a0, a1, b0, b1) in enumerate(allBlocks):
        out[a0:a1] = recv[slot * g:slot * g + (a1 - a0)]
        out[b0:b1] = recv[slot * g + (a1 - a0):slot * g + (a1 - a0) + (b1 - b0)]
    return out


def _backendDevice():
    import torch
    import torch.distributed as dist
    if dist.get_backend() == 'nccl':
        return torch.device('cuda', torch.cuda.current_device())
    return torch.device('cpu')


def broadcastBytes(payload, nbytes, src=0):
    """Broadcast ``nbytes`` bytes (a bytes object on ``src``, ignored elsewhere)."""
    rank, nranks = worldInfo()
    if nranks == 1:
        return bytes(payload)
    import torch
    import torch.distributed as dist
    if rank == src:
        buf = torch.tensor(list(bytes(payload)), dtype=torch.uint8)
    else:
        buf = torch.zeros(nbytes, dtype=torch.uint8)
    buf = buf.to(_backendDevice())
    dist.broadcast(buf, src=src)
    return bytes(buf.cpu().numpy().tobytes())


def allGatherRows(local, n):
    """Concatenate the per-rank row slices (numpy float64[m_rank]) into float64[n]."""
    rank, nranks = worldInfo()
    local = numpy.ascontiguousarray(local, numpy.float64)
    if nranks == 1:
        return local.copy()
    import torch
    import torch.distributed as dist
    chunk = rowChunk(n, nranks)
    dev = _backendDevice()
    send = torch.zeros(chunk, dtype=torch.float64)
    send[:local.shape[0]] = torch.from_numpy(local)
    send = send.to(dev)
    recv = torch.empty(chunk * nranks, dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send)
    return recv.cpu().numpy()[:n].copy()


def allGatherMatrixRows(localRows, n):
    """Gather row blocks float64[m_rank, ncols] into float64[n, ncols] on every rank."""
    rank, nranks = worldInfo()
    localRows = numpy.ascontiguousarray(localRows, numpy.float64)
    if nranks == 1:
        return localRows
    import torch
    import torch.distributed as dist
    chunk = rowChunk(n, nranks)
    ncols = localRows.shape[1]
    dev = _backendDevice()
    send = torch.zeros((chunk, ncols), dtype=torch.float64)
    send[:localRows.shape[0]] = torch.from_numpy(localRows)
    send = send.to(dev)
    recv = torch.empty((chunk * nranks, ncols), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send)
    return recv.cpu().numpy()[:n].copy()
