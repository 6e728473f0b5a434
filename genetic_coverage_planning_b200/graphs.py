"""Traversability graph on the device: PathMaker.computeTraversableGraph (pathmaker.py:119-155)
with Mappy.lineCollisionCheck (mappy.py:124-159) through `gcp_build_graph` (csrc/gcp_graph.cu,
SURVEY 8f N4).

The reference builds dense N x N displacement tables and, for every candidate edge, dilates the
whole map again and scans all of its obstacle pixels.  The kernels search a grid of waypoints and
scan only the pixels that can fail the reference's test.  Rows where a bearing lands within
1e-11 rad of a sector boundary come back flagged and are redone by `_row_host`, which evaluates the
reference's numpy expressions on that row (never triggered on the worlds of this repo: bearings of
integer displacements stay ~1e-9 rad away from odd multiples of pi/8)."""
import ctypes as C

import numpy as np

from . import _lib


def _sectors():
    sector_width = np.pi / 4                                               # pathmaker.py:125
    sectors = np.array([-7 / 2, -5 / 2, -3 / 2, -1 / 2, 1 / 2, 3 / 2, 5 / 2, 7 / 2])
    return sector_width * sectors                                          # :126-127


def _dilated(img, num_dilations):
    import cv2
    return cv2.dilate(np.asarray(img), np.ones((3, 3), np.uint8), iterations=num_dilations)   # mappy.py:136-138


def _collision_free(first, second, obstacles):
    """Mappy.lineCollisionCheck (mappy.py:124-159) on a precomputed obstacle list."""
    x1, y1, x2, y2 = first[0], first[1], second[0], second[1]
    a, b, c = y2 - y1, x2 - x1, x2 * y1 - y2 * x1
    if a == b and b == 0:
        return False
    dist = abs(a * obstacles[:, 0] - b * obstacles[:, 1] + c) / np.sqrt(a * a + b * b)
    prox = np.bitwise_not(np.bitwise_and(
        np.bitwise_or(np.bitwise_and(obstacles[:, 0] <= x2, obstacles[:, 0] <= x1),
                      np.bitwise_and(obstacles[:, 0] >= x2, obstacles[:, 0] >= x1)),
        np.bitwise_or(np.bitwise_and(obstacles[:, 1] <= y2, obstacles[:, 1] <= y1),
                      np.bitwise_and(obstacles[:, 1] >= y2, obstacles[:, 1] >= y1))))
    if dist[prox].size > 0:
        return not (min(dist[prox]) <= 1)
    return True


def _row_host(i, xy, scale, max_dist, obstacles):
    """One row of computeTraversableGraph with the reference's expressions (pathmaker.py:121-155)."""
    disp = np.array([xy[:, 0] - xy[i, 0], xy[:, 1] - xy[i, 1]])
    distances = np.linalg.norm(disp, axis=0) * scale
    angles = np.arctan2(disp[1], disp[0])
    sect = np.digitize(angles, _sectors())
    sect[sect == 8] = 0
    out = -np.ones(8, dtype=np.int32)
    for sector in range(8):
        cand = np.array(np.where(sect == sector)).flatten()
        if sector == 4:
            cand = cand[cand != i]
        if cand.size > 0:
            j = cand[np.argmin(distances[cand])]
            if distances[j] < max_dist and _collision_free(xy[i], xy[j], obstacles):
                out[sector] = j
    return out


def compute_traversable_graph_device(img, xy, scale, max_dist, hall_width, device=0, return_info=False):
    """CSR (row_ptr, col) of the graph the reference's PathMaker would build for waypoints `xy`
    ((row, col) pixels) on map `img` (g / 255)."""
    lib = _lib.load()
    xy64 = np.ascontiguousarray(xy, dtype=np.int64)
    N = len(xy64)
    occ = np.ascontiguousarray(np.asarray(img) != 0, dtype=np.uint8)
    H, W = occ.shape
    gp = _lib.GcpGraphParams()
    gp.scale = scale
    gp.max_dist = max_dist
    safety_buffer = (hall_width / 2) / 3                                   # pathmaker.py:18,153
    gp.num_dilations = int(safety_buffer / scale)                          # mappy.py:136
    gp.cell = int(np.ceil(max_dist / scale)) + 1
    sec = _sectors()
    for k in range(8):
        gp.sectors[k] = float(sec[k])
    gp.eps_angle = 1e-11
    ncx, ncy = (W + gp.cell - 1) // gp.cell, (H + gp.cell - 1) // gp.cell
    cy = np.clip(xy64[:, 0] // gp.cell, 0, ncy - 1)
    cx = np.clip(xy64[:, 1] // gp.cell, 0, ncx - 1)
    cid = cy * ncx + cx
    items = np.argsort(cid, kind='stable').astype(np.uint32)               # ascending index inside a cell
    cell_ptr = np.zeros(ncx * ncy + 1, dtype=np.int64)
    np.add.at(cell_ptr, cid + 1, 1)
    cell_ptr = np.cumsum(cell_ptr).astype(np.uint32)
    xy32 = np.ascontiguousarray(xy64, dtype=np.int32)
    nbr = -np.ones((max(N, 1), 8), dtype=np.int32)
    amb = np.zeros(max(N, 1), dtype=np.uint8)
    ms = C.c_float(0.0)
    rc = lib.gcp_build_graph(int(device), occ.ctypes.data, H, W, xy32.ctypes.data, N, cell_ptr.ctypes.data,
                             items.ctypes.data, C.addressof(gp), nbr.ctypes.data, amb.ctypes.data,
                             C.addressof(ms))
    if rc != 0:
        raise RuntimeError("gcp_build_graph: %s" % lib.gcp_last_error(None).decode())
    nbr = nbr[:N]
    flagged = np.nonzero(amb[:N])[0]
    if len(flagged):
        obstacles = np.array(np.nonzero(_dilated(img, gp.num_dilations))).T     # mappy.py:139
        for i in flagged:
            nbr[i] = _row_host(int(i), xy64, scale, max_dist, obstacles)
    srt = np.sort(np.where(nbr < 0, np.iinfo(np.int32).max, nbr), axis=1)  # row-major np.where order
    keep = srt != np.iinfo(np.int32).max
    row_ptr = np.concatenate([[0], np.cumsum(keep.sum(axis=1))]).astype(np.int64)
    col = srt[keep].astype(np.int64)
    if return_info:
        return row_ptr, col, {"kernel_ms": float(ms.value), "ambiguous_rows": int(len(flagged))}
    return row_ptr, col


def computeTraversableGraph(pather, device=0):
    """Drop-in for `pather.computeTraversableGraph()`: fills `_graph` (dense N x N float64 0/1) and
    `_crossover_cand` (findCrossoverCand, pathmaker.py:157-161)."""
    m = pather._mappy
    row_ptr, col = compute_traversable_graph_device(m._img, pather._XY, pather._scale, pather._max_dist,
                                                    pather._hall_width, device)
    N = len(pather._XY)
    g = np.zeros((N, N))
    src = np.repeat(np.arange(N), np.diff(row_ptr))
    g[src, col] = 1
    pather._graph = g
    pather._crossover_cand = np.where(np.diff(row_ptr) > 2)
