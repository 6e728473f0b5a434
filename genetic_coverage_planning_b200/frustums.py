"""Footprint polygons on the device: Mappy.computeFrustums (mappy.py:274-328) through
`gcp_compute_frustums` (csrc/gcp_frustum.cu).

The reference loops over all N x N waypoint pairs and, per edge, over every obstacle pixel of
the map in numpy.  The kernel takes one CTA per source waypoint and only the window of pixels
that can pass the max_view filter.  Everything the reference decides in floating point is
decided on the same fp64 values; the few quantities that depend on libm (the heading of an edge,
its sine / cosine, the ray angles and their sines / cosines) are computed HERE with the
reference's own numpy expressions and handed to the kernel, and edges where a device atan2
lands within 1e-11 rad of a bin boundary (or a rotated coordinate within 1e-9 px of x.5) come back
flagged and are redone with `synth.compute_frustums` (the windowed numpy restatement, bit-checked
against the reference's tables in tests/test_synth.py).  There is no CPU path for the bulk:
without the CUDA library the call raises."""
import ctypes as C

import numpy as np

from . import _lib, synth


def edge_headings(xy, row_ptr, col):
    """(theta, dist, cos, sin) per edge, each evaluated with the scalar numpy calls of the
    reference (mappy.py:286-287, gori_tools.py:184-186) once per distinct pixel offset."""
    xy = np.asarray(xy, dtype=np.int64)
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    col = np.asarray(col, dtype=np.int64)
    src = np.repeat(np.arange(len(xy), dtype=np.int64), np.diff(row_ptr))
    dr = xy[col, 0] - xy[src, 0]
    dc = xy[col, 1] - xy[src, 1]
    key = np.stack([dr, dc], 1)
    uniq, inv = np.unique(key, axis=0, return_inverse=True)
    inv = np.asarray(inv).reshape(-1)
    th = np.empty(len(uniq)); di = np.empty(len(uniq)); co = np.empty(len(uniq)); si = np.empty(len(uniq))
    for k in range(len(uniq)):
        a, b = uniq[k, 0], uniq[k, 1]                    # numpy int64 scalars, as in the reference
        t = np.arctan2(b, a)                             # :286
        th[k] = t
        di[k] = np.linalg.norm([b, a])                   # :287
        co[k] = np.cos(t)
        si[k] = np.sin(t)
    return th[inv], di[inv], co[inv], si[inv]


def compute_frustums_device(img, xy, row_ptr, col, map_params, device=0, return_info=False):
    """Same contract as synth.compute_frustums: (poly int32 [E, R+2, 2], theta [E], dist [E])."""
    lib = _lib.load()
    scale = map_params['scale']
    num_rays = int(map_params['num_rays'])
    min_view = map_params['min_view']
    max_view = map_params['max_view']
    view_angle = map_params['view_angle']
    xy64 = np.asarray(xy, dtype=np.int64)
    N, E = len(xy64), len(col)
    alpha = view_angle / num_rays                                     # :279
    ray_angles = np.arange(num_rays) * alpha - view_angle / 2.0       # :280
    z = np.vstack((max_view / scale * np.ones_like(ray_angles[None, :]), ray_angles[None, :]))   # :307
    ray_sin = np.ascontiguousarray(np.sin(z[1]))                      # :314
    ray_cos = np.ascontiguousarray(np.cos(z[1]))
    fp = _lib.GcpFrustumParams()
    fp.num_rays = num_rays
    fp.window = int(np.ceil(max_view / scale)) + 1
    fp.scale = scale
    fp.max_view = max_view
    fp.max_view_px = float(z[0, 0])
    fp.half_view = view_angle / 2
    fp.near_last[0] = float(np.sin(z[1, -1]) * min_view)              # :315-316
    fp.near_last[1] = float(np.cos(z[1, -1]) * min_view)
    fp.near_first[0] = float(np.sin(z[1, 0]) * min_view)              # :317-318
    fp.near_first[1] = float(np.cos(z[1, 0]) * min_view)
    fp.eps_angle = 1e-11
    fp.eps_round = 1e-9
    theta, dist, cos_t, sin_t = edge_headings(xy64, row_ptr, col)
    occ = np.ascontiguousarray(np.asarray(img) != 0, dtype=np.uint8)
    H, W = occ.shape
    xy32 = np.ascontiguousarray(xy64, dtype=np.int32)
    rp = np.ascontiguousarray(row_ptr, dtype=np.uint32)
    cl = np.ascontiguousarray(col, dtype=np.uint32)
    poly = np.zeros((E, num_rays + 2, 2), dtype=np.int32)
    amb = np.zeros(max(E, 1), dtype=np.uint8)
    ms = C.c_float(0.0)
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (ray_angles, ray_sin, ray_cos, theta, cos_t, sin_t)]
    rc = lib.gcp_compute_frustums(int(device), occ.ctypes.data, H, W, xy32.ctypes.data, rp.ctypes.data,
                                  cl.ctypes.data, N, E, C.addressof(fp), *[a.ctypes.data for a in arrs],
                                  poly.ctypes.data, amb.ctypes.data, C.addressof(ms))
    if rc != 0:
        raise RuntimeError("gcp_compute_frustums: %s" % lib.gcp_last_error(None).decode())
    flagged = np.nonzero(amb[:E])[0]
    if len(flagged):
        src = np.repeat(np.arange(N, dtype=np.int64), np.diff(np.asarray(row_ptr, dtype=np.int64)))
        nodes = np.unique(src[flagged])
        p2, _, _ = synth.compute_frustums(img, xy64, row_ptr, col, map_params, only_nodes=nodes)
        poly[flagged] = p2[flagged]
    if return_info:
        return poly, theta, dist, {"kernel_ms": float(ms.value), "ambiguous_edges": int(len(flagged))}
    return poly, theta, dist


def computeFrustums(mappy, graph, device=0):
    """Drop-in for `mappy.computeFrustums(graph)` (pather_driver.py:232): fills
    `_traverse_frustum` [N, N, R+2, 2] int32, `_traverse_angles`, `_traverse_dists` [N, N]."""
    g = np.asarray(graph)
    N = g.shape[0]
    src, dst = np.nonzero(g == 1)                                     # row-major == CSR order
    row_ptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(row_ptr, src + 1, 1)
    row_ptr = np.cumsum(row_ptr)
    mp = dict(scale=mappy._scale, num_rays=mappy._num_rays, min_view=mappy._min_view,
              max_view=mappy._max_view, view_angle=mappy._view_angle)
    poly, theta, dist = compute_frustums_device(mappy._img, mappy._all_waypoints, row_ptr, dst, mp, device)
    R2 = poly.shape[1]
    fr = np.zeros((N, N, R2, 2), dtype=np.int32)
    an = np.zeros((N, N))
    di = np.zeros((N, N))
    fr[src, dst] = poly
    an[src, dst] = theta
    di[src, dst] = dist
    mappy._traverse_frustum, mappy._traverse_angles, mappy._traverse_dists = fr, an, di
