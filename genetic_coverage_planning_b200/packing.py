"""One-time host packing of the reference's lookup tables into the device layout.

Input  : what `Mappy.computeFrustums` / `PathMaker` hold (reference mappy.py:274-328,
         pathmaker.py:157-173): map image, waypoint pixel coords, 0/1 traversability
         graph, per-edge 17-vertex view polygon, heading and length.
Output : `PackedWorld` - CSR graph, per-edge cost records, per-edge footprint
         *sub-tiles* aligned to a 64x64-pixel block grid, per-block map bitsets.

Why this is exact:
  * the reference paints `cv2.fillPoly(draw_map, [poly], 1, offset=center)` once per
    executed step (mappy.py:353-355) and then only sums the map.  Painting is
    idempotent, so the painted set of an organism is the union of per-edge painted
    sets; each per-edge set is rasterised here ONCE by the same cv2.fillPoly call
    (on a window of the map; `tests/test_packing.py` proves window == full map).
  * a step whose (from,to) is not an edge reads an all-zero polygon, which fillPoly
    paints as the single centre pixel (F3/F4): stored as the "null edge" of `from`.
  * turning cost depends only on the edge heading because prev_theta stays 0 (F5):
    it is evaluated here per edge with the reference's own scalar expression.

Block layout (device): block b = by*nbx + bx covers x in [64bx,64bx+64), y in
[64by,64by+64); 64 rows of one little-endian u64 (bit x&63).  A warp lane l owns
rows 2l,2l+1 (one 16-byte word).  A *quarter* is 16 rows = 128 B = one cache line.
A sub-tile stores only its non-empty quarters; `payload = (offset128 << 4) | qmask`.
"""
import os

import numpy as np

BLK = 64
_inv_2pi = 0.5 / np.pi


def _rad_wrap_pi(angle):
    # reference gori_tools.py:170-172 (restated)
    angle -= 2 * np.pi * np.floor((angle + np.pi) * _inv_2pi)
    return angle


def turn_cost_per_edge(theta, rho):
    """reference mappy.py:349,352 with prev_theta == 0, evaluated with numpy *scalar*
    arithmetic exactly as the reference does (array pow may take a SIMD path)."""
    theta = np.asarray(theta, dtype=np.float64)
    uniq, inv = np.unique(theta, return_inverse=True)
    out = np.empty(len(uniq), dtype=np.float64)
    half_pi = np.pi / 2
    for k in range(len(uniq)):
        delta = _rad_wrap_pi(np.float64(uniq[k] - 0))
        out[k] = rho * abs(delta / half_pi) ** rho
    return out[inv]


def safety_image(img, narrowest_hall, scale):
    """reference mappy.py:18,36-39 (same cv2.dilate call)."""
    import cv2
    num_dilations = int((narrowest_hall / 2) / scale)
    kernel = np.ones((3, 3), np.uint8)
    dil = cv2.dilate(np.ascontiguousarray(img, dtype=np.float64), kernel,
                     iterations=num_dilations)
    return 0.25 * img + 0.25 * dil


def _rows_to_u64(bits_u8):
    """[..., 64] array of 0/1 -> [...] uint64 with bit x = column x."""
    packed = np.packbits(np.ascontiguousarray(bits_u8, dtype=np.uint8), axis=-1,
                         bitorder='little')
    return packed.view('<u8').reshape(bits_u8.shape[:-1])


def _blockify(plane, nby, nbx):
    """[H,W] bool -> [nby*nbx, 64] u64 block-major bitset (zero padded)."""
    H, W = plane.shape
    pad = np.zeros((nby * BLK, nbx * BLK), dtype=np.uint8)
    pad[:H, :W] = plane
    rows = _rows_to_u64(pad.reshape(nby * BLK, nbx, BLK))          # [nby*64, nbx]
    return np.ascontiguousarray(rows.reshape(nby, BLK, nbx).transpose(0, 2, 1)
                                ).reshape(nby * nbx, BLK)


class PackedWorld(object):
    """Numpy arrays in exactly the layout gcp_upload_* expects (include/gcp.h)."""
    pass


def pack_world(img, xy, row_ptr, col, edge_poly, edge_theta, edge_dist, map_params,
               progress=False, workers=None):
    """workers: processes for the footprint rasterisation (None = all host cores for big worlds,
    1 for small ones); the result does not depend on it."""
    pw = PackedWorld()
    img = np.ascontiguousarray(img, dtype=np.float64)
    H, W = img.shape
    N = int(len(xy))
    E = int(len(col))
    g_u8 = np.rint(img * 255).astype(np.int64)
    if not np.array_equal(g_u8 / 255, img):
        raise ValueError("map must be an 8-bit gray image divided by 255 "
                         "(reference pather_driver.py:202)")
    nbx = (W + BLK - 1) // BLK
    nby = (H + BLK - 1) // BLK
    pw.H, pw.W, pw.N, pw.E, pw.nbx, pw.nby = H, W, N, E, nbx, nby
    pw.n_blocks = nbx * nby
    if pw.n_blocks >= (1 << 28):
        raise ValueError("map too large for 28-bit block ids")

    # ---- graph ----------------------------------------------------------
    pw.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
    pw.col = np.ascontiguousarray(col, dtype=np.uint32)
    pw.xy = np.ascontiguousarray(xy, dtype=np.int32)
    deg = np.diff(np.asarray(row_ptr, dtype=np.int64))
    pw.xover_cand = (deg > 2).astype(np.uint8)          # pathmaker.py:157-161
    pw.max_degree = int(deg.max()) if N else 0

    # ---- per-edge costs (+ N null edges with zero cost) --------------------
    rho = float(map_params['rho'])
    cost = np.zeros((E + N, 2), dtype=np.float64)
    cost[:E, 0] = edge_dist
    cost[:E, 1] = turn_cost_per_edge(edge_theta, rho)
    pw.edge_cost = cost
    pw.edge_theta = np.ascontiguousarray(edge_theta, dtype=np.float64)

    # ---- map planes --------------------------------------------------------
    safety = safety_image(img, map_params['narrowest_hall'], map_params['scale'])
    mask = np.logical_and(safety < 0.3, safety > 0)     # mappy.py:337
    zero_g = g_u8 == 0
    gray = np.logical_and(g_u8 > 0, g_u8 < 255)
    pw.blk_mask0 = _blockify(np.logical_and(mask, zero_g), nby, nbx)
    blk_mask_all = _blockify(mask, nby, nbx)            # host only: pre-masked footprint store
    pw.blk_free0 = _blockify(zero_g, nby, nbx)
    gy, gx = np.nonzero(gray)
    gb = (gy // BLK) * nbx + (gx // BLK)
    order = np.argsort(gb, kind='stable')
    gy, gx, gb = gy[order], gx[order], gb[order]
    ent = ((gy % BLK) * BLK + (gx % BLK)).astype(np.uint32)
    ent |= (255 - g_u8[gy, gx]).astype(np.uint32) << 12
    ent |= mask[gy, gx].astype(np.uint32) << 20
    pw.gray_ent = ent
    gp = np.zeros(pw.n_blocks + 1, dtype=np.int64)
    np.add.at(gp, gb + 1, 1)
    pw.gray_ptr = np.cumsum(gp).astype(np.uint32)
    pw.mask_size = int(mask.sum())                      # draw_map[buffer_mask].size
    pw.gsum_mask = int(g_u8[mask].sum())                # sum over mask of g
    pw.num_occluded = float(np.sum(img))                # mappy.py:25 (numpy pairwise)
    pw.map_size = int(img.size)
    pw.is_binary = bool(not gray.any())
    pw.blend = float(map_params['coverage_blend'])
    pw.solo_sep_thresh = int(map_params['solo_sep_thresh'])

    # ---- per-edge footprints ----------------------------------------------
    edge_poly = np.ascontiguousarray(edge_poly, dtype=np.int32)
    nv = edge_poly.shape[1] if E else 17
    reach = int(np.abs(edge_poly).max()) + 2 if E else 2
    src = np.repeat(np.arange(N, dtype=np.int64), deg)
    polys = np.concatenate([edge_poly, np.zeros((N, nv, 2), dtype=np.int32)]) if E else \
        np.zeros((N, nv, 2), dtype=np.int32)
    node = np.concatenate([src, np.arange(N, dtype=np.int64)])
    job = dict(polys=polys, cx=pw.xy[node, 1].astype(np.int64),      # center = (int(loc[1]), int(loc[0]))
               cy=pw.xy[node, 0].astype(np.int64), reach=reach, H=H, W=W, nbx=nbx,
               blk_mask_all=blk_mask_all)                            # host only: pre-masked footprint store
    EN = E + N
    if workers is None:
        workers = min(os.cpu_count() or 1, 32) if EN >= 20000 else 1
    if workers > 1:
        # every edge is rasterised independently: fan the edge range out over forked processes
        # (the tables are inherited, not pickled) and stitch the per-range stores together
        import multiprocessing as mp
        global _JOB
        _JOB = job
        bounds = np.linspace(0, EN, workers * 4 + 1).astype(np.int64)
        ranges = [(int(bounds[k]), int(bounds[k + 1])) for k in range(len(bounds) - 1) if bounds[k + 1] > bounds[k]]
        try:
            with mp.get_context("fork").Pool(workers) as pool:
                parts = pool.map(_rasterise_range, ranges)
        finally:
            _JOB = None
    else:
        parts = [_rasterise_edges(job, 0, EN, progress)]
    sub_cnt = np.concatenate([p_['sub_cnt'] for p_ in parts])
    m_sub_cnt = np.concatenate([p_['m_sub_cnt'] for p_ in parts])
    off128 = m_off128 = 0
    pay, m_pay = [], []
    for p_ in parts:                                     # payload = (offset128 << 4) | qmask: rebase the offsets
        pay.append(p_['sub_payload'] + (np.uint64(off128) << np.uint64(4)))
        m_pay.append(p_['m_sub_payload'] + (np.uint64(m_off128) << np.uint64(4)))
        off128 += p_['n_q']
        m_off128 += p_['m_n_q']
    if off128 >= (1 << 28):
        raise ValueError("footprint store exceeds 2^28 quarters (32 GiB)")
    sub_ptr = np.zeros(EN + 1, dtype=np.int64)
    np.cumsum(sub_cnt, out=sub_ptr[1:])
    m_sub_ptr = np.zeros(EN + 1, dtype=np.int64)
    np.cumsum(m_sub_cnt, out=m_sub_ptr[1:])
    pw.sub_ptr = sub_ptr.astype(np.uint32)
    pw.sub_block = np.concatenate([p_['sub_block'] for p_ in parts]).astype(np.uint32)
    pw.sub_payload = np.concatenate(pay).astype(np.uint32)
    pw.tile_data = np.concatenate([p_['tile'] for p_ in parts]).astype('<u8')
    pw.n_quarters = off128
    pw.edge_pixels = np.concatenate([p_['n_pix'] for p_ in parts])
    pw.max_sub_per_edge = int(sub_cnt.max()) if EN else 0
    if m_off128 >= (1 << 27):
        raise ValueError("masked footprint store exceeds 2^27 quarters")
    pw.m_sub_ptr = m_sub_ptr.astype(np.uint32)
    pw.m_sub_block = np.concatenate([p_['m_sub_block'] for p_ in parts]).astype(np.uint32)
    pw.m_sub_payload = np.concatenate(m_pay).astype(np.uint32)
    pw.m_tile_data = np.concatenate([p_['m_tile'] for p_ in parts]).astype('<u8')
    pw.m_n_quarters = m_off128
    return pw


_JOB = None


def _rasterise_range(r):
    return _rasterise_edges(_JOB, r[0], r[1], False)


def _rasterise_edges(job, e0, e1, progress=False):
    """Footprint sub-tiles of edges [e0, e1) (ids >= E are the null edges): both stores, with
    quarter offsets relative to this range."""
    import cv2
    polys, cx, cy, reach = job['polys'], job['cx'], job['cy'], job['reach']
    H, W, nbx, blk_mask_all = job['H'], job['W'], job['nbx'], job['blk_mask_all']
    kb = (2 * reach) // BLK + 2                         # canvas blocks per side
    canvas = np.zeros((kb * BLK, kb * BLK), dtype=np.uint8)
    n = e1 - e0
    sub_cnt = np.zeros(n, dtype=np.int64)
    m_sub_cnt = np.zeros(n, dtype=np.int64)
    n_pix = np.zeros(n, dtype=np.int32)
    sub_block, sub_payload, chunks = [], [], []
    m_sub_block, m_sub_payload, m_chunks = [], [], []
    off128 = m_off128 = 0
    it = range(e0, e1)
    if progress:
        from tqdm import tqdm
        it = tqdm(it, desc="rasterising footprints")
    for e in it:
        bx0 = (cx[e] - reach) // BLK                    # floor division, may be negative
        by0 = (cy[e] - reach) // BLK
        ox, oy = bx0 * BLK, by0 * BLK
        canvas[:] = 0
        # the reference draws on the map itself (mappy.py:355), so cv2 clips the polygon against the
        # map's borders - and a clipped edge is NOT the unclipped one with the outside pixels
        # removed (cv2 restarts its line walk at the clip point).  Draw into the part of the
        # canvas that lies inside the map: its borders coincide with the map's wherever the
        # footprint crosses them, and are out of reach elsewhere.
        x_lo, x_hi = max(0, -ox), min(kb * BLK, W - ox)
        y_lo, y_hi = max(0, -oy), min(kb * BLK, H - oy)
        if x_hi > x_lo and y_hi > y_lo:
            cv2.fillPoly(canvas[y_lo:y_hi, x_lo:x_hi], [polys[e]], 1,
                         offset=(int(cx[e] - ox - x_lo), int(cy[e] - oy - y_lo)))
        rows = _rows_to_u64(canvas.reshape(kb * BLK, kb, BLK))     # [kb*64, kb]
        n_sub = 0
        m_n_sub = 0
        for j in range(kb):
            for i in range(kb):
                blk = rows[j * BLK:(j + 1) * BLK, i]
                if not blk.any():
                    continue
                q_any = blk.reshape(4, 16).any(axis=1)
                qmask = int(q_any[0]) | int(q_any[1]) << 1 | int(q_any[2]) << 2 | int(q_any[3]) << 3
                sub_block.append((by0 + j) * nbx + (bx0 + i))
                sub_payload.append((off128 << 4) | qmask)
                chunks.append(blk.reshape(4, 16)[q_any].reshape(-1).copy())
                off128 += int(q_any.sum())
                n_sub += 1
                mblk = blk & blk_mask_all[(by0 + j) * nbx + (bx0 + i)]
                if mblk.any():
                    mq = mblk.reshape(4, 16).any(axis=1)
                    mqm = int(mq[0]) | int(mq[1]) << 1 | int(mq[2]) << 2 | int(mq[3]) << 3
                    m_sub_block.append((by0 + j) * nbx + (bx0 + i))
                    m_sub_payload.append((m_off128 << 4) | mqm)
                    m_chunks.append(mblk.reshape(4, 16)[mq].reshape(-1).copy())
                    m_off128 += int(mq.sum())
                    m_n_sub += 1
        n_pix[e - e0] = int(canvas.sum())
        sub_cnt[e - e0] = n_sub
        m_sub_cnt[e - e0] = m_n_sub
    z64 = np.zeros(0, dtype='<u8')
    return dict(sub_cnt=sub_cnt, m_sub_cnt=m_sub_cnt, n_pix=n_pix, n_q=off128, m_n_q=m_off128,
                sub_block=np.array(sub_block, dtype=np.int64), sub_payload=np.array(sub_payload, dtype=np.uint64),
                tile=np.concatenate(chunks) if chunks else z64,
                m_sub_block=np.array(m_sub_block, dtype=np.int64),
                m_sub_payload=np.array(m_sub_payload, dtype=np.uint64),
                m_tile=np.concatenate(m_chunks) if m_chunks else z64)


def pack_from_reference(mappy, pather, progress=False):
    """Build the packed world from live reference-API objects (Mappy, PathMaker)
    exactly as pather_driver.py:207-232 leaves them."""
    mp = {'scale': mappy._scale, 'narrowest_hall': mappy._hall_width, 'rho': mappy._rho,
          'solo_sep_thresh': mappy._solo_sep_thresh, 'coverage_blend': mappy._coverage_blend}
    if hasattr(pather, '_csr'):
        # CSR-backed objects (synthetic worlds: a dense N x N table cannot exist at 10^5 nodes)
        row_ptr, col = pather._csr
        return pack_world(mappy._img, pather._XY, row_ptr, col, mappy._edge_poly,
                          mappy._edge_theta, mappy._edge_dist, mp, progress=progress)
    g = np.asarray(pather._graph)
    src, dst = np.nonzero(g == 1)
    N = g.shape[0]
    row_ptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(row_ptr, src + 1, 1)
    row_ptr = np.cumsum(row_ptr)
    mp = {'scale': mappy._scale, 'narrowest_hall': mappy._hall_width, 'rho': mappy._rho,
          'solo_sep_thresh': mappy._solo_sep_thresh, 'coverage_blend': mappy._coverage_blend}
    return pack_world(mappy._img, pather._XY, row_ptr, dst,
                      mappy._traverse_frustum[src, dst], mappy._traverse_angles[src, dst],
                      mappy._traverse_dists[src, dst], mp, progress=progress)
