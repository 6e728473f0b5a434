"""Synthetic large worlds for the benchmark configs and a
windowed restatement of the reference's table builders that scales to 10^5..10^6 edges.

Everything here is one-time host setup (numpy / scipy / cv2), not the hot path:
  make_map            binary rooms-and-doors occupancy map (BSP), seeded
  place_waypoints     grid waypoints near walls, off the safety buffer
                      (role of PathMaker.smartlyPlaceDots, pathmaker.py:30-117)
  build_graph         nearest neighbour per octant within max_traverse_dist, collision
                      pruned (role of PathMaker.computeTraversableGraph, pathmaker.py:119-155;
                      same sector rule, :125-145)
  compute_frustums    per-edge view polygon / heading / length; windowed, CSR restatement of
                      Mappy.computeFrustums (mappy.py:274-328).  Because the reference drops
                      every obstacle pixel with distance >= max_view before using it (:294),
                      looking only at the (2R+1)^2 window around the waypoint is exactly
                      equivalent; tests/test_synth.py checks all 1,618 edges of the
                      reference's big map bit-for-bit.
  random_walk_population  vectorised heading-biased walks (role of PathMaker.makeMeAPath,
                      pathmaker.py:175-213) for seeding benchmark populations
"""
import numpy as np

DEFAULT_MAP_PARAMS = {          # pather_driver.py:66-88 (6-agent column)
    'scale': 0.15, 'narrowest_hall': 2.5, 'num_rays': 15, 'min_view': 0.2, 'max_view': 7,
    'view_angle': 69.4 * np.pi / 180, 'rho': 4.0, 'solo_sep_thresh': 10,
    'coverage_blend': 1.0}
DEFAULT_PATH_PARAMS = {'path_memory': 10, 'max_traverse_dist': 4.0, 'log_scale_weight': 0.7}


def make_map(size, seed=0, min_room=128, max_room=288, door=(56, 96), wall=(2, 4)):
    """Binary occupancy image (uint8 0/1, 1 = obstacle): BSP rooms, every wall has a door."""
    rng = np.random.RandomState(seed)
    H = W = int(size)
    img = np.zeros((H, W), dtype=np.uint8)
    t = int(wall[1])
    img[:t, :] = 1
    img[-t:, :] = 1
    img[:, :t] = 1
    img[:, -t:] = 1
    stack = [(0, 0, H, W)]
    while stack:
        r0, c0, r1, c1 = stack.pop()
        h, w = r1 - r0, c1 - c0
        if max(h, w) <= max_room:
            continue
        horizontal = h >= w if abs(h - w) > min_room // 2 else bool(rng.randint(2))
        span = h if horizontal else w
        if span < 2 * min_room:
            continue
        cut = int(rng.randint(min_room, span - min_room + 1))
        th = int(rng.randint(wall[0], wall[1] + 1))
        dlen = int(rng.randint(door[0], door[1] + 1))
        if horizontal:
            rr = r0 + cut
            img[rr:rr + th, c0:c1] = 1
            lo, hi = c0 + 12, c1 - 12 - dlen
            if hi > lo:
                dpos = int(rng.randint(lo, hi))
                img[rr:rr + th, dpos:dpos + dlen] = 0
            stack.append((r0, c0, rr, c1))
            stack.append((rr, c0, r1, c1))
        else:
            cc = c0 + cut
            img[r0:r1, cc:cc + th] = 1
            lo, hi = r0 + 12, r1 - 12 - dlen
            if hi > lo:
                dpos = int(rng.randint(lo, hi))
                img[dpos:dpos + dlen, cc:cc + th] = 0
            stack.append((r0, c0, r1, cc))
            stack.append((r0, cc, r1, c1))
    img[:t, :] = 1
    img[-t:, :] = 1
    img[:, :t] = 1
    img[:, -t:] = 1
    return img


def place_waypoints(img, map_params, stride=None, wall_factor=3.5):
    """Grid waypoints on free pixels whose distance to the nearest obstacle lies in
    [safety_buffer, wall_factor*safety_buffer] (the near-wall band the reference's pruning
    keeps, pathmaker.py:107-112; 3.5 instead of 2.5 keeps the band graph one component).  Returns int64 [N,2] (row, col)."""
    import cv2
    scale = map_params['scale']
    hall = map_params['narrowest_hall']
    if stride is None:
        stride = int(hall // scale)
    safety_px = (hall / 2) / scale
    dist = cv2.distanceTransform((img == 0).astype(np.uint8), cv2.DIST_L2, 5)
    H, W = img.shape
    rr, cc = np.mgrid[0:H:stride, 0:W:stride]
    rr, cc = rr.ravel(), cc.ravel()
    d = dist[rr, cc]
    keep = (d >= safety_px) & (d <= safety_px * wall_factor)
    return np.stack([rr[keep], cc[keep]], 1).astype(np.int64)


def build_graph(img, xy, map_params, path_params):
    """CSR (row_ptr, col) of the directed traversability graph."""
    import cv2
    from scipy.spatial import cKDTree
    scale = map_params['scale']
    max_px = path_params['max_traverse_dist'] / scale
    N = len(xy)
    tree = cKDTree(xy.astype(np.float64))
    pairs = tree.query_pairs(max_px * (1 - 1e-12), output_type='ndarray')
    if len(pairs) == 0:
        return np.zeros(N + 1, dtype=np.int64), np.zeros(0, dtype=np.int64)
    src = np.concatenate([pairs[:, 0], pairs[:, 1]])
    dst = np.concatenate([pairs[:, 1], pairs[:, 0]])
    disp = (xy[dst] - xy[src]).astype(np.float64)
    dist = np.sqrt(disp[:, 0] ** 2 + disp[:, 1] ** 2) * scale
    ok = dist < path_params['max_traverse_dist']                    # pathmaker.py:146
    src, dst, disp, dist = src[ok], dst[ok], disp[ok], dist[ok]
    ang = np.arctan2(disp[:, 1], disp[:, 0])                        # pathmaker.py:124
    sectors = (np.pi / 4) * np.array([-7 / 2, -5 / 2, -3 / 2, -1 / 2, 1 / 2, 3 / 2, 5 / 2, 7 / 2])
    sec = np.digitize(ang, sectors)
    sec[sec == 8] = 0                                               # :132-133
    # nearest candidate per (src, sector); ties -> lowest dst index like np.argmin (:143)
    order = np.lexsort((dst, dist, sec, src))
    src, dst, sec = src[order], dst[order], sec[order]
    first = np.ones(len(src), dtype=bool)
    first[1:] = (src[1:] != src[:-1]) | (sec[1:] != sec[:-1])
    src, dst = src[first], dst[first]
    # collision pruning against the map dilated by int((safety_buffer/3)/scale) (:151-154,
    # mappy.py:138-140): reject if any sample along the segment is within 1 px of it
    nd = int(((map_params['narrowest_hall'] / 2) / 3) / scale)
    dil = cv2.dilate(img.astype(np.uint8), np.ones((3, 3), np.uint8), iterations=nd + 1)
    ns = int(np.ceil(max_px)) + 1
    tt = np.linspace(0.0, 1.0, ns)[None, :]
    pr = np.rint(xy[src, 0:1] + (xy[dst, 0:1] - xy[src, 0:1]) * tt).astype(np.int64)
    pc = np.rint(xy[src, 1:2] + (xy[dst, 1:2] - xy[src, 1:2]) * tt).astype(np.int64)
    free = ~dil[pr, pc].astype(bool).any(axis=1)
    src, dst = src[free], dst[free]
    order = np.lexsort((dst, src))
    src, dst = src[order], dst[order]
    row_ptr = np.zeros(N + 1, dtype=np.int64)
    np.add.at(row_ptr, src + 1, 1)
    return np.cumsum(row_ptr), dst.astype(np.int64)


def _rot2d(theta):
    # gori_tools.py:184-186
    return np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])


def compute_frustums(img, xy, row_ptr, col, map_params, progress=False, only_nodes=None):
    """Windowed restatement of Mappy.computeFrustums (mappy.py:274-328) in CSR form.
    img: float64 map as the reference holds it (g/255).  Returns (poly int32 [E,R+2,2],
    theta [E], dist [E]).  only_nodes: fill only the out-edges of these source waypoints
    (frustums.py recomputes the edges its device kernel flags as ambiguous this way)."""
    scale = map_params['scale']
    num_rays = int(map_params['num_rays'])
    min_view = map_params['min_view']
    max_view = map_params['max_view']
    view_angle = map_params['view_angle']
    E = len(col)
    N = len(xy)
    poly = np.zeros((E, num_rays + 2, 2), dtype=np.int32)
    theta = np.zeros(E)
    dists = np.zeros(E)
    alpha = view_angle / num_rays                                     # :279
    ray_angles = np.arange(num_rays) * alpha - view_angle / 2.0       # :280
    H, W = img.shape
    R = int(np.ceil(max_view / scale)) + 1
    nz = img != 0
    it = range(N) if only_nodes is None else [int(v) for v in only_nodes]
    if progress:
        from tqdm import tqdm
        it = tqdm(it, desc="frustums")
    for frm in it:
        lo, hi = int(row_ptr[frm]), int(row_ptr[frm + 1])
        if lo == hi:
            continue
        wpt_loc = xy[frm]
        r0, r1 = max(0, int(wpt_loc[0]) - R), min(H, int(wpt_loc[0]) + R + 1)
        c0, c1 = max(0, int(wpt_loc[1]) - R), min(W, int(wpt_loc[1]) + R + 1)
        orr, occ = np.nonzero(nz[r0:r1, c0:c1])
        obstacles = np.array([orr + r0, occ + c0]) * scale             # :278 on the window
        displacements0 = np.array([obstacles[0] - wpt_loc[0] * scale,
                                   obstacles[1] - wpt_loc[1] * scale])  # :290-291
        distances0 = np.linalg.norm(displacements0, axis=0)            # :293
        dist_thresh = distances0 < max_view                            # :294
        distances0 = distances0[dist_thresh]
        displacements0 = displacements0[:, dist_thresh]
        base_angles = np.arctan2(displacements0[1], displacements0[0])
        for e in range(lo, hi):
            next_wpt = xy[col[e]]
            wpt_theta = np.arctan2(next_wpt[1] - wpt_loc[1], next_wpt[0] - wpt_loc[0])   # :286
            travel_cost = np.linalg.norm([next_wpt[1] - wpt_loc[1], next_wpt[0] - wpt_loc[0]])
            angles = base_angles - wpt_theta                            # :298
            angles[angles < -np.pi] += 2 * np.pi                        # :300-301
            angles[angles > np.pi] -= 2 * np.pi
            in_bounds = np.array(np.where(np.logical_and(angles < (view_angle / 2),
                                                         angles > -(view_angle / 2))))
            ang = angles[in_bounds]
            dis = distances0[in_bounds]
            dangles = np.digitize(ang, ray_angles)                      # :306
            z = np.vstack((max_view / scale * np.ones_like(ray_angles[None, :]),
                           ray_angles[None, :]))
            if dis.size:
                for ii in range(num_rays):                              # :308-312
                    sel = dis[dangles == ii + 1]
                    if sel.size:
                        z[0, ii] = min(sel) / scale
            pts = np.array([np.sin(z[1]) * z[0], np.cos(z[1]) * z[0]])  # :314
            pts = np.hstack((pts, np.array([[np.sin(z[1, -1]) * min_view],
                                            [np.cos(z[1, -1]) * min_view]]),
                             np.array([[np.sin(z[1, 0]) * min_view],
                                       [np.cos(z[1, 0]) * min_view]])))
            Rot = _rot2d(wpt_theta).T                                   # :319
            poly[e] = np.around(Rot.dot(pts)).astype('int32').T         # :320
            theta[e] = wpt_theta
            dists[e] = travel_cost
    return poly, theta, dists


def spread_starts(xy, n, size):
    """n start waypoints spread over the map (nearest waypoint to points on a ring)."""
    c = size / 2.0
    out = []
    for k in range(n):
        a = 2 * np.pi * k / max(n, 1)
        target = np.array([c + 0.3 * size * np.sin(a), c + 0.3 * size * np.cos(a)])
        out.append(int(np.argmin(((xy - target) ** 2).sum(1))))
    return out


def clustered_starts(xy, row_ptr, n, size):
    """n start waypoints next to each other near the map centre (the reference's own C1 set-up
    starts several agents from one waypoint, pather_driver.py:125-134): agents that start
    together share edges, so inter-agent loop closures - and feasible organisms - occur.
    Waypoints with at least 3 out-edges are preferred (walks need room to diverge)."""
    c = np.array([size / 2.0, size / 2.0])
    d2 = ((np.asarray(xy, dtype=np.float64) - c) ** 2).sum(1)
    deg = np.diff(np.asarray(row_ptr, dtype=np.int64))
    order = np.argsort(d2 + np.where(deg >= 3, 0.0, 1e18), kind='stable')
    k = max(1, (n + 1) // 2)                     # two agents per start waypoint
    return [int(order[a // 2 % k]) for a in range(n)]


def random_walk_population(xy, row_ptr, col, starts, n_org, walk_len, max_dna_len, seed=1,
                           path_memory=10, log_scale_weight=0.7, chunk=65536):
    """[n_org, A, max_dna_len] int32, -1 padded.  Each row is a heading-biased walk of
    walk_len steps from starts[a] that avoids its last `path_memory` nodes when it can
    (the rule of PathMaker.makeMeAPath, pathmaker.py:188-213), vectorised over walkers."""
    rng = np.random.RandomState(seed)
    N = len(xy)
    A = len(starts)
    deg = np.diff(row_ptr)
    D = int(deg.max())
    nbr = -np.ones((N, D), dtype=np.int64)
    for k in range(D):
        has = deg > k
        nbr[has, k] = col[row_ptr[:-1][has] + k]
    disp = xy[np.maximum(nbr, 0)] - xy[:, None, :]
    bearing = np.arctan2(disp[..., 1], disp[..., 0])            # pathmaker.py:179
    n_steps = min(walk_len, max_dna_len - 1)
    dna = -np.ones((n_org * A, max_dna_len), dtype=np.int32)
    start_all = np.tile(np.asarray(starts, dtype=np.int64), n_org)
    for s in range(0, n_org * A, chunk):
        e = min(n_org * A, s + chunk)
        Wn = e - s
        cur = start_all[s:e].copy()
        heading = rng.rand(Wn) * 2 * np.pi - np.pi               # :190 (never updated)
        hist = -np.ones((Wn, path_memory), dtype=np.int64)
        hist[:, 0] = cur
        dna[s:e, 0] = cur
        for t in range(n_steps):
            cand = nbr[cur]                                       # [Wn, D]
            valid = cand >= 0
            fresh = valid & ~(cand[:, :, None] == hist[:, None, :]).any(axis=2)
            use = np.where(fresh.any(axis=1)[:, None], fresh, valid)     # :199-207
            ang = bearing[cur] - heading[:, None]
            ang -= 2 * np.pi * np.floor((ang + np.pi) * (0.5 / np.pi))
            w = np.where(use, np.exp(-0.5 * (ang / log_scale_weight) ** 2), 0.0)
            dead = ~use.any(axis=1)
            w[dead, 0] = 1.0
            cdf = np.cumsum(w, axis=1)
            u = rng.rand(Wn) * cdf[:, -1]
            pick = (cdf < u[:, None]).sum(axis=1)
            pick = np.minimum(pick, D - 1)
            nxt = cand[np.arange(Wn), pick]
            nxt = np.where(dead | (nxt < 0), cur, nxt)
            hist[:, (t + 1) % path_memory] = nxt
            cur = nxt
            dna[s:e, t + 1] = cur
    return dna.reshape(n_org, A, max_dna_len)


class SynthWorld(object):
    """Arrays of one synthetic world, same field names as the oracle's World."""
    pass


def make_world(size=4096, seed=0, map_params=None, path_params=None, stride=None,
               progress=False, frustums="host"):
    """frustums: "host" = the windowed numpy restatement below, "device" = the CUDA kernel
    (frustums.compute_frustums_device; bit-identical, tests/test_gpu_frustums.py)."""
    mp = dict(DEFAULT_MAP_PARAMS if map_params is None else map_params)
    pp = dict(DEFAULT_PATH_PARAMS if path_params is None else path_params)
    w = SynthWorld()
    u8 = make_map(size, seed)
    w.img = u8.astype(np.float64)          # binary: g in {0,255} -> g/255 in {0,1}
    w.xy = place_waypoints(u8, mp, stride)
    w.row_ptr, w.col = build_graph(u8, w.xy, mp, pp)
    if frustums == "device":
        from . import frustums as _fr
        w.edge_poly, w.edge_theta, w.edge_dist, w.frustum_info = _fr.compute_frustums_device(
            w.img, w.xy, w.row_ptr, w.col, mp, return_info=True)
    else:
        w.edge_poly, w.edge_theta, w.edge_dist = compute_frustums(w.img, w.xy, w.row_ptr, w.col, mp,
                                                                  progress)
    w.map_params, w.path_params = mp, pp
    w.N, w.E = len(w.xy), len(w.col)
    return w


def random_walk_population_torch(xy, row_ptr, col, starts, n_org, walk_len, max_dna_len, seed=1,
                                 path_memory=10, log_scale_weight=0.7, device="cuda",
                                 chunk=1 << 20):
    """Same walk rule as random_walk_population, on torch tensors (benchmark SETUP only:
    it seeds synthetic populations on the GPU box in seconds instead of minutes).
    starts: A start waypoints shared by all organisms, or an [n_org, A] array of per-organism
    starts ("diverse" populations that touch the whole footprint store).
    Returns an int32 torch tensor [n_org, A, max_dna_len] on `device`."""
    import torch
    N = len(xy)
    starts = np.asarray(starts, dtype=np.int64)
    A = starts.shape[-1]
    deg = np.diff(row_ptr)
    D = int(deg.max())
    nbr = -np.ones((N, D), dtype=np.int64)
    for k in range(D):
        has = deg > k
        nbr[has, k] = col[row_ptr[:-1][has] + k]
    disp = xy[np.maximum(nbr, 0)] - xy[:, None, :]
    bearing = np.arctan2(disp[..., 1], disp[..., 0])
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(int(seed))
    nbr_t = torch.from_numpy(nbr).to(dev)
    bearing_t = torch.from_numpy(bearing).to(dev)
    n_steps = min(walk_len, max_dna_len - 1)
    total = n_org * A
    dna = torch.full((total, max_dna_len), -1, dtype=torch.int32, device=dev)
    if starts.ndim == 2:
        assert starts.shape[0] == n_org
        start_all = torch.from_numpy(starts.reshape(-1)).to(dev)
    else:
        start_all = torch.from_numpy(starts).to(dev).repeat(n_org)
    two_pi = 2 * np.pi
    for s in range(0, total, chunk):
        e = min(total, s + chunk)
        Wn = e - s
        cur = start_all[s:e].clone()
        heading = torch.rand(Wn, generator=g, device=dev, dtype=torch.float64) * two_pi - np.pi
        hist = torch.full((Wn, path_memory), -1, dtype=torch.int64, device=dev)
        hist[:, 0] = cur
        dna[s:e, 0] = cur.to(torch.int32)
        ar = torch.arange(Wn, device=dev)
        for t in range(n_steps):
            cand = nbr_t[cur]
            valid = cand >= 0
            fresh = valid & ~(cand[:, :, None] == hist[:, None, :]).any(dim=2)
            use = torch.where(fresh.any(dim=1)[:, None], fresh, valid)
            ang = bearing_t[cur] - heading[:, None]
            ang = ang - two_pi * torch.floor((ang + np.pi) * (0.5 / np.pi))
            w = torch.where(use, torch.exp(-0.5 * (ang / log_scale_weight) ** 2),
                            torch.zeros((), dtype=torch.float64, device=dev))
            dead = ~use.any(dim=1)
            w[dead, 0] = 1.0
            cdf = torch.cumsum(w, dim=1)
            u = torch.rand(Wn, generator=g, device=dev, dtype=torch.float64) * cdf[:, -1]
            pick = (cdf < u[:, None]).sum(dim=1).clamp_(max=D - 1)
            nxt = cand[ar, pick]
            nxt = torch.where(dead | (nxt < 0), cur, nxt)
            hist[:, (t + 1) % path_memory] = nxt
            cur = nxt
            dna[s:e, t + 1] = cur.to(torch.int32)
    return dna.reshape(n_org, A, max_dna_len)


class CsrMappy(object):
    """Mappy-shaped holder of a synthetic world's map-side tables (attribute names of the
    reference's Mappy, mappy.py:14-64, plus per-edge CSR tables instead of dense N x N)."""

    def __init__(self, world):
        mp = world.map_params
        self._img = world.img
        self.shape = world.img.shape
        self._scale = mp['scale']
        self._hall_width = mp['narrowest_hall']
        self._safety_buffer = self._hall_width / 2
        self._rho = mp['rho']
        self._solo_sep_thresh = mp['solo_sep_thresh']
        self._coverage_blend = mp['coverage_blend']
        self._all_waypoints = world.xy
        self._edge_poly, self._edge_theta, self._edge_dist = (world.edge_poly, world.edge_theta,
                                                              world.edge_dist)


class CsrPather(object):
    """PathMaker-shaped holder (pathmaker.py:15-24,157-173) with a CSR graph."""

    def __init__(self, world):
        self._XY = world.xy
        self._csr = (world.row_ptr, world.col)
        self._path_memory = world.path_params.get('path_memory', 10)
        self._log_scale_weight = world.path_params.get('log_scale_weight', 0.7)
        self._crossover_cand = (np.where(np.diff(world.row_ptr) > 2)[0],)

    @property
    def waypoint_locs(self):
        return self._XY
