"""CPU restatement (numpy + cv2) of the reference's fitness / ranking hot path.

TEST INFRASTRUCTURE.  This module is the ORACLE: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it.  The product
package (genetic_coverage_planning_b200) never does; it fails loudly when its CUDA
library is missing.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function below
against outputs of the unmodified reference (/root/reference/src, imported through
oracle/ref_shim.py) captured in tests/golden/*.npz by tests/golden/make_golden.py,
and - when /root/reference is present - against the live reference.  The reference
ships no tests or golden vectors of its own.

Third-party arithmetic the reference calls on this path and that is therefore
called here too, not re-derived: cv2.fillPoly (mappy.py:355), cv2.dilate
(mappy.py:38), np.sum pairwise order, np.linalg.eig (geneticalgorithm.py:193).
Versions pinned by the fixtures: cv2 4.13.0, numpy 2.3.5.

All file:line citations are into /root/reference/src.
"""
import numpy as np
import cv2

inv_2pi = 0.5 / np.pi


def rad_wrap_pi(angle):
    """gori_tools.py:170-172."""
    angle -= 2 * np.pi * np.floor((angle + np.pi) * inv_2pi)
    return angle


# ---------------------------------------------------------------------------
# mappy.py:330-370  Mappy.getCoverage
# ---------------------------------------------------------------------------
def get_coverage(world, dna, return_map=False, return_counts=False):
    """Faithful restatement incl. its cost structure (two map copies, mask rebuild,
    one fillPoly per executed step, two full-map reductions).

    Table reads `table[wpt, nxt]` follow numpy indexing: nxt == -1 addresses column
    N-1; a non-edge reads zeros (F4)."""
    dna = np.asarray(dna)
    num_agents = dna.shape[0]
    cover_map = np.copy(world.img)                                   # :335
    draw_map = np.copy(cover_map)                                    # :336
    buffer_mask = np.logical_and(world.safety_img < 0.3, world.safety_img > 0)  # :337
    rho = world.map_params['rho']
    blend = world.map_params['coverage_blend']
    zero_poly = np.zeros(world.edge_poly.shape[1:], dtype=np.int32)

    travel_cost = np.zeros(num_agents)
    turning_cost = np.zeros(num_agents)
    for agent in range(num_agents):
        prev_theta = 0                                               # :342 (never updated, F5)
        row = dna[agent]
        for idx, wpt in enumerate(row):
            if wpt == -1 or idx == len(row) - 1:                     # :344
                break
            wpt_loc = world.xy[wpt]
            e = world.edge_id(wpt, row[idx + 1])
            if e >= 0:
                wpt_theta = world.edge_theta[e]
                dist = world.edge_dist[e]
                frustum = world.edge_poly[e]
            else:
                wpt_theta = 0.0
                dist = 0.0
                frustum = zero_poly
            delta_theta = rad_wrap_pi(np.float64(wpt_theta - prev_theta))   # :349
            travel_cost[agent] += dist                                       # :351
            turning_cost[agent] += rho * abs(delta_theta / (np.pi / 2)) ** rho   # :352
            center = (int(wpt_loc[1]), int(wpt_loc[0]))                      # :353
            cv2.fillPoly(draw_map, [frustum], 1, offset=center)              # :355

    internal_coverage = (np.sum(draw_map) - world.num_occluded) / \
        (draw_map.size - world.num_occluded)                         # :358
    wall_coverage = np.sum(draw_map[buffer_mask]) / draw_map[buffer_mask].size   # :361
    coverage = blend * wall_coverage + (1 - blend) * internal_coverage          # :364
    if return_counts:
        covered = draw_map != world.img        # pixels whose value changed (g<255 and painted)
        painted = _painted_map(world, dna)
        return (-coverage, travel_cost, turning_cost,
                dict(n_cov_all=int(painted.sum()),
                     n_cov_mask=int((painted & buffer_mask).sum()),
                     changed=int(covered.sum())))
    if return_map:
        return -coverage, travel_cost, turning_cost, draw_map
    return -coverage, travel_cost, turning_cost


def _painted_map(world, dna):
    """Boolean map of every pixel any executed step's fillPoly touches."""
    canvas = np.zeros(world.shape, dtype=np.uint8)
    zero_poly = np.zeros(world.edge_poly.shape[1:], dtype=np.int32)
    for row in np.asarray(dna):
        for idx, wpt in enumerate(row):
            if wpt == -1 or idx == len(row) - 1:
                break
            e = world.edge_id(wpt, row[idx + 1])
            poly = world.edge_poly[e] if e >= 0 else zero_poly
            loc = world.xy[wpt]
            cv2.fillPoly(canvas, [poly], 1, offset=(int(loc[1]), int(loc[0])))
    return canvas.astype(bool)


def traversable(world, dna):
    """New engine output: 1 iff every executed step between two real
    genes is a graph edge (pather._graph[a, b] == 1).  The sentinel step of a padded
    row (next gene == -1, F3) is not a move and is not tested."""
    for row in np.asarray(dna):
        for idx, wpt in enumerate(row):
            if wpt == -1 or idx == len(row) - 1:
                break
            nxt = row[idx + 1]
            if nxt == -1:
                continue
            if world.edge_id(wpt, nxt) < 0:
                return 0
    return 1


# ---------------------------------------------------------------------------
# mappy.py:372-384  Mappy.getDuplicateWPs   (argsort canonicalised to stable, F9)
# ---------------------------------------------------------------------------
def get_duplicate_wps(wpt_sequence):
    unq, unq_idx, unq_cnt = np.unique(wpt_sequence, axis=0, return_inverse=True,
                                      return_counts=True)
    unq_idx = np.asarray(unq_idx).reshape(-1)
    cnt_mask = unq_cnt > 1
    dup_ids = unq[cnt_mask]
    cnt_idx, = np.nonzero(cnt_mask)
    idx_mask = np.isin(unq_idx, cnt_idx)
    idx_idx, = np.nonzero(idx_mask)
    srt_idx = np.argsort(unq_idx[idx_mask], kind='stable')
    dup_idx = np.split(idx_idx[srt_idx], np.cumsum(unq_cnt[cnt_mask])[:-1])
    return len(dup_ids), dup_idx


# ---------------------------------------------------------------------------
# mappy.py:386-437  Mappy.getLoopClosures
# ---------------------------------------------------------------------------
def get_loop_closures(world, dna):
    dna = np.asarray(dna)
    num_agents = dna.shape[0]
    thresh = world.map_params['solo_sep_thresh']
    path_splits = np.zeros(num_agents).astype(int)
    wps = dna[dna != -1]                                             # :390
    wpt_sequence = np.array([wps, np.roll(wps, -1)]).T               # :391
    path_split_idx = 0
    for agent in range(num_agents):                                  # :401-403
        path_split_idx += len(dna[agent][dna[agent] != -1])
        path_splits[agent] = path_split_idx
    lc_mat = np.zeros((num_agents, num_agents))
    if len(wps) == 0:
        return lc_mat
    num_dups, dup_idx = get_duplicate_wps(wpt_sequence)              # :405
    if num_dups > 0:
        for dup in dup_idx:
            if not np.isin(path_splits, dup).any():                  # :409 (F10)
                agent_lc = np.digitize(dup, path_splits)             # :410
                unq_lc = np.unique(agent_lc)
                if len(unq_lc) == 1:
                    if (abs(np.diff(dup)) > thresh).any():           # :413
                        lc_mat[unq_lc[0], unq_lc[0]] += 1
                elif len(unq_lc) == 2:
                    lc_mat[unq_lc[0], unq_lc[1]] += 1
                else:
                    for ii in range(len(unq_lc)):                    # :422-424
                        for jj in range(ii + 1, len(unq_lc)):
                            lc_mat[unq_lc[ii], unq_lc[jj]] += 1
    return lc_mat


def get_loop_closures_direct(world, dna):
    """Independent second statement used to cross-check the
    line-by-line one above: hash-group the (from,to) keys, no numpy set routines."""
    dna = np.asarray(dna)
    A = dna.shape[0]
    thresh = world.map_params['solo_sep_thresh']
    rows = [r[r != -1] for r in dna]
    wps = np.concatenate(rows) if A else np.zeros(0, dtype=int)
    n = len(wps)
    lc = np.zeros((A, A))
    if n == 0:
        return lc
    splits = np.cumsum([len(r) for r in rows])
    split_set = set(int(s) for s in splits)
    groups = {}
    for i in range(n):
        groups.setdefault((int(wps[i]), int(wps[(i + 1) % n])), []).append(i)
    for members in groups.values():
        if len(members) < 2:
            continue
        if any(m in split_set for m in members):
            continue
        owners = sorted(set(int(np.sum(splits <= m)) for m in members))
        if len(owners) == 1:
            if any(b - a > thresh for a, b in zip(members[:-1], members[1:])):
                lc[owners[0], owners[0]] += 1
        else:
            for x in range(len(owners)):
                for y in range(x + 1, len(owners)):
                    lc[owners[x], owners[y]] += 1
    return lc


def num_components(lc_mat):
    """geneticalgorithm.py:183-194 (Laplacian eigenvalue count)."""
    A = lc_mat.shape[0]
    solo = np.diagonal(lc_mat)
    combo = lc_mat - np.diag(solo)
    adjacency = ((combo + combo.T) >= 1).astype(int)
    degree = np.diag(np.sum(adjacency, axis=0))
    eig, _ = np.linalg.eig(degree - adjacency)
    return int(A - np.sum(eig > 1e-6))


# ---------------------------------------------------------------------------
# geneticalgorithm.py:177-201  Organism.calcObj
# ---------------------------------------------------------------------------
def calc_obj(world, dna, org_params=None, detail=False):
    op = org_params if org_params is not None else world.org_params
    coverage, travel_dist, turning_cost = get_coverage(world, dna)
    dist_a, turn_a = travel_dist, turning_cost
    travel_dist = np.sum(travel_dist)
    turning_cost = np.sum(turning_cost)
    lc_mat = get_loop_closures(world, dna)
    solo_lcs = np.diagonal(lc_mat)
    combo_lc_mat = lc_mat - np.diag(solo_lcs)
    combo_lcs = np.sum(combo_lc_mat)
    num_con = num_components(lc_mat)
    travel_cost = (travel_dist + turning_cost) * op['flight_time_scale']
    feasible = not ((solo_lcs < op['min_solo_lcs']).any()
                    or combo_lcs < op['min_comb_lcs'] or num_con > 1)
    raw_cov = coverage
    if not feasible:
        coverage = 0
    if detail:
        return dict(obj=[coverage, travel_cost], neg_coverage=raw_cov, dist=dist_a,
                    turn=turn_a, lc_mat=lc_mat, n_components=num_con,
                    feasible=int(feasible))
    return [coverage, travel_cost]


# ---------------------------------------------------------------------------
# geneticalgorithm.py:86-123  GeneticalGorithm.rackNstack  (maximin fitness)
# ---------------------------------------------------------------------------
def coverage_constraint(gen_num, gen_params):
    alpha = min(1, (gen_num / gen_params['coverage_aging']))          # :98
    c0 = -gen_params['coverage_constr_0']
    cf = -gen_params['coverage_constr_f']
    return (1 - alpha) * c0 + alpha * cf                              # :99-100


def scaled_objectives(objs, gen_num, gen_params):
    constr = coverage_constraint(gen_num, gen_params)
    objs = np.asarray(objs, dtype=np.float64)
    sc = objs.copy()
    sc[:, 1] = np.where(objs[:, 0] > constr, objs[:, 1] + 1000.0, objs[:, 1])   # :102-107
    return sc


def rack_n_stack_loop(objs, gen_num, gen_params):
    """Literal double loop (:109-120); O(N^2) python, small N only."""
    sc = scaled_objectives(objs, gen_num, gen_params)
    fits = []
    for ii in range(len(sc)):
        comp_min = []
        for jj in range(len(sc)):
            if ii == jj:
                continue
            comp_min.append(np.min([sc[ii, 0] - sc[jj, 0], sc[ii, 1] - sc[jj, 1]]))
        fits.append(np.max(comp_min))
    return np.array(fits)


def rack_n_stack(objs, gen_num, gen_params, block=1024):
    """Same arithmetic (fp64 sub / min / max are exact selections), vectorised in
    row blocks so N = 8,192 finishes in seconds."""
    sc = scaled_objectives(objs, gen_num, gen_params)
    n = len(sc)
    fit = np.empty(n)
    for s in range(0, n, block):
        e = min(n, s + block)
        d0 = sc[s:e, None, 0] - sc[None, :, 0]
        d1 = sc[s:e, None, 1] - sc[None, :, 1]
        m = np.minimum(d0, d1)
        m[np.arange(e - s), np.arange(s, e)] = -np.inf
        fit[s:e] = m.max(axis=1)
    return fit


def arena_order(fit):
    """gori_tools.py:232-236 with the sort canonicalised to stable (F15)."""
    return np.argsort(np.asarray(fit), kind='stable')


# ---------------------------------------------------------------------------
# gori_tools.py:207-229  lowVarSample  (indices instead of deep copies)
# ---------------------------------------------------------------------------
def low_var_sample_indices(fitnesses, pressure, r):
    """`r` is the single U(0, 1/M) draw (:217)."""
    log_w = -np.array(fitnesses)
    log_w = log_w - np.max(log_w)
    base = pressure + 1
    w = base ** log_w
    w = w / np.sum(w)
    M = len(fitnesses)
    c = w[0]
    i = 0
    out = []
    for m in range(M):
        u = r + m / M
        while u > c:
            i += 1
            c = c + w[i]
        out.append(i)
    return np.array(out)


def rack_n_stack_sweep(objs, gen_num, gen_params):
    """O(N log N) evaluation of the same maximin fitness for two objectives, used to pin the
    algorithm of the CUDA kernel k_maximin_sweep on the CPU.

    fit_i = max_{j != i} min(rn(a_i - a_j), rn(b_i - b_j)).  With gladiators sorted by a
    ascending, x_(k) = rn(a_i - a_(k)) is non-increasing in k and
    Y(k) = rn(b_i - min_{j <= k, j != i} b_(j)) is non-decreasing, and
    max_j min(x_j, y_j) = max_{k != pos(i)} min(x_(k), Y(k)); the maximum sits next to the
    crossing, found by bisection.  Only differences of actual elements are evaluated, so the
    result is bit-identical to the double loop (geneticalgorithm.py:109-120)."""
    sc = scaled_objectives(objs, gen_num, gen_params)
    n = len(sc)
    a, b = sc[:, 0] + 0.0, sc[:, 1] + 0.0          # +0.0 folds -0.0 into +0.0
    perm = np.argsort(a, kind='stable')
    sa, sb = a[perm], b[perm]
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)
    # prefix (min1, argmin1, min2) of b along the sorted order; min2 excludes argmin1
    m1 = np.empty(n); i1 = np.empty(n, dtype=np.int64); m2 = np.empty(n)
    cm1, ci1, cm2 = np.inf, -1, np.inf
    for k in range(n):
        v = sb[k]
        if v < cm1:
            cm1, ci1, cm2 = v, k, cm1
        elif v < cm2:
            cm2 = v
        m1[k], i1[k], m2[k] = cm1, ci1, cm2
    fit = np.empty(n)
    with np.errstate(invalid='ignore'):
        for i in range(n):
            p = inv[i]

            def xy(k):
                pm = m2[k] if i1[k] == p else m1[k]
                return a[i] - sa[k], b[i] - pm          # b_i - inf = -inf for an empty prefix

            lo, hi = -1, n - 1                          # largest k with x_(k) >= Y(k), or -1
            while lo < hi:
                mid = (lo + hi + 1) >> 1
                x, y = xy(mid)
                if x >= y:
                    lo = mid
                else:
                    hi = mid - 1
            best = -np.inf
            kl = lo - 1 if lo == p else lo
            kr = lo + 2 if lo + 1 == p else lo + 1
            for k in (kl, kr):
                if 0 <= k < n and k != p:
                    x, y = xy(k)
                    best = max(best, min(x, y))
            fit[i] = best
    return fit
