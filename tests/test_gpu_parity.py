"""GPU parity: CUDA path (through the C ABI) vs golden vectors of the unmodified
reference and vs the oracle on seeded inputs.  Integer outputs bit-exact; float costs
bit-exact (tolerance 0); -coverage within 1e-12 on the gray reference maps and
bit-exact on binary maps (SURVEY.md section 8a parity rules)."""
import os

import numpy as np
import pytest
import torch

import fitness_oracle as fo
from world import World

pytestmark = pytest.mark.gpu

COV_TOL_GRAY = 1e-12


def _engine(world, org_params=None, num_agents=None):
    from genetic_coverage_planning_b200 import packing
    from genetic_coverage_planning_b200.engine import FitnessEngine
    pw = packing.pack_world(world.img, world.xy, world.row_ptr, world.col, world.edge_poly,
                            world.edge_theta, world.edge_dist, world.map_params)
    op = org_params or world.org_params
    return FitnessEngine(pw, op, num_agents=num_agents)


@pytest.fixture(scope="module")
def eng_big(world_big):
    return _engine(world_big)


def _check_against_golden(eng, z, exact_cov=False):
    """Both evaluation paths against the reference's vectors: the staged general kernels (all
    pixel counters) and the objectives-only path (the fused kernel when coverage_blend == 1)."""
    dna = z['dna'].astype(np.int32)
    _check_result(eng.evaluate(dna, detail=True, pixel_counts=False), z, exact_cov, False, eng.org_params)
    general = eng.evaluate(dna)                      # fused kernel over the unmasked store, all pixel counters
    _check_result(general, z, exact_cov, True, eng.org_params)
    eng.set_option("fused_eval", 0)                  # the staged general kernels (k_coverage)
    try:
        staged = eng.evaluate(dna)
    finally:
        eng.set_option("fused_eval", 1)
    _check_result(staged, z, exact_cov, True, eng.org_params)
    for name in ("obj", "neg_cov", "dist", "turn", "lc_mat", "flags", "cov"):
        assert torch.equal(getattr(general, name), getattr(staged, name)), name
    obj_only = eng.evaluate(dna, detail=False)
    assert np.max(np.abs(obj_only.obj.cpu().numpy() - z['obj'])) <= (0 if exact_cov else COV_TOL_GRAY)


def _check_result(r, z, exact_cov, counters, op):
    dist, turn = r.dist.cpu().numpy(), r.turn.cpu().numpy()
    assert np.array_equal(dist, z['dist']), "per-agent travel distance not bit-exact"
    assert np.array_equal(turn, z['turn']), "per-agent turning cost not bit-exact"
    if counters:
        cov = r.cov.cpu().numpy()
        assert np.array_equal(cov[:, 5], z['n_painted'])
        assert np.array_equal(cov[:, 4], z['n_painted_mask'])
    assert np.array_equal(r.lc_mat.cpu().numpy(), z['lc'].astype(np.int32))
    flags = r.flags.cpu().numpy()
    assert np.array_equal(flags[:, 2], z['ncomp'])
    assert not flags[:, 3].any()
    obj = r.obj.cpu().numpy()
    assert np.array_equal(obj[:, 1], z['obj'][:, 1]), "travel cost objective not bit-exact"
    neg = r.neg_cov.cpu().numpy()
    if exact_cov:
        assert np.array_equal(neg, z['neg_cov'])
        assert np.array_equal(obj[:, 0], z['obj'][:, 0])
    else:
        assert np.max(np.abs(neg - z['neg_cov'])) <= COV_TOL_GRAY
        assert np.max(np.abs(obj[:, 0] - z['obj'][:, 0])) <= COV_TOL_GRAY
    # feasibility (geneticalgorithm.py:197-200) re-derived for EVERY organism from the reference's own
    # loop-closure matrix and component count, and - where the reference's output shows it (its
    # coverage is zeroed exactly when infeasible) - from its objectives
    lc_ref = z['lc'].astype(np.int64)
    solo = np.diagonal(lc_ref, axis1=1, axis2=2)
    combo = lc_ref.sum(axis=(1, 2)) - solo.sum(axis=1)
    feasible_ref = ~((solo < op['min_solo_lcs']).any(axis=1) | (combo < op['min_comb_lcs']) | (z['ncomp'] > 1))
    assert np.array_equal(flags[:, 0] == 1, feasible_ref)
    shown = z['neg_cov'] != 0
    assert np.array_equal(feasible_ref[shown], z['obj'][shown, 0] != 0)
    if 'feasible' in z:
        assert np.array_equal(flags[:, 0], z['feasible'])


def test_eval_golden_big(eng_big, golden_dir):
    _check_against_golden(eng_big, np.load(os.path.join(golden_dir, "eval_big.npz")))


def test_eval_golden_evolved_feasible(eng_big, world_small, golden_dir):
    """Populations evolved by the unmodified reference (make_golden_feasible.py): mostly FEASIBLE
    organisms, i.e. the branch that keeps the coverage objective, on both reference maps."""
    zb = np.load(os.path.join(golden_dir, "eval_big_evolved.npz"))
    zs = np.load(os.path.join(golden_dir, "eval_small_evolved.npz"))
    assert zb['feasible'].mean() > 0.5 and zs['feasible'].mean() > 0.5
    _check_against_golden(eng_big, zb)
    _check_against_golden(_engine(world_small), zs)


def test_edge_cases_golden(eng_big, golden_dir):
    _check_against_golden(eng_big, np.load(os.path.join(golden_dir, "edge_big.npz")))


def test_eval_golden_small(world_small, golden_dir):
    eng = _engine(world_small)
    _check_against_golden(eng, np.load(os.path.join(golden_dir, "eval_small.npz")))


def test_blend_golden(golden_dir):
    w = World.load_npz(os.path.join(golden_dir, "world_big.npz"))
    w.map_params['coverage_blend'] = 0.1
    eng = _engine(w)
    _check_against_golden(eng, np.load(os.path.join(golden_dir, "eval_big_blend01.npz")))


@pytest.mark.parametrize("A", [1, 3])
def test_agent_counts_golden(golden_dir, A):
    z = np.load(os.path.join(golden_dir, "eval_big_A%d.npz" % A))
    w = World.load_npz(os.path.join(golden_dir, "world_big.npz"))
    w.map_params['solo_sep_thresh'] = int(z['solo_sep_thresh'])
    op = dict(w.org_params)
    op.update(min_solo_lcs=int(z['min_solo_lcs']), min_comb_lcs=int(z['min_comb_lcs']),
              max_dna_len=int(z['max_dna_len']), start_idx=w.org_params['start_idx'][:A])
    eng = _engine(w, op, num_agents=A)
    _check_against_golden(eng, z)


# ---- synthetic BINARY world evaluated by the unmodified reference (make_golden_synth.py): the
#      fused kernel's binary instantiation and the staged kernels against the reference itself,
#      -coverage bit-exact
@pytest.fixture(scope="module")
def world_synth_ref(golden_dir):
    return World.load_npz(os.path.join(golden_dir, "world_synth.npz"))


def test_eval_golden_synth_binary(world_synth_ref, golden_dir):
    eng = _engine(world_synth_ref)
    _check_against_golden(eng, np.load(os.path.join(golden_dir, "eval_synth.npz")), exact_cov=True)
    _check_against_golden(eng, np.load(os.path.join(golden_dir, "edge_synth.npz")), exact_cov=True)


def test_agent_count_golden_synth_binary(world_synth_ref, golden_dir):
    z = np.load(os.path.join(golden_dir, "eval_synth_A3.npz"))
    w = World.load_npz(os.path.join(golden_dir, "world_synth.npz"))
    w.map_params['solo_sep_thresh'] = int(z['solo_sep_thresh'])
    op = dict(w.org_params)
    op.update(min_solo_lcs=int(z['min_solo_lcs']), min_comb_lcs=int(z['min_comb_lcs']),
              max_dna_len=int(z['max_dna_len']), start_idx=w.org_params['start_idx'][:3])
    eng = _engine(w, op, num_agents=3)
    _check_against_golden(eng, z, exact_cov=True)


def _random_dna(world, P, A, L, seed, p_nonedge=0.02):
    """Random walks on the graph with ragged lengths, a few teleports (non-edges),
    some full-length rows, some empty rows."""
    rng = np.random.RandomState(seed)
    deg = np.diff(world.row_ptr)
    dna = -np.ones((P, A, L), dtype=np.int32)
    for p in range(P):
        for a in range(A):
            n = int(rng.choice([0, 1, 2, L, rng.randint(3, L + 1)], p=[.02, .02, .02, .1, .84]))
            if n == 0:
                continue
            w = int(rng.randint(world.N))
            for i in range(n):
                dna[p, a, i] = w
                if rng.rand() < p_nonedge or deg[w] == 0:
                    w = int(rng.randint(world.N))
                else:
                    w = int(world.col[world.row_ptr[w] + rng.randint(deg[w])])
    return dna


def test_eval_vs_oracle_random(eng_big, world_big):
    A, L = eng_big.A, eng_big.L
    dna = _random_dna(world_big, 96, A, L, seed=5)
    r = eng_big.evaluate(dna)
    dist, turn, cov = r.dist.cpu().numpy(), r.turn.cpu().numpy(), r.cov.cpu().numpy()
    lc, obj, flags = r.lc_mat.cpu().numpy(), r.obj.cpu().numpy(), r.flags.cpu().numpy()
    neg = r.neg_cov.cpu().numpy()
    for k in range(len(dna)):
        d = dna[k].astype(int)
        c, dd, tt, cnt = fo.get_coverage(world_big, d, return_counts=True)
        assert np.array_equal(dd, dist[k]) and np.array_equal(tt, turn[k]), k
        assert cnt['n_cov_all'] == cov[k, 5] and cnt['n_cov_mask'] == cov[k, 4], k
        assert abs(c - neg[k]) <= COV_TOL_GRAY
        det = fo.calc_obj(world_big, d, detail=True)
        assert np.array_equal(det['lc_mat'].astype(np.int32), lc[k]), k
        assert det['n_components'] == flags[k, 2] and det['feasible'] == flags[k, 0], k
        assert det['obj'][1] == obj[k, 1]
        assert abs(det['obj'][0] - obj[k, 0]) <= COV_TOL_GRAY
        assert fo.traversable(world_big, d) == flags[k, 1], k


@pytest.mark.parametrize("A,L", [(1, 3), (1, 33), (2, 4), (3, 64), (5, 97), (7, 150), (8, 250), (9, 40), (16, 31), (32, 20)])
def test_shape_sweep_vs_oracle(world_synth_ref, A, L):
    """Odd genome shapes (runtime shared-memory layouts, 1..32 agents, rows of 2..250 genes) with ragged,
    empty, full and teleporting rows plus interior -1 holes: objectives-only kernel (fused), the staged
    kernels, the general kernels and int16 chromosomes against the oracle, bit-exact (binary map)."""
    w = world_synth_ref
    op = dict(w.org_params)
    op.update(max_dna_len=L, start_idx=list(w.org_params['start_idx'][:1]) * A, min_solo_lcs=0, min_comb_lcs=0)
    eng = _engine(w, op, num_agents=A)
    dna = _random_dna(w, 24, A, L, seed=100 * A + L, p_nonedge=0.05)
    rng = np.random.RandomState(A + L)
    for k in range(0, 24, 3):                                   # interior holes (mappy.py:343-345 vs :390)
        a, i = int(rng.randint(A)), int(rng.randint(L))
        dna[k, a, i] = -1
    full = eng.evaluate(dna)                                     # general kernels, all outputs
    fused = eng.evaluate(dna, detail=False)
    eng.set_option("fused_eval", 0)
    staged = eng.evaluate(dna, detail=False)
    eng.set_option("fused_eval", 1)
    i16 = eng.evaluate16(torch.from_numpy(dna.astype(np.int16)).to(eng.device))
    for r in (fused, staged, i16):
        assert torch.equal(r.obj, full.obj) and torch.equal(r.flags, full.flags)
    obj, flags, lc = full.obj.cpu().numpy(), full.flags.cpu().numpy(), full.lc_mat.cpu().numpy()
    ow = World(w.img, w.xy, w.row_ptr, w.col, w.edge_poly, w.edge_theta, w.edge_dist, w.map_params, op,
               name="sweep")
    for k in range(len(dna)):
        det = fo.calc_obj(ow, dna[k].astype(int), detail=True)
        assert det['obj'][0] == obj[k, 0] and det['obj'][1] == obj[k, 1], (A, L, k)
        assert np.array_equal(det['lc_mat'].astype(np.int32), lc[k]), (A, L, k)
        assert det['feasible'] == flags[k, 0] and det['n_components'] == flags[k, 2], (A, L, k)
        assert fo.traversable(ow, dna[k].astype(int)) == flags[k, 1], (A, L, k)


def test_shard_equivalence_and_host_path(eng_big, golden_dir):
    import torch
    z = np.load(os.path.join(golden_dir, "eval_big.npz"))
    dna = np.concatenate([z['dna'].astype(np.int32)] * 3)
    full = eng_big.evaluate(dna)
    parts = [eng_big.evaluate(dna[s:s + 50]) for s in range(0, len(dna), 50)]
    for name in ("obj", "dist", "turn", "cov", "lc_mat", "flags"):
        a = getattr(full, name).cpu().numpy()
        b = np.concatenate([getattr(p, name).cpu().numpy() for p in parts])
        assert np.array_equal(a, b), name
    obj_h, flags_h = eng_big.evaluate_host(torch.from_numpy(dna).pin_memory())
    assert np.array_equal(obj_h.numpy(), full.obj.cpu().numpy())
    assert np.array_equal(flags_h.numpy(), full.flags.cpu().numpy())


@pytest.mark.parametrize("algo", [1, 2])      # 1 = O(N^2) dominance tiles, 2 = sort-based sweep
def test_maximin_and_order_golden(eng_big, golden_dir, algo):
    import torch
    import ref_shim
    z = np.load(os.path.join(golden_dir, "rank.npz"))
    _, _, _, gp = ref_shim.driver_params(6)
    eng_big.set_option("rank_algo", algo)
    try:
        for gen in (0, 17, 80, 200):
            constr = fo.coverage_constraint(gen, gp)
            fit, sc = eng_big.maximin(torch.from_numpy(z['objs']), constr, return_scaled=True)
            assert np.array_equal(sc.cpu().numpy(), z['sc_gen%d' % gen])
            assert np.array_equal(fit.cpu().numpy(), z['fit_gen%d' % gen])
            order, nondom = eng_big.arena_order(fit)
            assert np.array_equal(order.cpu().numpy(), z['order_gen%d' % gen])
            assert np.array_equal(nondom.cpu().numpy() == 1, z['fit_gen%d' % gen] < 0)
        fit = eng_big.maximin(torch.from_numpy(z['syn_objs']), fo.coverage_constraint(40, gp))
        assert np.array_equal(fit.cpu().numpy(), z['syn_fit_gen40'])
    finally:
        eng_big.set_option("rank_algo", 0)


@pytest.mark.parametrize("algo", [0, 1, 2])
@pytest.mark.parametrize("n", [2, 3, 257, 5000, 20011])
def test_maximin_vs_oracle_large_with_ties(eng_big, algo, n):
    import torch
    import ref_shim
    _, _, _, gp = ref_shim.driver_params(6)
    rs = np.random.RandomState(2 + n)
    objs = np.stack([-rs.randint(0, 300, n) / 300.0, rs.rand(n) * 3 + 0.5], 1)
    objs[rs.rand(n) < 0.5, 0] = 0.0                      # infeasible organisms tie (F15)
    if n >= 5000:
        objs[100:140] = objs[200:240]                    # exact duplicates
        objs[300:400, 1] = objs[300, 1]                  # ties in the second objective
    if n > 3:
        objs[-1, 0] = -0.0                               # signed zero must not matter
    eng_big.set_option("rank_algo", algo)
    try:
        for gen in (0, 50):
            constr = fo.coverage_constraint(gen, gp)
            want = fo.rack_n_stack(objs, gen, gp)
            fit = eng_big.maximin(torch.from_numpy(objs), constr)
            assert np.array_equal(fit.cpu().numpy(), want)
            order, _ = eng_big.arena_order(fit)
            assert np.array_equal(order.cpu().numpy(), fo.arena_order(want))
            # row-sharded evaluation (multi-GPU layout) gives the same rows
            lo, hi = n // 4, max(n // 4 + 1, (2 * n) // 3)
            part = eng_big.maximin(torch.from_numpy(objs), constr, row_range=(lo, hi))
            assert np.array_equal(part.cpu().numpy()[lo:hi], want[lo:hi])
    finally:
        eng_big.set_option("rank_algo", 0)


def test_radix_order_is_stable_on_adversarial_keys(eng_big):
    """arena order = stable sort by (fit, index): negative / positive / zero / huge / tiny fits,
    long runs of equal keys, n not a multiple of the sort tile."""
    import torch
    rs = np.random.RandomState(11)
    n = 70001
    fit = np.concatenate([rs.randn(20000) * 1e-3, rs.randn(20000) * 1e6, np.zeros(5000),
                          -np.zeros(5000), np.full(10000, -0.25), rs.randint(-3, 3, 10001) * 0.5])
    assert len(fit) == n
    rs.shuffle(fit)
    for algo in (1, 2):
        eng_big.set_option("rank_algo", algo)
        order, nondom = eng_big.arena_order(torch.from_numpy(fit))
        eng_big.set_option("rank_algo", 0)
        assert np.array_equal(order.cpu().numpy(), np.argsort(fit, kind='stable'))
        assert np.array_equal(nondom.cpu().numpy() == 1, fit < 0)


@pytest.mark.parametrize("n", [1, 2, 2048, 2049, 70001, 302000, 700001, 1300000])
def test_radix_sort_one_launch_per_pass_equals_staged_passes(eng_big, n):
    """The default radix sort (<= 2^20 pairs: one launch per pass, tile offsets by decoupled look-back
    with ticketed tiles) and the histogram / scan / scatter variant give the stable (fit, index) order:
    one tile, a partial tile, more tiles than SMs can hold at once, and a size beyond the one-launch
    limit; heavy ties (the order inside a tie group is what stability is about)."""
    import torch
    rs = np.random.RandomState(n % 1000)
    fit = np.where(rs.rand(n) < 0.4, rs.randint(-4, 4, n) * 0.125, rs.randn(n))
    want = np.argsort(fit, kind='stable')
    for one in (1, 0):
        eng_big.set_option("sort_onesweep", one)
        try:
            order, nondom = eng_big.arena_order(torch.from_numpy(fit))
        finally:
            eng_big.set_option("sort_onesweep", 1)
        assert np.array_equal(order.cpu().numpy(), want), one
        assert np.array_equal(nondom.cpu().numpy() == 1, fit < 0)


# ---------------------------------------------------------------------------------------
# binary synthetic world: -coverage must be BIT-exact, and the objectives-only fast path
# (pre-masked footprint store) must agree with the general kernel and the oracle.
# ---------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def synth_setup():
    from genetic_coverage_planning_b200 import packing, synth, worlds
    from genetic_coverage_planning_b200.engine import FitnessEngine
    sw = synth.make_world(1024, seed=5)
    pw = packing.pack_world(sw.img, sw.xy, sw.row_ptr, sw.col, sw.edge_poly, sw.edge_theta,
                            sw.edge_dist, sw.map_params)
    starts = synth.spread_starts(sw.xy, 6, 1024)
    op = worlds.org_params_for(dict(L=200), starts)
    ow = World(sw.img, sw.xy, sw.row_ptr, sw.col, sw.edge_poly, sw.edge_theta, sw.edge_dist,
               sw.map_params, op, name="synth1024")
    eng = FitnessEngine(pw, op, num_agents=6)
    return sw, pw, ow, op, eng, starts


def test_binary_world_bit_exact_and_fast_path(synth_setup):
    from genetic_coverage_planning_b200 import synth
    sw, pw, ow, op, eng, starts = synth_setup
    assert pw.is_binary and eng.has_masked
    walks = synth.random_walk_population(sw.xy, sw.row_ptr, sw.col, starts, 40, 150, 200, seed=9)
    dna = np.concatenate([walks, _random_dna(ow, 40, 6, 200, seed=10)]).astype(np.int32)
    full = eng.evaluate(dna, detail=True)        # general kernel, all counters
    fast = eng.evaluate(dna, detail=False)       # pre-masked fast kernel
    assert np.array_equal(full.obj.cpu().numpy(), fast.obj.cpu().numpy())
    assert np.array_equal(full.flags.cpu().numpy(), fast.flags.cpu().numpy())
    obj, neg = full.obj.cpu().numpy(), full.neg_cov.cpu().numpy()
    cov = full.cov.cpu().numpy()
    for k in range(len(dna)):
        d = dna[k].astype(int)
        c, dd, tt, cnt = fo.get_coverage(ow, d, return_counts=True)
        assert c == neg[k], "binary map: -coverage must be bit-exact"
        assert cnt['n_cov_all'] == cov[k, 5] and cnt['n_cov_mask'] == cov[k, 4]
        o = fo.calc_obj(ow, d)
        assert o[0] == obj[k, 0] and o[1] == obj[k, 1]
    # stage API: masked kernel count == general kernel mask count
    edge_ids, step_info, dist, turn, flags = eng.path_costs(dna, with_step_info=True)
    cov_m, _ = eng.coverage_masked(step_info)
    cov_g, _ = eng.coverage(edge_ids)
    assert np.array_equal(cov_m.cpu().numpy()[:, 4], cov_g.cpu().numpy()[:, 4])


def test_fused_kernel_equals_staged_kernels(synth_setup):
    """The one-kernel fast path (gcp_fused.cu) against the staged path_costs / coverage_masked /
    loop_closures kernels and the general kernel: every output identical, ragged rows,
    non-edges, empty rows and out-of-range genes included."""
    from genetic_coverage_planning_b200 import synth
    sw, pw, ow, op, eng, starts = synth_setup
    walks = synth.random_walk_population(sw.xy, sw.row_ptr, sw.col, starts, 300, 150, 200, seed=19)
    dna = np.concatenate([walks, _random_dna(ow, 212, 6, 200, seed=20)]).astype(np.int32)
    dna[7, 2, 5] = sw.N + 3                          # out-of-range gene -> error flag, no crash
    dna[9, 0, 3] = -7
    fused = eng.evaluate(dna, detail=True, pixel_counts=False)
    eng.set_option("fused_eval", 0)
    try:
        staged = eng.evaluate(dna, detail=True, pixel_counts=False)
    finally:
        eng.set_option("fused_eval", 1)
    general = eng.evaluate(dna, detail=True, pixel_counts=True)          # fused kernel, unmasked store
    eng.set_option("fused_eval", 0)
    try:
        general_staged = eng.evaluate(dna, detail=True, pixel_counts=True)
    finally:
        eng.set_option("fused_eval", 1)
    ok = np.ones(len(dna), dtype=bool)
    ok[[7, 9]] = False
    for name in ("obj", "neg_cov", "dist", "turn", "lc_mat", "flags"):
        a, b, c, d = (getattr(r, name).cpu().numpy() for r in (fused, staged, general, general_staged))
        assert np.array_equal(a[ok], b[ok]), name
        assert np.array_equal(a[ok], c[ok]), name
        assert np.array_equal(a[ok], d[ok]), name
    assert np.array_equal(general.cov.cpu().numpy()[ok], general_staged.cov.cpu().numpy()[ok])
    assert fused.flags.cpu().numpy()[7, 3] & 1 and fused.flags.cpu().numpy()[9, 3] & 1
    for k in (0, 1, 300, 301, 400):
        o = fo.calc_obj(ow, dna[k].astype(int))
        got = fused.obj.cpu().numpy()[k]
        assert o[0] == got[0] and o[1] == got[1]


def test_sort_based_loop_closures(eng_big, world_big, golden_dir):
    """gcp_lcbig.cu (whole-population radix sort of edge keys, used when A*L outgrows shared
    memory) against the golden vectors of the reference and the shared-memory kernel."""
    z = np.load(os.path.join(golden_dir, "eval_big.npz"))
    e = np.load(os.path.join(golden_dir, "edge_big.npz"))
    dna = np.concatenate([z['dna'], e['dna'], _random_dna(world_big, 300, eng_big.A, eng_big.L, seed=31)]
                         ).astype(np.int32)
    want = eng_big.evaluate(dna)
    n = len(z['dna'])
    for mode in (1, 2):                      # 1: radix sort of edge keys, 2: hash-partitioned shared memory
        eng_big.set_option("force_lc_big", mode)
        try:
            got = eng_big.evaluate(dna)
        finally:
            eng_big.set_option("force_lc_big", 0)
        for name in ("obj", "neg_cov", "lc_mat", "flags"):
            assert np.array_equal(getattr(got, name).cpu().numpy(), getattr(want, name).cpu().numpy()), (mode, name)
        assert np.array_equal(got.lc_mat.cpu().numpy()[:n], z['lc'].astype(np.int32))


def test_many_agents_long_rows(synth_setup):
    """C4-style shape (16 agents x 640 genes = 10,240 genes per organism): staged kernels with
    block partitions + sort-based loop closures, bit-exact against the oracle."""
    from genetic_coverage_planning_b200 import synth, worlds
    from genetic_coverage_planning_b200.engine import FitnessEngine
    sw, pw, ow, op, eng, starts = synth_setup
    A, L = 16, 640
    st = synth.spread_starts(sw.xy, A, 1024)
    op2 = worlds.org_params_for(dict(L=L), st)
    eng2 = FitnessEngine(pw, op2, num_agents=A)
    dna = synth.random_walk_population(sw.xy, sw.row_ptr, sw.col, st, 5, 520, L, seed=12).astype(np.int32)
    r = eng2.evaluate(dna, detail=True, pixel_counts=False)       # partitioned-hash loop closures
    assert not r.flags.cpu().numpy()[:, 3].any()
    eng2.set_option("force_lc_big", 1)                            # radix-sort formulation
    r1 = eng2.evaluate(dna, detail=True, pixel_counts=False)
    eng2.set_option("force_lc_big", 0)
    assert np.array_equal(r1.lc_mat.cpu().numpy(), r.lc_mat.cpu().numpy())
    assert np.array_equal(r1.obj.cpu().numpy(), r.obj.cpu().numpy())
    ow2 = World(sw.img, sw.xy, sw.row_ptr, sw.col, sw.edge_poly, sw.edge_theta, sw.edge_dist,
                sw.map_params, op2, name="synth-16x640")
    for k in range(2):
        det = fo.calc_obj(ow2, dna[k].astype(int), detail=True)
        assert det['obj'][0] == r.obj.cpu().numpy()[k, 0]
        assert det['obj'][1] == r.obj.cpu().numpy()[k, 1]
        assert np.array_equal(det['lc_mat'].astype(np.int32), r.lc_mat.cpu().numpy()[k])


def test_host_pipeline_long_genomes_equals_device_call(synth_setup):
    """gcp_eval_host overlaps neighbouring chunks on three streams; the long-genome kernels keep
    bucket records in per-context scratch, one stretch per pipeline slot.  16 agents x 640 genes over
    enough organisms for the slots to wrap (chunks of 1,024 / 2,048 / 4,096 / 4,096 ...): the host call
    must return what one device-resident call returns."""
    import torch
    from genetic_coverage_planning_b200 import synth, worlds
    from genetic_coverage_planning_b200.engine import FitnessEngine
    sw, pw, ow, op, eng, starts = synth_setup
    A, L = 16, 640
    st = synth.spread_starts(sw.xy, A, 1024)
    eng2 = FitnessEngine(pw, worlds.org_params_for(dict(L=L), st), num_agents=A)
    base = synth.random_walk_population(sw.xy, sw.row_ptr, sw.col, st, 48, 520, L, seed=21).astype(np.int32)
    base[::5, 3, 200:] = -1                                       # ragged rows, different work per organism
    pick = np.random.default_rng(3).integers(0, len(base), 12000)
    dna = np.ascontiguousarray(base[pick])
    want = eng2.evaluate(dna)
    obj_h, flags_h = eng2.evaluate_host(torch.from_numpy(dna).pin_memory())
    assert np.array_equal(obj_h.numpy(), want.obj.cpu().numpy())
    assert np.array_equal(flags_h.numpy(), want.flags.cpu().numpy())
    one = eng2.evaluate(base)                                     # and each copy equals its original
    assert np.array_equal(want.obj.cpu().numpy(), one.obj.cpu().numpy()[pick])


def test_maximum_genome(synth_setup):
    """The largest chromosome the ABI accepts (32 agents x 2,047 genes = 65,504 <= 65,535):
    staged kernels, many hash partitions, bit-exact against the oracle."""
    from genetic_coverage_planning_b200 import synth, worlds
    from genetic_coverage_planning_b200.engine import FitnessEngine
    sw, pw, ow, op, eng, starts = synth_setup
    A, L = 32, 2047
    st = synth.spread_starts(sw.xy, A, 1024)
    op2 = worlds.org_params_for(dict(L=L), st)
    eng2 = FitnessEngine(pw, op2, num_agents=A)
    dna = synth.random_walk_population(sw.xy, sw.row_ptr, sw.col, st, 2, 1990, L, seed=13).astype(np.int32)
    dna[1, 5, 100:] = -1                                         # ragged
    dna[1, 9, :] = -1                                            # an empty row
    r = eng2.evaluate(dna, detail=True, pixel_counts=False)
    assert not r.flags.cpu().numpy()[:, 3].any()
    ow2 = World(sw.img, sw.xy, sw.row_ptr, sw.col, sw.edge_poly, sw.edge_theta, sw.edge_dist,
                sw.map_params, op2, name="synth-32x2047")
    for k in range(2):
        det = fo.calc_obj(ow2, dna[k].astype(int), detail=True)
        assert det['obj'][0] == r.obj.cpu().numpy()[k, 0]
        assert det['obj'][1] == r.obj.cpu().numpy()[k, 1]
        assert np.array_equal(det['lc_mat'].astype(np.int32), r.lc_mat.cpu().numpy()[k])
        assert det['n_components'] == r.flags.cpu().numpy()[k, 2]
    with pytest.raises(Exception):                               # one gene more is refused, not truncated
        FitnessEngine(pw, worlds.org_params_for(dict(L=2048), st), num_agents=A)


def test_int16_genes_equal_int32(synth_setup):
    """gcp_eval16 / gcp_eval_host16 (half-width chromosomes) give the int32 results bit for bit."""
    import torch
    from genetic_coverage_planning_b200 import synth
    sw, pw, ow, op, eng, starts = synth_setup
    walks = synth.random_walk_population(sw.xy, sw.row_ptr, sw.col, starts, 5000, 150, 200, seed=23)
    dna = np.concatenate([walks, _random_dna(ow, 121, 6, 200, seed=24)]).astype(np.int32)
    want = eng.evaluate(dna, detail=False)
    d16 = torch.from_numpy(dna.astype(np.int16))
    got = eng.evaluate16(d16.cuda())
    assert np.array_equal(got.obj.cpu().numpy(), want.obj.cpu().numpy())
    assert np.array_equal(got.flags.cpu().numpy(), want.flags.cpu().numpy())
    obj_h, flags_h = eng.evaluate_host(d16.pin_memory())
    assert np.array_equal(obj_h.numpy(), want.obj.cpu().numpy())
    assert np.array_equal(flags_h.numpy(), want.flags.cpu().numpy())
    obj_h, flags_h = eng.evaluate_host(torch.from_numpy(dna).pin_memory())
    assert np.array_equal(obj_h.numpy(), want.obj.cpu().numpy())


def test_out_of_range_genes_raise_the_error_flag(synth_setup):
    """A gene outside [-1, N) is reported in flags[:, 3] bit 0 by every evaluation path (fused int32 /
    int16 chromosomes, staged kernels) and raise_on_error() turns it into an exception; clean
    organisms of the same launch are untouched."""
    import torch
    from genetic_coverage_planning_b200 import synth
    from genetic_coverage_planning_b200.engine import GcpError
    sw, pw, ow, op, eng, starts = synth_setup
    dna = synth.random_walk_population(sw.xy, sw.row_ptr, sw.col, starts, 6, 150, 200, seed=31).astype(np.int32)
    clean = eng.evaluate(dna, detail=False)
    bad = dna.copy()
    bad[1, 2, 40] = sw.N + 5
    bad[3, 0, 7] = -7
    bad[4, 5, 199] = 32000                     # last gene of the chromosome (N < 32,000 here)
    assert sw.N < 32000
    want_err = np.array([0, 1, 0, 1, 1, 0])
    runs = {"fused32": eng.evaluate(bad, detail=False),
            "fused16": eng.evaluate16(torch.from_numpy(bad.astype(np.int16)).to(eng.device)),
            "staged": eng.evaluate(bad)}
    for name, r in runs.items():
        f = r.flags.cpu().numpy()
        assert np.array_equal(f[:, 3] & 1, want_err), name
        ok = want_err == 0
        assert np.array_equal(r.obj.cpu().numpy()[ok], clean.obj.cpu().numpy()[ok]), name
        with pytest.raises(GcpError):
            eng.raise_on_error(r)
    eng.raise_on_error(clean)


def test_host_call_narrows_int32_genes_on_the_way(synth_setup):
    """gcp_eval_host on a world whose ids fit int16: worker threads narrow the int32 chromosomes
    (saturating) into a pinned int16 copy while finished chunks go to the GPU.  Same objectives and
    flags as the device-resident call and as the plain int32 transfer, for 1 / 3 / auto threads, ragged
    population sizes, and genes far outside int16 (they must stay flagged, not wrap into valid ids)."""
    import torch
    from genetic_coverage_planning_b200 import synth
    sw, pw, ow, op, eng, starts = synth_setup
    base = synth.random_walk_population(sw.xy, sw.row_ptr, sw.col, starts, 40, 150, 200, seed=41).astype(np.int32)
    for P in (1, 37, 5000, 9731):
        dna = np.ascontiguousarray(base[np.random.default_rng(P).integers(0, len(base), P)])
        if P >= 37:
            dna[5, 1, 10] = 70000                   # 70000 & 0xFFFF = 4464 would be a valid waypoint
            dna[9, 0, 3] = -65537 + 2               # low half = 1
            dna[11, 2, 0] = -2
        want = eng.evaluate(dna, detail=False)
        wo, wf = want.obj.cpu().numpy(), want.flags.cpu().numpy()
        if P >= 37:
            assert wf[5, 3] & 1 and wf[9, 3] & 1 and wf[11, 3] & 1 and not wf[6, 3]
        pinned = torch.from_numpy(dna).pin_memory()
        for narrow, threads in ((1, 1), (1, 3), (1, 0), (0, 0)):
            eng.set_option("host_narrow", narrow)
            eng.set_option("host_threads", threads)
            try:
                obj_h, flags_h = eng.evaluate_host(pinned)
            finally:
                eng.set_option("host_narrow", -1)
                eng.set_option("host_threads", 0)
            ok = wf[:, 3] == 0                       # objectives of flagged organisms are unspecified
            assert np.array_equal(flags_h.numpy()[:, 3] & 1, wf[:, 3] & 1), (P, narrow, threads)
            assert np.array_equal(flags_h.numpy()[ok], wf[ok]), (P, narrow, threads)
            assert np.array_equal(obj_h.numpy()[ok], wo[ok]), (P, narrow, threads)


def test_binary_world_blend(synth_setup):
    """blend < 1 needs the internal-coverage counter: general kernel, still bit-exact."""
    sw, pw, ow, op, eng, starts = synth_setup
    from genetic_coverage_planning_b200 import synth
    dna = synth.random_walk_population(sw.xy, sw.row_ptr, sw.col, starts, 12, 150, 200, seed=4)
    try:
        eng.set_params(op, blend=0.1)
        ow.map_params['coverage_blend'] = 0.1
        r = eng.evaluate(dna.astype(np.int32), detail=False)
        obj = r.obj.cpu().numpy()
        for k in range(len(dna)):
            o = fo.calc_obj(ow, dna[k].astype(int))
            assert o[0] == obj[k, 0] and o[1] == obj[k, 1]
    finally:
        eng.set_params(op, blend=1.0)
        ow.map_params['coverage_blend'] = 1.0


def test_long_genomes_partition_fallback(synth_setup):
    """A*L far beyond the shared-memory pair store: the kernels must fall back to block
    partitions and stay exact (C4-style long paths)."""
    from genetic_coverage_planning_b200 import synth, worlds
    from genetic_coverage_planning_b200.engine import FitnessEngine
    sw, pw, ow, op, eng, starts = synth_setup
    A, L = 12, 640
    st = synth.spread_starts(sw.xy, A, 1024)
    op2 = worlds.org_params_for(dict(L=L), st)
    eng2 = FitnessEngine(pw, op2, num_agents=A)
    dna = synth.random_walk_population(sw.xy, sw.row_ptr, sw.col, st, 6, 600, L, seed=11)
    full = eng2.evaluate(dna.astype(np.int32), detail=True)
    fast = eng2.evaluate(dna.astype(np.int32), detail=False)
    assert np.array_equal(full.obj.cpu().numpy(), fast.obj.cpu().numpy())
    assert not full.flags.cpu().numpy()[:, 3].any()
    ow2 = World(sw.img, sw.xy, sw.row_ptr, sw.col, sw.edge_poly, sw.edge_theta, sw.edge_dist,
                sw.map_params, op2, name="synth-long")
    for k in range(3):
        det = fo.calc_obj(ow2, dna[k].astype(int), detail=True)
        assert det['obj'][0] == full.obj.cpu().numpy()[k, 0]
        assert det['obj'][1] == full.obj.cpu().numpy()[k, 1]
        assert np.array_equal(det['lc_mat'].astype(np.int32), full.lc_mat.cpu().numpy()[k])


@pytest.mark.parametrize("A,L,walk", [(3, 640, 600), (6, 640, 600), (6, 200, 150), (2, 97, 90)])
def test_general_fused_kernel_partitions_and_cta_sizes(synth_setup, A, L, walk):
    """The general fused kernel (all six pixel counters, blend 0.3) beyond its compile-time shape: 512
    threads with two hash partitions (3 x 640 genes), 1,024 threads with two partitions (6 x 640), the
    static 6 x 200 layout and a small runtime shape - every output equal to the staged general kernels,
    objectives equal to the oracle."""
    from genetic_coverage_planning_b200 import synth, worlds
    from genetic_coverage_planning_b200.engine import FitnessEngine
    sw, pw, ow, op, eng, starts = synth_setup
    st = synth.spread_starts(sw.xy, A, 1024)
    op2 = worlds.org_params_for(dict(L=L), st)
    op2.update(min_solo_lcs=0, min_comb_lcs=1)
    eng2 = FitnessEngine(pw, op2, num_agents=A)
    eng2.set_params(op2, blend=0.3)
    dna = synth.random_walk_population(sw.xy, sw.row_ptr, sw.col, st, 40, walk, L, seed=31 + A).astype(np.int32)
    fused = eng2.evaluate(dna)
    eng2.set_option("fused_eval", 0)
    try:
        staged = eng2.evaluate(dna)
    finally:
        eng2.set_option("fused_eval", 1)
    assert not fused.flags.cpu().numpy()[:, 3].any()
    for name in ("obj", "neg_cov", "dist", "turn", "lc_mat", "flags", "cov"):
        assert torch.equal(getattr(fused, name), getattr(staged, name)), name
    mp = dict(sw.map_params)
    mp['coverage_blend'] = 0.3
    ow2 = World(sw.img, sw.xy, sw.row_ptr, sw.col, sw.edge_poly, sw.edge_theta, sw.edge_dist, mp, op2,
                name="synth-general")
    obj = fused.obj.cpu().numpy()
    for k in range(2):
        o = fo.calc_obj(ow2, dna[k].astype(int))
        assert o[0] == obj[k, 0] and o[1] == obj[k, 1]


def test_prefetch_variants_of_the_fused_kernel(synth_setup):
    """cov_prefetch selects how the fused kernel requests footprint lines ahead of use when the store
    outgrows the L2 (bit 1: look-ahead inside the union stream, its own kernel variant): same bits."""
    from genetic_coverage_planning_b200 import synth
    sw, pw, ow, op, eng, starts = synth_setup
    dna = synth.random_walk_population(sw.xy, sw.row_ptr, sw.col, starts, 300, 150, 200, seed=77).astype(np.int32)
    want = eng.evaluate(dna, detail=True, pixel_counts=False)
    try:
        for pf in (1, 2, 3, 0):
            eng.set_option("cov_prefetch", pf)
            got = eng.evaluate(dna, detail=True, pixel_counts=False)
            for name in ("obj", "neg_cov", "dist", "turn", "lc_mat", "flags"):
                assert torch.equal(getattr(got, name), getattr(want, name)), (pf, name)
    finally:
        eng.set_option("cov_prefetch", -1)


def test_long_genome_split_kernels_edge_cases_and_fallback(synth_setup):
    """16 agents x 640 genes (beyond the fused kernel): the split form (bucket + partition kernels, the
    default), the monolithic kernels, the radix-sort loop closures and the split form with its overflow
    fallback forced must agree bit for bit - on walks, an all-empty organism, empty agent rows, interior
    -1 holes, a full row, teleporting rows and an out-of-range gene; two organisms against the oracle."""
    from genetic_coverage_planning_b200 import synth, worlds
    from genetic_coverage_planning_b200.engine import FitnessEngine
    sw, pw, ow, op, eng, starts = synth_setup
    A, L = 16, 640
    st = synth.spread_starts(sw.xy, A, 1024)
    op2 = worlds.org_params_for(dict(L=L), st)
    op2.update(min_solo_lcs=0, min_comb_lcs=1)
    eng2 = FitnessEngine(pw, op2, num_agents=A)
    dna = synth.random_walk_population(sw.xy, sw.row_ptr, sw.col, st, 24, 600, L, seed=41).astype(np.int32)
    rng = np.random.RandomState(5)
    dna[1] = -1                                           # nothing at all
    dna[2, 3] = -1; dna[2, 9] = -1                        # empty agent rows
    for a in range(A):                                    # interior holes: the step rule stops, the sequence goes on
        dna[3, a, rng.randint(5, 500)] = -1
    dna[4, 0] = rng.randint(0, sw.N, size=L)              # a full row of teleports (non-edges)
    dna[5, :, 100:] = -1                                  # short rows
    dna[6, 7, 20] = sw.N + 5                              # out of range: error flag
    ref = eng2.evaluate(dna, detail=True, pixel_counts=False)
    assert ref.flags.cpu().numpy()[6, 3] & 1
    ok = np.ones(len(dna), dtype=bool)
    ok[6] = False
    runs = {}
    for name, opts in (("monolithic", dict(cov_split=0, lc_split=0)), ("radix", dict(force_lc_big=1)),
                       ("fallback", dict(cov_split_cap=1 << 30, lc_split_slack=60))):
        for k, v in opts.items():
            eng2.set_option(k, v)
        try:
            runs[name] = eng2.evaluate(dna, detail=True, pixel_counts=False)
        finally:
            for k in opts:
                eng2.set_option(k, 1 if k in ("cov_split", "lc_split") else 0)
    for name, r in runs.items():
        for f in ("obj", "neg_cov", "dist", "turn", "lc_mat", "flags"):
            a, b = getattr(ref, f).cpu().numpy(), getattr(r, f).cpu().numpy()
            assert np.array_equal(a[ok], b[ok]), (name, f)
    ow2 = World(sw.img, sw.xy, sw.row_ptr, sw.col, sw.edge_poly, sw.edge_theta, sw.edge_dist,
                sw.map_params, op2, name="synth-16x640-edge")
    obj, lc = ref.obj.cpu().numpy(), ref.lc_mat.cpu().numpy()
    for k in (2, 3):
        det = fo.calc_obj(ow2, dna[k].astype(int), detail=True)
        assert det['obj'][0] == obj[k, 0] and det['obj'][1] == obj[k, 1], k
        assert np.array_equal(det['lc_mat'].astype(np.int32), lc[k]), k


def test_long_genome_split_kernels_on_a_gray_map(world_big):
    """The reference's big (gray) map with 12 agents x 640 genes: the partition kernels apply the gray
    pixel weights per bucket; split form == monolithic kernels bit for bit, objectives == oracle (1e-12)."""
    from genetic_coverage_planning_b200 import synth
    w = world_big
    A, L = 12, 640
    op = dict(w.org_params)
    op.update(max_dna_len=L, start_idx=(list(w.org_params['start_idx']) * 2)[:A], min_solo_lcs=0, min_comb_lcs=1)
    eng = _engine(w, op, num_agents=A)
    dna = synth.random_walk_population(w.xy, w.row_ptr, w.col, op['start_idx'], 12, 600, L, seed=3).astype(np.int32)
    split = eng.evaluate(dna, detail=True, pixel_counts=False)
    eng.set_option("cov_split", 0)
    eng.set_option("lc_split", 0)
    try:
        mono = eng.evaluate(dna, detail=True, pixel_counts=False)
    finally:
        eng.set_option("cov_split", 1)
        eng.set_option("lc_split", 1)
    assert not split.flags.cpu().numpy()[:, 3].any()
    for f in ("obj", "neg_cov", "dist", "turn", "lc_mat", "flags"):
        assert torch.equal(getattr(split, f), getattr(mono, f)), f
    ow = World(w.img, w.xy, w.row_ptr, w.col, w.edge_poly, w.edge_theta, w.edge_dist, w.map_params, op,
               name="big-12x640")
    obj = split.obj.cpu().numpy()
    for k in range(2):
        o = fo.calc_obj(ow, dna[k].astype(int))
        assert abs(o[0] - obj[k, 0]) <= COV_TOL_GRAY and o[1] == obj[k, 1]
