"""GPU: selection / variation kernels.  Deterministic parts are checked bit-exactly against
golden vectors of the reference (pruneUTurns, lowVarSample); the random parts (Philox streams
instead of numpy's MT19937) by the invariants and distributions SURVEY.md section 4 lists."""
import os
import pickle

import numpy as np
import pytest

import fitness_oracle as fo
from world import World

pytestmark = pytest.mark.gpu


def _op(world, **kw):
    op = dict(world.org_params)
    op.update(kw)
    return op


def _engine(world, op):
    from genetic_coverage_planning_b200 import packing
    from genetic_coverage_planning_b200.engine import FitnessEngine
    pw = packing.pack_world(world.img, world.xy, world.row_ptr, world.col, world.edge_poly,
                            world.edge_theta, world.edge_dist, world.map_params)
    eng = FitnessEngine(pw, op, num_agents=len(op['start_idx']))
    eng.set_vary_params(op, {'path_memory': 10, 'log_scale_weight': 0.7})
    return eng


def _edges(world):
    src = np.repeat(np.arange(world.N), np.diff(world.row_ptr))
    return set(zip(src.tolist(), world.col.tolist()))


def _valid(row):
    row = np.asarray(row)
    n = int((row != -1).sum())
    assert (row[:n] != -1).all() and (row[n:] == -1).all(), "row must be a valid prefix + -1 padding"
    return row[:n]


def test_prune_uturns_golden(world_big, golden_dir):
    import torch
    z = np.load(os.path.join(golden_dir, "prune.npz"))
    eng = _engine(world_big, _op(world_big, mutation_prob=0.0, muterpolate_prob=0.0))
    dna = torch.from_numpy(z['before'].astype(np.int32)).cuda()
    eng.mutate(dna, seed=1, gen=0)
    assert np.array_equal(dna.cpu().numpy(), z['after'].astype(np.int32))


def test_low_var_select_golden(world_big, golden_dir):
    import torch
    z = np.load(os.path.join(golden_dir, "rank.npz"))
    eng = _engine(world_big, world_big.org_params)
    fit = torch.from_numpy(z['fit_gen17'])
    for r, idx in zip(z['lowvar_r'], z['lowvar_idx']):
        got = eng.low_var_select(fit, float(z['gamma']), float(r)).cpu().numpy()
        assert np.array_equal(got, idx)
    # large population vs the oracle restatement
    rs = np.random.RandomState(0)
    fit = rs.randn(5000) * 0.3
    got = eng.low_var_select(torch.from_numpy(fit), 0.5, 0.5 / 5000).cpu().numpy()
    want = fo.low_var_sample_indices(fit, 0.5, 0.5 / 5000)
    assert np.array_equal(got, want)
    # the BASELINE population, random and heavily tied fitness: measured mismatch rate
    # of the parallel cumulate against the reference's sequential `c += w[i]` is 0
    # (profiles/probe_r02a.jsonl); demand equality
    for kind in ("randn", "ties"):
        rs = np.random.RandomState(65536)
        fit = rs.randn(65536) * 0.3
        if kind == "ties":
            fit[rs.rand(65536) < 0.7] = 0.25
        got = eng.low_var_select(torch.from_numpy(fit), 0.5, 0.37 / 65536).cpu().numpy()
        assert np.array_equal(got, fo.low_var_sample_indices(fit, 0.5, 0.37 / 65536)), kind


def _parents(world, eng, P, walk=150, seed=3):
    from genetic_coverage_planning_b200 import synth
    op = world.org_params
    return synth.random_walk_population(world.xy, world.row_ptr, world.col, op['start_idx'], P,
                                        walk, op['max_dna_len'], seed=seed).astype(np.int32)


def test_crossover_invariants(world_big):
    import torch
    w = world_big
    op = _op(w, crossover_prob=1.0)
    eng = _engine(w, op)
    P, A, L = 64, eng.A, eng.L
    par = _parents(w, eng, P)
    par[5, 2, 40:] = -1                              # a short row: child may need padMinDNA
    strong = torch.arange(P, dtype=torch.int32)
    kids = eng.crossover(par, strong, seed=11, gen=2).cpu().numpy()
    kids2 = eng.crossover(par, strong, seed=11, gen=2).cpu().numpy()
    assert np.array_equal(kids, kids2), "same seed/generation must reproduce"
    assert not np.array_equal(kids, eng.crossover(par, strong, seed=12, gen=2).cpu().numpy())
    cand = set(w.crossover_cand().tolist())
    edges = _edges(w)
    T, mn = op['crossover_time_thresh'], op['min_dna_len']
    n_cross = 0
    for k in range(P // 2):
        for a in range(A):
            pm, pd = _valid(par[k, a]), _valid(par[k + P // 2, a])
            c1, c2 = _valid(kids[2 * k, a]), _valid(kids[2 * k + 1, a])
            if np.array_equal(c1, pm) and np.array_equal(c2, pd):
                continue                              # no candidate pair existed
            n_cross += 1
            # find the crossover point: c1 = pm[:i] + pd[j:], c2 = pd[:j] + pm[i:]
            ok = False
            for i in range(1, len(pm)):
                if not np.array_equal(c1[:i], pm[:i]) or pm[i] not in cand:
                    continue
                for j in range(max(1, i - T + 1), min(len(pd), i + T)):
                    if pd[j] != pm[i]:
                        continue
                    e1 = np.concatenate([pm[:i], pd[j:]])[:L]
                    e2 = np.concatenate([pd[:j], pm[i:]])[:L]
                    ok1 = np.array_equal(c1[:len(e1)], e1) and (len(e1) >= mn and len(c1) == len(e1)
                                                                or len(e1) < mn and len(c1) == mn)
                    ok2 = np.array_equal(c2[:len(e2)], e2) and (len(e2) >= mn and len(c2) == len(e2)
                                                                or len(e2) < mn and len(c2) == mn)
                    if ok1 and ok2:
                        for c, e in ((c1, e1), (c2, e2)):     # padMinDNA tail walks on edges
                            for x in range(len(e) - 1, len(c) - 1):
                                assert (int(c[x]), int(c[x + 1])) in edges
                        ok = True
                        break
                if ok:
                    break
            assert ok, (k, a)
    assert n_cross > 0.5 * (P // 2) * A
    # crossover_prob = 0: children are the parents, paired (k, k+P/2) -> (2k, 2k+1)
    eng0 = _engine(w, _op(w, crossover_prob=0.0))
    kids0 = eng0.crossover(par, strong, seed=1, gen=0).cpu().numpy()
    assert np.array_equal(kids0[0::2], par[:P // 2]) and np.array_equal(kids0[1::2], par[P // 2:])


def test_mutation_invariants(world_big):
    import torch
    w = world_big
    eng = _engine(w, _op(w, mutation_prob=1.0, muterpolate_prob=0.0))
    P, A, L = 96, eng.A, eng.L
    par = _parents(w, eng, P, walk=120)
    dna = torch.from_numpy(par.copy()).cuda()
    eng.mutate(dna, seed=5, gen=1)
    out = dna.cpu().numpy()
    edges = _edges(w)
    changed = steps = off_graph = 0
    for p in range(P):
        for a in range(A):
            r0, r1 = _valid(par[p, a]), _valid(out[p, a])
            assert r1[0] == r0[0] and 2 <= len(r1) <= L
            changed += not np.array_equal(r0, r1)
            for x in range(len(r1) - 1):
                steps += 1
                off_graph += (int(r1[x]), int(r1[x + 1])) not in edges
    assert changed > 0.9 * P * A
    # Walks stay on the graph.  The only off-graph steps are the reference's own artefact:
    # pruneUTurns on OVERLAPPING u-turns (X A B A B -> X B, geneticalgorithm.py:346-357),
    # which its fitness tolerates; they are rare.
    assert off_graph < 0.005 * steps, (off_graph, steps)
    # mutation_prob = 0 and no u-turns in the input -> identity
    eng0 = _engine(w, _op(w, mutation_prob=0.0, muterpolate_prob=0.0))
    d0 = torch.from_numpy(out.copy()).cuda()
    eng0.mutate(d0, seed=5, gen=1)
    again = d0.cpu().numpy()
    for p in range(P):
        for a in range(A):
            r = _valid(out[p, a])
            if not ((r[:-2] == r[2:]).any()):
                assert np.array_equal(again[p, a], out[p, a])


def test_muterpolate_invariants(world_big):
    import torch
    w = world_big
    eng = _engine(w, _op(w, mutation_prob=0.0, muterpolate_prob=1.0))
    P, A = 64, eng.A
    par = _parents(w, eng, P, walk=150, seed=8)
    dna = torch.from_numpy(par.copy()).cuda()
    eng.mutate(dna, seed=9, gen=4)
    out = dna.cpu().numpy()
    nbrs = {n: set(w.col[w.row_ptr[n]:w.row_ptr[n + 1]].tolist()) for n in range(w.N)}
    shorter = 0
    for p in range(P):
        for a in range(A):
            r0, r1 = _valid(par[p, a]), _valid(out[p, a])
            assert r1[0] == r0[0] and len(r1) <= len(r0)
            shorter += len(r1) < len(r0)
            allowed = set(r0.tolist())
            for g in r0:
                allowed |= nbrs[int(g)]
            assert set(r1.tolist()) <= allowed      # shortcut nodes are neighbours of old ones
    assert shorter > 0.3 * P * A


def test_random_walks_and_statistics(world_big):
    from genetic_coverage_planning_b200 import synth
    w = world_big
    op = w.org_params
    eng = _engine(w, op)
    P, A, L = 256, eng.A, eng.L
    dna = eng.random_walks(P, op['start_idx'], 150, seed=3).cpu().numpy()
    assert np.array_equal(dna, eng.random_walks(P, op['start_idx'], 150, seed=3).cpu().numpy())
    edges = _edges(w)
    for p in range(0, P, 8):
        for a in range(A):
            r = _valid(dna[p, a])
            assert len(r) == 151 and r[0] == op['start_idx'][a]
            for x in range(150):
                assert (int(r[x]), int(r[x + 1])) in edges
    # same rule as the numpy restatement of makeMeAPath: compare simple statistics
    ref = synth.random_walk_population(w.xy, w.row_ptr, w.col, op['start_idx'], P, 150, L, seed=1)

    def stats(d):
        uniq = np.mean([len(np.unique(r[r >= 0])) for r in d.reshape(-1, L)])
        end = w.xy[d[:, :, 150]] - w.xy[d[:, :, 0]]
        return uniq, np.sqrt((end ** 2).sum(-1)).mean()
    u_dev, e_dev = stats(dna)
    u_ref, e_ref = stats(ref)
    assert abs(u_dev - u_ref) < 0.1 * u_ref, (u_dev, u_ref)
    assert abs(e_dev - e_ref) < 0.15 * e_ref, (e_dev, e_ref)


def test_generation_step_against_oracle(world_big):
    """One full generation on the device, every deterministic stage re-derived by the oracle."""
    import torch
    import ref_shim
    w = world_big
    op = w.org_params
    _, _, _, gp = ref_shim.driver_params(6)
    eng = _engine(w, op)
    P = 48
    par = torch.from_numpy(_parents(w, eng, P)).cuda()
    par_obj = eng.evaluate(par, detail=False).obj
    constr = fo.coverage_constraint(0, gp)
    par_fit = eng.maximin(par_obj, constr)
    kids = eng.select_vary(par, par_fit, gp['gamma'], seed=77, gen=0)
    kid_obj = eng.evaluate(kids, detail=False).obj
    glad_obj = torch.cat([par_obj, kid_obj])
    fit = eng.maximin(glad_obj, constr)
    order, nondom = eng.arena_order(fit)
    new_dna, new_obj, new_fit = eng.gather_survivors(order, par, kids, glad_obj, fit)
    want_fit = fo.rack_n_stack(glad_obj.cpu().numpy(), 0, gp)
    assert np.array_equal(fit.cpu().numpy(), want_fit)
    want_order = fo.arena_order(want_fit)
    assert np.array_equal(order.cpu().numpy(), want_order)
    glad_dna = np.concatenate([par.cpu().numpy(), kids.cpu().numpy()])
    assert np.array_equal(new_dna.cpu().numpy(), glad_dna[want_order[:P]])
    assert np.array_equal(new_obj.cpu().numpy(), glad_obj.cpu().numpy()[want_order[:P]])
    assert np.array_equal(new_fit.cpu().numpy(), want_fit[want_order[:P]])
    for k in (0, 7, 23):                          # children objectives vs oracle
        det = fo.calc_obj(w, kids[k].cpu().numpy().astype(int))
        got = kid_obj[k].cpu().numpy()
        assert det[1] == got[1] and abs(det[0] - got[0]) <= 1e-12


def test_fused_select_vary_equals_staged_kernels(world_big):
    """gcp_select_vary (lowVarSample + crossover with the constructor variation fused into the
    same kernel) must produce exactly what the staged calls produce: same Philox streams."""
    import torch
    w = world_big
    eng = _engine(w, w.org_params)
    P = 128
    par = torch.from_numpy(_parents(w, eng, P, seed=21)).cuda()
    rs = np.random.RandomState(3)
    fit = torch.from_numpy(rs.randn(P) * 0.2).cuda()
    for seed, gen in ((5, 0), (123456789012345, 9)):
        fused = eng.select_vary(par, fit, 0.5, seed=seed, gen=gen).cpu().numpy()
        m64 = (1 << 64) - 1
        z = (seed ^ (0x9E3779B97F4A7C15 * (gen + 1))) & m64
        z ^= z >> 30
        z = (z * 0xBF58476D1CE4E5B9) & m64
        z ^= z >> 27
        z = (z * 0x94D049BB133111EB) & m64
        z ^= z >> 31
        r0 = ((z >> 11) * (1.0 / 9007199254740992.0)) / P
        strong = eng.low_var_select(fit, 0.5, r0)
        kids = eng.crossover(par, strong, seed=seed, gen=gen)
        eng.mutate(kids, seed=seed, gen=gen, salt=0)
        assert np.array_equal(fused, kids.cpu().numpy())
        for row in fused.reshape(-1, fused.shape[-1]):
            _valid(row)


def test_long_rows_variation(world_big):
    """L = 640 rows (C4-style): shared-memory rows beyond the default 48 KB, invariants hold."""
    import torch
    w = world_big
    op = _op(w, max_dna_len=640, mutation_prob=1.0, muterpolate_prob=1.0, crossover_prob=1.0)
    eng = _engine(w, op)
    from genetic_coverage_planning_b200 import synth
    P = 32
    par = synth.random_walk_population(w.xy, w.row_ptr, w.col, op['start_idx'], P, 500, 640,
                                       seed=2).astype(np.int32)
    fit = torch.from_numpy(np.linspace(-1, 1, P)).cuda()
    kids = eng.select_vary(torch.from_numpy(par).cuda(), fit, 0.5, seed=1, gen=1).cpu().numpy()
    edges = _edges(w)
    steps = off = moved = 0
    for p in range(P):
        for a in range(eng.A):
            r = _valid(kids[p, a])
            assert 2 <= len(r) <= 640
            # pruneUTurns on overlapping u-turns at the head (A B A B -> B) may drop the start
            # waypoint: the reference's own artefact (geneticalgorithm.py:346-357), rare
            moved += r[0] != op['start_idx'][a]
            for x in range(len(r) - 1):
                steps += 1
                off += (int(r[x]), int(r[x + 1])) not in edges
    assert off < 0.02 * steps, (off, steps)       # muterpolate / prune artefacts only (F4)
    assert moved <= 0.05 * P * eng.A, moved


def test_facade_drop_in(tmp_path):
    """GeneticalGorithm / Organism facade on a small synthetic world (CSR-backed Mappy /
    PathMaker stand-ins): constructor, runEvolution history, views, rackNstack, pickle."""
    from genetic_coverage_planning_b200 import synth, worlds
    from genetic_coverage_planning_b200.geneticalgorithm import GeneticalGorithm, Organism
    sw = synth.make_world(768, seed=3)
    mappy, pather = synth.CsrMappy(sw), synth.CsrPather(sw)
    starts = synth.spread_starts(sw.xy, 3, 768)
    op = worlds.org_params_for(dict(L=120), starts)
    op.update(min_solo_lcs=0, min_comb_lcs=0)
    gp = dict(worlds.GEN_PARAMS_6)
    gp.update(gen_size=32, starting_path_len=80, num_agents=3, coverage_constr_0=0.0,
              org_params=op)
    pop = GeneticalGorithm(mappy, pather, gp, seed=5)
    assert len(pop._gen_parent) == 32 and len(pop._gen_parent_fit) == 32
    o0 = pop._gen_parent[0]
    assert o0._dna.shape == (3, 120) and len(o0._obj_val) == 2 and o0._max_dna_len == 120
    hist = pop.runEvolution(3)
    assert len(hist) == 3 and len(hist[0]) == 32 and len(hist[0][0]) == 2
    assert pop._gen_num == 3 and pop.RunEvolution.__func__ is pop.runEvolution.__func__
    # parents' stored objectives equal a fresh oracle evaluation of their DNA
    ow = World(sw.img, sw.xy, sw.row_ptr, sw.col, sw.edge_poly, sw.edge_theta, sw.edge_dist,
               sw.map_params, op, name="facade")
    for org in pop._gen_parent[:4]:
        want = fo.calc_obj(ow, org._dna)
        assert want[0] == org._obj_val[0] and want[1] == org._obj_val[1]
    # fitnesses are non-decreasing (arena order) and rackNstack is the oracle's maximin
    f = pop._gen_parent_fit
    assert all(f[i] <= f[i + 1] for i in range(len(f) - 1))
    fits = pop.rackNstack(pop._gen_parent)
    want = fo.rack_n_stack(np.array([o._obj_val for o in pop._gen_parent]), pop._gen_num,
                           dict(coverage_aging=gp['coverage_aging'],
                                coverage_constr_0=gp['coverage_constr_0'],
                                coverage_constr_f=gp['coverage_constr_f']))
    assert np.array_equal(np.array(fits), want)
    # single-organism API
    kid_a, kid_b = pop._gen_parent[0].crossover(pop._gen_parent[1])
    assert kid_a._dna.shape == (3, 120) and kid_a._dna[0, 0] == starts[0]
    assert kid_a.calcObj() == kid_a._obj_val
    org = Organism([r for r in pop._gen_parent[2]._dna_list], mappy, pather, op, _seed=1)
    assert org._dna[1, 0] == starts[1]
    # migration: continue a population that came from the reference's classes (duck-typed here:
    # plain objects carrying the attributes of geneticalgorithm.py:16-35 / :127-175)
    class _Ref(object):
        pass
    old = _Ref()
    for k in ("_G_sz", "_starting_path_len", "_num_agents", "_gamma", "_cov_constr_0",
              "_cov_constr_f", "_cov_aging", "_org_params", "_gen_num"):
        setattr(old, k, getattr(pop, k))
    old._gen_parent = []
    for o in pop._gen_parent:
        r = _Ref()
        r._dna, r._mappy, r._pather, r._obj_val = o._dna.copy(), mappy, pather, [0.0, 0.0]
        old._gen_parent.append(r)
    cont = GeneticalGorithm.from_population(old, seed=9)
    assert cont._gen_num == pop._gen_num
    assert [o._obj_val for o in cont._gen_parent] == [o._obj_val for o in pop._gen_parent]
    assert cont._gen_parent_fit == cont.rackNstack(cont._gen_parent)       # re-ranked among themselves
    assert len(cont.runEvolution(1)) == 1 and cont._gen_num == pop._gen_num + 1
    # pickle round trip (gori_tools.savePopulation / loadPopulation)
    fn = os.path.join(tmp_path, "pop.pkl")
    pickle.dump(pop, open(fn, "wb"))
    pop2 = pickle.load(open(fn, "rb"))
    assert pop2._gen_num == 3
    assert np.array_equal(pop2._gen_parent[5]._dna, pop._gen_parent[5]._dna)
    h2 = pop2.runEvolution(1)
    assert len(h2) == 1 and pop2._gen_num == 4


def test_evolution_matches_reference_statistics(world_small, golden_dir):
    """BASELINE config C1 end to end: the drop-in GeneticalGorithm on the reference's own small map
    (map_scaled.png + wilk_3 graph, 6 agents, pather_driver parameters) against per-generation
    statistics of the UNMODIFIED reference's runEvolution (tests/golden/make_golden_evolution.py).
    Random streams differ (Philox vs MT19937), so the comparison is statistical: means over
    three seeds, tolerances a few times the reference's own seed-to-seed spread."""
    import ref_shim
    from genetic_coverage_planning_b200 import synth
    from genetic_coverage_planning_b200.geneticalgorithm import GeneticalGorithm
    z = np.load(os.path.join(golden_dir, "evolution_small.npz"))
    ref = z['traj'].mean(axis=0)                      # [gen, stat]
    n_gen, G = int(z['n_gen']), int(z['gen_size'])
    w = world_small
    w.path_params = {'path_memory': 10, 'log_scale_weight': 0.7}
    mappy, pather = synth.CsrMappy(w), synth.CsrPather(w)
    _, _, op, gp = ref_shim.driver_params(6)
    op = dict(op)
    op['start_idx'] = [int(x) for x in z['start_idx']]
    gp = dict(gp)
    gp.update(gen_size=G, coverage_constr_0=0.0, org_params=op,
              starting_path_len=int(z['starting_path_len']))

    def stats(pop):
        obj = np.array([o._obj_val for o in pop._gen_parent], dtype=float)
        lens = np.array([[(row != -1).sum() for row in o._dna] for o in pop._gen_parent])
        return [(obj[:, 0] < 0).mean(), -obj[:, 0].min(), -obj[:, 0].mean(), obj[:, 1].mean(),
                lens.mean()]
    runs = []
    for seed in (11, 12, 13):
        pop = GeneticalGorithm(mappy, pather, gp, seed=seed)
        traj = [stats(pop)]
        for g in range(n_gen):
            pop.runEvolution(1)
            if g + 1 in (5, n_gen):
                traj.append(stats(pop))
        runs.append(traj)
    got = np.array(runs).mean(axis=0)                 # rows: gen 0, gen 5, gen n_gen
    FEAS, BEST, MEAN, COST, LEN = range(5)
    # initial population: random walks + constructor variation (firstGeneration)
    assert abs(got[0, FEAS] - ref[0, FEAS]) <= 0.2
    assert got[0, BEST] >= 0.95
    assert abs(got[0, LEN] - ref[0, LEN]) <= 0.08 * ref[0, LEN]
    assert abs(got[0, COST] - ref[0, COST]) <= 0.10 * ref[0, COST]
    # after 5 generations the constraint pressure has made the population feasible
    assert got[1, FEAS] >= 0.9 and ref[5, FEAS] >= 0.9
    assert got[1, MEAN] >= 0.9
    # end of the run
    assert got[2, FEAS] == 1.0
    assert got[2, BEST] >= ref[n_gen, BEST] - 0.002
    assert abs(got[2, MEAN] - ref[n_gen, MEAN]) <= 0.03
    assert abs(got[2, COST] - ref[n_gen, COST]) <= 0.12 * ref[n_gen, COST]
    assert abs(got[2, LEN] - ref[n_gen, LEN]) <= 0.12 * ref[n_gen, LEN]


def test_legacy_reference_pickle_migrates(golden_dir):
    """a population evolved and pickled by the UNMODIFIED reference
    (tests/golden/make_golden_legacy.py: cropped small map, graph from the reference's own
    computeTraversableGraph, 3 agents, 2 generations) opens without the reference, and the GPU
    evaluation of its parents reproduces the objectives the reference stored for them."""
    import os
    from genetic_coverage_planning_b200 import legacy
    z = np.load(os.path.join(golden_dir, "legacy_population.npz"))
    pop = legacy.load(os.path.join(golden_dir, "legacy_population.pkl.gz"), seed=1)
    assert pop._gen_num == int(z['gen_num']) and pop._G_sz == len(z['dna'])
    got = {np.asarray(o._dna).astype(np.int64).tobytes(): np.asarray(o._obj_val, dtype=float)
           for o in pop._gen_parent}
    for dna, obj in zip(z['dna'], z['obj']):
        mine = got[np.asarray(dna).astype(np.int64).tobytes()]
        assert mine[1] == obj[1]                               # travel cost: bit-exact
        assert abs(mine[0] - obj[0]) <= 1e-12                  # -coverage on a gray map
    # ranking of the migrated parents is the reference's rackNstack on the same objectives
    want = fo.rack_n_stack(np.array([o._obj_val for o in pop._gen_parent]), pop._gen_num,
                           dict(coverage_aging=pop._cov_aging, coverage_constr_0=-pop._cov_constr_0,
                                coverage_constr_f=-pop._cov_constr_f))
    assert np.array_equal(np.asarray(pop._gen_parent_fit), want)
    hist = pop.runEvolution(2)                                 # and evolution continues
    assert pop._gen_num == int(z['gen_num']) + 2 and len(hist) == 2 and len(hist[0]) == len(z['dna'])


# ---------------------------------------------------------------------------------------------
# Distributions of the Philox kernels against samples drawn from the UNMODIFIED reference
# (tests/golden/make_golden_variation_dist.py -> variation_dist.npz, config C1).  Two-sample
# chi-square on statistics that are functions of the output rows only; rare categories pooled.
# The seeds are fixed, so the tests are deterministic; the acceptance level (p > 1e-3 per
# statistic) is what an unbiased operator passes with wide margin and a biased one does not.
# ---------------------------------------------------------------------------------------------
def _chi2_p(a, b, min_count=10):
    """p-value of H0 'samples a and b (integer categories) come from one distribution'."""
    from scipy.stats import chi2
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    cats = np.union1d(a, b)
    ca = np.array([(a == c).sum() for c in cats], dtype=np.float64)
    cb = np.array([(b == c).sum() for c in cats], dtype=np.float64)
    rare = (ca + cb) < min_count
    if rare.any():
        ca = np.append(ca[~rare], ca[rare].sum())
        cb = np.append(cb[~rare], cb[rare].sum())
    keep = (ca + cb) > 0
    ca, cb = ca[keep], cb[keep]
    if len(ca) < 2:
        return 1.0
    na, nb = ca.sum(), cb.sum()
    stat = ((np.sqrt(nb / na) * ca - np.sqrt(na / nb) * cb) ** 2 / (ca + cb)).sum()
    return float(chi2.sf(stat, len(ca) - 1))


def _dist_helpers():
    import importlib.util
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden_variation_dist.py")
    spec = importlib.util.spec_from_file_location("mk_dist", p)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="module")
def dist_gold(golden_dir):
    return np.load(os.path.join(golden_dir, "variation_dist.npz"))


P_MIN = 1e-3


def test_walk_distribution_matches_reference(world_small, dist_gold):
    """makeMeAPath (pathmaker.py:188-213): node reached after 1 / 3 / 10 steps and the number of
    distinct waypoints of a 150-step walk, per start waypoint."""
    h = _dist_helpers()
    w = world_small
    op = w.org_params
    eng = _engine(w, op)
    ref = dist_gold["walks"].astype(np.int64)                         # [n, A, 151]
    got = eng.random_walks(4000, op['start_idx'], 150, seed=2024).cpu().numpy()[:, :, :151]
    assert (got[:, :, 150] >= 0).all()
    starts = np.array(op['start_idx'])
    for s in np.unique(starts):
        agents = np.nonzero(starts == s)[0]
        r = ref[:, agents].reshape(-1, 151)
        g = got[:, agents].reshape(-1, 151)
        for step in (1, 3, 10):
            p = _chi2_p(r[:, step], g[:, step])
            assert p > P_MIN, ("node after %d steps from %d" % (step, s), p)
        p = _chi2_p(h.n_distinct(r) // 4, h.n_distinct(g) // 4)
        assert p > P_MIN, ("distinct nodes from %d" % s, p)


def _replicate(rows, n):
    return np.repeat(np.asarray(rows, dtype=np.int32)[None], n, axis=0).copy()


def test_mutation_distribution_matches_reference(world_small, dist_gold):
    """Organism.mutation (geneticalgorithm.py:251-268) + pruneUTurns on one fixed organism:
    cut point (through the common prefix with the input) and new length, per agent."""
    import torch
    h = _dist_helpers()
    w = world_small
    base = dist_gold["base"].astype(np.int32)
    eng = _engine(w, _op(w, mutation_prob=1.0, muterpolate_prob=0.0))
    dna = torch.from_numpy(_replicate(base, 6000)).cuda()
    eng.mutate(dna, seed=31, gen=3)
    got = dna.cpu().numpy()
    ref = dist_gold["mutation"].astype(np.int32)
    for a in range(base.shape[0]):
        p = _chi2_p(h.common_prefix(ref[:, a], base[a]) // 4, h.common_prefix(got[:, a], base[a]) // 4)
        assert p > P_MIN, ("mutation prefix, agent %d" % a, p)
        p = _chi2_p(h.valid_len(ref[:, a]) // 5, h.valid_len(got[:, a]) // 5)
        assert p > P_MIN, ("mutation length, agent %d" % a, p)


def test_muterpolate_distribution_matches_reference(world_small, dist_gold):
    """Organism.muterpolate (geneticalgorithm.py:270-309) on one fixed organism: how much is cut
    out and where the first change sits, per agent."""
    import torch
    h = _dist_helpers()
    w = world_small
    base = dist_gold["base"].astype(np.int32)
    eng = _engine(w, _op(w, mutation_prob=0.0, muterpolate_prob=1.0))
    dna = torch.from_numpy(_replicate(base, 6000)).cuda()
    eng.mutate(dna, seed=32, gen=5)
    got = dna.cpu().numpy()
    ref = dist_gold["muterpolate"].astype(np.int32)
    for a in range(base.shape[0]):
        p = _chi2_p(h.valid_len(ref[:, a]) // 2, h.valid_len(got[:, a]) // 2)
        assert p > P_MIN, ("muterpolate length, agent %d" % a, p)
        p = _chi2_p(h.common_prefix(ref[:, a], base[a]) // 4, h.common_prefix(got[:, a], base[a]) // 4)
        assert p > P_MIN, ("muterpolate first change, agent %d" % a, p)


def test_crossover_distribution_matches_reference(world_small, dist_gold):
    """Organism.crossover / matchWaypt (geneticalgorithm.py:211-249,311-333) of one fixed pair:
    the cut pair (i, j) through child 1's common prefix with the mother and its length, per agent;
    crossover happens with the reference's probability."""
    import torch
    h = _dist_helpers()
    w = world_small
    base, mate = dist_gold["base"].astype(np.int32), dist_gold["mate"].astype(np.int32)
    eng = _engine(w, _op(w, mutation_prob=0.0, muterpolate_prob=0.0))
    K = 8000
    parents = torch.from_numpy(np.stack([base, mate])).cuda()
    strong = torch.cat([torch.zeros(K, dtype=torch.int32), torch.ones(K, dtype=torch.int32)]).cuda()
    kids = eng.crossover(parents, strong, seed=33, gen=7)
    eng.mutate(kids, seed=33, gen=7)                                  # constructor of the children: prune only
    kids = kids.cpu().numpy()
    got0, got1 = kids[0::2], kids[1::2]
    ref = dist_gold["crossover"].astype(np.int32)                     # [n, 2, A, L]
    for a in range(base.shape[0]):
        same_ref = (ref[:, 0, a] == base[a]).all(-1)
        same_got = (got0[:, a] == base[a]).all(-1)
        p = _chi2_p(same_ref, same_got, min_count=1)
        assert p > P_MIN, ("crossover frequency, agent %d" % a, p, same_ref.mean(), same_got.mean())
        p = _chi2_p(h.common_prefix(ref[:, 0, a], base[a]) // 3, h.common_prefix(got0[:, a], base[a]) // 3)
        assert p > P_MIN, ("cut index i, agent %d" % a, p)
        p = _chi2_p(h.valid_len(ref[:, 0, a]) // 3, h.valid_len(got0[:, a]) // 3)
        assert p > P_MIN, ("child 1 length, agent %d" % a, p)
        p = _chi2_p(h.common_prefix(ref[:, 1, a], mate[a]) // 3, h.common_prefix(got1[:, a], mate[a]) // 3)
        assert p > P_MIN, ("cut index j, agent %d" % a, p)
