"""GPU, BASELINE.json's full size (config C3: 4096^2 synthetic map, 6 agents x 200 genes,
population 65,536; ranking over 131,072 gladiators): the oracle cannot run at this size in
reasonable time, so parity is carried by size-independent properties of the domain plus a
handful of organisms checked against the oracle."""
import numpy as np
import pytest

import fitness_oracle as fo
from world import World

pytestmark = pytest.mark.gpu

P, A, L = 65536, 6, 200


@pytest.fixture(scope="module")
def c3():
    import torch
    from genetic_coverage_planning_b200 import synth, worlds
    from genetic_coverage_planning_b200.engine import FitnessEngine
    w, pw = worlds.build_synth_world(4096, seed=0)
    starts = synth.spread_starts(w.xy, A, 4096)
    op = worlds.org_params_for(dict(L=L), starts)
    eng = FitnessEngine(pw, op, num_agents=A)
    dna = synth.random_walk_population_torch(w.xy, w.row_ptr, w.col, starts, P, 150, L, seed=1,
                                             device=eng.device)
    torch.cuda.synchronize()
    return w, pw, op, eng, dna


def test_full_population_properties(c3):
    import torch
    w, pw, op, eng, dna = c3
    full = eng.evaluate(dna, detail=True, pixel_counts=False)            # fused kernel
    obj, flags, lc = full.obj, full.flags, full.lc_mat
    assert not bool(flags[:, 3].any())
    # shards == whole (multi-GPU partitioning is exact)
    parts = [eng.evaluate(dna[s:s + 8192], detail=False).obj for s in range(0, P, 8192)]
    assert torch.equal(torch.cat(parts), obj)
    # permuting the organisms permutes the results
    perm = torch.randperm(P, device=dna.device, generator=torch.Generator(device=dna.device).manual_seed(3))
    assert torch.equal(eng.evaluate(dna[perm].contiguous(), detail=False).obj, obj[perm])
    # int16 chromosomes, host path, staged kernels, general kernels: same objectives
    assert torch.equal(eng.evaluate16(dna.to(torch.int16)).obj, obj)
    eng.set_option("fused_eval", 0)
    try:
        assert torch.equal(eng.evaluate(dna, detail=False).obj, obj)
    finally:
        eng.set_option("fused_eval", 1)
    gen = eng.evaluate(dna[:16384], detail=True, pixel_counts=True)      # all six counters
    assert torch.equal(gen.obj, obj[:16384]) and torch.equal(gen.lc_mat, lc[:16384])
    # permuting the AGENTS of an organism keeps the painted set, the total cost up to fp64
    # summation order, and permutes the loop-closure matrix symmetrically (lc + lc^T)
    ap = [3, 0, 5, 1, 4, 2]
    sub = dna[:4096]
    r0 = eng.evaluate(sub, detail=True, pixel_counts=False)
    r1 = eng.evaluate(sub[:, ap].contiguous(), detail=True, pixel_counts=False)
    assert torch.equal(r1.neg_cov, r0.neg_cov)
    assert torch.equal(r1.dist, r0.dist[:, ap]) and torch.equal(r1.turn, r0.turn[:, ap])
    s0 = (r0.lc_mat + r0.lc_mat.transpose(1, 2))[:, ap][:, :, ap]
    s1 = r1.lc_mat + r1.lc_mat.transpose(1, 2)
    off = ~torch.eye(A, dtype=torch.bool, device=dna.device)
    assert torch.equal(s1[:, off], s0[:, off])
    # union bounds: coverage of the organism >= coverage of any single agent alone, and
    # <= the sum over agents (popcount of a union)
    single = []
    for a in range(A):
        d = torch.full_like(sub, -1)
        d[:, a] = sub[:, a]
        single.append(-eng.evaluate(d, detail=True, pixel_counts=False).neg_cov)
    single = torch.stack(single, 1)
    cov = -r0.neg_cov
    assert bool((cov >= single.max(1).values).all()) and bool((cov <= single.sum(1) + 1e-12).all())
    # a few organisms against the oracle, bit-exact
    ow = World(w.img, w.xy, w.row_ptr, w.col, w.edge_poly, w.edge_theta, w.edge_dist, w.map_params,
               op, name="c3")
    for k in (0, 31337, P - 1):
        det = fo.calc_obj(ow, dna[k].cpu().numpy().astype(int), detail=True)
        o = obj[k].cpu().numpy()
        assert det['obj'][0] == o[0] and det['obj'][1] == o[1]
        assert np.array_equal(det['lc_mat'].astype(np.int32), lc[k].cpu().numpy())


def test_full_size_ranking_properties(c3):
    import torch
    w, pw, op, eng, dna = c3
    obj = eng.evaluate(dna, detail=False).obj
    rs = np.random.RandomState(0)
    kid = obj.clone()
    kid[:, 0] = torch.from_numpy(-rs.randint(0, 4000, P) / 4000.0).to(obj.device)   # feasible-looking mix
    kid[torch.from_numpy(rs.rand(P) < 0.4).to(obj.device), 0] = 0.0                 # infeasible ties (F15)
    glad = torch.cat([obj, kid])                                                   # 131,072 gladiators
    for constr in (-0.3, -0.9):
        eng.set_option("rank_algo", 2)
        fit_sweep = eng.maximin(glad, constr)
        order_sort, nondom = eng.arena_order(fit_sweep)
        eng.set_option("rank_algo", 1)
        try:
            fit_tiles = eng.maximin(glad, constr)                                  # the O(N^2) double loop
            order_count, _ = eng.arena_order(fit_tiles)
        finally:
            eng.set_option("rank_algo", 0)
        assert torch.equal(fit_sweep, fit_tiles)
        assert torch.equal(order_sort, order_count)
        # stable ascending order by (fit, index)
        f = fit_sweep[order_sort.long()]
        assert bool((f[1:] >= f[:-1]).all())
        tie = f[1:] == f[:-1]
        assert bool((order_sort[1:][tie] > order_sort[:-1][tie]).all())
        assert torch.equal(nondom == 1, fit_sweep < 0)
        # the oracle on a random subset of rows (each row needs all 131,072 columns)
        sc = fo.scaled_objectives(glad.cpu().numpy(), 0, dict(coverage_aging=80, coverage_constr_0=-constr,
                                                               coverage_constr_f=0.95))
        for i in rs.randint(0, 2 * P, 24):
            m = np.minimum(sc[i, 0] - sc[:, 0], sc[i, 1] - sc[:, 1])
            m[i] = -np.inf
            assert m.max() == float(fit_sweep[i])


_ORACLE_WORLD = None


def _oracle_eval(dnas):
    return [fo.calc_obj(_ORACLE_WORLD, d.astype(int), detail=True) for d in dnas]


def _oracle_many(ow, dnas):
    """the oracle over all host cores (forked workers inherit the world)"""
    import multiprocessing as mp
    import os
    global _ORACLE_WORLD
    _ORACLE_WORLD = ow
    ow.safety_img
    ow.edge_id(0, 0)
    n = max(1, min(os.cpu_count() or 1, 32))
    chunks = [c for c in np.array_split(np.asarray(dnas), n) if len(c)]
    with mp.get_context("fork").Pool(len(chunks)) as pool:
        out = pool.map(_oracle_eval, chunks)
    return [d for part in out for d in part]


def test_random_organisms_of_the_headline_population_against_oracle(c3):
    """384 organisms drawn at random from the 65,536 of the benchmark population, plus 128 taken
    from a population after 30 generations on the device (clustered starts and the loop-closure
    constraints of the bench's generation leg: feasible organisms, shorter and mutated paths):
    objectives, loop-closure matrices, components and feasibility bit-exact against the oracle
    (binary map: -coverage exact)."""
    import torch
    from genetic_coverage_planning_b200 import synth, worlds
    from genetic_coverage_planning_b200.distributed import ShardedEvolution
    w, pw, op, eng, dna = c3
    rs = np.random.RandomState(12)
    idx = np.sort(rs.choice(P, 384, replace=False))
    sample = dna[torch.from_numpy(idx).to(dna.device)].contiguous()
    op2 = dict(op, min_solo_lcs=0, min_comb_lcs=1)
    n_feasible = 0
    try:
        eng.set_params(op2)
        eng.set_vary_params(op2, w.path_params)
        cst = synth.clustered_starts(w.xy, w.row_ptr, A, 4096)
        sub = synth.random_walk_population_torch(w.xy, w.row_ptr, w.col, cst, 4096, 150, L, seed=9,
                                                 device=eng.device)
        evo = ShardedEvolution(eng, worlds.GEN_PARAMS_6, 4096)
        evo.load(sub, 0)
        for g in range(30):
            evo.step(5, g)
        evolved = evo.gather_parents()[:128].contiguous()
        for pop, params in ((evolved, op2), (sample, op)):
            eng.set_params(params)
            ow = World(w.img, w.xy, w.row_ptr, w.col, w.edge_poly, w.edge_theta, w.edge_dist,
                       w.map_params, params, name="c3")
            r = eng.evaluate(pop, detail=True, pixel_counts=False)
            obj, lc, flags = r.obj.cpu().numpy(), r.lc_mat.cpu().numpy(), r.flags.cpu().numpy()
            want = _oracle_many(ow, pop.cpu().numpy())
            for k, det in enumerate(want):
                assert det['obj'][0] == obj[k, 0] and det['obj'][1] == obj[k, 1], k
                assert np.array_equal(det['lc_mat'].astype(np.int32), lc[k]), k
                assert det['feasible'] == flags[k, 0] and det['n_components'] == flags[k, 2], k
                n_feasible += det['feasible']
    finally:
        eng.set_params(op)
    assert n_feasible >= 64, "the evolved part must exercise the feasible branch (%d)" % n_feasible
