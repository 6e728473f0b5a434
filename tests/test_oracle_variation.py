"""The variation oracle (oracle/variation_oracle.py) against the unmodified reference: replaying
the seeds of tests/golden/make_golden_variation_rng.py through the same numpy calls must
reproduce the reference's paths, organisms, children and generations EXACTLY."""
import importlib.util
import os

import numpy as np
import pytest

import variation_oracle as vo

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("variation_cases", os.path.join(HERE, "golden", "variation_cases.py"))
_vc = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_vc)
CASES = {k: getattr(_vc, k) for k in ("PATH_CASES", "ORG_CASES", "XOVER_SEEDS", "EVO_CASES")}


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "variation_rng.npz"))


def _gen_params(world):
    return {'gen_size': 100, 'starting_path_len': 150, 'num_agents': 6, 'gamma': 0.5,
            'coverage_constr_0': 0.3, 'coverage_constr_f': 0.95, 'coverage_aging': 80,
            'org_params': dict(world.org_params)}


def test_norm_logpdf_is_scipys():
    scipy_stats = pytest.importorskip("scipy.stats")
    rs = np.random.RandomState(0)
    x = np.concatenate([rs.uniform(-np.pi, np.pi, 2000), [0.0, np.pi, -np.pi]])
    for scale in (0.7, 1.0, 0.05):
        assert np.array_equal(vo.norm_logpdf(x, scale), scipy_stats.norm.logpdf(x, scale=scale))


def test_make_me_a_path_equals_reference(world_small, gold):
    g = vo.Graph(world_small)
    for k, (seed, n, start, prev_len) in enumerate(CASES["PATH_CASES"]):
        np.random.seed(seed)
        if prev_len is not None:
            prev = vo.make_me_a_path(g, prev_len, start)
            assert np.array_equal(prev, gold["path%d_prev" % k])
            path = vo.make_me_a_path(g, n, prev[-1], prev)
        else:
            path = vo.make_me_a_path(g, n, start)
        assert np.array_equal(path, gold["path%d" % k]), k
        assert len(path) == n + 1


def test_organism_constructor_equals_reference(world_small, gold):
    g = vo.Graph(world_small)
    for k, (seed, pm, pmu) in enumerate(CASES["ORG_CASES"]):
        op = dict(world_small.org_params)
        op['mutation_prob'], op['muterpolate_prob'] = pm, pmu
        np.random.seed(seed)
        paths = [vo.make_me_a_path(g, 150, s) for s in op['start_idx']]
        assert np.array_equal(np.array(paths), gold["org%d_in" % k])
        org = vo.Organism([q.copy() for q in paths], world_small, op, g)
        assert np.array_equal(org.dna, gold["org%d_dna" % k]), k
        assert np.array_equal(np.array(org.obj_val, dtype=np.float64), gold["org%d_obj" % k]), k
    # the cases exercise every branch: the variation changed the input paths where it was on
    assert not np.array_equal(gold["org1_dna"][:, :151], gold["org1_in"])
    assert not np.array_equal(gold["org2_dna"][:, :151], gold["org2_in"])


def test_crossover_equals_reference(world_small, gold):
    g = vo.Graph(world_small)
    n_swapped = 0
    for k, seed in enumerate(CASES["XOVER_SEEDS"]):
        op = dict(world_small.org_params)
        if k == 3:
            op['min_dna_len'] = 170
        np.random.seed(seed)
        mom = vo.Organism([vo.make_me_a_path(g, 150, s) for s in op['start_idx']], world_small, op, g)
        dad = vo.Organism([vo.make_me_a_path(g, 150, s) for s in op['start_idx']], world_small, op, g)
        assert np.array_equal(mom.dna, gold["xo%d_mom" % k]) and np.array_equal(dad.dna, gold["xo%d_dad" % k])
        kids = mom.crossover(dad)
        assert np.array_equal(kids[0].dna, gold["xo%d_kid0" % k]), k
        assert np.array_equal(kids[1].dna, gold["xo%d_kid1" % k]), k
        assert np.array_equal(np.array([kids[0].obj_val, kids[1].obj_val], dtype=np.float64), gold["xo%d_obj" % k])
        n_swapped += int(not np.array_equal(kids[0].dna[:, :30], mom.dna[:, :30]) or
                         not np.array_equal(kids[0].dna, mom.dna))
    assert n_swapped > 0


def test_generations_equal_reference(world_small, gold):
    for k, (seed, G, n_gen) in enumerate(CASES["EVO_CASES"]):
        gp = _gen_params(world_small)
        gp['gen_size'] = G
        gp['coverage_constr_0'] = 0.0
        np.random.seed(seed)
        pop = vo.Population(world_small, gp)
        assert np.array_equal(np.array([o.dna for o in pop.parents]), gold["evo%d_dna0" % k])
        assert np.array_equal(np.array(pop.fit), gold["evo%d_fit0" % k])
        for gen in range(n_gen):
            h = pop.run_evolution(1)
            assert np.array_equal(np.array([o.dna for o in pop.parents]), gold["evo%d_dna" % k][gen]), (k, gen)
            assert np.array_equal(np.array(pop.fit), gold["evo%d_fit" % k][gen]), (k, gen)
            assert np.array_equal(np.array(h[0], dtype=np.float64), gold["evo%d_hist" % k][gen])


@pytest.mark.ref
def test_live_reference_generation_matches_oracle(world_small):
    """With /root/reference present: one more seed, compared live (not through a fixture)."""
    import ref_shim
    if not ref_shim.available():
        pytest.skip("reference not present")
    ga, mappy_m, pm_m, got = ref_shim.load()
    m, p, (mp, pp, op, gp) = ref_shim.build_world("small", 6)
    gp = dict(gp)
    gp.update(gen_size=6, coverage_constr_0=0.0, org_params=dict(op))
    with ref_shim.stable_argsort():
        np.random.seed(77)
        ref = ga.GeneticalGorithm(m, p, gp)
        ref_hist = ref.runEvolution(2)
    np.random.seed(77)
    pop = vo.Population(world_small, gp)
    hist = pop.run_evolution(2)
    assert np.array_equal(np.array([o._dna for o in ref._gen_parent]), np.array([o.dna for o in pop.parents]))
    assert np.array_equal(np.array(ref_hist, dtype=np.float64), np.array(hist, dtype=np.float64))
