"""Pin the oracle (oracle/fitness_oracle.py) to outputs of the unmodified reference
captured in tests/golden/*.npz (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import fitness_oracle as fo
from world import World


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def _check_eval(world, z, op=None, n=None):
    op = op or world.org_params
    dnas = z['dna'].astype(int)
    n = len(dnas) if n is None else min(n, len(dnas))
    for k in range(n):
        dna = dnas[k]
        c, d, t, cnt = fo.get_coverage(world, dna, return_counts=True)
        assert c == z['neg_cov'][k]                      # bit-exact fp64
        assert np.array_equal(d, z['dist'][k])
        assert np.array_equal(t, z['turn'][k])
        assert cnt['n_cov_all'] == z['n_painted'][k]
        assert cnt['n_cov_mask'] == z['n_painted_mask'][k]
        lc = fo.get_loop_closures(world, dna)
        assert np.array_equal(lc, z['lc'][k])
        assert np.array_equal(fo.get_loop_closures_direct(world, dna), z['lc'][k])
        assert fo.num_components(lc) == z['ncomp'][k]
        obj = fo.calc_obj(world, dna, op)
        assert np.array_equal(np.array(obj, dtype=float), z['obj'][k])


def test_eval_big(world_big, golden_dir):
    _check_eval(world_big, _load(golden_dir, "eval_big.npz"), n=16)


def test_eval_small(world_small, golden_dir):
    _check_eval(world_small, _load(golden_dir, "eval_small.npz"), n=12)


def test_edge_cases(world_big, golden_dir):
    _check_eval(world_big, _load(golden_dir, "edge_big.npz"))


def test_blend(world_big, golden_dir):
    w = World.load_npz(os.path.join(golden_dir, "world_big.npz"))
    w.map_params['coverage_blend'] = 0.1
    _check_eval(w, _load(golden_dir, "eval_big_blend01.npz"), n=6)


@pytest.mark.parametrize("A", [1, 3])
def test_agent_counts(golden_dir, A):
    z = _load(golden_dir, "eval_big_A%d.npz" % A)
    w = World.load_npz(os.path.join(golden_dir, "world_big.npz"))
    w.map_params['solo_sep_thresh'] = int(z['solo_sep_thresh'])
    op = dict(w.org_params)
    op.update(min_solo_lcs=int(z['min_solo_lcs']), min_comb_lcs=int(z['min_comb_lcs']),
              max_dna_len=int(z['max_dna_len']))
    _check_eval(w, z, op, n=8)


def test_rank(golden_dir, world_big):
    z = _load(golden_dir, "rank.npz")
    _, _, _, gp = __import__("ref_shim").driver_params(6)
    for gen in (0, 17, 80, 200):
        sc = fo.scaled_objectives(z['objs'], gen, gp)
        assert np.array_equal(sc, z['sc_gen%d' % gen])
        fit = fo.rack_n_stack(z['objs'], gen, gp)
        assert np.array_equal(fit, z['fit_gen%d' % gen])
        assert np.array_equal(fo.rack_n_stack_loop(z['objs'][:20], gen, gp),
                              fo.rack_n_stack(z['objs'][:20], gen, gp))
        assert np.array_equal(fo.arena_order(fit), z['order_gen%d' % gen])
    assert np.array_equal(fo.rack_n_stack(z['syn_objs'], 40, gp), z['syn_fit_gen40'])
    for r, idx in zip(z['lowvar_r'], z['lowvar_idx']):
        assert np.array_equal(fo.low_var_sample_indices(z['fit_gen17'], float(z['gamma']), r), idx)


def test_sweep_maximin_equals_double_loop(golden_dir):
    """The O(N log N) restatement (pins the algorithm of k_maximin_sweep)
    is bit-identical to the reference's double loop: golden fits, heavy ties, duplicates."""
    z = _load(golden_dir, "rank.npz")
    _, _, _, gp = __import__("ref_shim").driver_params(6)
    for gen in (0, 17, 80, 200):
        assert np.array_equal(fo.rack_n_stack_sweep(z['objs'], gen, gp), z['fit_gen%d' % gen])
    assert np.array_equal(fo.rack_n_stack_sweep(z['syn_objs'], 40, gp), z['syn_fit_gen40'])
    for seed, n in enumerate([2, 3, 17, 400, 1500]):
        rs = np.random.RandomState(seed)
        objs = np.stack([-rs.randint(0, 30, n) / 30.0, rs.randint(0, 50, n) / 7.0 + 0.5], 1)
        objs[rs.rand(n) < 0.5, 0] = 0.0                      # infeasible organisms tie (F15)
        if n > 300:
            objs[100:140] = objs[200:240]                    # exact duplicates
        for gen in (0, 50, 100):
            assert np.array_equal(fo.rack_n_stack_sweep(objs, gen, gp),
                                  fo.rack_n_stack(objs, gen, gp))


# ---- synthetic BINARY world (the benchmark's world type), evaluated by the unmodified reference on
#      dense tables it built itself (tests/golden/make_golden_synth.py)
@pytest.fixture(scope="module")
def world_synth(golden_dir):
    return World.load_npz(os.path.join(golden_dir, "world_synth.npz"))


def test_synth_world_is_binary(world_synth):
    assert set(np.unique(world_synth.img)) <= {0.0, 1.0}


def test_eval_synth(world_synth, golden_dir):
    _check_eval(world_synth, _load(golden_dir, "eval_synth.npz"), n=12)


def test_edge_cases_synth(world_synth, golden_dir):
    _check_eval(world_synth, _load(golden_dir, "edge_synth.npz"))


def test_agent_count_synth(world_synth, golden_dir):
    z = _load(golden_dir, "eval_synth_A3.npz")
    w = World.load_npz(os.path.join(golden_dir, "world_synth.npz"))
    w.map_params['solo_sep_thresh'] = int(z['solo_sep_thresh'])
    op = dict(w.org_params)
    op.update(min_solo_lcs=int(z['min_solo_lcs']), min_comb_lcs=int(z['min_comb_lcs']),
              max_dna_len=int(z['max_dna_len']))
    _check_eval(w, z, op, n=8)


def test_eval_evolved_feasible(world_big, world_small, golden_dir):
    """reference-evolved, mostly feasible populations (tests/golden/make_golden_feasible.py)"""
    zb, zs = _load(golden_dir, "eval_big_evolved.npz"), _load(golden_dir, "eval_small_evolved.npz")
    assert zb['feasible'].sum() >= 12 and zs['feasible'].sum() >= 15
    _check_eval(world_big, zb, n=12)
    _check_eval(world_small, zs, n=12)
    for w, z in ((world_big, zb), (world_small, zs)):
        for k in range(12):
            assert fo.calc_obj(w, z['dna'][k].astype(int), detail=True)['feasible'] == z['feasible'][k]
