"""CPU, this container only (needs /root/reference): a population evolved and pickled by the B200
facade on a GPU box (tests/golden/facade_population.pkl.gz, made by make_facade_pickle.py) is opened
WITHOUT a GPU and handed to the unmodified reference:

  * got.getObjValsList / the attribute reads of got.plotty see the reference's data shapes;
  * the reference's own Organism.calcObj re-evaluates every GPU-evolved organism (children of the
    device crossover / mutation kernels, not hand-made DNA) and returns the objectives the GPU stored;
  * GeneticalGorithm.rackNstack of the reference writes the scaled objectives the GPU stored;
  * got.generateWaypointData (the savePath export) runs on the facade's Organism views.
"""
import gzip
import os
import pickle

import numpy as np
import pytest

import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference not present")


@pytest.fixture(scope="module")
def loaded(golden_dir):
    with gzip.open(os.path.join(golden_dir, "facade_population.pkl.gz"), "rb") as f:
        pop = pickle.loads(f.read())
    assert '_host_state' in pop.__dict__ and '_eng' not in pop.__dict__      # nothing touched the GPU
    ga, mappy_m, pm_m, got = ref_shim.load()
    m, p, params = ref_shim.build_cropped_world(3)
    return pop, ga, got, m, p


def test_facade_pickle_opens_without_gpu_and_reads_like_the_reference(loaded):
    pop, ga, got, m, p = loaded
    objs = np.array(got.getObjValsList(pop))                 # gori_tools.py:248-250, unmodified
    assert objs.shape == (pop._G_sz, 2) and pop._gen_num == 5
    assert len(pop._gen_parent_fit) == pop._G_sz
    o = pop._gen_parent[0]
    assert o._dna.shape == (3, o._max_dna_len) and o._dna.dtype.kind == 'i'
    # the world in the pickle is the one the reference builds from the same data
    assert np.array_equal(o._pather._graph, p._graph) and np.array_equal(o._pather._XY, p._XY)
    assert np.array_equal(o._mappy._traverse_frustum, m._traverse_frustum)
    assert '_host_state' in pop.__dict__                     # still no device state


def test_reference_recomputes_the_gpu_objectives_and_fitness(loaded):
    pop, ga, got, m, p = loaded
    shells = []
    for view in pop._gen_parent:
        org = ga.Organism.__new__(ga.Organism)               # bare shell: no constructor mutation
        org._mappy, org._dna, org._num_agents = m, np.asarray(view._dna), view._dna.shape[0]
        org._min_solo_lcs, org._min_comb_lcs = view._min_solo_lcs, view._min_comb_lcs
        org._ft_scale = view._ft_scale
        org._obj_val_sc = [0, 0]                             # geneticalgorithm.py:173
        with ref_shim.stable_argsort():                      # SURVEY F9 canonical loop closures
            org._obj_val = org.calcObj()                     # geneticalgorithm.py:177-201, unmodified
        assert float(org._obj_val[1]) == view._obj_val[1]                    # travel cost: bit-exact
        assert abs(float(org._obj_val[0]) - view._obj_val[0]) <= 1e-12       # -coverage (gray map)
        shells.append(org)
    ref_pop = ga.GeneticalGorithm.__new__(ga.GeneticalGorithm)
    ref_pop._cov_constr_0, ref_pop._cov_constr_f = pop._cov_constr_0, pop._cov_constr_f
    ref_pop._cov_aging, ref_pop._gen_num = pop._cov_aging, pop._gen_num
    for s, view in zip(shells, pop._gen_parent):
        s._obj_val = list(view._obj_val)                     # rank exactly what the GPU ranked
    ref_pop.rackNstack(shells)                               # geneticalgorithm.py:86-123, unmodified: writes _obj_val_sc
    # (the stored fitnesses come from the arena of parents + children, geneticalgorithm.py:72-84, so
    # they cannot be recomputed from the survivors alone; they must be in arena order)
    fits = pop._gen_parent_fit
    assert all(fits[i] <= fits[i + 1] for i in range(len(fits) - 1))
    for s, view in zip(shells, pop._gen_parent):
        assert [float(v) for v in s._obj_val_sc] == view._obj_val_sc


def test_reference_export_runs_on_facade_views(loaded):
    pop, ga, got, m, p = loaded
    view = pop._gen_parent[0]
    view._mappy, view._pather = m, p                         # live reference objects, as in a user's session
    with ref_shim.stable_argsort():
        wps, lcs, offsets = got.generateWaypointData(view, 1.5, 5, 0.5, prune=False)   # gori_tools.py:393-421
    assert len(wps) == 3
    for a in range(3):
        n = int((view._dna[a] != -1).sum())
        assert wps[a].shape == (n, 4)
        assert np.allclose(wps[a][:, :2], p._XY[view._dna[a][:n]] * view._scale * [-1, 1])
