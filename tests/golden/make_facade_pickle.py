#!/usr/bin/env python
"""Runs on a GPU box (no reference needed): evolve a population with the B200 facade on the world
stored in the reference-made legacy pickle (cropped small map) and pickle it the way
got.savePopulation does -> gpurun_out/facade_population.pkl.gz (committed as
tests/golden/facade_population.pkl.gz).  tests/test_reference_consumers.py opens that file in a
container that has the reference but no GPU and lets the unmodified reference read and re-evaluate it.

  gpurun -- python tests/golden/make_facade_pickle.py
"""
import gzip
import os
import pickle
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main():
    from genetic_coverage_planning_b200 import legacy
    from genetic_coverage_planning_b200.geneticalgorithm import GeneticalGorithm
    old = legacy.load_population(os.path.join(HERE, "legacy_population.pkl.gz"))
    first = old._gen_parent[0]
    gp = {'gen_size': 16, 'starting_path_len': 40, 'num_agents': 3, 'gamma': 0.5,
          'coverage_constr_0': 0.0, 'coverage_constr_f': 0.95, 'coverage_aging': 80,
          'org_params': dict(first._org_params)}
    pop = GeneticalGorithm(first._mappy, first._pather, gp, seed=3)
    pop.runEvolution(5)
    raw = pickle.dumps(pop)
    pickle.loads(raw)                                   # round trip on the box
    out = os.path.join(ROOT, "gpurun_out", "facade_population.pkl.gz")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with gzip.open(out, "wb", compresslevel=9) as f:
        f.write(raw)
    print("facade population: P=%d gen %d raw %.1f MB gz %.2f MB" %
          (pop._G_sz, pop._gen_num, len(raw) / 1e6, os.path.getsize(out) / 1e6))
    print([[round(v, 6) for v in o._obj_val] for o in pop._gen_parent[:4]])


if __name__ == "__main__":
    main()
