#!/usr/bin/env python
"""Golden vectors for the RANDOM part of the variation path, produced by the unmodified
reference under fixed np.random seeds: PathMaker.makeMeAPath, the Organism constructor
(mutation / muterpolate / pruneUTurns), Organism.crossover (matchWaypt, padMinDNA) and whole
generations of GeneticalGorithm.runEvolution (lowVarSample, arena) on config C1
(data/map_scaled.png + wilk_3 graph, 6 agents).  np.argsort is forced to kind='stable' while the
reference runs, as for every other fixture.

oracle/variation_oracle.py replays the same seeds through the same numpy calls and must
reproduce every array below EXACTLY (tests/test_oracle_variation.py).

Run from the repo root: python tests/golden/make_golden_variation_rng.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

sys.path.insert(0, HERE)
from variation_cases import EVO_CASES, ORG_CASES, PATH_CASES, XOVER_SEEDS  # noqa: E402


def main():
    ga, mappy_m, pm_m, got = ref_shim.load()
    m, p, (mp, pp, op, gp) = ref_shim.build_world("small", 6)
    out = {}
    with ref_shim.stable_argsort():
        # ---- makeMeAPath -------------------------------------------------------------
        for k, (seed, n, start, prev_len) in enumerate(PATH_CASES):
            np.random.seed(seed)
            prev = None
            if prev_len is not None:
                prev = p.makeMeAPath(prev_len, start)
                out["path%d_prev" % k] = prev
                path = p.makeMeAPath(n, prev[-1], prev)
            else:
                path = p.makeMeAPath(n, start)
            out["path%d" % k] = path
        # ---- Organism constructor ------------------------------------------------------
        for k, (seed, pm, pmu) in enumerate(ORG_CASES):
            o = dict(op)
            o['mutation_prob'], o['muterpolate_prob'] = pm, pmu
            np.random.seed(seed)
            paths = [p.makeMeAPath(gp['starting_path_len'], s) for s in o['start_idx']]
            org = ga.Organism([q.copy() for q in paths], m, p, o)
            out["org%d_in" % k] = np.array(paths)
            out["org%d_dna" % k] = org._dna.copy()
            out["org%d_obj" % k] = np.array(org._obj_val, dtype=np.float64)
        # ---- crossover -------------------------------------------------------------------
        for k, seed in enumerate(XOVER_SEEDS):
            o = dict(op)
            if k == 3:
                o['min_dna_len'] = 170        # padMinDNA must extend short children
            np.random.seed(seed)
            mom = ga.Organism([p.makeMeAPath(150, s) for s in o['start_idx']], m, p, o)
            dad = ga.Organism([p.makeMeAPath(150, s) for s in o['start_idx']], m, p, o)
            kids = mom.crossover(dad)
            out["xo%d_mom" % k], out["xo%d_dad" % k] = mom._dna.copy(), dad._dna.copy()
            out["xo%d_kid0" % k], out["xo%d_kid1" % k] = kids[0]._dna.copy(), kids[1]._dna.copy()
            out["xo%d_obj" % k] = np.array([kids[0]._obj_val, kids[1]._obj_val], dtype=np.float64)
        # ---- whole generations ----------------------------------------------------------
        for k, (seed, G, n_gen) in enumerate(EVO_CASES):
            g2 = dict(gp)
            g2['gen_size'] = G
            g2['coverage_constr_0'] = 0.0       # accept every first-generation candidate (:50)
            g2['org_params'] = dict(op)
            np.random.seed(seed)
            pop = ga.GeneticalGorithm(m, p, g2)
            out["evo%d_dna0" % k] = np.array([o._dna for o in pop._gen_parent])
            out["evo%d_fit0" % k] = np.array(pop._gen_parent_fit, dtype=np.float64)
            dnas, fits, hists = [], [], []
            for _ in range(n_gen):
                h = pop.runEvolution(1)
                dnas.append(np.array([o._dna for o in pop._gen_parent]))
                fits.append(np.array(pop._gen_parent_fit, dtype=np.float64))
                hists.append(np.array(h[0], dtype=np.float64))
            out["evo%d_dna" % k] = np.array(dnas)
            out["evo%d_fit" % k] = np.array(fits)
            out["evo%d_hist" % k] = np.array(hists)
    out = {k: (v.astype(np.int16) if v.dtype.kind == 'i' else v) for k, v in out.items()}
    np.savez_compressed(os.path.join(HERE, "variation_rng.npz"), **out)
    print("variation_rng golden: %d arrays" % len(out))


if __name__ == "__main__":
    main()
