#!/usr/bin/env python
"""Feasible-rich evaluation fixtures: populations EVOLVED by the unmodified reference
(GeneticalGorithm.runEvolution under the aged coverage constraint) on its two maps, then every
parent re-evaluated by the reference with all intermediates (same record format as
make_golden.py).  Seeded random walks are almost all infeasible (loop-closure constraints,
geneticalgorithm.py:197-200); after a dozen generations the penalty has made the population
feasible, so these fixtures exercise the branch that keeps the coverage objective.

  eval_big_evolved.npz    config C2 world (big map, 6 agents), 24 parents after 14 generations
  eval_small_evolved.npz  config C1 world (small map, 6 agents), 30 parents after 12 generations

Run from the repo root: python tests/golden/make_golden_feasible.py   (needs /root/reference)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from make_golden import VERSIONS, eval_records  # noqa: E402


def main():
    ga, mappy_m, pm_m, got = ref_shim.load()
    for config, G, n_gen, seed in (("big", 24, 14, 5), ("small", 30, 12, 6)):
        m, p, (mp, pp, op, gp) = ref_shim.build_world(config, 6)
        gp = dict(gp)
        gp.update(gen_size=G, coverage_constr_0=0.0, org_params=op)
        with ref_shim.stable_argsort():
            np.random.seed(seed)
            pop = ga.GeneticalGorithm(m, p, gp)
            pop.runEvolution(n_gen)
        dnas = np.array([o._dna for o in pop._gen_parent])
        stored = np.array([[float(v) for v in o._obj_val] for o in pop._gen_parent])
        rec = eval_records(ga, m, p, op, dnas)
        assert np.array_equal(rec['obj'], stored), "re-evaluation differs from the evolved objectives"
        feasible = ~((rec['obj'][:, 0] == 0) & (rec['neg_cov'] != 0))
        np.savez_compressed(os.path.join(HERE, "eval_%s_evolved.npz" % config), dna=dnas.astype(np.int16),
                            feasible=feasible.astype(np.uint8), generations=n_gen, versions=VERSIONS, **rec)
        print(config, "evolved: %d organisms, %d feasible, best coverage %.3f"
              % (len(dnas), int(feasible.sum()), -rec['obj'][:, 0].min()))


if __name__ == "__main__":
    main()
