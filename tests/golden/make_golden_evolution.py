#!/usr/bin/env python
"""Run the UNMODIFIED reference GA (geneticalgorithm.GeneticalGorithm.runEvolution) on its own
small map (BASELINE config C1: data/map_scaled.png + wilk_3 graph, 6 agents) and record the
per-generation population statistics -> tests/golden/evolution_small.npz.

The GPU facade draws from Philox streams instead of numpy's MT19937, so trajectories cannot be
equal draw for draw; tests/test_gpu_variation.py::test_evolution_matches_reference_statistics
compares these statistics (variation parity is statistical).

  python tests/golden/make_golden_evolution.py      (this container only: needs /root/reference)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

GEN_SIZE, N_GEN, SEEDS = 40, 30, (0, 1, 2)


def stats(pop):
    obj = np.array([o._obj_val for o in pop._gen_parent], dtype=float)
    lens = np.array([[(row != -1).sum() for row in o._dna] for o in pop._gen_parent])
    return [float((obj[:, 0] < 0).mean()), float(-obj[:, 0].min()), float(-obj[:, 0].mean()),
            float(obj[:, 1].mean()), float(lens.mean())]


def main():
    ga, mappy_m, pm_m, got = ref_shim.load()
    m, p, (mp, pp, op, gp) = ref_shim.build_world("small", 6)
    gp = dict(gp)
    gp.update(gen_size=GEN_SIZE, coverage_constr_0=0.0)     # skip the 466 s rejection loop
    gp['org_params'] = op
    out = []
    for seed in SEEDS:
        np.random.seed(seed)
        t0 = time.time()
        pop = ga.GeneticalGorithm(m, p, gp)
        traj = [stats(pop)]
        for g in range(N_GEN):
            pop.runEvolution(1)
            traj.append(stats(pop))
        out.append(traj)
        print("seed %d: %.0fs, final %s" % (seed, time.time() - t0, traj[-1]), flush=True)
    np.savez_compressed(os.path.join(HERE, "evolution_small.npz"), traj=np.array(out),
                        columns=np.array(["frac_feasible", "best_coverage", "mean_coverage",
                                          "mean_travel_cost", "mean_valid_len"]),
                        gen_size=GEN_SIZE, n_gen=N_GEN, seeds=np.array(SEEDS),
                        start_idx=np.array(op['start_idx']), starting_path_len=gp['starting_path_len'])


if __name__ == "__main__":
    main()
