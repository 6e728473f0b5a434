#!/usr/bin/env python
"""Round-2 probe: which start layout / constraint values give the C3 generation leg a
population that is not ~100 % infeasible (the timed generations ran on all-zero
coverage objectives)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    from genetic_coverage_planning_b200 import synth, worlds
    from genetic_coverage_planning_b200.engine import FitnessEngine
    from genetic_coverage_planning_b200.distributed import ShardedEvolution
    cfg = dict(worlds.CONFIGS["C3"])
    w, pw = worlds.build_synth_world(cfg['size'], seed=0)
    A, L = 6, 200
    d2 = ((w.xy - np.array([2048.0, 2048.0])) ** 2).sum(1)
    deg = np.diff(w.row_ptr)
    order = np.argsort(d2 + np.where(deg >= 3, 0, 1e18), kind='stable')
    layouts = {"one_node": [int(order[0])] * 6,
               "two_nodes": [int(order[k // 3]) for k in range(6)],
               "three_nodes": [int(order[k // 2]) for k in range(6)],
               "spread": synth.spread_starts(w.xy, A, 4096)}
    eng = None
    for name, starts in layouts.items():
        op = worlds.org_params_for(cfg, starts)
        if eng is None:
            eng = FitnessEngine(pw, op, num_agents=A)
        eng.set_vary_params(op, w.path_params)
        dna = synth.random_walk_population_torch(w.xy, w.row_ptr, w.col, starts, 16384, 150, L, seed=3,
                                                 device=eng.device)
        for (ms, mc) in ((2, 5), (1, 3), (0, 1)):
            op2 = dict(op, min_solo_lcs=ms, min_comb_lcs=mc)
            eng.set_params(op2)
            r = eng.evaluate(dna, detail=True, pixel_counts=False)
            feas = float(r.flags[:, 0].double().mean().item())
            cov = float((-r.neg_cov).mean().item())
            for c0, cf in ((0.3, 0.95), (0.0, 0.0)):
                gp = dict(worlds.GEN_PARAMS_6, coverage_constr_0=c0, coverage_constr_f=cf)
                evo = ShardedEvolution(eng, gp, 16384)
                evo.load(dna, 0)
                traj = []
                for g in range(60):
                    evo.step(11, g)
                    if g in (9, 19, 39, 59):
                        traj.append(float((evo.obj[:, 0] != 0).double().mean().item()))
                lens = float((evo.parents != -1).double().sum(-1).mean().item())
                print(json.dumps({"starts": name, "min_solo": ms, "min_comb": mc, "constr": [c0, cf],
                                  "feasible_walks": feas, "mean_cov_walks": cov,
                                  "feasible_after_10_20_40_60": traj, "mean_len_after": lens}), flush=True)
                del evo


if __name__ == "__main__":
    main()
