#!/usr/bin/env python
"""Stage times of one generation (CUDA events inside libgcp, gcp_stage_ms) at several population
sizes: python profiles/exp_gen_stages.py 65536 8192"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    from genetic_coverage_planning_b200 import synth, worlds
    from genetic_coverage_planning_b200.engine import FitnessEngine
    from genetic_coverage_planning_b200.distributed import ShardedEvolution
    cfg = dict(worlds.CONFIGS["C3"])
    w, pw = worlds.build_synth_world(cfg['size'], seed=0)
    cst = synth.clustered_starts(w.xy, w.row_ptr, 6, cfg['size'])
    op = worlds.org_params_for(cfg, cst)
    op.update(min_solo_lcs=0, min_comb_lcs=1)
    eng = FitnessEngine(pw, op, num_agents=6)
    eng.set_vary_params(op, w.path_params)
    names = ["path_costs", "eval", "loop_closures", "maximin", "arena_order", "select", "crossover", "mutate"]
    for P in [int(a) for a in sys.argv[1:]] or [65536]:
        dna = synth.random_walk_population_torch(w.xy, w.row_ptr, w.col, cst, P, 150, 200, seed=5, device=eng.device)
        evo = ShardedEvolution(eng, worlds.GEN_PARAMS_6, P)
        evo.load(dna, 0)
        for g in range(40):
            evo.step(4321, g)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for g in range(40, 50):
            evo.step(4321, g)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        eng.set_timing(True)
        acc = {}
        for g in range(50, 55):
            evo.step(4321, g)
            torch.cuda.synchronize()
            for k, n in enumerate(names):
                v = eng.stage_ms(k)
                if v >= 0:
                    acc[n] = acc.get(n, 0.0) + v / 5
        eng.set_timing(False)
        print(json.dumps({"P": P, "ms_per_generation": ms, "stages_ms": acc}), flush=True)
        del evo


if __name__ == "__main__":
    main()
