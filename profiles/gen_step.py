#!/usr/bin/env python
"""A few whole generations of the sharded evolution loop on one GPU (for ncu launch lists /
full captures of the variation kernels):  python profiles/gen_step.py [P] [n_gen] [n_pre]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    from genetic_coverage_planning_b200 import synth, worlds
    from genetic_coverage_planning_b200.engine import FitnessEngine
    from genetic_coverage_planning_b200.distributed import ShardedEvolution
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    n_gen = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    n_pre = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    cfg = dict(worlds.CONFIGS["C3"])
    w, pw = worlds.build_synth_world(cfg['size'], seed=0)
    cst = synth.clustered_starts(w.xy, w.row_ptr, 6, cfg['size'])
    op = worlds.org_params_for(cfg, cst)
    op.update(min_solo_lcs=0, min_comb_lcs=1)
    eng = FitnessEngine(pw, op, num_agents=6)
    eng.set_vary_params(op, w.path_params)
    dna = synth.random_walk_population_torch(w.xy, w.row_ptr, w.col, cst, P, 150, 200, seed=5, device=eng.device)
    evo = ShardedEvolution(eng, worlds.GEN_PARAMS_6, P)
    evo.load(dna, 0)
    for g in range(n_pre):
        evo.step(4321, g)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for g in range(n_pre, n_pre + n_gen):
        evo.step(4321, g)
    b.record()
    torch.cuda.synchronize()
    print("P=%d: %.3f ms / generation" % (P, a.elapsed_time(b) / n_gen))


if __name__ == "__main__":
    main()
