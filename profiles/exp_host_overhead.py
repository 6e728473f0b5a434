#!/usr/bin/env python
"""Is the generation loop host-bound at small populations?  CPU time to ENQUEUE one generation vs
the GPU time it takes: python profiles/exp_host_overhead.py 8192 65536"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
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
for P in [int(a) for a in sys.argv[1:]] or [8192]:
    dna = synth.random_walk_population_torch(w.xy, w.row_ptr, w.col, cst, P, 150, 200, seed=5, device=eng.device)
    evo = ShardedEvolution(eng, worlds.GEN_PARAMS_6, P)
    evo.load(dna, 0)
    for g in range(40):
        evo.step(4321, g)
    torch.cuda.synchronize()
    n = 100
    l0 = eng.launch_count()
    t0 = time.perf_counter()
    for g in range(40, 40 + n):
        evo.step(4321, g)
    t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(json.dumps({"P": P, "enqueue_ms_per_generation": 1e3 * t_cpu / n, "wall_ms_per_generation": 1e3 * t_all / n,
                      "launches_per_generation": (eng.launch_count() - l0) / n}))
    del evo
