#!/usr/bin/env python
"""A few launches of the general evaluation (all pixel counters) on C3, for ncu:
    python profiles/general_step.py [P] [reps] [blend]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from genetic_coverage_planning_b200 import synth, worlds
from genetic_coverage_planning_b200.engine import FitnessEngine
P = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
blend = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
cfg = worlds.CONFIGS["C3"]
sw, pw = worlds.build_synth_world(cfg["size"])
starts = synth.spread_starts(sw.xy, cfg["agents"], cfg["size"] // 4)
op = worlds.org_params_for(cfg, starts)
pw.blend = blend
eng = FitnessEngine(pw, op, num_agents=cfg["agents"])
eng.set_vary_params(op, {'path_memory': 10, 'log_scale_weight': 0.7})
dna = eng.random_walks(P, starts, cfg["walk"], seed=0)
for _ in range(reps):
    r = eng.evaluate(dna, detail=True, pixel_counts=True)
torch.cuda.synchronize()
print("ok", float(r.obj.sum().item()))
